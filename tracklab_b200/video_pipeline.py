"""Frame-major device pipeline for one video: frames -> detector rows -> (ReID features) -> tracker rows.

This is the B200 shape of OfflineTrackingEngine.video_loop (/root/reference/tracklab/engine/offline.py:10-35) for the
detect -> ReID -> associate stream (north_star): the reference runs the video module-major (detector over all frames, then
ReID over all detections, then the tracker frame by frame, each through DataFrames and DataLoader workers); here the video
is consumed in batches of frames and the three stages of consecutive batches overlap on three CUDA streams:

    s_copy : H2D copy of batch k+1 (frames in pinned host memory)         [only when streaming from the host]
    s_det  : letterbox -> YOLOX (CUDA graph) -> decode/NMS -> tk_pack_detections_ex (rows appended at a device cursor)
    s_reid : tk_crop_resize_norm over the rows of batch k -> ReID backbone (CUDA graph per crop bucket) -> features
    s_trk  : ONE whole-batch tracker launch (tk_bytetrack_run / tk_ocsort_run / tk_strongsort_run / tk_bpbreid_run)

The tracker ALWAYS consumes the detector's own rows (``det.dets`` / ``det.offsets``, written by tk_pack_detections_ex) unless
the caller injects rows explicitly (oracle-detector mode of SURVEY.md Appendix C). The only host involvement inside a video
is one 8-byte read of the detector's row cursor per batch, needed to size the ReID launch; it is taken one batch late (the
detector of batch k+1 is already enqueued), so the device never idles on it. DataFrames are materialised once per video by the
module layer (tracklab_b200/modules.py), not per frame.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .detector import YoloxDetectorDevice, synth_weights_path
from .device_trackers import BpbreidStrongSortDevice, ByteTrackDevice, OCSortDevice, StrongSortDevice


@dataclass
class VideoResult:
    """Device-side result of one video (nothing synchronised)."""
    out_rows: torch.Tensor        # float64 [cap, cols]
    out_fc: torch.Tensor          # int32 [1, F] rows per frame
    out_count: torch.Tensor       # int32 [1]
    det_cursor: torch.Tensor      # int32 [2] = {detector rows, frames}
    n_frames: int
    features: torch.Tensor | None = None


@dataclass
class HostResult:
    rows: np.ndarray              # float64 [R, cols]
    frame: np.ndarray             # int32 [R] frame index of every row
    det_rows: int                 # rows the detector emitted
    det_table: np.ndarray | None = None   # float64 [N,7] detector rows (only when asked for)
    det_offsets: np.ndarray | None = None


class DetectReidTrackPipeline:
    """detector (+ ReID) + tracker of one video stream; see the module docstring for the schedule."""

    def __init__(self, detector: YoloxDetectorDevice, tracker, batch: int, reid=None, rows_cap: int | None = None):
        self.det, self.trk, self.reid, self.batch = detector, tracker, reid, batch
        self.dev = detector.device
        self.kind = {ByteTrackDevice: "bytetrack", OCSortDevice: "ocsort", StrongSortDevice: "strongsort",
                     BpbreidStrongSortDevice: "bpbreid"}[type(tracker)]
        if self.kind in ("strongsort", "bpbreid") and reid is None:
            raise _lib.TrackKernError(f"{self.kind} needs a ReID stage")
        if self.kind == "bpbreid" and not detector.rows_ltwh:
            raise _lib.TrackKernError("the part-based tracker reads [l,t,w,h] rows: build the detector with rows_ltwh=True")
        if self.kind != "bpbreid" and detector.rows_ltwh:
            raise _lib.TrackKernError(f"{self.kind} reads [l,t,r,b] rows: build the detector with rows_ltwh=False")
        self.s_copy = torch.cuda.Stream(device=self.dev)
        self.s_det = torch.cuda.Stream(device=self.dev)
        self.s_reid = torch.cuda.Stream(device=self.dev) if reid is not None else None
        self.s_trk = torch.cuda.Stream(device=self.dev, priority=-1)   # short latency-bound launches: do not queue behind convolutions
        self.stage = None
        self.kernel_events = []  # (name, start, end, frames) CUDA events recorded on the launching stream
        self.launches = 0
        self.rows_cap = rows_cap or detector.dets.shape[0]
        self.feats = None
        self.vis = None
        if reid is not None:
            E = reid.feature_dim
            self.feats = torch.zeros((self.rows_cap, E), dtype=torch.float32, device=self.dev)
            if self.kind == "bpbreid":
                self.vis = torch.ones((self.rows_cap, tracker.n_parts), dtype=torch.float32, device=self.dev)
        self._host_cursor = None

    # ---- scheduling helpers ------------------------------------------------------------------------
    def _schedule(self, F: int, host: bool):
        """Frame ranges of the detector batches. Streaming from the host the video is PCIe-bound, so what is left after the
        last frame has arrived (one detector batch + its ReID + its tracker chunk) is pure tail: the last full batch is drained
        as a few halving batches (each replaying its own CUDA graph) so that only a small one remains after the last copy."""
        B = self.batch
        bounds = [(f0, min(F, f0 + B)) for f0 in range(0, F, B)]
        if not host or len(bounds) < 2 or bounds[-1][1] - bounds[-1][0] != B or B < 32:
            return bounds
        f0, f1 = bounds.pop()
        rem, sizes = B, []
        while rem > 12:
            s = rem // 2
            sizes.append(rem - s)
            rem = s
        sizes.append(rem)
        self.det.tail_sizes.update(sizes)
        for s in sizes:
            bounds.append((f0, f0 + s))
            f0 += s
        return bounds

    def _staging(self, like: torch.Tensor, slots: int):
        shape = (slots, self.batch) + tuple(like.shape[1:])
        if self.stage is None or self.stage.shape != shape:
            self.stage = torch.empty(shape, dtype=torch.uint8, device=self.dev)
        return self.stage

    # ---- one video -----------------------------------------------------------------------------------
    @torch.no_grad()
    def run_video(self, frames: torch.Tensor, tracker_dets: torch.Tensor | None = None,
                  tracker_offsets: torch.Tensor | None = None, out_rows: torch.Tensor | None = None,
                  time_kernels: bool = False) -> VideoResult:
        """frames uint8 [F,H,W,3], on the device or in pinned host memory.

        tracker_dets/tracker_offsets (device, float64[N,7] / int32[F+1]): oracle-detector mode — rows the ReID stage and the
        tracker consume INSTEAD of the detector's (the detector still runs). Default (None): the connected chain, the tracker
        consumes the rows the detector appended at its device cursor."""
        F = frames.shape[0]
        host = not frames.is_cuda
        det, trk, reid = self.det, self.trk, self.reid
        cur = torch.cuda.current_stream(self.dev)
        streams = [s for s in (self.s_copy, self.s_det, self.s_reid, self.s_trk) if s is not None]
        for s in streams:
            s.wait_stream(cur)
        with torch.cuda.stream(self.s_det):
            det.reset()
        with torch.cuda.stream(self.s_trk):
            trk.reset()
        own = tracker_dets is None
        if not own and reid is not None and tracker_dets.shape[0] > self.rows_cap:
            raise _lib.TrackKernError(f"{tracker_dets.shape[0]} injected rows exceed rows_cap={self.rows_cap}")
        t_dets = det.dets if own else tracker_dets
        t_offs = det.offsets if own else tracker_offsets
        cols = trk.COLS if self.kind == "bpbreid" else 8
        if out_rows is None:
            cap = (2 if self.kind == "strongsort" else 1) * t_dets.shape[0]
            out_rows = torch.empty((cap, cols), dtype=torch.float64, device=self.dev)
        out_start = torch.zeros(1, dtype=torch.int32, device=self.dev)
        out_count = torch.zeros(1, dtype=torch.int32, device=self.dev)
        out_fc = torch.zeros((1, F), dtype=torch.int32, device=self.dev)
        n_slots = 3 if reid is not None else 2       # a staging slot is busy until the crops of its batch are gathered
        stage = self._staging(frames, n_slots) if host else None
        bounds = self._schedule(F, host)
        if reid is not None and own:
            if self._host_cursor is None or self._host_cursor.shape[0] < len(bounds) + 1:
                self._host_cursor = torch.zeros((len(bounds) + 1, 2), dtype=torch.int32, pin_memory=True)
        free_ev = [None] * n_slots
        pending = None     # (index, f0, f1, batch frames, detector-done event, cursor event)
        row0_host = 0
        tgen_offs = tracker_offsets.cpu().numpy() if (reid is not None and not own) else None

        def finish(p):
            """ReID + tracker of a batch whose detector stage is already enqueued (and, own rows, whose row count is read)."""
            nonlocal row0_host
            i, f0, f1, batch, det_done, cur_ev = p
            n = f1 - f0
            slot_done = det_done
            if reid is not None:
                if own:
                    cur_ev.synchronize()                   # 8-byte cursor of batch i (batch i+1 is already running)
                    r0, r1 = row0_host, int(self._host_cursor[i, 0])
                    row0_host = r1
                    fidx = det.frame_of_row
                else:
                    r0, r1 = int(tgen_offs[f0]), int(tgen_offs[f1])
                    fidx = self._frame_index(tracker_offsets, f0, f1, r0, r1)
                if r1 > self.rows_cap:
                    raise _lib.TrackKernError(f"{r1} detector rows exceed rows_cap={self.rows_cap}")
                self.s_reid.wait_event(det_done)
                with torch.cuda.stream(self.s_reid):
                    if r1 > r0:
                        if time_kernels:
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record(self.s_reid)
                        self._reid_rows(batch, t_dets, fidx, r0, r1)
                        if time_kernels:
                            e1.record(self.s_reid)
                            self.kernel_events.append(("reid_stage", e0, e1, r1 - r0))
                    slot_done = torch.cuda.Event()
                    slot_done.record(self.s_reid)
                self.s_trk.wait_event(slot_done)
            else:
                self.s_trk.wait_event(det_done)   # tracker batch k reads the rows detector batch k wrote
            if host:
                free_ev[i % n_slots] = slot_done
            with torch.cuda.stream(self.s_trk):
                offs = t_offs[f0:f1 + 1].unsqueeze(0)
                if time_kernels:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(self.s_trk)
                nr = t_dets.shape[0]
                if self.kind == "strongsort":
                    _, fc, _ = trk.run(t_dets, offs, self.feats[:nr], out_rows=out_rows, out_start=out_start, out_count=out_count)
                elif self.kind == "bpbreid":
                    _, fc, _ = trk.run(t_dets, offs, self.feats[:nr].view(nr, trk.n_parts, trk.feature_dim), self.vis[:nr],
                                       out_rows=out_rows, out_start=out_start, out_count=out_count)
                else:
                    _, fc, _ = trk.run(t_dets, offs, out_rows=out_rows, out_start=out_start, out_count=out_count)
                if time_kernels:
                    e1.record(self.s_trk)
                    self.kernel_events.append((self.kind + "_video_kernel", e0, e1, n))
                out_fc[:, f0:f1] = fc
            self.launches += 1

        for i, (f0, f1) in enumerate(bounds):
            n = f1 - f0
            if host:
                slot = i % n_slots
                with torch.cuda.stream(self.s_copy):
                    if free_ev[slot] is not None:
                        self.s_copy.wait_event(free_ev[slot])
                    stage[slot, :n].copy_(frames[f0:f1], non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(self.s_copy)
                self.s_det.wait_event(ready)
                batch = stage[slot, :n]
            else:
                batch = frames[f0:f1]
            with torch.cuda.stream(self.s_det):
                if time_kernels:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(self.s_det)
                det.detect_batch(batch)
                if time_kernels:
                    e1.record(self.s_det)
                    self.kernel_events.append(("detect_batch", e0, e1, n))
                done = torch.cuda.Event()
                done.record(self.s_det)
                cur_ev = None
                if reid is not None and own:
                    self._host_cursor[i].copy_(det.cursor, non_blocking=True)
                    cur_ev = torch.cuda.Event()
                    cur_ev.record(self.s_det)
            self.launches += 1 + det.launches_per_batch   # letterbox + epilogues/nms/pack of the (graph-replayed) batch; convolutions are cuDNN
            if pending is not None:
                finish(pending)
            pending = (i, f0, f1, batch, done, cur_ev)
            if reid is None:          # nothing to wait for on the host: chain the tracker immediately
                finish(pending)
                pending = None
        if pending is not None:
            finish(pending)
        for s in streams:
            cur.wait_stream(s)
        return VideoResult(out_rows, out_fc, out_count, det.cursor, F, self.feats)

    # ---- ReID of the rows [r0, r1) of one batch ------------------------------------------------------
    def _frame_index(self, offsets, f0, f1, r0, r1):
        """Oracle-detector mode: batch-local image index of the injected rows r0..r1 (the detector's own rows carry it already)."""
        idx = torch.searchsorted(offsets[f0 + 1:f1 + 1].contiguous(), torch.arange(r0, r1, device=self.dev, dtype=torch.int32), right=True)
        full = torch.zeros(r1, dtype=torch.int32, device=self.dev)
        full[r0:r1] = idx.to(torch.int32)
        return full

    def _reid_rows(self, batch_frames, dets, frame_of_row, r0, r1):
        reid = self.reid
        n0 = kernel_launches()
        if self.kind == "bpbreid":
            out = reid.features(batch_frames, dets[r0:r1], frame_of_row[r0:r1], ltwh_rows=True)
        else:
            out = reid.features(batch_frames, dets[r0:r1], frame_of_row[r0:r1])
        self.feats[r0:r1] = out
        self.launches += kernel_launches() - n0

    # ---- results --------------------------------------------------------------------------------------
    def check_status(self):
        self.det.check_status()
        self.trk.check_status()

    def results_to_host(self, res: VideoResult, with_detections: bool = False) -> HostResult:
        """One D2H read of the video's result rows (what the module layer turns into DataFrame columns)."""
        n = int(res.out_count.item())
        rows = res.out_rows[:n].cpu().numpy()
        fc = res.out_fc[0].cpu().numpy()
        cur = res.det_cursor.cpu().numpy()
        out = HostResult(rows, np.repeat(np.arange(len(fc), dtype=np.int32), fc), int(cur[0]))
        if with_detections:
            out.det_table = self.det.dets[:int(cur[0])].cpu().numpy()
            out.det_offsets = self.det.offsets[:res.n_frames + 1].cpu().numpy()
        return out


def kernel_launches() -> int:
    from . import kernels
    return kernels.LAUNCHES


DetectTrackPipeline = DetectReidTrackPipeline   # round-1 name (detector + IoU tracker, no ReID stage)


# ---- BASELINE.json configurations ----------------------------------------------------------------------------------
CONFIGS = {
    # configs[1]: YOLOX-s + IoU-only ByteTrack (byte_track.yaml)
    "config2": dict(variant="s", tracker="bytetrack", reid=None,
                    hyper=dict(track_thresh=0.6, match_thresh=0.8, track_buffer=30, frame_rate=30)),
    # configs[2]: YOLOX-m + ResNet-50 ReID (256x128 crops, 2048-d) + cosine/IoU association with Kalman gating (strong_sort.yaml)
    "config3": dict(variant="m", tracker="strongsort", reid="resnet50",
                    hyper=dict(max_dist=0.1594374041012136, max_iou_dist=0.5431835667667874, max_age=40, max_unmatched_preds=0,
                               n_init=3, nn_budget=100, mc_lambda=0.995, ema_alpha=0.8962157769329083)),
    # configs[2], part-based flavour: the ResNet-50 feature exposed as embeddings [D,1,2048] / visibility 1 (bpbreid_strong_sort.yaml)
    "config3_bpbreid": dict(variant="m", tracker="bpbreid", reid="resnet50",
                            hyper=dict(ema_alpha=0.9, mc_lambda=0.995, max_dist=0.5, max_iou_distance=0.8, max_age=300, n_init=0,
                                       min_bbox_confidence=0.0, max_kalman_prediction_without_update=7)),
    "config2_ocsort": dict(variant="s", tracker="ocsort", reid=None,
                           hyper=dict(det_thresh=0.0, max_age=50, min_hits=1, iou_threshold=0.22136877277096445, delta_t=1,
                                      asso_func="giou", inertia=0.3941737016672115, use_byte=False)),
}


def build_pipeline(config: str = "config3", device="cuda:0", batch: int = 20, frames_cap: int = 512, image_size=(1920, 1080),
                   weights="auto", reid_precision: str = "bf16", min_confidence: float = 0.4, max_per_frame: int = 96,
                   ctas_per_video: int = 32, detector_kwargs=None, reid_model=None, detector_model=None,
                   use_graphs: bool = True) -> DetectReidTrackPipeline:
    """Assemble the product pipeline of one BASELINE configuration. ``weights``: "auto" = weights/yolox_<variant>_synth.pt
    when present (the detector trained on the synthetic generator), else seeded random weights with calibrated heads;
    None = seeded; a path = that file (missing -> error). ``use_graphs=False``: eager launches of the same kernels (for ncu launch
    lists; the timed bench always runs the CUDA graphs)."""
    cfg = CONFIGS[config]
    if not use_graphs:
        detector_kwargs = dict(detector_kwargs or {}, use_graph=False)
    dev = torch.device(device)
    rows_cap = frames_cap * max_per_frame
    w = synth_weights_path(cfg["variant"]) if weights == "auto" else weights
    det = YoloxDetectorDevice(cfg["variant"], device=dev, batch=batch, frames_cap=frames_cap, dets_cap=rows_cap,
                              weights=w if detector_model is None else None, model=detector_model,
                              rows_ltwh=(cfg["tracker"] == "bpbreid"), **(detector_kwargs or {}))
    reid = None
    if cfg["reid"] is not None:
        from .reid import ReidStageDevice
        reid = ReidStageDevice(device=dev, arch=cfg["reid"], precision=reid_precision, model=reid_model, use_graphs=use_graphs)
    W, H = image_size
    if cfg["tracker"] == "bytetrack":
        trk = ByteTrackDevice(**cfg["hyper"], min_confidence=min_confidence, cap_tracks=256, cap_dets=128, device=dev)
    elif cfg["tracker"] == "ocsort":
        trk = OCSortDevice(**cfg["hyper"], min_confidence=min_confidence, cap_tracks=256, cap_dets=128, device=dev)
    elif cfg["tracker"] == "strongsort":
        trk = StrongSortDevice(reid.feature_dim, **cfg["hyper"], min_confidence=min_confidence, image_size=(W, H),
                               ctas_per_video=ctas_per_video, cap_tracks=160, cap_dets=128, device=dev)
    else:
        trk = BpbreidStrongSortDevice(1, reid.feature_dim, **cfg["hyper"], ctas_per_video=ctas_per_video, cap_tracks=1024,
                                      cap_dets=128, device=dev)
    pipe = DetectReidTrackPipeline(det, trk, batch, reid=reid, rows_cap=rows_cap)
    pipe.config = config
    return pipe
