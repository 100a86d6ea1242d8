"""Frame-major device pipeline for one video: frames -> detector rows -> tracker rows.

This is the B200 shape of OfflineTrackingEngine.video_loop (/root/reference/tracklab/engine/offline.py:10-35)
for the detect -> associate path: frames are consumed in batches (from HBM, or from pinned host memory with
the H2D copy overlapped on its own stream), the detector stage runs as one CUDA-graph replay per batch and
the tracker consumes each batch's rows with ONE kernel launch per batch on a second stream, chained by
events — the host never waits inside a video. DataFrames are materialised once per video by the module
layer (tracklab_b200/modules.py), not per frame.
"""
from __future__ import annotations

import numpy as np
import torch

from .detector import YoloxDetectorDevice
from .device_trackers import ByteTrackDevice


class DetectTrackPipeline:
    def __init__(self, detector: YoloxDetectorDevice, tracker: ByteTrackDevice, batch: int):
        self.det, self.trk, self.batch = detector, tracker, batch
        self.dev = detector.device
        self.s_copy = torch.cuda.Stream(device=self.dev)
        self.s_det = torch.cuda.Stream(device=self.dev)
        self.s_trk = torch.cuda.Stream(device=self.dev)
        self.stage = None
        self.kernel_events = []  # (name, start, end) CUDA events recorded on the launching stream
        self.launches = 0

    def _schedule(self, F: int, host: bool):
        """Frame ranges of the detector batches. Streaming from the host the video is PCIe-bound, so what is left after the
        last frame has arrived (one detector batch + its tracker chunk) is pure tail: the last full batch is drained as a few
        halving batches (each replaying its own CUDA graph) so that only a small one remains after the last copy."""
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

    def _staging(self, like: torch.Tensor):
        shape = (2, self.batch) + tuple(like.shape[1:])
        if self.stage is None or self.stage.shape != shape:
            self.stage = torch.empty(shape, dtype=torch.uint8, device=self.dev)
        return self.stage

    @torch.no_grad()
    def run_video(self, frames: torch.Tensor, tracker_dets: torch.Tensor | None = None,
                  tracker_offsets: torch.Tensor | None = None, out_rows: torch.Tensor | None = None,
                  time_kernels: bool = False):
        """frames uint8 [F,H,W,3], on the device or in pinned host memory.

        tracker_dets/tracker_offsets (device, float64[N,7] / int32[F+1]): rows the tracker consumes. When None
        the tracker consumes the detector's own rows. Returns device tensors
        (out_rows float64[.,8], out_frame_count int32[F], out_count int32[1], det_cursor int32[2])."""
        F = frames.shape[0]
        host = not frames.is_cuda
        B = self.batch
        det, trk = self.det, self.trk
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_copy, self.s_det, self.s_trk):
            s.wait_stream(cur)
        with torch.cuda.stream(self.s_det):
            det.reset()
        with torch.cuda.stream(self.s_trk):
            trk.reset()
        own = tracker_dets is None
        t_dets = det.dets if own else tracker_dets
        t_offs = det.offsets if own else tracker_offsets
        if out_rows is None:
            out_rows = torch.empty((t_dets.shape[0], 8), dtype=torch.float64, device=self.dev)
        out_start = torch.zeros(1, dtype=torch.int32, device=self.dev)
        out_count = torch.zeros(1, dtype=torch.int32, device=self.dev)
        out_fc = torch.zeros((1, F), dtype=torch.int32, device=self.dev)
        stage = self._staging(frames) if host else None
        free_ev = [None, None]
        for i, (f0, f1) in enumerate(self._schedule(F, host)):
            n = f1 - f0
            if host:
                slot = i & 1
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
                if host:
                    free_ev[i & 1] = done
            self.launches += 1 + det.launches_per_batch   # letterbox + epilogues/nms/pack of the (graph-replayed) batch; convolutions are cuDNN
            self.s_trk.wait_event(done)   # tracker batch k depends on detector batch k (true data dependency when own=True)
            with torch.cuda.stream(self.s_trk):
                offs = t_offs[f0:f1 + 1].unsqueeze(0)
                if time_kernels:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(self.s_trk)
                _, fc, _ = trk.run(t_dets, offs, out_rows=out_rows, out_start=out_start, out_count=out_count)
                if time_kernels:
                    e1.record(self.s_trk)
                    self.kernel_events.append(("bytetrack_video_kernel", e0, e1, n))
                out_fc[:, f0:f1] = fc
            self.launches += 1
        cur.wait_stream(self.s_det)
        cur.wait_stream(self.s_trk)
        return out_rows, out_fc, out_count, det.cursor

    def results_to_host(self, out_rows, out_fc, out_count):
        """One D2H read of the video's result rows (what the module layer turns into DataFrame columns)."""
        n = int(out_count.item())
        rows = out_rows[:n].cpu().numpy()
        fc = out_fc[0].cpu().numpy()
        return rows, np.repeat(np.arange(len(fc), dtype=np.int32), fc)
