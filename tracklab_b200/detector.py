"""Device-resident YOLOX detector stage: frames (uint8, HBM) -> tracker-input rows (float64, HBM).

letterbox (tk_letterbox_u8) -> YOLOX-s/m in PyTorch bf16 channels-last under a CUDA graph
-> decode + NMS (tk_yolox_nms) -> wrapper row packing (tk_pack_detections).
Stands in for RTMLibDetector.process (/root/reference/tracklab/wrappers/bbox_detector/rtmlib_api.py:27-46)
for a whole batch of frames at once; nothing returns to the host until the video (or chunk) is done.
"""
from __future__ import annotations

import math
import os

import torch

from . import _lib, kernels
from .nets.yolox import build_yolox


WEIGHTS_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights")


def synth_weights_path(variant: str):
    """weights/yolox_<variant>_synth.pt (tools/train_synth_detector.py: the restated YOLOX trained on the synthetic generator so
    that it localises the synthetic targets) or None when the file is not there."""
    p = os.path.join(WEIGHTS_DIR, f"yolox_{variant}_synth.pt")
    return p if os.path.isfile(p) else None


def load_yolox_weights(variant: str, path, num_classes: int = 1):
    """state_dict file of tracklab_b200.nets.yolox.YOLOX -> module. A path that does not exist is an error (never a silent
    fall back to random weights)."""
    if not os.path.isfile(str(path)):
        raise _lib.TrackKernError(f"detector weights {path!r} not found")
    blob = torch.load(str(path), map_location="cpu")
    sd = blob.get("state_dict", blob)
    m = build_yolox(variant, num_classes, 0)
    m.load_state_dict({k: v.float() for k, v in sd.items()})
    return m.eval()


class YoloxDetectorDevice:
    def __init__(self, variant="s", device="cuda:0", batch=32, input_size=640, dtype=torch.bfloat16,
                 score_thr=0.7, nms_thr=0.45, max_per_image=256, num_classes=1, seed=1234,
                 frames_cap=4096, dets_cap=1 << 18, use_graph=True, model=None, fused=True, weights=None, rows_ltwh=False):
        if not torch.cuda.is_available():
            raise _lib.TrackKernError("YoloxDetectorDevice needs a CUDA device (no CPU path)")
        _lib.load()
        self.device = torch.device(device)
        self.batch, self.size, self.dtype = batch, input_size, dtype
        self.score_thr, self.nms_thr, self.max_per_image = score_thr, nms_thr, max_per_image
        if model is None and weights is not None:
            model = load_yolox_weights(variant, weights, num_classes)
        self.trained = weights is not None
        self.rows_ltwh = bool(rows_ltwh)
        self.model = (model if model is not None else build_yolox(variant, num_classes, seed))
        self.model = self.model.to(self.device).to(dtype).to(memory_format=torch.channels_last).eval()
        for p in self.model.parameters():
            p.requires_grad_(False)
        self.fused = None
        self.use_tc = os.environ.get("TK_NO_TC", "0") != "1"     # 1x1 layers on the hand-written tcgen05 GEMM (A/B switch for profiling)
        self.use_fused = bool(fused) and dtype == torch.bfloat16
        if self.use_fused:
            from .nets.yolox_fused import YoloxFused
            self._fused_cls = YoloxFused
            self.fused = YoloxFused(self.model, self.device, use_tc=self.use_tc)
            # Focus-unfolded input written by the letterbox kernel; channels 12.. are zero padding (32-channel pitch, see yolox_fused)
            self.x = torch.zeros((batch, self._fused_cls.STEM_IN, input_size // 2, input_size // 2), dtype=dtype,
                                 device=self.device).contiguous(memory_format=torch.channels_last)
        else:
            self.x = torch.empty((batch, 3, input_size, input_size), dtype=dtype, device=self.device,
                                 memory_format=torch.channels_last)
        self.status = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self.cursor = torch.zeros((2,), dtype=torch.int32, device=self.device)      # {next row, next frame}
        self.dets = torch.zeros((dets_cap, 7), dtype=torch.float64, device=self.device)
        self.offsets = torch.zeros((frames_cap + 1,), dtype=torch.int32, device=self.device)
        self.frame_of_row = torch.zeros((dets_cap,), dtype=torch.int32, device=self.device)   # batch-local image index of every row
        self.ratio = None
        self.geom = None
        self.graph = None
        self.use_graph = use_graph
        self.tail_sizes = set()     # batch sizes < batch that get their own graph (set by the streaming pipeline's drain schedule)
        self._tail_graphs = {}
        self.pred = None
        self.nms_out = None
        self.time_kernels = False
        self.launches_per_batch = 0
        self.kernel_events = []
        self.variant = variant
        torch.backends.cudnn.benchmark = True

    # ---- synthetic-weight calibration ------------------------------------------------------------
    @torch.no_grad()
    def calibrate(self, frames: torch.Tensor, target_per_image: float = 120.0, spread: float = 2.5):
        """Random-init heads emit near-constant logits, so nothing (or everything) clears score_thr.
        Rescale the obj/cls prediction convolutions per pyramid level to unit-variance logits times
        ``spread`` and pick one bias so that about ``target_per_image`` anchors pass the threshold — this
        gives decode+NMS a realistic candidate load with synthetic weights."""
        x, _ = kernels.letterbox(frames[: self.batch], self.size, self.dtype, swap_rb=True, channels_last=True)
        raw = self.model(x).float()   # calibration statistics come from the plain module (same weights)
        A = raw.shape[1]
        n8, n16 = (self.size // 8) ** 2, (self.size // 16) ** 2
        bounds = [(0, n8), (n8, n8 + n16), (n8 + n16, A)]
        for lvl, (a0, a1) in enumerate(bounds):
            for conv, col in ((self.model.obj_preds[lvl], slice(4, 5)), (self.model.cls_preds[lvl], slice(5, None))):
                v = raw[:, a0:a1, col]
                mu, sd = v.mean().item(), max(v.std().item(), 1e-6)
                w = conv.weight.data.float() * (spread / sd)
                b = (conv.bias.data.float() - mu) * (spread / sd)
                conv.weight.data.copy_(w.to(conv.weight.dtype))
                conv.bias.data.copy_(b.to(conv.bias.dtype))
        raw = self.model(x).float()
        lo, hi = -12.0, 6.0
        for _ in range(30):
            mid = 0.5 * (lo + hi)
            sc = torch.sigmoid(raw[..., 4] + mid)[..., None] * torch.sigmoid(raw[..., 5:] + mid)
            n = (sc > self.score_thr).sum().item() / raw.shape[0]
            if n > target_per_image:
                hi = mid
            else:
                lo = mid
        shift = 0.5 * (lo + hi)
        for lvl in range(3):
            self.model.obj_preds[lvl].bias.data += shift
            self.model.cls_preds[lvl].bias.data += shift
        self.graph = None
        self._tail_graphs = {}
        if self.use_fused:
            self.fused = self._fused_cls(self.model, self.device, use_tc=self.use_tc)
        return shift

    # ---- one batch ---------------------------------------------------------------------------------
    def _net(self, x):
        return self.fused(x) if self.use_fused else self.model(x)

    def _forward_post(self, W, H):
        n0 = kernels.LAUNCHES
        self._forward_post_impl(W, H)
        self.launches_per_batch = kernels.LAUNCHES - n0   # libtrackkern launches of one batch (replayed by the CUDA graph)

    def _forward_post_impl(self, W, H, x=None):
        self.pred = self._net(self.x if x is None else x)
        self.nms_out = kernels.yolox_nms(self.pred, self.ratio, self.size, logits=True, score_thr=self.score_thr,
                                         nms_thr=self.nms_thr, max_out=self.max_per_image, status=self.status)
        boxes, scores, cls, count, _ = self.nms_out
        kernels.pack_detections(boxes, scores, cls, count, W, H, self.cursor, self.dets, self.offsets, self.status,
                                keep_class=0, fixed_conf=1.0, category_id=1.0, ltwh=self.rows_ltwh, frame_of_row=self.frame_of_row)

    def reset(self):
        self.cursor.zero_()
        self.status.zero_()

    @torch.no_grad()
    def detect_batch(self, frames: torch.Tensor):
        """frames uint8 [B<=batch, H, W, 3] on the device; appends rows/offsets at the device cursor."""
        B, H, W, _ = frames.shape
        assert B <= self.batch
        if B < self.batch and self.use_graph and self.use_fused and B in self.tail_sizes:
            # pipeline drain (video_pipeline.py): a few smaller batches at the end of a streamed video, each with its own graph
            ent = self._tail_graphs.get((B, H, W))
            if ent is None:
                x = torch.zeros((B, self._fused_cls.STEM_IN, self.size // 2, self.size // 2), dtype=self.dtype,
                                device=self.device).contiguous(memory_format=torch.channels_last)
                _, self.ratio = kernels.letterbox(frames, self.size, self.dtype, swap_rb=True, out=x, focus16=True)
                # the warm-up / capture runs append at the live cursor and may trip the sticky capacity bit near the end of
                # the buffers: cursor AND status are restored afterwards (each run restarts at the saved cursor)
                cur, st = self.cursor.clone(), self.status.clone()
                s = torch.cuda.Stream(device=self.device)
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    for _ in range(3):
                        self.cursor.copy_(cur)
                        self._forward_post_impl(W, H, x)
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
                self.cursor.copy_(cur)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._forward_post_impl(W, H, x)
                self.cursor.copy_(cur); self.status.copy_(st)
                ent = (g, x)
                self._tail_graphs[(B, H, W)] = ent
            g, x = ent
            _, self.ratio = kernels.letterbox(frames, self.size, self.dtype, swap_rb=True, out=x, focus16=True)
            g.replay()
            return
        if B < self.batch:   # ragged tail: run eagerly on a view (rare: once per video)
            if self.use_fused:
                x = torch.zeros((B, self._fused_cls.STEM_IN, self.size // 2, self.size // 2), dtype=self.dtype,
                                device=self.device).contiguous(memory_format=torch.channels_last)
                _, ratio = kernels.letterbox(frames, self.size, self.dtype, swap_rb=True, out=x, focus16=True)
            else:
                x, ratio = kernels.letterbox(frames, self.size, self.dtype, swap_rb=True, channels_last=True)
            pred = self._net(x)
            boxes, scores, cls, count, _ = kernels.yolox_nms(pred, ratio, self.size, True, self.score_thr, self.nms_thr,
                                                             self.max_per_image, status=self.status)
            kernels.pack_detections(boxes, scores, cls, count, W, H, self.cursor, self.dets, self.offsets, self.status,
                                    ltwh=self.rows_ltwh, frame_of_row=self.frame_of_row)
            return
        if self.time_kernels:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _, self.ratio = kernels.letterbox(frames, self.size, self.dtype, swap_rb=True, out=self.x, focus16=self.use_fused)
        if self.time_kernels:
            e1.record()
            self.kernel_events.append(("letterbox_kernel", e0, e1, B))
        if not self.use_graph:
            self._forward_post(W, H)
            return
        if self.graph is None or self.geom != (H, W):
            self.geom = (H, W)
            cur, st = self.cursor.clone(), self.status.clone()
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):   # warm-up (cuDNN autotune) outside capture; every run restarts at the saved cursor
                    self.cursor.copy_(cur)
                    self._forward_post(W, H)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.cursor.copy_(cur)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._forward_post(W, H)
            self.cursor.copy_(cur); self.status.copy_(st)
        self.graph.replay()

    def detect_into(self, frames: torch.Tensor) -> int:
        """Convenience (tests/smoke): reset, detect one batch, synchronise, return the number of rows."""
        self.reset()
        self.detect_batch(frames)
        torch.cuda.synchronize()
        self.check_status()
        return int(self.cursor[0].item())

    def check_status(self):
        st = int(self.status.item())
        if st:
            raise _lib.TrackKernError("detector post-processing: " + _lib.status_text(st))

    @property
    def flops_per_frame(self) -> float:
        return {"s": 26.8e9, "m": 73.8e9}.get(getattr(self, "variant", "s"), math.nan)
