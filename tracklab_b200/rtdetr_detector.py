"""Device-resident RT-DETR detector stage: frames (uint8, HBM) -> per-image detection rows (float64, HBM).

tk_resize_frames_u8 (Pillow-exact 640x640 resize + 1/255 rescale) -> transformers RTDetrForObjectDetection on the GPU
(fp32, or bf16 autocast) -> tk_rtdetr_decode (sigmoid, top-Q, box decode, threshold, class filter, sanitise, ltwh).
Stands in for RTDetr.process (/root/reference/tracklab/wrappers/bbox_detector/transformers_api.py:31-54) for a whole
batch of frames; nothing returns to the host until the caller reads the rows.
"""
from __future__ import annotations

import torch

from . import _lib, kernels
from .nets.rtdetr import build_rtdetr, calibrate_person_bias


class RTDetrDetectorDevice:
    def __init__(self, device="cuda:0", min_confidence=0.4, precision="bf16", model=None, seed=1234, input_size=640,
                 keep_label=0, num_labels=80, use_graphs=True):
        if not torch.cuda.is_available():
            raise _lib.TrackKernError("RTDetrDetectorDevice needs a CUDA device (no CPU path)")
        _lib.load()
        self.device = torch.device(device)
        self.model = (model if model is not None else build_rtdetr(seed, num_labels)).to(self.device).eval()
        self.min_confidence, self.precision, self.size, self.keep_label = float(min_confidence), precision, input_size, keep_label
        self.use_graphs = use_graphs and precision == "bf16"
        self._graphs = {}   # batch size -> (graph, static input, logits, boxes); the transformers forward is ~1500 small launches

    @torch.no_grad()
    def calibrate(self, frames: torch.Tensor, per_image: int = 40):
        """Synthetic weights only (nets/rtdetr.py): set the class-0 bias so that ~per_image detections pass on ``frames``."""
        x = kernels.resize_frames(frames, (self.size, self.size), torch.float32, 1.0 / 255.0)
        return calibrate_person_bias(self.model, x, per_image, self.min_confidence)

    def _forward_graphed(self, pixel_values: torch.Tensor):
        n = pixel_values.shape[0]
        entry = self._graphs.get(n)
        if entry is None:
            try:
                x = pixel_values.clone()
                s = torch.cuda.Stream(device=self.device)
                s.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(s), torch.autocast(device_type="cuda", dtype=torch.bfloat16):
                    for _ in range(2):
                        self.model(pixel_values=x)
                torch.cuda.current_stream(self.device).wait_stream(s)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g), torch.autocast(device_type="cuda", dtype=torch.bfloat16):
                    out = self.model(pixel_values=x)
                    lg, bx = out.logits.float().contiguous(), out.pred_boxes.float().contiguous()
                entry = (g, x, lg, bx)
            except Exception as e:   # a host-synchronising op inside the third-party forward: fall back to eager launches
                torch.cuda.synchronize(self.device)
                self.use_graphs = False
                self.graph_error = f"{type(e).__name__}: {e}"
                return None
            self._graphs[n] = entry
        g, x, lg, bx = entry
        x.copy_(pixel_values)
        g.replay()
        return lg, bx

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor):
        """pixel_values float32 [n,3,S,S] on the device -> (logits float32 [n,Q,C], boxes float32 [n,Q,4])."""
        if self.use_graphs:
            r = self._forward_graphed(pixel_values)
            if r is not None:
                return r
        if self.precision == "bf16":
            with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
                out = self.model(pixel_values=pixel_values)
        else:   # parity mode: fp32 without TF32
            tf32_c, tf32_m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            try:
                out = self.model(pixel_values=pixel_values)
            finally:
                torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32_c, tf32_m
        return out.logits.float().contiguous(), out.pred_boxes.float().contiguous()

    @torch.no_grad()
    def detect_batch(self, frames: torch.Tensor):
        """frames uint8 [n,H,W,3] (device) -> (rows float64 [n,Q,6] = [l,t,w,h,score,query], counts int32 [n]); rows of an image
        are in descending score order, the first counts[i] are valid."""
        n, H, W, _ = frames.shape
        x = kernels.resize_frames(frames, (self.size, self.size), torch.float32, 1.0 / 255.0)
        logits, boxes = self.forward(x)
        return kernels.rtdetr_decode(logits, boxes, (W, H), self.min_confidence, self.keep_label)
