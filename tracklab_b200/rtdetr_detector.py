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
                 keep_label=0, num_labels=80):
        if not torch.cuda.is_available():
            raise _lib.TrackKernError("RTDetrDetectorDevice needs a CUDA device (no CPU path)")
        _lib.load()
        self.device = torch.device(device)
        self.model = (model if model is not None else build_rtdetr(seed, num_labels)).to(self.device).eval()
        self.min_confidence, self.precision, self.size, self.keep_label = float(min_confidence), precision, input_size, keep_label

    @torch.no_grad()
    def calibrate(self, frames: torch.Tensor, per_image: int = 40):
        """Synthetic weights only (nets/rtdetr.py): set the class-0 bias so that ~per_image detections pass on ``frames``."""
        x = kernels.resize_frames(frames, (self.size, self.size), torch.float32, 1.0 / 255.0)
        return calibrate_person_bias(self.model, x, per_image, self.min_confidence)

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor):
        """pixel_values float32 [n,3,S,S] on the device -> (logits float32 [n,Q,C], boxes float32 [n,Q,4])."""
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
