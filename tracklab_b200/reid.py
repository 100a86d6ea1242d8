"""Device-resident ReID stage: frames + detection rows (HBM) -> appearance features float32 [N, E] (HBM).

crop gather (tk_crop_resize_norm, PIL-exact, written in the stem's space-to-depth layout) -> ResNet-50 in bf16 channels-last
(cuDNN convolutions with fused bias/ReLU/residual, libtrackkern pooling kernels, one CUDA graph per crop-count bucket). Stands in for the in-tracker ReID forward of the StrongSORT plugin
(/root/reference/plugins/track/strong_sort/strong_sort.py:135-145, reid_multibackend.py:184-237) for all detections of a
batch of frames at once; the features feed tk_strongsort_run.
"""
from __future__ import annotations

import torch

from . import _lib, kernels
from .nets.resnet_reid import ResNet50ReID, build_resnet50_reid


def build_reid_model(arch: str = "resnet50", seed: int = 1234):
    """Architectures the StrongSORT plugin's factory can name (deep/reid_model_factory.py:122-127 parses it from the weights file
    name): resnet50 (2048-d), osnet_x1_0 / osnet_ibn_x1_0 (512-d, the YAML default)."""
    if arch == "resnet50":
        return build_resnet50_reid(seed)
    if arch in ("osnet_x1_0", "osnet_ibn_x1_0"):
        from .nets.osnet_reid import build_osnet_reid
        return build_osnet_reid(seed, ibn="ibn" in arch)
    raise _lib.TrackKernError(f"ReID architecture {arch!r} is not built (resnet50, osnet_x1_0, osnet_ibn_x1_0)")


class _GraphedBackbone:
    """Generic bf16 channels-last executor for a ReID module: crop count rounded up to a bucket, one CUDA graph per bucket."""

    BUCKET = 64

    def __init__(self, model, device, crop_hw=(256, 128), use_graphs=True):
        self.model = model.to(device).to(torch.bfloat16).to(memory_format=torch.channels_last).eval()
        self.device, self.crop_hw, self.use_graphs = device, crop_hw, use_graphs
        self._graphs, self._pool = {}, None

    def bucket(self, n):
        return max(self.BUCKET, (n + self.BUCKET - 1) // self.BUCKET * self.BUCKET)

    def input_buffer(self, n):
        nb = self.bucket(n)
        e = self._graphs.get(nb)
        if e is not None:
            return e[1]
        return torch.zeros((nb, 3, *self.crop_hw), dtype=torch.bfloat16, device=self.device).contiguous(memory_format=torch.channels_last)

    @torch.no_grad()
    def __call__(self, x, n_valid):
        if not self.use_graphs:
            return self.model(x).float()[:n_valid]
        nb = x.shape[0]
        e = self._graphs.get(nb)
        if e is None:
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                for _ in range(2):
                    self.model(x)
            torch.cuda.current_stream(self.device).wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self._pool):
                out = self.model(x).float()
            if self._pool is None:
                self._pool = g.pool()
            e = (g, x, out)
            self._graphs[nb] = e
        g, xs, out = e
        if x.data_ptr() != xs.data_ptr():
            xs.copy_(x)
        g.replay()
        return out[:n_valid]


class ReidStageDevice:
    def __init__(self, device="cuda:0", max_crops=2048, seed=1234, model=None, fused=True, precision="bf16", legacy=False,
                 use_graphs=True, arch="resnet50"):
        if not torch.cuda.is_available():
            raise _lib.TrackKernError("ReidStageDevice needs a CUDA device (no CPU path)")
        _lib.load()
        self.device = torch.device(device)
        self.model = (model if model is not None else build_reid_model(arch, seed)).to(self.device).eval()
        self.feature_dim = self.model.feature_dim
        self.max_crops = max_crops
        self.fused = None
        self.generic = None
        self.precision = precision
        if precision == "fp32":     # parity mode: plain fp32 module, TF32 off (features within ~1e-5 of the CPU reference)
            self.model = self.model.float()
        elif not isinstance(self.model, ResNet50ReID):   # OSNet flavours: framework kernels under a CUDA graph
            self.generic = _GraphedBackbone(self.model, self.device, use_graphs=use_graphs)
        elif fused:
            from .nets.resnet_fused import ResNet50Fused
            self.fused = ResNet50Fused(self.model, self.device, legacy=legacy, use_graphs=use_graphs)
        else:
            self.model = self.model.to(torch.bfloat16).to(memory_format=torch.channels_last)
        torch.backends.cudnn.benchmark = True

    @torch.no_grad()
    def features(self, frames: torch.Tensor, dets: torch.Tensor, det_frame: torch.Tensor, ltwh_rows: bool = False) -> torch.Tensor:
        """frames uint8 [F,H,W,3], dets float64 [N,7], det_frame int32 [N] (frame index of each row) -> float32 [N,E].
        ltwh_rows: rows are the detector's [l,t,w,h,..] and the crop follows the ReID wrapper's rule (kernels.crop_resize_norm)."""
        N = dets.shape[0]
        out = torch.empty((N, self.feature_dim), dtype=torch.float32, device=self.device)
        for i in range(0, N, self.max_crops):
            j = min(N, i + self.max_crops)
            if self.precision == "fp32":
                x = kernels.crop_resize_norm(frames, dets[i:j], det_frame[i:j], out_dtype=torch.float32, ltwh_rows=ltwh_rows)
                tf32 = torch.backends.cudnn.allow_tf32
                torch.backends.cudnn.allow_tf32 = False
                try:
                    out[i:j] = self.model(x)
                finally:
                    torch.backends.cudnn.allow_tf32 = tf32
            elif self.generic is not None:
                buf = self.generic.input_buffer(j - i)
                kernels.crop_resize_norm(frames, dets[i:j], det_frame[i:j], out_dtype=torch.bfloat16, channels_last=True, out=buf, ltwh_rows=ltwh_rows)
                out[i:j] = self.generic(buf, j - i)
            elif self.fused is not None and not self.fused.legacy:
                buf = self.fused.input_buffer(j - i)   # s2d16 stem layout, crop count rounded up to the bucket
                kernels.crop_resize_norm(frames, dets[i:j], det_frame[i:j], s2d16_out=buf, ltwh_rows=ltwh_rows)
                out[i:j] = self.fused(buf, n_valid=j - i)
            elif self.fused is not None:
                x = kernels.crop_resize_norm(frames, dets[i:j], det_frame[i:j], out_dtype=torch.bfloat16, channels_last=True,
                                             pad_channels_to=8, ltwh_rows=ltwh_rows)
                out[i:j] = self.fused(x)
            else:
                x = kernels.crop_resize_norm(frames, dets[i:j], det_frame[i:j], out_dtype=torch.bfloat16, channels_last=True, ltwh_rows=ltwh_rows)
                out[i:j] = self.model(x).float()
        return out
