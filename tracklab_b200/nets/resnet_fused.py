"""Fused channels-last bf16 executor for ResNet50ReID (GPU only). Same weights as ``nets.resnet_reid.ResNet50ReID.forward``:

  * every convolution runs with its bias and ReLU (and the bottleneck's residual add) inside the cuDNN call
    (``torch.cudnn_convolution_relu`` / ``torch.cudnn_convolution_add_relu``): no separate epilogue pass over HBM;
  * the 7x7 / stride-2 stem on RGB is evaluated as a 4x4 / stride-1 convolution on the 2x2 space-to-depth crop
    (12 -> 16 channels, zero border written by ``tk_crop_resize_norm`` in its TK_CROP_LAYOUT_S2D16 layout): identical sums,
    ~5x faster in cuDNN than 7x7/2 on a 3 -> 8 channel input (tools/probe_stem.py);
  * ``tk_maxpool3x3s2_nhwc`` and ``tk_avgpool_nhwc`` replace the framework's pooling passes;
  * the crop count is rounded up to a bucket so that cuDNN sees a bounded set of shapes, and each bucket's forward is
    captured once into a CUDA graph (about 60 launches per replay).

``legacy=True`` keeps the first version (plain cuDNN convolutions + ``tk_bias_act_nhwc`` epilogues on an 8-channel input) for
A/B timing.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import kernels
from .resnet_reid import ConvBias, ResNet50ReID

BUCKET = 64


def stem_weight_s2d16(w: torch.Tensor) -> torch.Tensor:
    """[64,3,7,7] stride-2 pad-3 weights -> [64,16,4,4] stride-1 pad-0 weights on the s2d16 input (zero where no tap maps)."""
    w2 = torch.zeros((w.shape[0], 16, 4, 4), dtype=w.dtype, device=w.device)
    for ky in range(7):
        ay, py = divmod(ky - 3, 2)           # ky - 3 = 2 * ay + py
        for kx in range(7):
            ax, px = divmod(kx - 3, 2)
            c0 = (py * 2 + px) * 3
            w2[:, c0:c0 + 3, ay + 2, ax + 2] = w[:, :, ky, kx]
    return w2


class ResNet50Fused:
    def __init__(self, model: ResNet50ReID, device, legacy: bool = False, use_graphs: bool = True, crop_hw=(256, 128),
                 use_tc: bool | None = None):
        import os
        self.device = torch.device(device)
        # 1x1 stride-1 layers (conv1 / conv3 of every bottleneck: 32 of the 53 convolutions) on the hand-written tcgen05 GEMM with
        # bias + ReLU (+ residual) fused (csrc/conv1x1_tc.cu) instead of cuDNN's fused convolution; TK_RESNET_TC=0/1 overrides
        self.use_tc = (os.environ.get("TK_RESNET_TC", "0") == "1") if use_tc is None else use_tc
        self.tc_layers = 0
        self.model = model
        self.legacy = legacy
        self.use_graphs = use_graphs and not legacy
        self.crop_hw = crop_hw
        self._c = {}
        for mod in model.modules():
            if isinstance(mod, ConvBias):
                w = mod.conv.weight.detach().to(self.device, torch.float32)
                if w.shape[1] == 3:
                    # legacy: pad RGB to 8 input channels (zero weights); fused: space-to-depth stem
                    w = F.pad(w, (0, 0, 0, 0, 0, 5)) if legacy else stem_weight_s2d16(w)
                b32 = mod.conv.bias.detach().to(self.device, torch.float32).contiguous()
                self._c[id(mod)] = (w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last), b32, b32.to(torch.bfloat16),
                                    tuple(mod.conv.stride), tuple(mod.conv.padding))
        # a bottleneck's projection shortcut has no activation of its own: its bias is folded into conv3's
        # (relu(conv3(y) + b3 + down(x) + bd) = relu(conv3(y) + (b3 + bd) + down_nobias(x))), so no separate bias pass remains
        self._b3, self._b3f, self._w2d = {}, {}, {}
        for mod in model.modules():
            if isinstance(mod, ConvBias) and tuple(mod.conv.kernel_size) == (1, 1) and tuple(mod.conv.stride) == (1, 1):
                w = self._c[id(mod)][0]
                self._w2d[id(mod)] = w.reshape(w.shape[0], w.shape[1]).contiguous()
        for mod in model.modules():
            if hasattr(mod, "conv3") and hasattr(mod, "down"):
                b = self._c[id(mod.conv3)][1]
                if mod.down is not None:
                    b = b + self._c[id(mod.down)][1]
                self._b3[id(mod)] = b.to(torch.bfloat16)
                self._b3f[id(mod)] = b.contiguous()
        self._graphs = {}   # bucket size -> (graph, static input, static output)
        self._pool = None   # one private memory pool shared by all buckets' graphs (they never run concurrently)

    # ---- first version: cuDNN convolution + libtrackkern epilogue ------------------------------------------------
    def _conv_legacy(self, x, m, act, residual=None):
        w, b, _, stride, pad = self._c[id(m)]
        y = F.conv2d(x, w, None, stride, pad)
        return kernels.bias_act(y, b, y, 0, act, residual)

    @torch.no_grad()
    def forward_legacy(self, x8: torch.Tensor) -> torch.Tensor:
        """x8: [N,8,256,128] bf16 channels-last (RGB in channels 0..2, zeros elsewhere) -> float32 [N,2048]."""
        m = self.model
        x = self._conv_legacy(x8, m.conv1, 2)
        x = F.max_pool2d(x, 3, 2, 1)
        for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
            for blk in layer:
                idt = x if blk.down is None else self._conv_legacy(x, blk.down, 0)
                y = self._conv_legacy(x, blk.conv1, 2)
                y = self._conv_legacy(y, blk.conv2, 2)
                x = self._conv_legacy(y, blk.conv3, 3, residual=idt)
        return x.float().mean(dim=(2, 3))

    # ---- fused version ----------------------------------------------------------------------------------------------
    def _relu_conv(self, x, m, residual=None, stem=False, bias=None):
        w, _, b16, stride, pad = self._c[id(m)]
        if bias is not None:
            b16 = bias
        if stem:
            stride, pad = (1, 1), (0, 0)
        if self.use_tc and id(m) in self._w2d:
            self.tc_layers += 1
            b32 = self._c[id(m)][1] if bias is None else self._b3f_cur
            return kernels.conv1x1_bias_act(x, self._w2d[id(m)], b32, act=2 if residual is None else 3, residual=residual)
        if residual is None:
            return torch.cudnn_convolution_relu(x, w, b16, stride, pad, (1, 1), 1)
        return torch.cudnn_convolution_add_relu(x, w, residual, 1.0, b16, stride, pad, (1, 1), 1)

    def _body(self, xs: torch.Tensor) -> torch.Tensor:
        m = self.model
        x = self._relu_conv(xs, m.conv1, stem=True)
        x = kernels.maxpool3x3s2(x)
        for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
            for blk in layer:
                if blk.down is None:
                    idt = x
                else:
                    w, _, _, stride, pad = self._c[id(blk.down)]
                    idt = F.conv2d(x, w, None, stride, pad)
                y = self._relu_conv(x, blk.conv1)
                y = self._relu_conv(y, blk.conv2)
                self._b3f_cur = self._b3f[id(blk)]
                x = self._relu_conv(y, blk.conv3, residual=idt, bias=self._b3[id(blk)])
        return kernels.avgpool(x)

    def input_buffer(self, n_crops: int) -> torch.Tensor:
        """Zero-initialised s2d16 stem input for ``n_crops`` rounded up to the bucket (the crop kernel fills the interior)."""
        nb = self.bucket(n_crops)
        entry = self._graphs.get(nb)
        if entry is not None:
            return entry[1]
        h, w = self.crop_hw
        return torch.zeros((nb, 16, h // 2 + 3, w // 2 + 3), dtype=torch.bfloat16, device=self.device).contiguous(memory_format=torch.channels_last)

    @staticmethod
    def bucket(n: int) -> int:
        return max(BUCKET, (n + BUCKET - 1) // BUCKET * BUCKET)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor, n_valid: int | None = None) -> torch.Tensor:
        """Legacy: x = [N,8,H,W]. Fused: x = the buffer returned by ``input_buffer`` (first ``n_valid`` crops filled)."""
        if self.legacy:
            return self.forward_legacy(x)
        nb = x.shape[0]
        n_valid = nb if n_valid is None else n_valid
        if not self.use_graphs:
            return self._body(x)[:n_valid]
        entry = self._graphs.get(nb)
        if entry is None:
            # warm-up on a side stream (cuDNN autotuning must happen outside the capture), then capture
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                for _ in range(2):
                    self._body(x)
            torch.cuda.current_stream(self.device).wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self._pool):
                out = self._body(x)
            if self._pool is None:
                self._pool = g.pool()
            entry = (g, x, out)
            self._graphs[nb] = entry
        g, x_static, out = entry
        if x.data_ptr() != x_static.data_ptr():
            x_static.copy_(x)
        g.replay()
        return out[:n_valid]
