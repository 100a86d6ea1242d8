"""Fused channels-last executor for ResNet50ReID (GPU only): cuDNN convolutions without bias + one libtrackkern epilogue
pass per convolution (bias + ReLU, or bias + residual + ReLU for the bottleneck tail). Same weights/arithmetic as
``nets.resnet_reid.ResNet50ReID.forward``."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import kernels
from .resnet_reid import ConvBias, ResNet50ReID


class ResNet50Fused:
    def __init__(self, model: ResNet50ReID, device):
        self.device = torch.device(device)
        self.model = model
        self._c = {}
        for mod in model.modules():
            if isinstance(mod, ConvBias):
                w = mod.conv.weight.detach().to(self.device, torch.float32)
                if w.shape[1] == 3:   # pad RGB to 8 input channels (zero weights) so cuDNN needs no NHWC padding pass
                    w = F.pad(w, (0, 0, 0, 0, 0, 5))
                self._c[id(mod)] = (w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last),
                                    mod.conv.bias.detach().to(self.device, torch.float32).contiguous(), mod.conv.stride, mod.conv.padding)

    def _conv(self, x, m, act, residual=None):
        w, b, stride, pad = self._c[id(m)]
        y = F.conv2d(x, w, None, stride, pad)
        return kernels.bias_act(y, b, y, 0, act, residual)

    @torch.no_grad()
    def __call__(self, x8: torch.Tensor) -> torch.Tensor:
        """x8: [N,8,256,128] bf16 channels-last (RGB in channels 0..2, zeros elsewhere) -> float32 [N,2048]."""
        m = self.model
        x = self._conv(x8, m.conv1, 2)
        x = F.max_pool2d(x, 3, 2, 1)
        for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
            for blk in layer:
                idt = x if blk.down is None else self._conv(x, blk.down, 0)
                y = self._conv(x, blk.conv1, 2)
                y = self._conv(y, blk.conv2, 2)
                x = self._conv(y, blk.conv3, 3, residual=idt)
        return x.float().mean(dim=(2, 3))
