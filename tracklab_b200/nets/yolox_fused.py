"""Fused channels-last executor for the YOLOX module (GPU only).

Same weights and arithmetic as ``nets.yolox.YOLOX.forward`` (the parity test compares both), different
schedule: every convolution runs in cuDNN WITHOUT bias, and one libtrackkern epilogue pass
(tk_bias_act_nhwc) applies bias + SiLU (+ the bottleneck residual) and writes straight into the channel slice
of the concat buffer that consumes it — so the separate bias-add, SiLU, torch.cat, max-pool and up-sampling
kernels of the eager graph (~80 % of its device time on B200) disappear. The Focus space-to-depth is produced
by the letterbox kernel itself (layout "focus16").
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import kernels
from .yolox import YOLOX, Bottleneck, ConvAct, CSPLayer


class _Conv:
    """One ConvAct as (bf16 channels-last weight, fp32 bias, stride, padding)."""

    def __init__(self, m: ConvAct, device, pad_in_to: int | None = None):
        w = m.conv.weight.detach().to(device=device, dtype=torch.float32)
        if pad_in_to is not None and w.shape[1] < pad_in_to:
            w = F.pad(w, (0, 0, 0, 0, 0, pad_in_to - w.shape[1]))
        self.w = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        self.b = m.conv.bias.detach().to(device=device, dtype=torch.float32).contiguous()
        self.stride = m.conv.stride
        self.padding = m.conv.padding
        self.cout = self.w.shape[0]
        # 1x1 stride-1 layers run on the hand-written tcgen05 GEMM with the epilogue fused in (csrc/conv1x1_tc.cu)
        self.w2d = None
        if tuple(m.conv.kernel_size) == (1, 1) and tuple(self.stride) == (1, 1) and self.w.shape[1] % 8 == 0 and self.cout % 16 == 0:
            self.w2d = self.w.reshape(self.cout, self.w.shape[1]).contiguous()
        # 3x3 stride-1 layers: the implicit-GEMM tcgen05 kernel (csrc/conv3x3_tc.cu); stride-2 layers stay in cuDNN
        self.tc3 = (tuple(m.conv.kernel_size) == (3, 3) and tuple(self.stride) == (1, 1) and tuple(self.padding) == (1, 1)
                    and self.w.shape[1] % 16 == 0 and self.cout % 16 == 0)


class YoloxFused:
    # input channels of the stem convolution: the 12 Focus channels zero-padded to 32 (with 16 cuDNN picks an sm_80 kernel,
    # 425 us per 50 frames; with 32 an sm_100 one, 164 us — tools/probe_yolox_stem.py)
    STEM_IN = 32

    def __init__(self, model: YOLOX, device, use_tc: bool = True, use_tc3: bool | None = None):
        import os
        self.device = torch.device(device)
        self.use_tc = use_tc
        # 3x3 layers on the implicit-GEMM tcgen05 kernel: opt-in (TK_TC3=1). Measured on B200 (profiles/r02_conv_microbench.md) it is
        # correct but slower than cuDNN's native sm_100 3x3 kernels + the epilogue pass at the detector's channel counts
        # (per-tap TMA boxes with 64..256-byte rows are request-bound), so the default keeps cuDNN for 3x3.
        self.use_tc3 = (use_tc and os.environ.get("TK_TC3", "0") == "1") if use_tc3 is None else use_tc3
        self.tc_layers = 0        # 1x1 layers of the last forward that ran on the tcgen05 path
        self.nc = model.num_classes
        dev = self.device
        self._convs = {}
        self.model = model

        def reg(m, **kw):
            self._convs[id(m)] = _Conv(m, dev, **kw)

        for mod in model.modules():
            if isinstance(mod, ConvAct):
                reg(mod)
        self._convs[id(model.stem.conv)] = _Conv(model.stem.conv, dev, pad_in_to=self.STEM_IN)
        # prediction convolutions keep their bias inside cuDNN (4 / 1 / nc output channels, no activation):
        # reg + obj share their input, so they are one 5-channel convolution.
        self.pred_ro = []
        self.pred_cls = []
        for k in range(3):
            w = torch.cat([model.reg_preds[k].weight, model.obj_preds[k].weight], 0).detach()
            b = torch.cat([model.reg_preds[k].bias, model.obj_preds[k].bias], 0).detach()
            self.pred_ro.append((w.to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last), b.to(dev, torch.bfloat16)))
            self.pred_cls.append((model.cls_preds[k].weight.detach().to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last),
                                  model.cls_preds[k].bias.detach().to(dev, torch.bfloat16)))

    # ---- building blocks -------------------------------------------------------------------------
    def _conv(self, x, m: ConvAct, dst=None, dst_off=0, residual=None, res_off=0):
        c = self._convs[id(m)]
        if self.use_tc and c.w2d is not None and x.shape[1] == c.w2d.shape[1]:
            self.tc_layers += 1
            return kernels.conv1x1_bias_act(x, c.w2d, c.b, dst=dst, dst_offset=dst_off, act=1, residual=residual, res_offset=res_off)
        if self.use_tc3 and c.tc3 and x.shape[1] == c.w.shape[1]:
            self.tc_layers += 1
            return kernels.conv3x3_bias_act(x, c.w, c.b, dst=dst, dst_offset=dst_off, act=1, residual=residual, res_offset=res_off)
        y = F.conv2d(x, c.w, None, c.stride, c.padding)
        return kernels.bias_act(y, c.b, y if dst is None else dst, dst_off, 1, residual, res_off)

    def _new(self, B, C, H, W):
        return torch.empty((B, C, H, W), dtype=torch.bfloat16, device=self.device, memory_format=torch.channels_last)

    def _csp(self, x, layer: CSPLayer, dst=None, dst_off=0):
        B, _, H, W = x.shape
        hid = self._convs[id(layer.conv1)].cout
        cat = self._new(B, 2 * hid, H, W)
        x1 = self._conv(x, layer.conv1)
        n = len(layer.m)
        for k, b in enumerate(layer.m):
            t = self._conv(x1, b.conv1)
            last = k == n - 1
            x1 = self._conv(t, b.conv2, dst=cat if last else None, dst_off=0, residual=x1 if b.add else None)
        self._conv(x, layer.conv2, dst=cat, dst_off=hid)
        out = self._conv(cat, layer.conv3, dst=dst, dst_off=dst_off)
        return out

    def _copy_into(self, src, dst, off):
        zero = self._zeros(src.shape[1])
        return kernels.bias_act(src, zero, dst, off, 0)

    def _zeros(self, c):
        z = getattr(self, "_zero_bias", None)
        if z is None or z.numel() < c:
            self._zero_bias = torch.zeros(max(c, 1024), dtype=torch.float32, device=self.device)
        return self._zero_bias

    # ---- forward -----------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, x16: torch.Tensor) -> torch.Tensor:
        """x16: [B,16,S/2,S/2] bf16 channels-last produced by kernels.letterbox(..., focus16=True)."""
        m = self.model
        B = x16.shape[0]
        self.tc_layers = 0
        x = self._conv(x16, m.stem.conv)
        x = self._conv(x, m.dark2[0]); x = self._csp(x, m.dark2[1])
        x = self._conv(x, m.dark3[0]); d3 = self._csp(x, m.dark3[1])
        x = self._conv(d3, m.dark4[0]); d4 = self._csp(x, m.dark4[1])
        x = self._conv(d4, m.dark5[0])
        # SPP bottleneck
        spp = m.dark5[1]
        x = self._conv(x, spp.conv1)
        hid = x.shape[1]
        cat4 = self._new(B, 4 * hid, x.shape[2], x.shape[3])
        kernels.spp_pool(x, cat4)
        x = self._conv(cat4, spp.conv2)
        d5 = self._csp(x, m.dark5[2])

        c3, c4 = d3.shape[1], d4.shape[1]
        H4, W4, H3, W3 = d4.shape[2], d4.shape[3], d3.shape[2], d3.shape[3]
        cat_n4 = self._new(B, 2 * c4, d5.shape[2], d5.shape[3])     # [bu_conv1(pan1) | f0]
        cat_p4 = self._new(B, 2 * c4, H4, W4)                       # [up(f0) | d4]
        cat_n3 = self._new(B, 2 * c3, H4, W4)                       # [bu_conv2(pan2) | f1]
        cat_p3 = self._new(B, 2 * c3, H3, W3)                       # [up(f1) | d3]
        self._conv(d5, m.lateral_conv0, dst=cat_n4, dst_off=c4)     # f0
        kernels.upsample2x(cat_n4, cat_p4, 0, src_offset=c4, channels=c4)
        self._copy_into(d4, cat_p4, c4)
        p4 = self._csp(cat_p4, m.C3_p4)
        self._conv(p4, m.reduce_conv1, dst=cat_n3, dst_off=c3)      # f1
        kernels.upsample2x(cat_n3, cat_p3, 0, src_offset=c3, channels=c3)
        self._copy_into(d3, cat_p3, c3)
        pan2 = self._csp(cat_p3, m.C3_p3)
        self._conv(pan2, m.bu_conv2, dst=cat_n3, dst_off=0)
        pan1 = self._csp(cat_n3, m.C3_n3)
        self._conv(pan1, m.bu_conv1, dst=cat_n4, dst_off=0)
        pan0 = self._csp(cat_n4, m.C3_n4)

        outs = []
        for k, f in enumerate((pan2, pan1, pan0)):
            s = self._conv(f, m.stems[k])
            c = self._conv(self._conv(s, m.cls_convs[k][0]), m.cls_convs[k][1])
            r = self._conv(self._conv(s, m.reg_convs[k][0]), m.reg_convs[k][1])
            ro = F.conv2d(r, self.pred_ro[k][0], self.pred_ro[k][1])
            cl = F.conv2d(c, self.pred_cls[k][0], self.pred_cls[k][1])
            o = torch.cat((ro, cl), dim=1)                           # [B, 5+nc, h, w]
            outs.append(o.permute(0, 2, 3, 1).reshape(B, -1, 5 + self.nc))
        return torch.cat(outs, dim=1).contiguous()                   # [B, 8400, 5+nc]
