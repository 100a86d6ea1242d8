"""YOLOX-s / YOLOX-m in plain PyTorch (inference form: BatchNorm folded into the convolution bias).

The reference runs these detectors as third-party ONNX graphs through rtmlib + onnxruntime
(/root/reference/tracklab/wrappers/bbox_detector/rtmlib_api.py:19-30,
/root/reference/tracklab/configs/modules/bbox_detector/yolox_rtmlib_s.yaml:4-7); neither the graphs nor
the runtime exist offline, so the architecture is restated from the YOLOX paper's model family
(CSPDarknet + PAFPN + decoupled head; depth/width 0.33/0.50 for -s, 0.67/0.75 for -m) with seeded
random weights. Output: raw head tensor [B, 8400, 5+nc] = (reg xywh, obj logit,
cls logits) — sigmoid, grid decode and NMS run in libtrackkern (tk_yolox_nms).
"""
from __future__ import annotations

import torch
import torch.nn as nn


class ConvAct(nn.Module):
    """Conv2d(+folded BN) + SiLU."""

    def __init__(self, cin, cout, k=1, s=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, s, (k - 1) // 2, bias=True)
        self.act = nn.SiLU(inplace=True)

    def forward(self, x):
        return self.act(self.conv(x))


class Bottleneck(nn.Module):
    def __init__(self, cin, cout, shortcut=True, expansion=1.0):
        super().__init__()
        hid = int(cout * expansion)
        self.conv1 = ConvAct(cin, hid, 1)
        self.conv2 = ConvAct(hid, cout, 3)
        self.add = shortcut and cin == cout

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        return y + x if self.add else y


class CSPLayer(nn.Module):
    def __init__(self, cin, cout, n=1, shortcut=True):
        super().__init__()
        hid = int(cout * 0.5)
        self.conv1 = ConvAct(cin, hid, 1)
        self.conv2 = ConvAct(cin, hid, 1)
        self.conv3 = ConvAct(2 * hid, cout, 1)
        self.m = nn.Sequential(*[Bottleneck(hid, hid, shortcut, 1.0) for _ in range(n)])

    def forward(self, x):
        return self.conv3(torch.cat((self.m(self.conv1(x)), self.conv2(x)), dim=1))


class SPPBottleneck(nn.Module):
    def __init__(self, cin, cout, ks=(5, 9, 13)):
        super().__init__()
        hid = cin // 2
        self.conv1 = ConvAct(cin, hid, 1)
        self.pools = nn.ModuleList([nn.MaxPool2d(k, 1, k // 2) for k in ks])
        self.conv2 = ConvAct(hid * (len(ks) + 1), cout, 1)

    def forward(self, x):
        x = self.conv1(x)
        return self.conv2(torch.cat([x] + [p(x) for p in self.pools], dim=1))


class Focus(nn.Module):
    def __init__(self, cin, cout, k=3):
        super().__init__()
        self.conv = ConvAct(cin * 4, cout, k)

    def forward(self, x):
        tl, bl = x[..., ::2, ::2], x[..., 1::2, ::2]
        tr, br = x[..., ::2, 1::2], x[..., 1::2, 1::2]
        return self.conv(torch.cat((tl, bl, tr, br), dim=1))


class YOLOX(nn.Module):
    def __init__(self, depth=0.33, width=0.50, num_classes=1):
        super().__init__()
        bc = int(width * 64)
        bd = max(round(depth * 3), 1)
        self.num_classes = num_classes
        self.stem = Focus(3, bc)
        self.dark2 = nn.Sequential(ConvAct(bc, bc * 2, 3, 2), CSPLayer(bc * 2, bc * 2, bd))
        self.dark3 = nn.Sequential(ConvAct(bc * 2, bc * 4, 3, 2), CSPLayer(bc * 4, bc * 4, bd * 3))
        self.dark4 = nn.Sequential(ConvAct(bc * 4, bc * 8, 3, 2), CSPLayer(bc * 8, bc * 8, bd * 3))
        self.dark5 = nn.Sequential(ConvAct(bc * 8, bc * 16, 3, 2), SPPBottleneck(bc * 16, bc * 16),
                                   CSPLayer(bc * 16, bc * 16, bd, shortcut=False))
        c3, c4, c5 = bc * 4, bc * 8, bc * 16
        self.up = nn.Upsample(scale_factor=2, mode="nearest")
        self.lateral_conv0 = ConvAct(c5, c4, 1)
        self.C3_p4 = CSPLayer(2 * c4, c4, bd, False)
        self.reduce_conv1 = ConvAct(c4, c3, 1)
        self.C3_p3 = CSPLayer(2 * c3, c3, bd, False)
        self.bu_conv2 = ConvAct(c3, c3, 3, 2)
        self.C3_n3 = CSPLayer(2 * c3, c4, bd, False)
        self.bu_conv1 = ConvAct(c4, c4, 3, 2)
        self.C3_n4 = CSPLayer(2 * c4, c5, bd, False)
        hc = int(256 * width)
        self.stems = nn.ModuleList([ConvAct(c, hc, 1) for c in (c3, c4, c5)])
        self.cls_convs = nn.ModuleList([nn.Sequential(ConvAct(hc, hc, 3), ConvAct(hc, hc, 3)) for _ in range(3)])
        self.reg_convs = nn.ModuleList([nn.Sequential(ConvAct(hc, hc, 3), ConvAct(hc, hc, 3)) for _ in range(3)])
        self.cls_preds = nn.ModuleList([nn.Conv2d(hc, num_classes, 1) for _ in range(3)])
        self.reg_preds = nn.ModuleList([nn.Conv2d(hc, 4, 1) for _ in range(3)])
        self.obj_preds = nn.ModuleList([nn.Conv2d(hc, 1, 1) for _ in range(3)])

    def forward(self, x):
        x = self.stem(x)
        x = self.dark2(x)
        d3 = self.dark3(x)
        d4 = self.dark4(d3)
        d5 = self.dark5(d4)
        f0 = self.lateral_conv0(d5)
        p4 = self.C3_p4(torch.cat((self.up(f0), d4), 1))
        f1 = self.reduce_conv1(p4)
        pan2 = self.C3_p3(torch.cat((self.up(f1), d3), 1))
        pan1 = self.C3_n3(torch.cat((self.bu_conv2(pan2), f1), 1))
        pan0 = self.C3_n4(torch.cat((self.bu_conv1(pan1), f0), 1))
        outs = []
        for k, f in enumerate((pan2, pan1, pan0)):
            s = self.stems[k](f)
            c = self.cls_preds[k](self.cls_convs[k](s))
            rf = self.reg_convs[k](s)
            o = torch.cat((self.reg_preds[k](rf), self.obj_preds[k](rf), c), 1)  # [B, 5+nc, h, w]
            outs.append(o.flatten(2))
        return torch.cat(outs, dim=2).permute(0, 2, 1).contiguous()  # [B, 8400, 5+nc]


_VARIANTS = {"s": (0.33, 0.50), "m": (0.67, 0.75), "tiny": (0.33, 0.375), "l": (1.0, 1.0)}


def build_yolox(variant="s", num_classes=1, seed=1234, prior_prob=None):
    """Seeded random-init model (torch.manual_seed(seed), default init).
    ``prior_prob`` sets the obj/cls prediction biases like YOLOX's ``initialize_biases`` so a random-weight
    network can be steered to emit a realistic number of above-threshold candidates."""
    depth, width = _VARIANTS[variant]
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = YOLOX(depth, width, num_classes)
    if prior_prob is not None:
        import math
        b = -math.log((1 - prior_prob) / prior_prob)
        for conv in list(m.cls_preds) + list(m.obj_preds):
            nn.init.constant_(conv.bias, b)
    torch.random.set_rng_state(gen_state)
    return m.eval()
