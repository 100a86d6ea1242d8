"""OSNet (x1.0 and the IBN variant) ReID backbone in inference form (BatchNorm folded into the preceding convolution / linear
layer), 512-d feature — the default weights of the StrongSORT configuration are ``osnet_ibn_x1_0_msmt17.pt``
(/root/reference/tracklab/configs/modules/track/strong_sort.yaml:8).

Architecture of the vendored torchreid model
(/root/reference/plugins/track/strong_sort/deep/models/osnet.py:27-60 ConvLayer, :63-160 1x1 / 3x3 / LightConv3x3,
:166-226 ChannelGate, :229-286 OSBlock, :292-446 OSNet with layers [2,2,2], channels [64,256,384,512], eval forward =
featuremaps -> global average pool -> fc (Linear + BatchNorm1d + ReLU); :585-601 osnet_ibn_x1_0: InstanceNorm in the stem
and after the residual add of the first stage). ``from_reference_state_dict`` folds a reference ``state_dict``;
tests/test_reid_backbone_cpu.py checks both give the same features in the build container.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _fold(w, bn_w, bn_b, mean, var, eps=1e-5):
    s = bn_w / torch.sqrt(var + eps)
    return w * s.reshape(-1, *([1] * (w.dim() - 1))), bn_b - mean * s


class ConvAct(nn.Module):
    """conv (+ folded BN) [+ ReLU]."""

    def __init__(self, cin, cout, k, stride=1, groups=1, relu=True, bias=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, groups=groups, bias=bias)
        self.relu = relu

    def forward(self, x):
        x = self.conv(x)
        return F.relu(x) if self.relu else x


class LightConv3x3(nn.Module):
    """1x1 linear (no norm) + depthwise 3x3 with the folded BN + ReLU (osnet.py:134-160)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 1, bias=False)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, groups=cout, bias=True)

    def forward(self, x):
        return F.relu(self.conv2(self.conv1(x)))


class ChannelGate(nn.Module):
    def __init__(self, c, reduction=16):
        super().__init__()
        self.fc1 = nn.Conv2d(c, c // reduction, 1)
        self.fc2 = nn.Conv2d(c // reduction, c, 1)

    def forward(self, x):
        g = torch.sigmoid(self.fc2(F.relu(self.fc1(x.mean(dim=(2, 3), keepdim=True)))))
        return x * g


class OSBlock(nn.Module):
    def __init__(self, cin, cout, IN=False):
        super().__init__()
        mid = cout // 4
        self.conv1 = ConvAct(cin, mid, 1)
        self.conv2a = LightConv3x3(mid, mid)
        self.conv2b = nn.Sequential(*[LightConv3x3(mid, mid) for _ in range(2)])
        self.conv2c = nn.Sequential(*[LightConv3x3(mid, mid) for _ in range(3)])
        self.conv2d = nn.Sequential(*[LightConv3x3(mid, mid) for _ in range(4)])
        self.gate = ChannelGate(mid)
        self.conv3 = ConvAct(mid, cout, 1, relu=False)
        self.downsample = ConvAct(cin, cout, 1, relu=False) if cin != cout else None
        self.IN = nn.InstanceNorm2d(cout, affine=True) if IN else None

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        x1 = self.conv1(x)
        x2 = self.gate(self.conv2a(x1)) + self.gate(self.conv2b(x1)) + self.gate(self.conv2c(x1)) + self.gate(self.conv2d(x1))
        out = self.conv3(x2) + idt
        if self.IN is not None:
            out = self.IN(out)
        return F.relu(out)


class OSNetReID(nn.Module):
    feature_dim = 512

    def __init__(self, ibn: bool = False, channels=(64, 256, 384, 512)):
        super().__init__()
        self.ibn = ibn
        c = channels
        if ibn:   # InstanceNorm depends on the input: it cannot be folded
            self.conv1 = nn.Sequential(nn.Conv2d(3, c[0], 7, 2, 3, bias=False), nn.InstanceNorm2d(c[0], affine=True), nn.ReLU())
        else:
            self.conv1 = ConvAct(3, c[0], 7, 2)
        self.maxpool = nn.MaxPool2d(3, 2, 1)

        def stage(cin, cout, reduce, IN=False):
            layers = [OSBlock(cin, cout, IN=IN), OSBlock(cout, cout, IN=IN)]
            if reduce:
                layers.append(nn.Sequential(ConvAct(cout, cout, 1), nn.AvgPool2d(2, 2)))
            return nn.Sequential(*layers)

        self.conv2 = stage(c[0], c[1], True, IN=ibn)
        self.conv3 = stage(c[1], c[2], True)
        self.conv4 = stage(c[2], c[3], False)
        self.conv5 = ConvAct(c[3], c[3], 1)
        self.fc = nn.Linear(c[3], self.feature_dim)

    def forward(self, x):
        x = self.maxpool(self.conv1(x))
        x = self.conv5(self.conv4(self.conv3(self.conv2(x))))
        return F.relu(self.fc(x.mean(dim=(2, 3))))

    @torch.no_grad()
    def from_reference_state_dict(self, sd, eps=1e-5):
        sd = {k: v.float() for k, v in sd.items() if torch.is_tensor(v)}

        def bn(prefix):
            return sd[prefix + ".weight"], sd[prefix + ".bias"], sd[prefix + ".running_mean"], sd[prefix + ".running_var"]

        def conv_bn(dst: ConvAct, p):            # ConvLayer / Conv1x1 / Conv1x1Linear: p.conv + p.bn
            w, b = _fold(sd[p + ".conv.weight"], *bn(p + ".bn"), eps)
            dst.conv.weight.copy_(w); dst.conv.bias.copy_(b)

        def light(dst: LightConv3x3, p):
            dst.conv1.weight.copy_(sd[p + ".conv1.weight"])
            w, b = _fold(sd[p + ".conv2.weight"], *bn(p + ".bn"), eps)
            dst.conv2.weight.copy_(w); dst.conv2.bias.copy_(b)

        def block(dst: OSBlock, p):
            conv_bn(dst.conv1, p + ".conv1")
            light(dst.conv2a, p + ".conv2a")
            for name, n in (("conv2b", 2), ("conv2c", 3), ("conv2d", 4)):
                for i in range(n):
                    light(getattr(dst, name)[i], f"{p}.{name}.{i}")
            for fc in ("fc1", "fc2"):
                getattr(dst.gate, fc).weight.copy_(sd[f"{p}.gate.{fc}.weight"]); getattr(dst.gate, fc).bias.copy_(sd[f"{p}.gate.{fc}.bias"])
            conv_bn(dst.conv3, p + ".conv3")
            if dst.downsample is not None:
                conv_bn(dst.downsample, p + ".downsample")
            if dst.IN is not None:
                dst.IN.weight.copy_(sd[p + ".IN.weight"]); dst.IN.bias.copy_(sd[p + ".IN.bias"])

        if self.ibn:
            self.conv1[0].weight.copy_(sd["conv1.conv.weight"])
            self.conv1[1].weight.copy_(sd["conv1.bn.weight"]); self.conv1[1].bias.copy_(sd["conv1.bn.bias"])
        else:
            conv_bn(self.conv1, "conv1")
        for name in ("conv2", "conv3", "conv4"):
            st = getattr(self, name)
            block(st[0], f"{name}.0"); block(st[1], f"{name}.1")
            if len(st) == 3:
                conv_bn(st[2][0], f"{name}.2.0")
        conv_bn(self.conv5, "conv5")
        w, b = sd["fc.0.weight"], sd["fc.0.bias"]
        s = sd["fc.1.weight"] / torch.sqrt(sd["fc.1.running_var"] + eps)
        self.fc.weight.copy_(w * s[:, None]); self.fc.bias.copy_((b - sd["fc.1.running_mean"]) * s + sd["fc.1.bias"])
        return self


def build_osnet_reid(seed=1234, ibn=False):
    """Seeded random weights (no network): kaiming-normal convolutions / N(0, 0.01) linear layers as the reference's
    _init_params (osnet.py:392-413), identity BatchNorm statistics folded in."""
    st = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = OSNetReID(ibn=ibn)
    k = 1.0 / (1.0 + 1e-5) ** 0.5
    for mod in m.modules():
        if isinstance(mod, nn.Conv2d):
            nn.init.kaiming_normal_(mod.weight, mode="fan_out", nonlinearity="relu")
            if mod.bias is not None:
                nn.init.zeros_(mod.bias)
        elif isinstance(mod, nn.Linear):
            nn.init.normal_(mod.weight, 0, 0.01)
            mod.weight.data.mul_(k)
            nn.init.zeros_(mod.bias)
    torch.random.set_rng_state(st)
    return m.eval()
