"""ResNet-50 ReID backbone in inference form (BatchNorm folded into the convolution bias), 2048-d pooled feature.

Architecture of the vendored torchreid model the StrongSORT plugin builds
(/root/reference/plugins/track/strong_sort/deep/models/resnet.py:104-163 Bottleneck, :166-361 ResNet, :425-437 resnet50:
layers [3,4,6,3], stride on the 3x3 convolution, last_stride 2, eval forward = global average pool, no fc).
``from_reference_state_dict`` folds a reference ``state_dict`` (conv + BN) into this form; tests/test_reid_backbone_cpu.py
checks both give the same features in the build container.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class ConvBias(nn.Module):
    def __init__(self, cin, cout, k, stride=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, bias=True)

    def forward(self, x):
        return self.conv(x)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin, planes, stride=1, downsample=False):
        super().__init__()
        self.conv1 = ConvBias(cin, planes, 1)
        self.conv2 = ConvBias(planes, planes, 3, stride)
        self.conv3 = ConvBias(planes, planes * 4, 1)
        self.down = ConvBias(cin, planes * 4, 1, stride) if downsample else None

    def forward(self, x):
        idt = x if self.down is None else self.down(x)
        y = torch.relu(self.conv1(x))
        y = torch.relu(self.conv2(y))
        return torch.relu(self.conv3(y) + idt)


class ResNet50ReID(nn.Module):
    feature_dim = 2048

    def __init__(self):
        super().__init__()
        self.conv1 = ConvBias(3, 64, 7, 2)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        cfg = [(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)]
        layers, cin = [], 64
        for planes, n, stride in cfg:
            blocks = [Bottleneck(cin, planes, stride, downsample=True)]
            cin = planes * 4
            blocks += [Bottleneck(cin, planes) for _ in range(n - 1)]
            layers.append(nn.Sequential(*blocks))
        self.layer1, self.layer2, self.layer3, self.layer4 = layers

    def forward(self, x):
        x = self.maxpool(torch.relu(self.conv1(x)))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return x.mean(dim=(2, 3))

    @torch.no_grad()
    def from_reference_state_dict(self, sd, eps=1e-5):
        """Fold (conv, bn) pairs of a torchreid/torchvision-style ResNet-50 state_dict into this module."""
        def fold(dst: ConvBias, conv_key, bn_key):
            w = sd[conv_key + ".weight"].float()
            g, b = sd[bn_key + ".weight"].float(), sd[bn_key + ".bias"].float()
            m, v = sd[bn_key + ".running_mean"].float(), sd[bn_key + ".running_var"].float()
            s = g / torch.sqrt(v + eps)
            dst.conv.weight.copy_(w * s[:, None, None, None])
            dst.conv.bias.copy_(b - m * s)

        fold(self.conv1, "conv1", "bn1")
        for li, layer in enumerate((self.layer1, self.layer2, self.layer3, self.layer4), start=1):
            for bi, blk in enumerate(layer):
                p = f"layer{li}.{bi}"
                fold(blk.conv1, p + ".conv1", p + ".bn1")
                fold(blk.conv2, p + ".conv2", p + ".bn2")
                fold(blk.conv3, p + ".conv3", p + ".bn3")
                if blk.down is not None:
                    fold(blk.down, p + ".downsample.0", p + ".downsample.1")
        return self


def build_resnet50_reid(seed=1234):
    """Seeded random weights (no network): kaiming-normal convolutions as the reference's _init_params (resnet.py:304-326),
    identity BatchNorm statistics folded in."""
    st = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = ResNet50ReID()
    for mod in m.modules():
        if isinstance(mod, nn.Conv2d):
            nn.init.kaiming_normal_(mod.weight, mode="fan_out", nonlinearity="relu")
            mod.weight.data.mul_(1.0 / (1.0 + 1e-5) ** 0.5)
            nn.init.zeros_(mod.bias)
    torch.random.set_rng_state(st)
    return m.eval()
