"""Helpers to run the UNMODIFIED reference plugins on the same weights as the device path (test infrastructure)."""
import os
import tempfile

import torch


def resnet50_reference_state_dict(mine, eps: float = 1e-5):
    """tracklab_b200.nets.resnet_reid.ResNet50ReID (BatchNorm folded) -> state_dict of the reference's vendored ResNet-50
    (/root/reference/plugins/track/strong_sort/deep/models/resnet.py: conv + BatchNorm) holding the same function: identity
    BatchNorm statistics, the folded bias in the BatchNorm bias, conv weight scaled by sqrt(1 + eps)."""
    sd = {}

    def put(conv_key, bn_key, m):
        sd[conv_key + ".weight"] = m.conv.weight.detach() * (1.0 + eps) ** 0.5
        c = m.conv.weight.shape[0]
        sd[bn_key + ".weight"], sd[bn_key + ".bias"] = torch.ones(c), m.conv.bias.detach().clone()
        sd[bn_key + ".running_mean"], sd[bn_key + ".running_var"] = torch.zeros(c), torch.ones(c)
        sd[bn_key + ".num_batches_tracked"] = torch.tensor(0)

    put("conv1", "bn1", mine.conv1)
    for li, layer in enumerate((mine.layer1, mine.layer2, mine.layer3, mine.layer4), start=1):
        for bi, blk in enumerate(layer):
            q = f"layer{li}.{bi}"
            put(q + ".conv1", q + ".bn1", blk.conv1); put(q + ".conv2", q + ".bn2", blk.conv2); put(q + ".conv3", q + ".bn3", blk.conv3)
            if blk.down is not None:
                put(q + ".downsample.0", q + ".downsample.1", blk.down)
    return sd


def reference_strongsort(hyper, seed: int = 1234):
    """The unmodified StrongSORT plugin (crops + ReID + association inside ``update(dets, img)``,
    /root/reference/plugins/track/strong_sort/strong_sort.py:23-85) on CPU with the ResNet-50 weights of the device path.
    Needs oracle.ref_env.install() to have run."""
    from pathlib import Path

    from strong_sort.strong_sort import StrongSORT

    from tracklab_b200.nets.resnet_reid import build_resnet50_reid
    tmp = Path(tempfile.mkdtemp()) / "resnet50_synth.pt"     # the plugin's factory reads the architecture from the file name
    torch.save(resnet50_reference_state_dict(build_resnet50_reid(seed)), tmp)
    return StrongSORT(tmp, torch.device("cpu"), False, **hyper)
