"""RT-DETR (r50vd) detector backbone: the Hugging Face ``RTDetrForObjectDetection`` module the reference wrapper runs
(/root/reference/tracklab/wrappers/bbox_detector/transformers_api.py:15-22), instantiated from its config with seeded random
weights because ``from_pretrained`` needs the network. The architecture (ResNet-50vd backbone, hybrid encoder, 6-layer
deformable decoder, 300 queries) is transformers' own code on both arms of the parity tests; only the class-0 ("person")
bias of the last decoder head is raised (+2.0: the random head scores every query ~0.07, the 87th percentile then sits
at the reference's 0.4 threshold) so that synthetic frames produce a realistic ~40 class-0 detections.
"""
from __future__ import annotations

import torch


def build_rtdetr(seed: int = 1234, num_labels: int = 80, state_dict=None):
    """Seeded synthetic weights. transformers' default initialisation (normal, std 0.01-0.02, zero score heads, tiny norm
    scales) lets the signal decay to ~1e-9 before the query selection, where LayerNorm then amplifies rounding noise: outputs
    would not depend on the image and would differ between devices. Convolutions / linear layers are therefore re-drawn
    variance-preserving (kaiming / 1/sqrt(fan_in)) with identity normalisation layers; the class-0 row of the last decoder
    head is widened so that scores spread. Use ``calibrate_person_bias`` on real frames to get a realistic detection count."""
    from transformers import RTDetrConfig, RTDetrForObjectDetection
    st = torch.random.get_rng_state()
    torch.manual_seed(seed)
    model = RTDetrForObjectDetection(RTDetrConfig(num_labels=num_labels)).eval()
    if state_dict is not None:
        model.load_state_dict(state_dict)
    else:
        with torch.no_grad():
            for mod in model.modules():
                if isinstance(mod, torch.nn.Conv2d):
                    torch.nn.init.kaiming_normal_(mod.weight, mode="fan_in", nonlinearity="relu")
                elif isinstance(mod, torch.nn.Linear):
                    mod.weight.normal_(0.0, 1.0 / mod.in_features ** 0.5)
                elif "BatchNorm" in type(mod).__name__ or isinstance(mod, torch.nn.LayerNorm):   # identity statistics / affine
                    mod.weight.fill_(1.0)
                    mod.bias.zero_()
                    if hasattr(mod, "running_mean"):
                        mod.running_mean.zero_()
                        mod.running_var.fill_(1.0)
            head = model.model.decoder.class_embed[-1]
            head.bias.fill_(-4.0)
            head.weight[0] *= 8.0
    torch.random.set_rng_state(st)
    for p in model.parameters():
        p.requires_grad_(False)
    return model


@torch.no_grad()
def calibrate_person_bias(model, pixel_values: torch.Tensor, per_image: int = 40, threshold: float = 0.4) -> float:
    """Synthetic weights only: shift the class-0 bias of the last decoder head so that about ``per_image`` of the queries score
    above ``threshold`` on ``pixel_values`` (float32 [n,3,S,S] on the model's device). Returns the bias that was set."""
    head = model.model.decoder.class_embed[-1]
    l0 = model(pixel_values=pixel_values).logits[..., 0].float().flatten()
    q = 1.0 - per_image / model.config.num_queries
    head.bias[0] += float(torch.logit(torch.tensor(threshold))) - float(torch.quantile(l0, q))
    return float(head.bias[0])
