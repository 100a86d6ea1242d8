"""Stand-in for the two box converters of ultralytics.utils.ops (8.3.143)."""
import numpy as np
import torch


def _empty_like(x):
    return torch.empty_like(x) if isinstance(x, torch.Tensor) else np.empty_like(x)


def xyxy2xywh(x):
    y = _empty_like(x)
    y[..., 0] = (x[..., 0] + x[..., 2]) / 2
    y[..., 1] = (x[..., 1] + x[..., 3]) / 2
    y[..., 2] = x[..., 2] - x[..., 0]
    y[..., 3] = x[..., 3] - x[..., 1]
    return y


def xywh2xyxy(x):
    y = _empty_like(x)
    xy = x[..., :2]
    wh = x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y
