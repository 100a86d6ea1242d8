"""Pairwise box-overlap oracles (test infrastructure)."""
import numpy as np


def iou_xyxy(a, b):
    """Plain IoU, boxes x1y1x2y2, broadcast [N,M] — /root/reference/plugins/track/oc_sort/association.py:5-21."""
    a = np.asarray(a)[:, None, :]
    b = np.asarray(b)[None, :, :]
    iw = np.maximum(0.0, np.minimum(a[..., 2], b[..., 2]) - np.maximum(a[..., 0], b[..., 0]))
    ih = np.maximum(0.0, np.minimum(a[..., 3], b[..., 3]) - np.maximum(a[..., 1], b[..., 1]))
    inter = iw * ih
    return inter / ((a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
                    + (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1]) - inter)


def giou_xyxy(a, b):
    """(GIoU+1)/2 — /root/reference/plugins/track/oc_sort/association.py:24-55."""
    a = np.asarray(a)[:, None, :]
    b = np.asarray(b)[None, :, :]
    iw = np.maximum(0.0, np.minimum(a[..., 2], b[..., 2]) - np.maximum(a[..., 0], b[..., 0]))
    ih = np.maximum(0.0, np.minimum(a[..., 3], b[..., 3]) - np.maximum(a[..., 1], b[..., 1]))
    inter = iw * ih
    iou = inter / ((a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
                   + (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1]) - inter)
    cw = np.maximum(a[..., 2], b[..., 2]) - np.minimum(a[..., 0], b[..., 0])
    ch = np.maximum(a[..., 3], b[..., 3]) - np.minimum(a[..., 1], b[..., 1])
    assert (cw > 0).all() and (ch > 0).all()
    hull = cw * ch
    g = iou - (hull - inter) / hull
    return (g + 1.0) / 2.0


def iou_plus1_f32(a_tlbr, b_tlbr):
    """IoU with +1-pixel extents in float32 — /root/reference/plugins/track/byte_track/matching.py:182-218.

    The reference is a scalar double loop over np.float32 values; every operation below is the
    same IEEE float32 operation applied elementwise, in the same order, so the bits are identical.
    """
    a = np.ascontiguousarray(a_tlbr, dtype=np.float32).reshape(-1, 4)
    b = np.ascontiguousarray(b_tlbr, dtype=np.float32).reshape(-1, 4)
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    if out.size == 0:
        return out
    one = np.float32(1)
    area_b = (b[:, 2] - b[:, 0] + one) * (b[:, 3] - b[:, 1] + one)
    area_a = (a[:, 2] - a[:, 0] + one) * (a[:, 3] - a[:, 1] + one)
    iw = np.minimum(a[:, None, 2], b[None, :, 2]) - np.maximum(a[:, None, 0], b[None, :, 0]) + one
    ih = np.minimum(a[:, None, 3], b[None, :, 3]) - np.maximum(a[:, None, 1], b[None, :, 1]) + one
    ok = (iw > 0) & (ih > 0)
    inter = iw * ih
    ua = area_a[:, None] + area_b[None, :] - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        val = inter / ua
    out[ok] = val[ok]
    return out


def iou_tlwh_one_to_many(box, cands):
    """IoU of one tlwh box against ``cands[M,4]`` (tlwh), no +1 —
    /root/reference/plugins/track/strong_sort/sort/iou_matching.py:7-39.

    Elementwise restatement: a product over a length-2 axis is one multiply, so the bits match the
    reference's ``prod(axis=1)`` formulation. dtype follows NumPy promotion (f64 track box vs
    f32 detection boxes -> f64), as in the reference.
    """
    cands = np.asarray(cands)
    x0 = np.maximum(box[0], cands[:, 0])
    y0 = np.maximum(box[1], cands[:, 1])
    x1 = np.minimum(box[0] + box[2], cands[:, 0] + cands[:, 2])
    y1 = np.minimum(box[1] + box[3], cands[:, 1] + cands[:, 3])
    iw = np.maximum(0.0, x1 - x0)
    ih = np.maximum(0.0, y1 - y0)
    inter = iw * ih
    return inter / (box[2] * box[3] + cands[:, 2] * cands[:, 3] - inter)
