"""YOLOX post-processing oracle (test infrastructure): decode, threshold, class-aware NMS, wrapper rows.

Follows rtmlib 0.0.13 ``YOLOX.postprocess`` / ``multiclass_nms`` / ``nms`` (un-vendored third party behind
/root/reference/tracklab/wrappers/bbox_detector/rtmlib_api.py:30; SURVEY.md §3.2 [3P-memory] — parity
unpinned by the reference tree, restated from the published YOLOX demo post-processing) and the wrapper
/root/reference/tracklab/wrappers/bbox_detector/rtmlib_api.py:31-46 (clip, ltwh, conf = 1.0, running id).
All arithmetic float32 like the onnxruntime outputs it operates on.
"""
import numpy as np


def grids_and_strides(size=640, strides=(8, 16, 32)):
    gs, ss = [], []
    for s in strides:
        n = size // s
        xv, yv = np.meshgrid(np.arange(n), np.arange(n))
        gs.append(np.stack((xv, yv), 2).reshape(-1, 2))
        ss.append(np.full((n * n, 1), s))
    return np.concatenate(gs, 0).astype(np.float32), np.concatenate(ss, 0).astype(np.float32)


def _nms(boxes, scores, thr):
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = scores.argsort()[::-1]
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(i)
        xx1 = np.maximum(x1[i], x1[order[1:]])
        yy1 = np.maximum(y1[i], y1[order[1:]])
        xx2 = np.minimum(x2[i], x2[order[1:]])
        yy2 = np.minimum(y2[i], y2[order[1:]])
        w = np.maximum(0.0, xx2 - xx1 + 1)
        h = np.maximum(0.0, yy2 - yy1 + 1)
        inter = w * h
        ovr = inter / (areas[i] + areas[order[1:]] - inter)
        order = order[np.where(ovr <= thr)[0] + 1]
    return keep


def yolox_postprocess(pred, ratio, size=640, score_thr=0.7, nms_thr=0.45):
    """pred float32[A, 5+nc] (obj/cls already sigmoid-ed). Returns (boxes f32[K,4] xyxy, scores f32[K], cls int[K])
    sorted by score descending within the concatenation of classes."""
    pred = np.asarray(pred, dtype=np.float32).copy()
    grids, strides = grids_and_strides(size)
    pred[:, :2] = (pred[:, :2] + grids) * strides
    pred[:, 2:4] = np.exp(pred[:, 2:4]) * strides
    boxes = pred[:, :4]
    scores = pred[:, 4:5] * pred[:, 5:]
    xyxy = np.ones_like(boxes)
    xyxy[:, 0] = boxes[:, 0] - boxes[:, 2] / 2.0
    xyxy[:, 1] = boxes[:, 1] - boxes[:, 3] / 2.0
    xyxy[:, 2] = boxes[:, 0] + boxes[:, 2] / 2.0
    xyxy[:, 3] = boxes[:, 1] + boxes[:, 3] / 2.0
    xyxy /= np.float32(ratio)
    ob, osc, ocl = [], [], []
    for c in range(scores.shape[1]):
        sc = scores[:, c]
        m = sc > score_thr
        if m.sum() == 0:
            continue
        vb, vs = xyxy[m], sc[m]
        keep = _nms(vb, vs, nms_thr)
        ob.append(vb[keep]); osc.append(vs[keep]); ocl.append(np.full(len(keep), c))
    if not ob:
        return np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), np.zeros((0,), np.int64)
    b, s, c = np.concatenate(ob), np.concatenate(osc), np.concatenate(ocl)
    o = np.argsort(-s, kind="stable")
    return b[o], s[o], c[o]


def wrapper_rows(boxes_xyxy, width, height, first_id, conf=1.0, category_id=1.0):
    """rtmlib_api.py:31-46 + oc_sort_api.py:33-47: clip (coordinates.py:270-295), ltwh, back to ltrb; float32 boxes."""
    rows = []
    for k, bb in enumerate(np.asarray(boxes_xyxy, dtype=np.float32)):
        bb = bb.copy()
        bb[0] = max(0, min(bb[0], width - 2))
        bb[1] = max(0, min(bb[1], height - 2))
        bb[2] = max(1, min(bb[2], width - 1))
        bb[3] = max(1, min(bb[3], height - 1))
        ltwh = np.array([bb[0], bb[1], bb[2] - bb[0], bb[3] - bb[1]])
        ltrb = np.array([ltwh[0], ltwh[1], ltwh[0] + ltwh[2], ltwh[1] + ltwh[3]])
        rows.append(np.array([*ltrb, conf, category_id, first_id + k]))
    return np.asarray(rows, dtype=np.float64).reshape(-1, 7)
