"""RT-DETR post-processing oracle (test infrastructure; never imported by tracklab_b200).

Restates transformers' ``RTDetrImageProcessor.post_process_object_detection`` (use_focal_loss=True; third-party, present in
the image as a pip package, pinned by tests/test_rtdetr_cpu.py against the package itself) followed by the reference
wrapper's loop /root/reference/tracklab/wrappers/bbox_detector/transformers_api.py:37-53 (label 0 only, running id) and
/root/reference/tracklab/utils/coordinates.py:270-295,318-328 (sanitize_bbox_ltrb + ltrb_to_ltwh on the float32 box).
"""
import numpy as np


def sigmoid_f32(x):
    x = np.asarray(x, dtype=np.float32)
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(-x))).astype(np.float32)


def post_process(logits, boxes, image_wh, threshold, keep_label=0):
    """logits float32 [Q,C], boxes float32 [Q,4] (cxcywh relative) -> rows float64 [k,6] = [l,t,w,h,score,query] by descending
    score (ties: ascending flat index)."""
    logits = np.asarray(logits, dtype=np.float32)
    boxes = np.asarray(boxes, dtype=np.float32)
    Q, C = logits.shape
    W, H = np.float32(image_wh[0]), np.float32(image_wh[1])
    half = np.float32(0.5)
    corners = np.stack([boxes[:, 0] - half * boxes[:, 2], boxes[:, 1] - half * boxes[:, 3],
                        boxes[:, 0] + half * boxes[:, 2], boxes[:, 1] + half * boxes[:, 3]], axis=1) * np.array([W, H, W, H], dtype=np.float32)
    scores = sigmoid_f32(logits).reshape(-1)
    order = np.lexsort((np.arange(scores.size), -scores.astype(np.float64)))[:Q]   # top-Q, score desc, flat index asc
    rows = []
    for idx in order:
        s = scores[idx]
        label, q = idx % C, idx // C
        if not (s > np.float32(threshold)) or (keep_label >= 0 and label != keep_label):
            continue
        b = corners[q].copy()   # float32 array, sanitised in place like coordinates.py:288-292
        b[0] = max(0, min(b[0], int(image_wh[0]) - 2)); b[1] = max(0, min(b[1], int(image_wh[1]) - 2))
        b[2] = max(1, min(b[2], int(image_wh[0]) - 1)); b[3] = max(1, min(b[3], int(image_wh[1]) - 1))
        ltwh = np.array([b[0], b[1], b[2] - b[0], b[3] - b[1]])
        rows.append([*ltwh.astype(np.float64), float(s), float(q if keep_label >= 0 else idx)])
    return np.asarray(rows, dtype=np.float64).reshape(-1, 6)
