"""HOTA (Luiten et al. 2020) for one sequence (test infrastructure; never imported by tracklab_b200).

Restates HOTA.eval_sequence of the TrackEval fork vendored in the reference
(/root/reference/plugins/eval/PoseTrack21/posetrack21/posetrack21/trackeval/metrics/hota.py:28-154,205-221; the pinned
sn-trackeval 0.4.0 the evaluator uses is not in the tree): DetA / AssA / LocA / HOTA over alpha = 0.05 .. 0.95, box IoU as the
similarity. Pinned to that class by tests/test_oracle_cpu.py (build container) and by tests/golden/hota_case.npz.
"""
import numpy as np
from scipy.optimize import linear_sum_assignment

ALPHAS = np.arange(0.05, 0.99, 0.05)


def iou_ltwh(a, b):
    """a [N,4], b [M,4] ltwh -> IoU [N,M]."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    ax1, ay1, ax2, ay2 = a[:, 0], a[:, 1], a[:, 0] + a[:, 2], a[:, 1] + a[:, 3]
    bx1, by1, bx2, by2 = b[:, 0], b[:, 1], b[:, 0] + b[:, 2], b[:, 1] + b[:, 3]
    iw = np.clip(np.minimum(ax2[:, None], bx2[None]) - np.maximum(ax1[:, None], bx1[None]), 0, None)
    ih = np.clip(np.minimum(ay2[:, None], by2[None]) - np.maximum(ay1[:, None], by1[None]), 0, None)
    inter = iw * ih
    union = (a[:, 2] * a[:, 3])[:, None] + (b[:, 2] * b[:, 3])[None] - inter
    return np.where(union > 0, inter / np.maximum(union, 1e-300), 0.0)


def box_ious_xywh(b1, b2):
    """_BaseDataset._calculate_box_ious(box_format='xywh') of the vendored TrackEval fork
    (/root/reference/plugins/eval/PoseTrack21/posetrack21/posetrack21/trackeval/datasets/_base_dataset.py:244-282), the similarity
    its MOT dataset class feeds HOTA with (posetrack_mot.py:479)."""
    eps = np.finfo("float").eps
    b1, b2 = np.array(b1, dtype=float).reshape(-1, 4), np.array(b2, dtype=float).reshape(-1, 4)
    b1[:, 2] = b1[:, 0] + b1[:, 2]; b1[:, 3] = b1[:, 1] + b1[:, 3]
    b2[:, 2] = b2[:, 0] + b2[:, 2]; b2[:, 3] = b2[:, 1] + b2[:, 3]
    mn = np.minimum(b1[:, None, :], b2[None, :, :]); mx = np.maximum(b1[:, None, :], b2[None, :, :])
    inter = np.maximum(mn[..., 2] - mx[..., 0], 0) * np.maximum(mn[..., 3] - mx[..., 1], 0)
    a1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    a2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    union = a1[:, None] + a2[None, :] - inter
    inter[a1 <= 0 + eps, :] = 0
    inter[:, a2 <= 0 + eps] = 0
    inter[union <= 0 + eps] = 0
    union[union <= 0 + eps] = 1
    return inter / union


def hota_from_boxes(gt_boxes, gt_ids, gt_off, tr_boxes, tr_ids, tr_off):
    """Frame-major xywh boxes + contiguous ids + offsets -> hota_sequence (the inputs of tk_hota_sequence)."""
    g, t, s = [], [], []
    for f in range(len(gt_off) - 1):
        a, b = slice(gt_off[f], gt_off[f + 1]), slice(tr_off[f], tr_off[f + 1])
        g.append(np.asarray(gt_ids[a], dtype=int)); t.append(np.asarray(tr_ids[b], dtype=int))
        s.append(box_ious_xywh(gt_boxes[a], tr_boxes[b]))
    return hota_sequence(g, t, s)


def hota_sequence(gt_ids, tr_ids, sims):
    """gt_ids / tr_ids: per frame int arrays of contiguous ids (0..n-1); sims: per frame [n_gt_t, n_tr_t] similarities.
    Returns dict of arrays over ALPHAS (HOTA, DetA, AssA, LocA, HOTA_TP/FN/FP) like the reference's res."""
    nA = len(ALPHAS)
    n_gt = 1 + max((int(g.max()) for g in gt_ids if len(g)), default=-1)
    n_tr = 1 + max((int(t.max()) for t in tr_ids if len(t)), default=-1)
    res = {k: np.zeros(nA) for k in ("HOTA_TP", "HOTA_FN", "HOTA_FP", "LocA", "AssA", "AssRe", "AssPr")}
    num_tr, num_gt = sum(len(t) for t in tr_ids), sum(len(g) for g in gt_ids)
    if num_tr == 0 or num_gt == 0:
        res["HOTA_FN" if num_tr == 0 else "HOTA_FP"] = (num_gt if num_tr == 0 else num_tr) * np.ones(nA)
        res["LocA"] = np.ones(nA)
        return _final(res)
    pot = np.zeros((n_gt, n_tr)); gcnt = np.zeros((n_gt, 1)); tcnt = np.zeros((1, n_tr))
    for g, t, s in zip(gt_ids, tr_ids, sims):
        den = s.sum(0)[None, :] + s.sum(1)[:, None] - s
        si = np.zeros_like(s)
        m = den > 0 + np.finfo("float").eps
        si[m] = s[m] / den[m]
        pot[g[:, None], t[None, :]] += si
        gcnt[g] += 1
        tcnt[0, t] += 1
    align = pot / (gcnt + tcnt - pot)
    mc = [np.zeros_like(pot) for _ in ALPHAS]
    for g, t, s in zip(gt_ids, tr_ids, sims):
        if len(g) == 0:
            res["HOTA_FP"] += len(t)
            continue
        if len(t) == 0:
            res["HOTA_FN"] += len(g)
            continue
        rows, cols = linear_sum_assignment(-(align[g[:, None], t[None, :]] * s))
        for a, alpha in enumerate(ALPHAS):
            ok = s[rows, cols] >= alpha - np.finfo("float").eps
            r, c = rows[ok], cols[ok]
            n = len(r)
            res["HOTA_TP"][a] += n; res["HOTA_FN"][a] += len(g) - n; res["HOTA_FP"][a] += len(t) - n
            if n > 0:
                res["LocA"][a] += sum(s[r, c])
                mc[a][g[r], t[c]] += 1
    for a in range(nA):
        m = mc[a]
        res["AssA"][a] = np.sum(m * (m / np.maximum(1, gcnt + tcnt - m))) / np.maximum(1, res["HOTA_TP"][a])
        res["AssRe"][a] = np.sum(m * (m / np.maximum(1, gcnt))) / np.maximum(1, res["HOTA_TP"][a])
        res["AssPr"][a] = np.sum(m * (m / np.maximum(1, tcnt))) / np.maximum(1, res["HOTA_TP"][a])
    res["LocA"] = np.maximum(1e-10, res["LocA"]) / np.maximum(1e-10, res["HOTA_TP"])
    return _final(res)


def _final(res):
    res["DetRe"] = res["HOTA_TP"] / np.maximum(1, res["HOTA_TP"] + res["HOTA_FN"])
    res["DetPr"] = res["HOTA_TP"] / np.maximum(1, res["HOTA_TP"] + res["HOTA_FP"])
    res["DetA"] = res["HOTA_TP"] / np.maximum(1, res["HOTA_TP"] + res["HOTA_FN"] + res["HOTA_FP"])
    res["HOTA"] = np.sqrt(res["DetA"] * res["AssA"])
    return res


def hota_of_tracker_rows(video, rows, frames, id_col=4, det_col=7):
    """HOTA of tracker rows (one row per reported detection, ``det_col`` = detection id = row of ``video.dets``) against the
    generator's identities (``video.gt_identity``; false positives are not ground truth). Boxes: the detections themselves on
    the ground-truth side, the tracker's output boxes (x1y1x2y2 in columns 0..3) on the other."""
    gt_ids, tr_ids, sims = [], [], []
    uniq = {v: i for i, v in enumerate(np.unique(rows[:, id_col]))}
    first_det = int(video.dets[0, 6]) if len(video.dets) else 0
    for f in range(video.n_frames):
        sl = slice(video.offsets[f], video.offsets[f + 1])
        ident = video.gt_identity[sl]
        keep = ident >= 0
        d = video.dets[sl][keep]
        gbox = np.column_stack([d[:, 0], d[:, 1], d[:, 2] - d[:, 0], d[:, 3] - d[:, 1]]) if len(d) else np.zeros((0, 4))
        r = rows[frames == f]
        tbox = np.column_stack([r[:, 0], r[:, 1], r[:, 2] - r[:, 0], r[:, 3] - r[:, 1]]) if len(r) else np.zeros((0, 4))
        gt_ids.append(ident[keep].astype(int))
        tr_ids.append(np.array([uniq[v] for v in r[:, id_col]], dtype=int))
        sims.append(iou_ltwh(gbox, tbox))
    return hota_sequence(gt_ids, tr_ids, sims)
