"""BPBReID-StrongSORT per-frame association oracle (test infrastructure; never imported by tracklab_b200).

Restates
  /root/reference/plugins/track/bpbreid_strong_sort/strong_sort.py:53-141            (update, output rule, detection filter)
  /root/reference/plugins/track/bpbreid_strong_sort/sort/tracker.py:92-99,123-167,242-333,409-441 (predict, update,
        strong_sort_matching, matching information, _initiate_track)
  /root/reference/plugins/track/bpbreid_strong_sort/sort/track.py:66-92,112-195      (Track life cycle, visibility-aware EMA)
  /root/reference/plugins/track/bpbreid_strong_sort/sort/kalman_filter.py:47-227      (all noise proportional to the height)
  /root/reference/plugins/track/bpbreid_strong_sort/sort/nn_matching.py:99-135,171-209 (part-based distance, gallery)
  /root/reference/plugins/track/bpbreid_strong_sort/sort/linear_assignment.py:11-73,132-175
  /root/reference/plugins/track/bpbreid_strong_sort/sort/iou_matching.py:42-78
and the wrapper /root/reference/tracklab/wrappers/track/bpbreid_strong_sort_api.py:73-118 (frames without rows are skipped;
no confidence filter besides ``min_bbox_confidence``). Both matching strategies are restated (``strong_sort_matching``, the
reference YAML and what the device kernel implements, and the single-stage ``bot_sort_matching`` of tracker.py:335-363 with
``_full_cost_metric`` :169-240) for ``motion_criterium="iou"``.

PARITY UNPINNED for the appearance term: ``compute_distance_matrix_using_bp_features`` lives in the un-vendored torchreid
git dependency; it is restated (SURVEY.md §8c [3P-memory]) as the visibility-weighted mean over parts of the Euclidean
distance between L2-normalised part embeddings, halved (nn_matching.py:133). The goldens are produced by the unmodified
plugin running on the same restatement (oracle/ref_shims_torchreid).
"""
import numpy as np
import scipy.linalg
from scipy.optimize import linear_sum_assignment

from .boxes_np import iou_tlwh_one_to_many

W_POS, W_VEL = 1.0 / 20, 1.0 / 160
INFTY = 1e5
CHI2_4 = 9.4877

_F = np.eye(8, 8)
for _i in range(4):
    _F[_i, 4 + _i] = 1.0
_H = np.eye(4, 8)


def kf_initiate(z):  # kalman_filter.py:47-72
    mean = np.r_[z, np.zeros_like(z)]
    std = [2 * W_POS * z[3]] * 4 + [10 * W_VEL * z[3]] * 4
    return mean, np.diag(np.square(std))


def kf_predict(mean, cov):  # kalman_filter.py:74-104
    std_pos = [W_POS * mean[3]] * 4
    std_vel = [W_VEL * mean[3]] * 4
    q = np.diag(np.square(np.r_[std_pos, std_vel]))
    return np.dot(_F, mean), np.linalg.multi_dot((_F, cov, _F.T)) + q


def kf_project(mean, cov, confidence=0.0):  # kalman_filter.py:106-136
    std = [W_POS * mean[3]] * 4
    std = [(1 - confidence) * x for x in std]
    return np.dot(_H, mean), np.linalg.multi_dot((_H, cov, _H.T)) + np.diag(np.square(std))


def kf_update(mean, cov, z, confidence=0.0):  # kalman_filter.py:138-166
    pm, pc = kf_project(mean, cov, confidence)
    chol, lower = scipy.linalg.cho_factor(pc, lower=True, check_finite=False)
    gain = scipy.linalg.cho_solve((chol, lower), np.dot(cov, _H.T).T, check_finite=False).T
    return mean + np.dot(z - pm, gain.T), cov - np.linalg.multi_dot((gain, pc, gain.T))


def kf_gating(mean, cov, zs):  # kalman_filter.py:168-227 (only_position=False, 'maha')
    pm, pc = kf_project(mean, cov)
    d = zs - pm
    chol = np.linalg.cholesky(pc)
    z = scipy.linalg.solve_triangular(chol, d.T, lower=True, check_finite=False, overwrite_b=True)
    return np.sum(z * z, axis=0)


def part_distance(track_feat, track_vis, det_feats, det_vis):
    """nn_matching.py:99-135 on the restated torchreid function: float32 [D] distances of one track to D detections."""
    def norm(x):
        n = np.sqrt((x * x).sum(axis=-1, keepdims=True))
        return x / np.maximum(n, np.float32(1e-12))
    a = norm(track_feat.astype(np.float32))            # [K,E]
    b = norm(det_feats.astype(np.float32))             # [D,K,E]
    diff = a[None] - b
    d = np.sqrt((diff * diff).sum(axis=-1))            # [D,K]
    w = track_vis.astype(np.float32)[None] * det_vis.astype(np.float32)
    return ((d * w).sum(axis=1) / w.sum(axis=1)) / 2


class _Det:
    def __init__(self, det_id, ltwh, conf, feat, vis):
        self.id, self.ltwh, self.confidence = det_id, np.asarray(ltwh, dtype=float), float(conf)
        self.feat, self.vis = np.asarray(feat, dtype=np.float32), np.asarray(vis)
        self.matched_with = None

    def xyah(self):
        r = self.ltwh.copy()
        r[:2] += r[2:] / 2
        r[2] /= r[3]
        return r


class _Trk:
    def __init__(self, det, tid, cls, conf, n_init):
        self.id, self.cls, self.conf = tid, int(cls), conf
        self.hits, self.age, self.tsu = 1, 1, 0
        self.feat, self.vis = det.feat, det.vis
        self.mean, self.cov = kf_initiate(det.xyah())
        self.last_det = det
        self.state = "c" if self.hits >= n_init else "t"
        self.last_pred_ltwh = None

    def ltwh(self):
        m = self.mean[:4].copy()
        w = m[2] * m[3]
        return np.array([m[0] - w / 2, m[1] - m[3] / 2, w, m[3]])


def _min_cost(cost, max_distance, t_idx, d_idx):  # linear_assignment.py:11-73
    thr = cost.copy()
    thr[thr > max_distance] = max_distance + 1e-5
    rows, cols = linear_sum_assignment(thr)
    un_d = [d for c, d in enumerate(d_idx) if c not in cols]
    un_t = [t for r, t in enumerate(t_idx) if r not in rows]
    pairs = []
    for r, c in zip(rows, cols):
        if thr[r, c] > max_distance:
            un_t.append(t_idx[r])
            un_d.append(d_idx[c])
        else:
            pairs.append((t_idx[r], d_idx[c]))
    return pairs, un_t, un_d


class BpbreidStrongSortOracle:
    def __init__(self, ema_alpha=0.9, mc_lambda=0.995, max_dist=0.5, max_iou_distance=0.8, max_age=300, n_init=0,
                 min_bbox_confidence=0.0, max_kalman_prediction_without_update=7, matching_strategy="strong_sort_matching",
                 gating_thres_factor=1.0, w_kfgd=1.0, w_reid=1.0, w_st=1.0):
        self.strategy, self.gtf, self.w = matching_strategy, gating_thres_factor, (w_kfgd, w_reid, w_st)
        self.alpha, self.lam, self.max_dist, self.max_iou = ema_alpha, mc_lambda, max_dist, max_iou_distance
        self.max_age, self.n_init, self.min_conf, self.max_pred = max_age, n_init, min_bbox_confidence, max_kalman_prediction_without_update
        self.tracks, self.next_id = [], 1

    def update(self, ids, ltwh, feats, vis, confs, classes):
        """Returns rows float64 [M, 14] = [track_id, kf_ltwh(4), pred_kf_ltwh(4) (NaN for births), matched_code (0 none, 1 'R',
        2 'S'), matched_dist, hits, age, det_id] for confirmed tracks updated in this frame."""
        dets = [_Det(ids[i], ltwh[i], c, feats[i], vis[i]) for i, c in enumerate(confs)]
        dets = [d for d in dets if d.confidence > self.min_conf]
        # tracker.py:139-152 indexes the UNFILTERED class/confidence arrays with the filtered position; kept as is.
        cls_of = {id(d): classes[j] for j, d in enumerate(dets)}
        conf_of = {id(d): confs[j] for j, d in enumerate(dets)}
        for t in self.tracks:  # tracker.predict (track.py:128-135)
            if t.tsu < self.max_pred:
                t.mean, t.cov = kf_predict(t.mean, t.cov)
            t.age += 1
            t.tsu += 1
        if len(dets) > 0 and self.strategy == "bot_sort_matching":
            self._update_bot_sort(dets, cls_of, conf_of)
        elif len(dets) > 0:
            confirmed = [i for i, t in enumerate(self.tracks) if t.state == "c"]
            unconfirmed = [i for i, t in enumerate(self.tracks) if t.state != "c"]
            all_d = list(range(len(dets)))
            cost_a = None
            if len(confirmed) == 0:
                pairs_a, un_d = [], all_d
            else:
                df, dv = np.stack([d.feat for d in dets]), np.stack([d.vis for d in dets])
                cost_a = np.zeros((len(confirmed), len(dets)))
                zs = np.asarray([d.xyah() for d in dets])
                for r, k in enumerate(confirmed):
                    cost_a[r, :] = part_distance(self.tracks[k].feat, self.tracks[k].vis, df, dv)
                for r, k in enumerate(confirmed):  # gate_cost_matrix (linear_assignment.py:132-175)
                    g = kf_gating(self.tracks[k].mean, self.tracks[k].cov, zs)
                    cost_a[r, g > CHI2_4] = INFTY
                    cost_a[r] = self.lam * cost_a[r] + (1 - self.lam) * g
                pairs_a, _, un_d = _min_cost(cost_a, self.max_dist, confirmed, all_d)
            un_t_a = list(set(confirmed) - set(k for k, _ in pairs_a))
            cand = unconfirmed + [k for k in un_t_a if self.tracks[k].tsu == 1]
            un_t_a = [k for k in un_t_a if self.tracks[k].tsu != 1]
            cost_b, un_d_a = None, list(un_d)
            if len(un_d) == 0 or len(cand) == 0:
                pairs_b, un_t_b = [], cand
            else:
                cost_b = np.zeros((len(cand), len(un_d)))  # iou_cost (iou_matching.py:42-78)
                boxes = np.asarray([dets[i].ltwh for i in un_d])
                for r, k in enumerate(cand):
                    cost_b[r, :] = 1.0 - iou_tlwh_one_to_many(self.tracks[k].ltwh(), boxes)
                pairs_b, un_t_b, un_d = _min_cost(cost_b, self.max_iou, cand, un_d)
            # add_matching_information (tracker.py:409-421)
            for i, d in enumerate(dets):
                d.matched_with = None
            if cost_a is not None:
                m = {dd: tt for tt, dd in pairs_a}
                for i, di in enumerate(all_d):
                    dets[di].matched_with = ("R", cost_a[confirmed.index(m[di]), i]) if di in m else None
            if cost_b is not None:
                m = {dd: tt for tt, dd in pairs_b}
                for i, di in enumerate(un_d_a):
                    dets[di].matched_with = ("S", cost_b[cand.index(m[di]), i]) if di in m else None
            self._apply(dets, pairs_a + pairs_b, list(set(un_t_a + un_t_b)), un_d, cls_of, conf_of)
        rows = []
        for t in self.tracks:  # strong_sort.py:96-120
            if t.state != "c" or t.tsu > 0:
                continue
            d = t.last_det
            code, dist = (0, np.nan) if d.matched_with is None else ((1 if d.matched_with[0] == "R" else 2), d.matched_with[1])
            pred = t.last_pred_ltwh if t.last_pred_ltwh is not None else np.full(4, np.nan)
            rows.append([t.id, *t.ltwh(), *pred, code, dist, t.hits, t.age, float(d.id)])
        return np.asarray(rows, dtype=np.float64).reshape(-1, 14)

    def _update_bot_sort(self, dets, cls_of, conf_of):
        """tracker.py:335-363 + _full_cost_metric :169-240: one stage over ALL tracks on (w_kfgd * pos + w_reid * app + w_st * st) /
        sum(w), pos = sqrt(gating distance) / (sqrt(chi2) * gating_thres_factor); entries are voided where the position OR the
        appearance gate fails — np.logical_or(pos_gate, app_gate, st_gate) uses its third argument as the OUTPUT array, so the
        spatio-temporal gate is never applied (reproduced)."""
        w_k, w_r, w_s = self.w
        idx, all_d = list(range(len(self.tracks))), list(range(len(dets)))
        for d in dets:
            d.matched_with = None
        if not idx:
            return self._apply(dets, [], [], all_d, cls_of, conf_of)
        zs = np.asarray([d.xyah() for d in dets])
        df, dv = np.stack([d.feat for d in dets]), np.stack([d.vis for d in dets])
        boxes = np.asarray([d.ltwh for d in dets])
        thr = np.sqrt(CHI2_4)
        pos, app, st = (np.empty((len(idx), len(dets))) for _ in range(3))
        for r, k in enumerate(idx):
            t = self.tracks[k]
            pos[r] = np.sqrt(kf_gating(t.mean, t.cov, zs)) / (thr * self.gtf)
            app[r] = part_distance(t.feat, t.vis, df, dv)
            st[r] = 1.0 - iou_tlwh_one_to_many(t.ltwh(), boxes)
        cost = (w_k * pos + w_r * app + st * w_s) / (w_k + w_r + w_s)
        pos_gate = pos > 1.0 if w_k > 0 else np.zeros_like(pos, dtype=bool)
        app_gate = app > self.max_dist if w_r > 0 else np.zeros_like(app, dtype=bool)
        st_gate = st > self.max_iou if w_s > 0 else np.zeros_like(st, dtype=bool)
        if w_k > 0:
            cost[np.logical_or(pos_gate, app_gate)] = INFTY
        elif w_s > 0:
            cost[np.logical_or(app_gate, st_gate)] = INFTY
        else:
            cost[app_gate] = INFTY
        pairs, _, un_d = _min_cost(cost, self.max_dist, idx, all_d)
        un_t = list(set(idx) - set(k for k, _ in pairs))
        m = {dd: tt for tt, dd in pairs}
        for i, di in enumerate(all_d):
            dets[di].matched_with = ("R", cost[idx.index(m[di]), i]) if di in m else None
        self._apply(dets, pairs, un_t, un_d, cls_of, conf_of)

    def _apply(self, dets, pairs, unmatched_tracks, un_d, cls_of, conf_of):
        for k, j in pairs:  # Track.update (track.py:137-174)
            t, d = self.tracks[k], dets[j]
            t.conf, t.cls, t.last_det = conf_of[id(d)], int(cls_of[id(d)]), d
            t.last_pred_ltwh = t.ltwh()
            t.mean, t.cov = kf_update(t.mean, t.cov, d.xyah(), d.confidence)
            tv, dv_ = t.vis, d.vis
            xor = np.logical_xor(tv, dv_)
            et = (tv * dv_) * np.float32(self.alpha) + xor * tv
            ed = (tv * dv_) * np.float32(1 - self.alpha) + xor * dv_
            sm = np.expand_dims(et, 1) * t.feat + np.expand_dims(ed, 1) * d.feat
            sm[np.logical_and(et == 0.0, ed == 0.0)] = 1
            t.feat, t.vis = sm, np.maximum(tv, dv_)
            t.hits += 1
            t.tsu = 0
            if t.state == "t" and t.hits >= self.n_init:
                t.state = "c"
        for k in unmatched_tracks:  # mark_missed (track.py:181-187)
            t = self.tracks[k]
            if t.state == "t":
                t.state = "d"
            elif t.tsu > self.max_age:
                t.state = "d"
        for j in un_d:
            d = dets[j]
            self.tracks.append(_Trk(d, self.next_id, cls_of[id(d)], conf_of[id(d)], self.n_init))
            self.next_id += 1
        self.tracks = [t for t in self.tracks if t.state != "d"]

    def run_video(self, dets, offsets, feats, vis):
        """dets float64 [N,7] wrapper rows (l,t,r,b,conf,cls,id); feats float32 [N,K,E]; vis [N,K]."""
        out, fr = [], []
        for f in range(len(offsets) - 1):
            sl = slice(offsets[f], offsets[f + 1])
            d = dets[sl]
            if len(d) == 0:
                continue
            ltwh = np.column_stack([d[:, 0], d[:, 1], d[:, 2] - d[:, 0], d[:, 3] - d[:, 1]])
            r = self.update(d[:, 6], ltwh, feats[sl], vis[sl], d[:, 4], np.zeros(len(d)))
            out.append(r)
            fr.append(np.full(len(r), f, dtype=np.int32))
        if not out:
            return np.zeros((0, 14)), np.zeros((0,), dtype=np.int32)
        return np.concatenate(out), np.concatenate(fr)
