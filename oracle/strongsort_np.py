"""StrongSORT (DeepSORT lineage) per-frame association oracle (test infrastructure; never imported by tracklab_b200).

Restates, with externally supplied appearance features (the in-tracker ReID forward of strong_sort.py:135-145 is a
separate stage in this repo):
  /root/reference/plugins/track/strong_sort/strong_sort.py:41-85,88-121            (update, box conversions, output rule)
  /root/reference/plugins/track/strong_sort/sort/tracker.py:53-59,80-115,151-193   (predict, update, _match, _initiate_track)
  /root/reference/plugins/track/strong_sort/sort/track.py:65-95,97-108,245-322     (Track life cycle, EMA feature)
  /root/reference/plugins/track/strong_sort/sort/kalman_filter.py:47-214            (x/y/a/h-scaled noise, confidence-scaled R)
  /root/reference/plugins/track/strong_sort/sort/nn_matching.py:30-49,73-91,127-161 (cosine metric, gallery with budget)
  /root/reference/plugins/track/strong_sort/sort/linear_assignment.py:11-72,75-128,131-174 (matching, gating, fusion)
  /root/reference/plugins/track/strong_sort/sort/iou_matching.py:7-82               (IoU cost)
and the wrapper filter /root/reference/tracklab/wrappers/track/strong_sort_api.py:66-93 (``ecc`` off: camera compensation is
out of scope). ``max_unmatched_preds`` must be 0 as in the reference YAML (strong_sort.yaml:19, q10).

Quirks kept: float32 detections (sort/detection.py:34-36) make ``initiate`` return a float32 mean AND covariance; the
gallery keeps the last ``budget`` EMA features of every confirmed track (one appended per frame, nn_matching.py:127-142)
and the appearance distance is the minimum over it in float32; output boxes are ``int()``-truncated and clipped, tracks
are reported while ``time_since_update <= 1`` with the detection id of their LAST update (strong_sort.py:70-82).
"""
import numpy as np
import scipy.linalg
from scipy.optimize import linear_sum_assignment

from .boxes_np import iou_tlwh_one_to_many

W_POS, W_VEL = 1.0 / 20, 1.0 / 160
INFTY = 1e5
CHI2_4 = 9.4877
TENTATIVE, CONFIRMED, DELETED = 1, 2, 3

_F = np.eye(8, 8)
for _i in range(4):
    _F[_i, 4 + _i] = 1.0
_H = np.eye(4, 8)


def kf_initiate(z):  # kalman_filter.py:47-78
    mean = np.r_[z, np.zeros_like(z)]
    std = [2 * W_POS * z[0], 2 * W_POS * z[1], 1 * z[2], 2 * W_POS * z[3],
           10 * W_VEL * z[0], 10 * W_VEL * z[1], 0.1 * z[2], 10 * W_VEL * z[3]]
    return mean, np.diag(np.square(std))


def kf_predict(mean, cov):  # kalman_filter.py:80-112
    std_pos = [W_POS * mean[0], W_POS * mean[1], 1 * mean[2], W_POS * mean[3]]
    std_vel = [W_VEL * mean[0], W_VEL * mean[1], 0.1 * mean[2], W_VEL * mean[3]]
    q = np.diag(np.square(np.r_[std_pos, std_vel]))
    return np.dot(_F, mean), np.linalg.multi_dot((_F, cov, _F.T)) + q


def kf_project(mean, cov, confidence=0.0):  # kalman_filter.py:114-144
    std = [W_POS * mean[3], W_POS * mean[3], 1e-1, W_POS * mean[3]]
    std = [(1 - confidence) * x for x in std]
    r = np.diag(np.square(std))
    return np.dot(_H, mean), np.linalg.multi_dot((_H, cov, _H.T)) + r


def kf_update(mean, cov, z, confidence=0.0):  # kalman_filter.py:146-174
    pm, pc = kf_project(mean, cov, confidence)
    chol, lower = scipy.linalg.cho_factor(pc, lower=True, check_finite=False)
    gain = scipy.linalg.cho_solve((chol, lower), np.dot(cov, _H.T).T, check_finite=False).T
    return mean + np.dot(z - pm, gain.T), cov - np.linalg.multi_dot((gain, pc, gain.T))


def kf_gating(mean, cov, zs):  # kalman_filter.py:176-214
    pm, pc = kf_project(mean, cov)
    d = zs - pm
    chol = np.linalg.cholesky(pc)
    z = scipy.linalg.solve_triangular(chol, d.T, lower=True, check_finite=False, overwrite_b=True)
    return np.sum(z * z, axis=0)


class _Det:  # sort/detection.py:33-52
    def __init__(self, tlwh, conf, feat):
        self.tlwh = np.asarray(tlwh, dtype=np.float32)
        self.confidence = float(conf)
        self.feature = np.asarray(feat, dtype=np.float32)

    def xyah(self):
        r = self.tlwh.copy()
        r[:2] += r[2:] / 2
        r[2] /= r[3]
        return r


class _Trk:  # sort/track.py
    def __init__(self, z, tid, cls, conf, feat, det_id):
        self.id, self.cls, self.conf, self.det_id = tid, int(cls), conf, det_id
        self.hits, self.age, self.tsu, self.state = 1, 1, 0, TENTATIVE
        feat /= np.linalg.norm(feat)  # in place on the detection's array (track.py:84)
        self.feat = feat
        self.mean, self.cov = kf_initiate(z)

    def tlwh(self):
        r = self.mean[:4].copy()
        r[2] *= r[3]
        r[:2] -= r[2:] / 2
        return r

    def tlbr(self):  # track.py:114-126
        r = self.tlwh()
        r[2:] = r[:2] + r[2:]
        return r


def _min_cost(cost, max_distance, t_idx, d_idx):  # linear_assignment.py:11-72 (cost already built)
    cost[cost > max_distance] = max_distance + 1e-5
    rows, cols = linear_sum_assignment(cost)
    un_d = [d for c, d in enumerate(d_idx) if c not in cols]
    un_t = [t for r, t in enumerate(t_idx) if r not in rows]
    pairs = []
    for r, c in zip(rows, cols):
        if cost[r, c] > max_distance:
            un_t.append(t_idx[r])
            un_d.append(d_idx[c])
        else:
            pairs.append((t_idx[r], d_idx[c]))
    return pairs, un_t, un_d


class StrongSortOracle:
    def __init__(self, max_dist=0.1594374041012136, max_iou_dist=0.5431835667667874, max_age=40, max_unmatched_preds=0,
                 n_init=3, nn_budget=100, mc_lambda=0.995, ema_alpha=0.8962157769329083, min_confidence=0.4,
                 image_size=(1920, 1080)):
        assert max_unmatched_preds == 0, "only the reference configuration (max_unmatched_preds: 0) is restated"
        self.max_dist, self.max_iou_dist, self.max_age, self.n_init = max_dist, max_iou_dist, max_age, n_init
        self.budget, self.lam, self.alpha, self.min_confidence = nn_budget, mc_lambda, ema_alpha, min_confidence
        self.width, self.height = image_size
        self.tracks, self.samples, self.next_id = [], {}, 1

    # ---- metric (nn_matching.py) ------------------------------------------------------------------
    def _appearance(self, t_idx, feats):
        cost = np.zeros((len(t_idx), len(feats)))
        b = feats / np.linalg.norm(feats, axis=1, keepdims=True)
        for r, k in enumerate(t_idx):
            x = np.asarray(self.samples[self.tracks[k].id])
            a = x / np.linalg.norm(x, axis=1, keepdims=True)
            cost[r, :] = (1.0 - np.dot(a, b.T)).min(axis=0)
        return cost

    # ---- camera motion compensation (tracker.py:66-68, track.py:216-239) ------------------------------
    def camera_update(self, warp):
        """warp: the 2x3 float32 matrix Track.ECC returned for (previous frame, this frame), or None ("ecc transform failed" /
        no previous frame). The reference recomputes it per track; it only depends on the frame pair."""
        if warp is None or not np.all(np.isfinite(warp)):
            return
        a, b = np.asarray(warp, dtype=np.float32).reshape(2, 3)
        matrix = np.array([a, b, [0, 0, 1]]).tolist()
        eye = np.eye(3)
        matrix = matrix if np.linalg.norm(eye - matrix) < 100 else eye
        for t in self.tracks:
            x1, y1, x2, y2 = t.tlbr()
            x1_, y1_, _ = matrix @ np.array([x1, y1, 1]).T
            x2_, y2_, _ = matrix @ np.array([x2, y2, 1]).T
            w, h = x2_ - x1_, y2_ - y1_
            cx, cy = x1_ + w / 2, y1_ + h / 2
            t.mean[:4] = [cx, cy, w / h, h]

    # ---- one frame (strong_sort.py:41-85) -----------------------------------------------------------
    def update(self, dets7, feats):
        dets7 = np.asarray(dets7, dtype=np.float64).reshape(-1, 7)
        xyxy = dets7[:, :4]
        xywh = np.empty_like(xyxy)
        xywh[..., 0] = (xyxy[..., 0] + xyxy[..., 2]) / 2
        xywh[..., 1] = (xyxy[..., 1] + xyxy[..., 3]) / 2
        xywh[..., 2] = xyxy[..., 2] - xyxy[..., 0]
        xywh[..., 3] = xyxy[..., 3] - xyxy[..., 1]
        tlwh = xywh.copy()
        tlwh[:, 0] = xywh[:, 0] - xywh[:, 2] / 2.0
        tlwh[:, 1] = xywh[:, 1] - xywh[:, 3] / 2.0
        conf, cls, ids = dets7[:, 4], dets7[:, 5], dets7[:, 6]
        dets = [_Det(tlwh[i], c, np.array(feats[i], dtype=np.float32)) for i, c in enumerate(conf)]

        for t in self.tracks:  # tracker.predict (tracker.py:53-59, track.py:245-249)
            t.mean, t.cov = kf_predict(t.mean, t.cov)
            t.age += 1
            t.tsu += 1

        # ---- _match (tracker.py:151-187)
        confirmed = [i for i, t in enumerate(self.tracks) if t.state == CONFIRMED]
        unconfirmed = [i for i, t in enumerate(self.tracks) if t.state != CONFIRMED]
        all_d = list(range(len(dets)))
        if len(all_d) == 0 or len(confirmed) == 0:
            pairs_a, un_d = [], all_d
        else:
            cost = self._appearance(confirmed, np.array([d.feature for d in dets]))
            zs = np.asarray([d.xyah() for d in dets])
            for r, k in enumerate(confirmed):  # gate_cost_matrix (linear_assignment.py:131-174)
                g = kf_gating(self.tracks[k].mean, self.tracks[k].cov, zs)
                cost[r, g > CHI2_4] = INFTY
                cost[r] = self.lam * cost[r] + (1 - self.lam) * g
            pairs_a, _, un_d = _min_cost(cost, self.max_dist, confirmed, all_d)
        un_t_a = list(set(confirmed) - set(k for k, _ in pairs_a))
        cand = unconfirmed + [k for k in un_t_a if self.tracks[k].tsu == 1]
        un_t_a = [k for k in un_t_a if self.tracks[k].tsu != 1]
        if len(un_d) == 0 or len(cand) == 0:
            pairs_b, un_t_b = [], cand
        else:
            cost = np.zeros((len(cand), len(un_d)))  # iou_cost (iou_matching.py:42-82)
            boxes = np.asarray([dets[i].tlwh for i in un_d])
            for r, k in enumerate(cand):
                if self.tracks[k].tsu > 1:
                    cost[r, :] = INFTY
                    continue
                cost[r, :] = 1.0 - iou_tlwh_one_to_many(self.tracks[k].tlwh(), boxes)
            pairs_b, un_t_b, un_d = _min_cost(cost, self.max_iou_dist, cand, un_d)
        pairs = pairs_a + pairs_b
        un_t = list(set(un_t_a + un_t_b))

        # ---- Tracker.update (tracker.py:80-115)
        for k, j in pairs:  # Track.update (track.py:272-301)
            t, d = self.tracks[k], dets[j]
            t.conf, t.cls = conf[j], int(cls[j])
            t.mean, t.cov = kf_update(t.mean, t.cov, d.xyah(), d.confidence)
            f = d.feature / np.linalg.norm(d.feature)
            s = self.alpha * t.feat + (1 - self.alpha) * f
            s /= np.linalg.norm(s)
            t.feat = s
            t.hits += 1
            t.tsu = 0
            if t.state == TENTATIVE and t.hits >= self.n_init:
                t.state = CONFIRMED
            t.det_id = ids[j]
        for k in un_t:  # mark_missed (track.py:303-309)
            t = self.tracks[k]
            if t.state == TENTATIVE or t.tsu > self.max_age:
                t.state = DELETED
        for j in un_d:
            d = dets[j]
            self.tracks.append(_Trk(d.xyah(), self.next_id, cls[j].item(), conf[j].item(), d.feature, ids[j].item()))
            self.next_id += 1
        self.tracks = [t for t in self.tracks if t.state != DELETED]
        active = [t.id for t in self.tracks if t.state == CONFIRMED]
        for t in self.tracks:  # partial_fit (nn_matching.py:127-142)
            if t.state != CONFIRMED:
                continue
            self.samples.setdefault(t.id, []).append(t.feat)
            self.samples[t.id] = self.samples[t.id][-self.budget:]
        self.samples = {k: self.samples[k] for k in active}

        rows = []
        for t in self.tracks:  # strong_sort.py:70-82,110-121
            if t.state != CONFIRMED or t.tsu > 1:
                continue
            x, y, w, h = t.tlwh()
            x1, x2 = max(int(x), 0), min(int(x + w), self.width - 1)
            y1, y2 = max(int(y), 0), min(int(y + h), self.height - 1)
            rows.append([x1, y1, x2, y2, t.id, t.cls, t.conf, t.det_id])
        return np.asarray(rows, dtype=np.float64).reshape(-1, 8)

    def run_video(self, dets, offsets, feats, warps=None):
        """Wrapper semantics of strong_sort_api.py:59-93; ``feats`` float32 [N,E] aligned with ``dets`` rows; ``warps`` [F,6]
        (cfg.ecc): the ECC matrix of (frame f-1, frame f), NaN rows = none."""
        out, fr = [], []
        for f in range(len(offsets) - 1):
            if warps is not None:
                self.camera_update(warps[f])
            d = dets[offsets[f]:offsets[f + 1]]
            if len(d) == 0:
                continue
            keep = d[:, 4] > self.min_confidence
            r = self.update(d[keep], feats[offsets[f]:offsets[f + 1]][keep])
            out.append(r)
            fr.append(np.full(len(r), f, dtype=np.int32))
        if not out:
            return np.zeros((0, 8)), np.zeros((0,), dtype=np.int32)
        return np.concatenate(out), np.concatenate(fr)
