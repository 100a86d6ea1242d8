"""BoT-SORT per-frame association oracle (test infrastructure; never imported by tracklab_b200). SURVEY.md 8f-2.

Restates
  /root/reference/plugins/track/bot_sort/bot_sort.py:15-240,243-485,487-545  (STrack, BoTSORT.update, list helpers)
  /root/reference/plugins/track/bot_sort/matching.py:37-48,72-89,127-195,198-233 (assignment, IoU / embedding distances, fusions)
  /root/reference/plugins/track/bot_sort/kalman_filter.py:23-268            (xywh filter)
and the wrapper filter /root/reference/tracklab/wrappers/track/bot_sort_api.py:63-65.
The in-tracker ReID network (`_get_features`, bot_sort.py:487-505) and the camera-motion estimator (`GMC.apply`, gmc.py) are INPUTS:
per-detection embeddings (float32) and one 2x3 warp per frame.

Behaviour of the plugin that decides ids / values and is kept on purpose:
  q1  detections are built from the CENTRE-form boxes but stored in the field the filter reads as tlwh (bot_sort.py:281,315-316), so the
      filter state is the centre shifted by half a box; xywh2xyxy on output (:469) undoes it. The low-score detections go through
      `tlbr_to_tlwh` on a centre-form box (:389-390), i.e. their width / height are (w - cx, h - cy).
  q2  a freshly initiated mean AND covariance are float32 (`_tlwh` is float32, kalman_filter.py:55-86 keeps NumPy scalars); the
      process noise of `multi_predict` is evaluated in float32 only when every track of the pool is still float32 (np.asarray of the
      list promotes otherwise) - i.e. in the second frame; `multi_gmc` / `update` move a track to float64.
  q3  the first association is the JDE fusion: lambda * cosine distance + (1 - lambda) * squared Mahalanobis distance, infeasible
      where the latter exceeds chi2inv95[4] (matching.py:165-176); the IoU / appearance minimum only serves the unconfirmed tracks.
  q4  tracks removed for age stay in `lost` one more frame (the removed list is extended after the subtraction, bot_sort.py:449-451).
"""
import numpy as np
import scipy.linalg
from scipy.spatial.distance import cdist

from .assign_np import lapjv_extended
from .boxes_np import iou_plus1_f32

NEW, TRACKED, LOST, REMOVED = 0, 1, 2, 4          # basetrack.py:5-10
CHI2INV95_4 = 9.4877
W_POS, W_VEL = 1.0 / 20, 1.0 / 160
_F = np.eye(8, 8)
for _i in range(4):
    _F[_i, 4 + _i] = 1.0
_H = np.eye(4, 8)


# ---- xywh Kalman filter (kalman_filter.py) ------------------------------------------------------------------------------------
def kf_initiate(z):  # :55-86 (z float32 -> float32 mean and covariance)
    mean = np.r_[z, np.zeros_like(z)]
    std = [2 * W_POS * z[2], 2 * W_POS * z[3], 2 * W_POS * z[2], 2 * W_POS * z[3],
           10 * W_VEL * z[2], 10 * W_VEL * z[3], 10 * W_VEL * z[2], 10 * W_VEL * z[3]]
    return mean, np.diag(np.square(std))


def kf_multi_predict(mean, cov):  # :155-192
    std_pos = [W_POS * mean[:, 2], W_POS * mean[:, 3], W_POS * mean[:, 2], W_POS * mean[:, 3]]
    std_vel = [W_VEL * mean[:, 2], W_VEL * mean[:, 3], W_VEL * mean[:, 2], W_VEL * mean[:, 3]]
    sqr = np.square(np.r_[std_pos, std_vel]).T
    q = np.asarray([np.diag(sqr[i]) for i in range(len(mean))])
    mean = np.dot(mean, _F.T)
    left = np.dot(_F, cov).transpose((1, 0, 2))
    return mean, np.dot(left, _F.T) + q


def kf_project(mean, cov):  # :126-153
    std = [W_POS * mean[2], W_POS * mean[3], W_POS * mean[2], W_POS * mean[3]]
    return np.dot(_H, mean), np.linalg.multi_dot((_H, cov, _H.T)) + np.diag(np.square(std))


def kf_update(mean, cov, z):  # :194-226
    pm, pc = kf_project(mean, cov)
    chol, lower = scipy.linalg.cho_factor(pc, lower=True, check_finite=False)
    gain = scipy.linalg.cho_solve((chol, lower), np.dot(cov, _H.T).T, check_finite=False).T
    return mean + np.dot(z - pm, gain.T), cov - np.linalg.multi_dot((gain, pc, gain.T))


def kf_gating(mean, cov, zs):  # :228-268, metric='maha'
    pm, pc = kf_project(mean, cov)
    d = zs - pm
    chol = np.linalg.cholesky(pc)
    z = scipy.linalg.solve_triangular(chol, d.T, lower=True, check_finite=False, overwrite_b=True)
    return np.sum(z * z, axis=0)


class _Rec:
    """One STrack (bot_sort.py:15-240)."""

    def __init__(self, box, score, cls, feat, det_id):
        self.box32 = np.asarray(box, dtype=np.float32)
        self.mean = self.cov = None
        self.activated = False
        self.state = NEW
        self.cls = -1
        self.cls_hist = []
        self.update_cls(cls, score)
        self.score = score
        self.tracklet_len = 0
        self.smooth = self.curr = None
        if feat is not None:
            self.update_features(feat)
        self.det_id = det_id
        self.track_id = 0
        self.frame_id = self.start_frame = 0

    def update_features(self, feat):  # :43-51 (float32, in place on the caller's array)
        feat /= np.linalg.norm(feat)
        self.curr = feat
        if self.smooth is None:
            self.smooth = feat
        else:
            self.smooth = 0.9 * self.smooth + (1 - 0.9) * feat
        self.smooth /= np.linalg.norm(self.smooth)

    def update_cls(self, cls, score):  # :53-71
        if len(self.cls_hist) > 0:
            max_freq, found = 0, False
            for c in self.cls_hist:
                if cls == c[0]:
                    c[1] += score
                    found = True
                if c[1] > max_freq:
                    max_freq = c[1]
                    self.cls = c[0]
            if not found:
                self.cls_hist.append([cls, score])
                self.cls = cls
        else:
            self.cls_hist.append([cls, score])
            self.cls = cls

    def tlwh(self):  # :169-178
        if self.mean is None:
            return self.box32.copy()
        out = self.mean[:4].copy()
        out[:2] -= out[2:] / 2
        return out

    def tlbr(self):  # :181-187
        out = self.tlwh().copy()
        out[2:] += out[:2]
        return out

    def to_xywh(self):  # :209-218,220-221
        out = np.asarray(self.tlwh()).copy()
        out[:2] += out[2:] / 2
        return out


def _iou_dist(a, b):  # matching.py:72-89,198-233
    return 1 - iou_plus1_f32([t.tlbr() for t in a], [t.tlbr() for t in b])


def _emb_dist(tracks, dets):  # matching.py:127-145
    cost = np.zeros((len(tracks), len(dets)), dtype=np.float32)
    if cost.size == 0:
        return cost
    df = np.asarray([t.curr for t in dets], dtype=np.float32)
    tf = np.asarray([t.smooth for t in tracks], dtype=np.float32)
    return np.maximum(0.0, cdist(tf, df, "cosine"))


def _fuse_motion(cost, tracks, dets, lambda_):  # matching.py:165-176
    if cost.size == 0:
        return cost
    zs = np.asarray([d.to_xywh() for d in dets])
    for row, t in enumerate(tracks):
        g = kf_gating(t.mean, t.cov, zs)
        cost[row, g > CHI2INV95_4] = np.inf
        cost[row] = lambda_ * cost[row] + (1 - lambda_) * g
    return cost


def _fuse_score(cost, dets):  # matching.py:187-195
    if cost.size == 0:
        return cost
    sim = 1 - cost
    scores = np.array([d.score for d in dets])
    scores = np.expand_dims(scores, axis=0).repeat(cost.shape[0], axis=0)
    return 1 - sim * scores


def _assign(cost, thresh):  # matching.py:37-48
    if cost.size == 0:
        return [], list(range(cost.shape[0])), list(range(cost.shape[1]))
    x, y = lapjv_extended(cost, cost_limit=thresh)
    pairs = [(i, int(j)) for i, j in enumerate(x) if j >= 0]
    return pairs, list(np.where(x < 0)[0]), list(np.where(y < 0)[0])


def _union(a, b):  # joint_stracks, bot_sort.py:507-519
    seen, out = set(), []
    for t in a:
        seen.add(t.track_id)
        out.append(t)
    for t in b:
        if t.track_id not in seen:
            seen.add(t.track_id)
            out.append(t)
    return out


def _minus(a, b):  # sub_stracks, :522-530
    keep = {}
    for t in a:
        keep[t.track_id] = t
    for t in b:
        if keep.get(t.track_id, 0):
            del keep[t.track_id]
    return list(keep.values())


class BotSortOracle:
    def __init__(self, track_high_thresh=0.45, new_track_thresh=0.6, track_buffer=30, match_thresh=0.8, proximity_thresh=0.5,
                 appearance_thresh=0.25, cmc_method="sparseOptFlow", frame_rate=30, lambda_=0.985, min_confidence=0.4):
        self.high, self.new_thresh, self.match_thresh = track_high_thresh, new_track_thresh, match_thresh
        self.prox, self.app, self.lambda_ = proximity_thresh, appearance_thresh, lambda_
        self.max_time_lost = int(frame_rate / 30.0 * track_buffer)
        self.min_confidence = min_confidence
        self.tracked, self.lost, self.removed = [], [], []
        self.frame_id = 0
        self._next = 0                     # BaseTrack.clear_count() in every BoTSORT() (bot_sort.py:262)

    def _activate(self, t):  # :108-124
        self._next += 1
        t.track_id = self._next
        z = np.asarray(t.box32).copy()
        z[:2] += z[2:] / 2
        t.mean, t.cov = kf_initiate(z)
        t.tracklet_len = 0
        t.state = TRACKED
        if self.frame_id == 1:
            t.activated = True
        t.frame_id = t.start_frame = self.frame_id

    def _correct(self, t, det, reactivate):  # update :142-165 / re_activate :126-140
        t.mean, t.cov = kf_update(t.mean, t.cov, det.to_xywh())
        if det.curr is not None:
            t.update_features(det.curr)
        t.tracklet_len = 0 if reactivate else t.tracklet_len + 1
        t.state = TRACKED
        t.activated = True
        t.frame_id = self.frame_id
        t.score = det.score
        t.update_cls(det.cls, det.score)
        t.det_id = det.det_id

    @staticmethod
    def _gmc(tracks, H):  # multi_gmc :93-106
        if len(tracks) > 0:
            mm = np.asarray([t.mean.copy() for t in tracks])
            cc = np.asarray([t.cov for t in tracks])
            R8 = np.kron(np.eye(4, dtype=float), H[:2, :2])
            tr = H[:2, 2]
            for i, (m, c) in enumerate(zip(mm, cc)):
                m = R8.dot(m)
                m[:2] += tr
                tracks[i].mean, tracks[i].cov = m, R8.dot(c).dot(R8.transpose())

    def update(self, dets, embs, warp):
        """dets float64 [D,7] after the wrapper filter, embs float32 [D,E], warp float64 [2,3]; returns float64 [M,8]."""
        dets = np.asarray(dets, dtype=np.float64).reshape(-1, 7)
        self.frame_id += 1
        activated, refind, lost_now, removed_now = [], [], [], []
        xyxy = dets[:, :4]
        xywh = np.empty_like(xyxy)
        xywh[..., 0] = (xyxy[..., 0] + xyxy[..., 2]) / 2
        xywh[..., 1] = (xyxy[..., 1] + xyxy[..., 3]) / 2
        xywh[..., 2] = xyxy[..., 2] - xyxy[..., 0]
        xywh[..., 3] = xyxy[..., 3] - xyxy[..., 1]
        conf, cls, ids = dets[:, 4], dets[:, 5], dets[:, 6]
        high = conf > self.high
        second = np.logical_and(conf > 0.1, conf < self.high)
        feats = np.asarray(embs, dtype=np.float32)[high]
        d_high = [_Rec(b, s, c, f.copy(), i) for b, s, c, f, i in zip(xywh[high], conf[high], cls[high], feats, ids[high])]
        unconfirmed = [t for t in self.tracked if not t.activated]
        confirmed = [t for t in self.tracked if t.activated]

        pool = _union(confirmed, self.lost)
        if pool:
            mm = np.asarray([t.mean.copy() for t in pool])
            cc = np.asarray([t.cov for t in pool])
            for i, t in enumerate(pool):
                if t.state != TRACKED:
                    mm[i][6] = 0
                    mm[i][7] = 0
            mm, cc = kf_multi_predict(mm, cc)
            for t, m_, c_ in zip(pool, mm, cc):
                t.mean, t.cov = m_, c_
        H = np.asarray(warp, dtype=np.float64)
        self._gmc(pool, H)
        self._gmc(unconfirmed, H)

        cost = _fuse_motion(_emb_dist(pool, d_high), pool, d_high, self.lambda_)
        pairs, u_trk, u_det = _assign(cost, self.match_thresh)
        for it, idet in pairs:
            t = pool[it]
            if t.state == TRACKED:
                self._correct(t, d_high[idet], False)
                activated.append(t)
            else:
                self._correct(t, d_high[idet], True)
                refind.append(t)

        # second association with the low-score boxes (q1: tlbr_to_tlwh on a centre-form box)
        d_low = []
        for b, s, c, i in zip(xywh[second], conf[second], cls[second], ids[second]):
            bb = np.asarray(b).copy()
            bb[2:] -= bb[:2]
            d_low.append(_Rec(bb, s, c, None, i))
        rest = [pool[i] for i in u_trk if pool[i].state == TRACKED]
        pairs, u_trk2, _ = _assign(_iou_dist(rest, d_low), 0.5)
        for it, idet in pairs:
            t = rest[it]
            if t.state == TRACKED:
                self._correct(t, d_low[idet], False)
                activated.append(t)
            else:
                self._correct(t, d_low[idet], True)
                refind.append(t)
        for it in u_trk2:
            t = rest[it]
            if t.state != LOST:
                t.state = LOST
                lost_now.append(t)

        # unconfirmed tracks (bot_sort.py:412-432)
        left = [d_high[i] for i in u_det]
        iou_d = _iou_dist(unconfirmed, left)
        mask = iou_d > self.prox
        iou_d = _fuse_score(iou_d, left)
        emb_d = _emb_dist(unconfirmed, left) / 2.0
        emb_d[emb_d > self.app] = 1.0
        emb_d[mask] = 1.0
        cost = np.minimum(iou_d, emb_d)
        pairs, u_unc, u_det = _assign(cost, 0.7)
        for it, idet in pairs:
            self._correct(unconfirmed[it], left[idet], False)
            activated.append(unconfirmed[it])
        for it in u_unc:
            unconfirmed[it].state = REMOVED
            removed_now.append(unconfirmed[it])

        for i in u_det:  # births :434-441
            if left[i].score < self.new_thresh:
                continue
            self._activate(left[i])
            activated.append(left[i])

        for t in self.lost:  # :443-447
            if self.frame_id - t.frame_id > self.max_time_lost:
                t.state = REMOVED
                removed_now.append(t)
        self.tracked = [t for t in self.tracked if t.state == TRACKED]
        self.tracked = _union(self.tracked, activated)
        self.tracked = _union(self.tracked, refind)
        self.lost = _minus(self.lost, self.tracked)
        self.lost.extend(lost_now)
        self.lost = _minus(self.lost, self.removed)
        self.removed.extend(removed_now)
        pd = _iou_dist(self.tracked, self.lost)   # remove_duplicate_stracks :533-545
        da, db = [], []
        for p, q in zip(*np.where(pd < 0.15)):
            tp = self.tracked[p].frame_id - self.tracked[p].start_frame
            tq = self.lost[q].frame_id - self.lost[q].start_frame
            if tp > tq:
                db.append(q)
            else:
                da.append(p)
        self.tracked = [t for i, t in enumerate(self.tracked) if i not in da]
        self.lost = [t for i, t in enumerate(self.lost) if i not in db]

        rows = []
        for t in self.tracked:
            if not t.activated:
                continue
            b = t.tlwh()
            rows.append([b[0] - b[2] / 2, b[1] - b[3] / 2, b[0] + b[2] / 2, b[1] + b[3] / 2, t.track_id, t.cls, t.score, t.det_id])
        return np.asarray(rows, dtype=np.float64).reshape(-1, 8)

    def run_video(self, dets, offsets, embeddings, warps):
        """Wrapper semantics of bot_sort_api.py:60-66 (frames without detections are skipped)."""
        out, fr = [], []
        for f in range(len(offsets) - 1):
            sl = slice(offsets[f], offsets[f + 1])
            d = dets[sl]
            if len(d) == 0:
                continue
            keep = d[:, 4] > self.min_confidence
            r = self.update(d[keep], embeddings[sl][keep], warps[f])
            out.append(r)
            fr.append(np.full(len(r), f, dtype=np.int32))
        if not out:
            return np.zeros((0, 8)), np.zeros((0,), dtype=np.int32)
        return np.concatenate(out), np.concatenate(fr)
