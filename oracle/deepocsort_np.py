"""Deep OC-SORT per-frame association oracle (test infrastructure; never imported by tracklab_b200). SURVEY.md 8f-1.

Restates
  /root/reference/plugins/track/deep_oc_sort/ocsort.py:22-93,96-304,324-542  (helpers, KalmanBoxTracker, OCSort.update)
  /root/reference/plugins/track/deep_oc_sort/association.py:202-212,263-360  (linear_assignment, adaptive weighting, associate)
  /root/reference/plugins/track/deep_oc_sort/kalmanfilter.py:340-379,383-481,483-569 (predict, freeze, affine correction, ORU, update)
and the wrapper filter /root/reference/tracklab/wrappers/track/deep_oc_sort_api.py:63-67.
The in-tracker ReID network (ocsort.py:560-571) and the camera-motion estimator (cmc.py) are INPUTS here: per-detection embeddings
(float32) and one 2x3 affine per frame, exactly what `_get_features` / `CMCComputer.compute_affine` hand to `update`.

Behaviour of the reference that decides ids and is therefore kept operation by operation:
  q1  `linear_assignment` lists `[y[i], i] for i in x` WITHOUT dropping unassigned rows (association.py:207): every detection the
      solver left unassigned contributes the pair `[y[-1], -1]`, i.e. (detection of the LAST tracker, tracker -1). NumPy's negative
      indexing then re-validates that pair against the IoU threshold, so the last tracker is updated once more per unassigned
      detection (KF update, hit counters, embedding EMA); when that re-validation fails, `-1` enters the unmatched lists. (Rows stay
      unassigned only when detections outnumber trackers, and then every tracker - the last one included - holds a detection in the
      optimum of the extension, so `y[-1]` itself is never -1; the code below would handle it like NumPy does anyway.)
  q2  `last_observation` and `observations[age]` are the same array object, and both are row views of that frame's detection array:
      `apply_affine_correction` (ocsort.py:252-268) therefore warps the most recent observation TWICE whenever it is at most
      `delta_t` frames old.
  q3  `frame_count` is never incremented by `update` (only by the unused `update_public`), so `frame_count <= min_hits` always
      holds and every track updated in the frame is reported; the reported confidence is the one of the detection that created the track.
  q4  the ORU replay (kalmanfilter.py:432-481) reads the (x, y, w, h) measurements of the 8-d filter as (x, y, s, r), runs with the
      filter's default R = I4 / Q = I8, and the measurement noise of the real update is computed from the state BEFORE the replay.
  q5  the first-round VDC term is multiplied by the class column and the first round always uses plain IoU (as in OC-SORT).
  q6  the embedding EMA runs in float32 on torch tensors (`alpha * emb + (1 - alpha) * det_emb`, then `/= np.linalg.norm`), the
      appearance matrix is a float32 NumPy matmul, the adaptive weights are float32.
"""
import numpy as np
import torch

from .assign_np import lapjv_extended
from .ocsort_np import ASSO, _direction
from .boxes_np import iou_xyxy

_F8 = np.array([[1, 0, 0, 0, 1, 0, 0, 0], [0, 1, 0, 0, 0, 1, 0, 0], [0, 0, 1, 0, 0, 0, 1, 0], [0, 0, 0, 1, 0, 0, 0, 1],
                [0, 0, 0, 0, 1, 0, 0, 0], [0, 0, 0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 0, 0, 1]])
_H8 = np.array([[1, 0, 0, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0, 0, 0]])
_I8 = np.eye(8)


def process_noise(w, h, p=1 / 20, v=1 / 160):  # ocsort.py:82-86
    return np.diag(((p * w) ** 2, (p * h) ** 2, (p * w) ** 2, (p * h) ** 2, (v * w) ** 2, (v * h) ** 2, (v * w) ** 2, (v * h) ** 2))


def measurement_noise(w, h, m=1 / 20):  # ocsort.py:89-93
    wv, hv = (m * w) ** 2, (m * h) ** 2
    return np.diag((wv, hv, wv, hv))


def box_to_z(b):  # ocsort.py:48-53
    w = b[2] - b[0]
    h = b[3] - b[1]
    return np.array([b[0] + w / 2.0, b[1] + h / 2.0, w, h]).reshape((4, 1))


def x_to_box(x):  # ocsort.py:56-58
    cx, cy, w, h = x.reshape(-1)[:4]
    return np.array([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2]).reshape(1, 4)


class _Trk:
    """KalmanBoxTracker(new_kf=True) with its KalmanFilterNew (ocsort.py:96-304, kalmanfilter.py)."""

    def __init__(self, bbox5, cls, delta_t, emb, det_id, uid):
        self.cls, self.conf = cls, bbox5[-1]
        z = box_to_z(bbox5)
        w, h = z.reshape(-1)[2:4]
        self.P = process_noise(w, h)
        self.P[:4, :4] *= 4
        self.P[4:, 4:] *= 100
        self.x = np.zeros((8, 1))
        self.x[:4] = z
        self.tsu, self.id, self.hits, self.streak, self.age = 0, uid, 0, 0, 0
        self.last_obs = np.array([-1, -1, -1, -1, -1])
        self.obs = {}
        self.velocity = None
        self.delta_t = delta_t
        self.emb = emb
        self.frozen_flag = False            # KalmanBoxTracker.frozen
        self.det_id = det_id
        # KalmanFilterNew bookkeeping: history of measurements, observed flag, frozen copy, last measurement
        self.hist = []
        self.observed = False
        self.saved = None                   # dict(x, P, n_hist, last_meas) = attr_saved
        self.last_meas = None

    # ---- KalmanFilterNew -----------------------------------------------------------------------------
    def _kf_predict(self, Q):  # kalmanfilter.py:340-379
        self.x = np.dot(_F8, self.x)
        self.P = 1.0 * np.dot(np.dot(_F8, self.P), _F8.T) + Q

    def _kf_correct(self, z, R):  # kalmanfilter.py:531-569
        y = z - np.dot(_H8, self.x)
        pht = np.dot(self.P, _H8.T)
        s = np.dot(_H8, pht) + R
        si = np.linalg.inv(s)
        k = np.dot(pht, si)
        self.x = self.x + np.dot(k, y)
        ikh = _I8 - np.dot(k, _H8)
        self.P = np.dot(np.dot(ikh, self.P), ikh.T) + np.dot(np.dot(k, R), k.T)

    def _kf_update(self, z, R):  # kalmanfilter.py:483-569
        self.hist.append(z)
        if z is None:
            if self.observed:            # freeze (deepcopy of the state as of now; last_measurement = history_obs[-2])
                self.last_meas = self.hist[-2]
                self.saved = dict(x=self.x.copy(), P=self.P.copy(), n=len(self.hist), last_meas=self.last_meas.copy())
            self.observed = False
            return
        if not self.observed:
            self._replay()
        self.observed = True
        self._kf_correct(z, np.eye(4) if R is None else R)

    def _replay(self):  # unfreeze, kalmanfilter.py:432-481 (q4)
        if self.saved is None:
            return
        full = self.hist
        s = self.saved
        self.x, self.P, self.last_meas = s["x"], s["P"], s["last_meas"]
        self.saved = None                # the restored dict carries attr_saved = None of the time of the freeze
        self.hist = full[:s["n"] - 1]
        self.observed = True
        seen = [i for i, d in enumerate(full) if d is not None]
        i1, i2 = seen[-2], seen[-1]
        x1, y1, s1, r1 = self.last_meas
        w1, h1 = np.sqrt(s1 * r1), np.sqrt(s1 / r1)
        x2, y2, s2, r2 = full[i2]
        w2, h2 = np.sqrt(s2 * r2), np.sqrt(s2 / r2)
        gap = i2 - i1
        dx, dy, dw, dh = (x2 - x1) / gap, (y2 - y1) / gap, (w2 - w1) / gap, (h2 - h1) / gap
        for i in range(gap):
            xx, yy = x1 + (i + 1) * dx, y1 + (i + 1) * dy
            ww, hh = w1 + (i + 1) * dw, h1 + (i + 1) * dh
            vz = np.array([xx, yy, ww * hh, ww / float(hh)]).reshape((4, 1))
            self.hist.append(vz)
            self._kf_correct(vz, np.eye(4))
            if i != gap - 1:
                self._kf_predict(np.eye(8))

    def _kf_affine(self, m, t):  # kalmanfilter.py:388-405 (new_kf branch)
        big = np.kron(np.eye(4, dtype=float), m)
        self.x = big @ self.x
        self.x[:2] += t
        self.P = big @ self.P @ big.T
        if not self.observed and self.saved is not None:
            s = self.saved
            s["x"] = big @ s["x"]
            s["x"][:2] += t
            s["P"] = big @ s["P"] @ big.T
            s["last_meas"][:2] = m @ s["last_meas"][:2] + t
            s["last_meas"][2:] = m @ s["last_meas"][2:]

    # ---- KalmanBoxTracker ----------------------------------------------------------------------------
    def affine(self, A):  # ocsort.py:252-271 (q2: last_obs may be the same object as obs[age - dt])
        m, t = A[:, :2], A[:, 2].reshape(2, 1)
        if self.last_obs.sum() > 0:
            ps = self.last_obs[:4].reshape(2, 2).T
            ps = m @ ps + t
            self.last_obs[:4] = ps.T.reshape(-1)
        for dt in range(self.delta_t, -1, -1):
            if self.age - dt in self.obs:
                ps = self.obs[self.age - dt][:4].reshape(2, 2).T
                ps = m @ ps + t
                self.obs[self.age - dt][:4] = ps.T.reshape(-1)
        self._kf_affine(m, t)

    def predict(self):  # ocsort.py:273-299
        if self.x[2] + self.x[6] <= 0:
            self.x[6] = 0
        if self.x[3] + self.x[7] <= 0:
            self.x[7] = 0
        if self.frozen_flag:
            self.x[6] = self.x[7] = 0
        self._kf_predict(process_noise(self.x[2, 0], self.x[3, 0]))
        self.age += 1
        if self.tsu > 0:
            self.streak = 0
        self.tsu += 1
        return x_to_box(self.x)

    def update(self, bbox5, cls, det_id=None):  # ocsort.py:203-244
        if bbox5 is not None:
            self.frozen_flag = False
            self.cls = cls
            if self.last_obs.sum() >= 0:
                prev = None
                for dt in range(self.delta_t, 0, -1):
                    if self.age - dt in self.obs:
                        prev = self.obs[self.age - dt]
                        break
                if prev is None:
                    prev = self.last_obs
                self.velocity = _direction(prev, bbox5)
            self.last_obs = bbox5            # the same object goes into obs (q2)
            self.obs[self.age] = bbox5
            self.tsu = 0
            self.hits += 1
            self.streak += 1
            R = measurement_noise(self.x[2, 0], self.x[3, 0])        # from the state before a possible ORU replay (q4)
            self._kf_update(box_to_z(bbox5), R)
        else:
            self._kf_update(None, None)
            self.frozen_flag = True
        if det_id is not None:
            self.det_id = det_id

    def update_emb(self, emb, alpha):  # ocsort.py:246-248 (q6: torch float32 tensors, NumPy norm)
        self.emb = alpha * self.emb + (1 - alpha) * emb
        self.emb /= np.linalg.norm(self.emb)


def _k_prev(obs, age, k):  # ocsort.py:22-30
    if len(obs) == 0:
        return [-1, -1, -1, -1, -1]
    for i in range(k):
        if age - (k - i) in obs:
            return obs[age - (k - i)]
    return obs[max(obs.keys())]


def linear_assignment(cost):  # association.py:202-207 (q1: no `if i >= 0`)
    x, y = lapjv_extended(cost)
    return np.array([[y[i], i] for i in x])


def aw_max_metric(emb_cost, w_emb0, bottom):  # association.py:263-288
    w_emb = np.full_like(emb_cost, w_emb0)
    for i in range(emb_cost.shape[0]):
        inds = np.argsort(-emb_cost[i])
        if len(inds) < 2:
            continue
        if emb_cost[i, inds[0]] == 0:
            rw = 0
        else:
            rw = 1 - max((emb_cost[i, inds[1]] / emb_cost[i, inds[0]]) - bottom, 0) / (1 - bottom)
        w_emb[i] *= rw
    for j in range(emb_cost.shape[1]):
        inds = np.argsort(-emb_cost[:, j])
        if len(inds) < 2:
            continue
        if emb_cost[inds[0], j] == 0:
            cw = 0
        else:
            cw = 1 - max((emb_cost[inds[1], j] / emb_cost[inds[0], j]) - bottom, 0) / (1 - bottom)
        w_emb[:, j] *= cw
    return w_emb * emb_cost


def associate(dets, trks, iou_thr, velocities, prev_obs, vdc_weight, emb_cost, w_emb, aw_off, aw_param):
    """association.py:291-360. ``dets`` [D,6] = x1,y1,x2,y2,score,cls; emb_cost float32 [D,T] (NumPy) or None."""
    if len(trks) == 0:
        return np.empty((0, 2), dtype=int), np.arange(len(dets)), np.empty((0, 5), dtype=int)
    pt = prev_obs[..., np.newaxis]
    cx1, cy1 = (dets[:, 0] + dets[:, 2]) / 2.0, (dets[:, 1] + dets[:, 3]) / 2.0
    cx2, cy2 = (pt[:, 0] + pt[:, 2]) / 2.0, (pt[:, 1] + pt[:, 3]) / 2.0
    dx, dy = cx1 - cx2, cy1 - cy2
    norm = np.sqrt(dx ** 2 + dy ** 2) + 1e-6
    X, Y = dx / norm, dy / norm
    iy = np.repeat(velocities[:, 0][:, np.newaxis], Y.shape[1], axis=1)
    ix = np.repeat(velocities[:, 1][:, np.newaxis], X.shape[1], axis=1)
    ang = np.arccos(np.clip(ix * X + iy * Y, a_min=-1, a_max=1))
    ang = (np.pi / 2.0 - np.abs(ang)) / np.pi
    valid = np.ones(prev_obs.shape[0])
    valid[np.where(prev_obs[:, 4] < 0)] = 0
    iou = iou_xyxy(dets, trks)
    scores = np.repeat(dets[:, -1][:, np.newaxis], trks.shape[0], axis=1)
    valid = np.repeat(valid[:, np.newaxis], X.shape[1], axis=1)
    vdc = ((valid * ang) * vdc_weight).T * scores
    if min(iou.shape) > 0:
        a = (iou > iou_thr).astype(np.int32)
        if a.sum(1).max() == 1 and a.sum(0).max() == 1:
            pairs = np.stack(np.where(a), axis=1)
        else:
            if emb_cost is None:
                emb_cost = 0
            else:
                emb_cost = np.array(emb_cost, copy=True)
                emb_cost[iou <= 0] = 0
                if not aw_off:
                    emb_cost = aw_max_metric(emb_cost, w_emb, aw_param)
                else:
                    emb_cost *= w_emb
            pairs = linear_assignment(-(iou + vdc + emb_cost))
    else:
        pairs = np.empty(shape=(0, 2))
    un_d = [d for d in range(len(dets)) if d not in pairs[:, 0]]
    un_t = [t for t in range(len(trks)) if t not in pairs[:, 1]]
    keep = []
    for m in pairs:
        if iou[m[0], m[1]] < iou_thr:
            un_d.append(m[0])
            un_t.append(m[1])
        else:
            keep.append(m.reshape(1, 2))
    keep = np.empty((0, 2), dtype=int) if len(keep) == 0 else np.concatenate(keep, axis=0)
    return keep, np.array(un_d), np.array(un_t)


class DeepOCSortOracle:
    def __init__(self, det_thresh=0, max_age=30, min_hits=3, iou_threshold=0.3, delta_t=3, asso_func="iou", inertia=0.2,
                 w_association_emb=0.75, alpha_fixed_emb=0.95, aw_param=0.5, embedding_off=False, cmc_off=False, aw_off=False,
                 new_kf_off=False, min_confidence=0.4):
        if new_kf_off:
            raise NotImplementedError("the oracle restates the default 8-d filter (new_kf_off=False)")
        self.det_thresh, self.max_age, self.min_hits, self.iou_threshold = det_thresh, max_age, min_hits, iou_threshold
        self.delta_t, self.asso, self.inertia = delta_t, ASSO[asso_func], inertia
        self.w_emb, self.alpha_fixed, self.aw_param = w_association_emb, alpha_fixed_emb, aw_param
        self.embedding_off, self.cmc_off, self.aw_off = embedding_off, cmc_off, aw_off
        self.min_confidence = min_confidence
        self.trackers = []
        self.frame_count = 0                  # never incremented (q3)
        self._uid = 0

    def update(self, dets7, embs, affine):
        """dets7 float64 [D,7] after the wrapper filter, embs float32 [D,E] (torch CPU tensor or array), affine float64 [2,3] or
        None. Returns float64 [M,8] = x1,y1,x2,y2,id+1,cls,conf,det_id."""
        dets = np.asarray(dets7, dtype=np.float64).reshape(-1, 7)
        dets = dets[dets[:, 4] > self.det_thresh]                 # boolean mask: a fresh array per frame (q2)
        if self.embedding_off or dets.shape[0] == 0:
            dets_embs = np.ones((dets.shape[0], 1))
        else:
            dets_embs = torch.as_tensor(np.asarray(embs, dtype=np.float32))[torch.as_tensor(np.asarray(dets7)[:, 4] > self.det_thresh)]
        if not self.cmc_off and affine is not None:
            for t in self.trackers:
                t.affine(np.asarray(affine, dtype=np.float64))
        trust = (dets[:, 4] - self.det_thresh) / (1 - self.det_thresh)
        af = self.alpha_fixed
        dets_alpha = af + (1 - af) * (1 - trust)

        trks = np.zeros((len(self.trackers), 5))
        trk_embs, dead = [], []
        for t, row in enumerate(trks):
            pos = self.trackers[t].predict()[0]
            row[:] = [pos[0], pos[1], pos[2], pos[3], 0]
            if np.any(np.isnan(pos)):
                dead.append(t)
            else:
                trk_embs.append(self.trackers[t].emb)
        trks = np.ma.compress_rows(np.ma.masked_invalid(trks))
        trk_embs = np.vstack(trk_embs) if len(trk_embs) > 0 else np.array(trk_embs)
        for t in reversed(dead):
            self.trackers.pop(t)
        vel = np.array([t.velocity if t.velocity is not None else np.array((0, 0)) for t in self.trackers])
        last = np.array([t.last_obs for t in self.trackers])
        kobs = np.array([_k_prev(t.obs, t.age, self.delta_t) for t in self.trackers])

        if self.embedding_off or dets.shape[0] == 0 or trk_embs.shape[0] == 0:
            emb1 = None
        else:
            emb1 = np.asarray(dets_embs) @ trk_embs.T              # float32 NumPy matmul (q6)
        matched, un_d, un_t = associate(dets[:, :-1], trks, self.iou_threshold, vel, kobs, self.inertia, emb1, self.w_emb,
                                        self.aw_off, self.aw_param)
        for m in matched:
            self.trackers[m[1]].update(dets[m[0], :5], dets[m[0], 5], dets[m[0], 6])
            self.trackers[m[1]].update_emb(dets_embs[m[0]], alpha=dets_alpha[m[0]])

        if un_d.shape[0] > 0 and un_t.shape[0] > 0:              # OCR, ocsort.py:476-508
            left = np.array(self.asso(dets[un_d, :-1], last[un_t]))
            if left.max() > self.iou_threshold:
                gd, gt = [], []
                for m in linear_assignment(-left):
                    di, ti = un_d[m[0]], un_t[m[1]]
                    if left[m[0], m[1]] < self.iou_threshold:
                        continue
                    self.trackers[ti].update(dets[di, :5], dets[di, 5], dets[di, 6])
                    self.trackers[ti].update_emb(dets_embs[di], alpha=dets_alpha[di])
                    gd.append(di)
                    gt.append(ti)
                un_d = np.setdiff1d(un_d, np.array(gd))
                un_t = np.setdiff1d(un_t, np.array(gt))

        for ti in un_t:
            self.trackers[ti].update(None, None)
        for di in un_d:
            self.trackers.append(_Trk(dets[di, :5], dets[di, 5], self.delta_t, dets_embs[di], dets[di, 6], self._uid))
            self._uid += 1

        rows = []
        i = len(self.trackers)
        for t in reversed(self.trackers):                         # ocsort.py:519-536
            d = x_to_box(t.x)[0] if t.last_obs.sum() < 0 else t.last_obs[:4]
            if t.tsu < 1 and (t.streak >= self.min_hits or self.frame_count <= self.min_hits):
                rows.append(np.concatenate((d, [t.id + 1], [t.cls], [t.conf], [t.det_id])).reshape(1, -1))
            i -= 1
            if t.tsu > self.max_age:
                self.trackers.pop(i)
        return np.concatenate(rows) if rows else np.empty((0, 8))

    def run_video(self, dets, offsets, embeddings, affines=None):
        """Wrapper semantics of deep_oc_sort_api.py:60-67 (frames without detections are skipped: no update call, no CMC call)."""
        out, fr = [], []
        for f in range(len(offsets) - 1):
            sl = slice(offsets[f], offsets[f + 1])
            d = dets[sl]
            if len(d) == 0:
                continue
            keep = d[:, 4] > self.min_confidence
            r = self.update(d[keep], embeddings[sl][keep], None if affines is None else affines[f])
            out.append(r)
            fr.append(np.full(len(r), f, dtype=np.int32))
        if not out:
            return np.zeros((0, 8)), np.zeros((0,), dtype=np.int32)
        return np.concatenate(out), np.concatenate(fr)
