"""OC-SORT per-frame association oracle (test infrastructure; never imported by tracklab_b200).

Restates
  /root/reference/plugins/track/oc_sort/ocsort.py:10-54,57-169,183-334   (helpers, KalmanBoxTracker, OCSort.update)
  /root/reference/plugins/track/oc_sort/association.py:175-195,242-298    (VDC cost, associate)
  /root/reference/plugins/track/oc_sort/kalmanfilter.py:339-379,383-434,437-526 (predict, ORU freeze/unfreeze, update)
and the wrapper filter /root/reference/tracklab/wrappers/track/oc_sort_api.py:50-56.

The filter is written as functions over a small state record; the observation-centric re-update
(ORU) is kept as "history list + frozen (x, P, history length)" which is what the reference's
``deepcopy(self.__dict__)`` round trip amounts to (freeze happens only on an observed->unobserved
transition, so the saved dict never nests). NumPy calls are the same as the reference's so the
float64 results are bit-identical.

Quirks kept: the first-round VDC term is multiplied by the CLASS column
(association.py:262 receives ``dets[:, :-1]``), the first round always uses plain IoU, and a
thresholded IoU matrix that is already a partial permutation skips the solver.
"""
import numpy as np

from .assign_np import lapjv_extended
from .boxes_np import giou_xyxy, iou_xyxy

_F = np.array([[1, 0, 0, 0, 1, 0, 0], [0, 1, 0, 0, 0, 1, 0], [0, 0, 1, 0, 0, 0, 1], [0, 0, 0, 1, 0, 0, 0],
               [0, 0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 0, 1]])
_H = np.array([[1, 0, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0, 0]])
_I7 = np.eye(7)


def _make_rqp():  # ocsort.py:80-84
    R, P, Q = np.eye(4), np.eye(7), np.eye(7)
    R[2:, 2:] *= 10.0
    P[4:, 4:] *= 1000.0
    P *= 10.0
    Q[-1, -1] *= 0.01
    Q[4:, 4:] *= 0.01
    return R, P, Q


_R, _P0, _Q = _make_rqp()


def diou_xyxy(a, b):
    """association.py:58-95."""
    a = np.asarray(a)[:, None, :]
    b = np.asarray(b)[None, :, :]
    iw = np.maximum(0.0, np.minimum(a[..., 2], b[..., 2]) - np.maximum(a[..., 0], b[..., 0]))
    ih = np.maximum(0.0, np.minimum(a[..., 3], b[..., 3]) - np.maximum(a[..., 1], b[..., 1]))
    inter = iw * ih
    iou = inter / ((a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
                   + (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1]) - inter)
    inner = ((a[..., 0] + a[..., 2]) / 2.0 - (b[..., 0] + b[..., 2]) / 2.0) ** 2 \
        + ((a[..., 1] + a[..., 3]) / 2.0 - (b[..., 1] + b[..., 3]) / 2.0) ** 2
    outer = (np.maximum(a[..., 2], b[..., 2]) - np.minimum(a[..., 0], b[..., 0])) ** 2 \
        + (np.maximum(a[..., 3], b[..., 3]) - np.minimum(a[..., 1], b[..., 1])) ** 2
    return (iou - inner / outer + 1) / 2.0


def ciou_xyxy(a, b):
    """association.py:97-147."""
    a = np.asarray(a)[:, None, :]
    b = np.asarray(b)[None, :, :]
    iw = np.maximum(0.0, np.minimum(a[..., 2], b[..., 2]) - np.maximum(a[..., 0], b[..., 0]))
    ih = np.maximum(0.0, np.minimum(a[..., 3], b[..., 3]) - np.maximum(a[..., 1], b[..., 1]))
    inter = iw * ih
    iou = inter / ((a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
                   + (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1]) - inter)
    inner = ((a[..., 0] + a[..., 2]) / 2.0 - (b[..., 0] + b[..., 2]) / 2.0) ** 2 \
        + ((a[..., 1] + a[..., 3]) / 2.0 - (b[..., 1] + b[..., 3]) / 2.0) ** 2
    outer = (np.maximum(a[..., 2], b[..., 2]) - np.minimum(a[..., 0], b[..., 0])) ** 2 \
        + (np.maximum(a[..., 3], b[..., 3]) - np.minimum(a[..., 1], b[..., 1])) ** 2
    w1, h1 = a[..., 2] - a[..., 0], a[..., 3] - a[..., 1] + 1.0
    w2, h2 = b[..., 2] - b[..., 0], b[..., 3] - b[..., 1] + 1.0
    at = np.arctan(w2 / h2) - np.arctan(w1 / h1)
    v = (4 / (np.pi ** 2)) * (at ** 2)
    alpha = v / ((1 - iou) + v)
    return (iou - inner / outer - alpha * v + 1) / 2.0


def ct_dist_xyxy(a, b):
    """association.py:150-171."""
    a = np.asarray(a)[:, None, :]
    b = np.asarray(b)[None, :, :]
    d = np.sqrt(((a[..., 0] + a[..., 2]) / 2.0 - (b[..., 0] + b[..., 2]) / 2.0) ** 2
                + ((a[..., 1] + a[..., 3]) / 2.0 - (b[..., 1] + b[..., 3]) / 2.0) ** 2)
    d = d / d.max()
    return d.max() - d


ASSO = {"iou": iou_xyxy, "giou": giou_xyxy, "diou": diou_xyxy, "ciou": ciou_xyxy, "ct_dist": ct_dist_xyxy}


def box_to_z(b):  # ocsort.py:21-33
    w = b[2] - b[0]
    h = b[3] - b[1]
    return np.array([b[0] + w / 2.0, b[1] + h / 2.0, w * h, w / float(h + 1e-6)]).reshape((4, 1))


def x_to_box(x):  # ocsort.py:36-46 (score=None branch)
    w = np.sqrt(x[2] * x[3])
    h = x[2] / w
    return np.array([x[0] - w / 2.0, x[1] - h / 2.0, x[0] + w / 2.0, x[1] + h / 2.0]).reshape((1, 4))


def _direction(b1, b2):  # ocsort.py:49-54
    cx1, cy1 = (b1[0] + b1[2]) / 2.0, (b1[1] + b1[3]) / 2.0
    cx2, cy2 = (b2[0] + b2[2]) / 2.0, (b2[1] + b2[3]) / 2.0
    speed = np.array([cy2 - cy1, cx2 - cx1])
    return speed / (np.sqrt((cy2 - cy1) ** 2 + (cx2 - cx1) ** 2) + 1e-6)


class _Trk:
    """KalmanBoxTracker + its KalmanFilterNew (ocsort.py:57-169, kalmanfilter.py:109-527)."""

    def __init__(self, bbox5, cls, delta_t, det_id, uid):
        self.x = np.zeros((7, 1))
        self.P = _P0.copy()
        self.x[:4] = box_to_z(bbox5)
        self.tsu = 0
        self.id = uid
        self.hits = 0
        self.streak = 0
        self.age = 0
        self.conf = bbox5[-1]
        self.cls = cls
        self.last_obs = np.array([-1, -1, -1, -1, -1])
        self.obs = {}
        self.velocity = None
        self.delta_t = delta_t
        self.det_id = det_id
        # ORU bookkeeping
        self.hist = []
        self.observed = False
        self.frozen = None

    # -- KalmanFilterNew ---------------------------------------------------------------
    def _kf_predict(self):  # kalmanfilter.py:339-379
        self.x = np.dot(_F, self.x)
        self.P = 1.0 * np.dot(np.dot(_F, self.P), _F.T) + _Q

    def _kf_correct(self, z):  # kalmanfilter.py:488-526 (z already (4,1))
        y = z - np.dot(_H, self.x)
        pht = np.dot(self.P, _H.T)
        s = np.dot(_H, pht) + _R
        si = np.linalg.inv(s)
        k = np.dot(pht, si)
        self.x = self.x + np.dot(k, y)
        ikh = _I7 - np.dot(k, _H)
        self.P = np.dot(np.dot(ikh, self.P), ikh.T) + np.dot(np.dot(k, _R), k.T)

    def _kf_update(self, z):  # kalmanfilter.py:437-526
        self.hist.append(z)
        if z is None:
            if self.observed:  # freeze(): state as of now, history includes the None just appended
                self.frozen = (self.x.copy(), self.P.copy(), len(self.hist))
            self.observed = False
            return
        if not self.observed and self.frozen is not None:
            self._replay()
        self.observed = True
        self._kf_correct(z)

    def _replay(self):  # unfreeze(), kalmanfilter.py:390-434
        full = self.hist
        x, P, n = self.frozen
        self.x, self.P, self.frozen = x, P, None
        self.hist = full[:n - 1]
        self.observed = True
        seen = [i for i, d in enumerate(full) if d is not None]
        i1, i2 = seen[-2], seen[-1]
        x1, y1, s1, r1 = (float(v) for v in full[i1].reshape(-1))
        x2, y2, s2, r2 = (float(v) for v in full[i2].reshape(-1))
        w1, h1 = np.sqrt(s1 * r1), np.sqrt(s1 / r1)
        w2, h2 = np.sqrt(s2 * r2), np.sqrt(s2 / r2)
        gap = i2 - i1
        dx, dy, dw, dh = (x2 - x1) / gap, (y2 - y1) / gap, (w2 - w1) / gap, (h2 - h1) / gap
        for i in range(gap):
            xx, yy = x1 + (i + 1) * dx, y1 + (i + 1) * dy
            ww, hh = w1 + (i + 1) * dw, h1 + (i + 1) * dh
            vz = np.array([xx, yy, ww * hh, ww / float(hh)]).reshape((4, 1))
            self.hist.append(vz)
            self._kf_correct(vz)
            if i != gap - 1:
                self._kf_predict()

    # -- KalmanBoxTracker --------------------------------------------------------------
    def predict(self):  # ocsort.py:150-163
        if (self.x[6] + self.x[2]) <= 0:
            self.x[6] *= 0.0
        self._kf_predict()
        self.age += 1
        if self.tsu > 0:
            self.streak = 0
        self.tsu += 1
        return x_to_box(self.x)

    def update(self, bbox5, cls, det_id=None):  # ocsort.py:103-148
        if bbox5 is not None:
            self.conf = bbox5[-1]
            self.cls = cls
            if self.last_obs.sum() >= 0:
                prev = None
                for i in range(self.delta_t):
                    dt = self.delta_t - i
                    if self.age - dt in self.obs:
                        prev = self.obs[self.age - dt]
                        break
                if prev is None:
                    prev = self.last_obs
                self.velocity = _direction(prev, bbox5)
            self.last_obs = bbox5
            self.obs[self.age] = bbox5
            self.tsu = 0
            self.hits += 1
            self.streak += 1
            self._kf_update(box_to_z(bbox5))
        else:
            self._kf_update(None)
        if det_id is not None:
            self.det_id = det_id


def _k_prev(obs, age, k):  # ocsort.py:10-18
    if len(obs) == 0:
        return [-1, -1, -1, -1, -1]
    for i in range(k):
        if age - (k - i) in obs:
            return obs[age - (k - i)]
    return obs[max(obs.keys())]


def _solve(cost):  # association.py:187-191: rows of [det, trk]
    x, y = lapjv_extended(cost)
    return np.array([[y[i], i] for i in x if i >= 0])


def associate(dets, trks, iou_thr, velocities, prev_obs, vdc_weight):
    """association.py:242-298. ``dets`` is [D,6] = x1,y1,x2,y2,score,cls (so ``[:, -1]`` is the class)."""
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
            pairs = _solve(-(iou + vdc))
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


class OCSortOracle:
    def __init__(self, det_thresh=0, max_age=50, min_hits=1, iou_threshold=0.22136877277096445, delta_t=1,
                 asso_func="giou", inertia=0.3941737016672115, use_byte=False, min_confidence=0.4):
        self.det_thresh, self.max_age, self.min_hits = det_thresh, max_age, min_hits
        self.iou_threshold, self.delta_t, self.inertia, self.use_byte = iou_threshold, delta_t, inertia, use_byte
        self.asso = ASSO[asso_func]
        self.min_confidence = min_confidence
        self.trackers = []
        self.frame_count = 0
        self._uid = 0  # KalmanBoxTracker.count is reset by OCSort.__init__ (ocsort.py:201)

    def _spawn(self, bbox5, cls, det_id):
        t = _Trk(bbox5, cls, self.delta_t, det_id, self._uid)
        self._uid += 1
        return t

    def update(self, dets7):
        """dets7 float64[D,7] after the wrapper filter; returns float64[M,8] = x1,y1,x2,y2,id+1,cls,conf,det_id."""
        self.frame_count += 1
        out = np.asarray(dets7, dtype=np.float64).reshape(-1, 7)
        conf = out[:, 4]
        second = out[np.logical_and(conf > 0.1, conf < self.det_thresh)]
        dets = out[conf > self.det_thresh]

        trks = np.zeros((len(self.trackers), 5))
        dead = []
        for t, row in enumerate(trks):
            pos = self.trackers[t].predict()[0]
            row[:] = [pos[0], pos[1], pos[2], pos[3], 0]
            if np.any(np.isnan(pos)):
                dead.append(t)
        trks = np.ma.compress_rows(np.ma.masked_invalid(trks))
        for t in reversed(dead):
            self.trackers.pop(t)
        vel = np.array([t.velocity if t.velocity is not None else np.array((0, 0)) for t in self.trackers])
        last = np.array([t.last_obs for t in self.trackers])
        kobs = np.array([_k_prev(t.obs, t.age, self.delta_t) for t in self.trackers])

        matched, un_d, un_t = associate(dets[:, :-1], trks, self.iou_threshold, vel, kobs, self.inertia)
        for m in matched:
            self.trackers[m[1]].update(dets[m[0], :5], dets[m[0], 5], dets[m[0], 6])

        if self.use_byte and len(second) > 0 and un_t.shape[0] > 0:  # ocsort.py:264-282
            left = np.array(self.asso(second[:, :-1], trks[un_t]))
            if left.max() > self.iou_threshold:
                gone = []
                for m in _solve(-left):
                    di, ti = m[0], un_t[m[1]]
                    if left[m[0], m[1]] < self.iou_threshold:
                        continue
                    self.trackers[ti].update(second[di, :5], second[di, 5], second[di, 6])
                    gone.append(ti)
                un_t = np.setdiff1d(un_t, np.array(gone))

        if un_d.shape[0] > 0 and un_t.shape[0] > 0:  # OCR, ocsort.py:284-306
            left = np.array(self.asso(dets[un_d, :-1], last[un_t]))
            if left.max() > self.iou_threshold:
                gd, gt = [], []
                for m in _solve(-left):
                    di, ti = un_d[m[0]], un_t[m[1]]
                    if left[m[0], m[1]] < self.iou_threshold:
                        continue
                    self.trackers[ti].update(dets[di, :5], dets[di, 5], dets[di, 6])
                    gd.append(di)
                    gt.append(ti)
                un_d = np.setdiff1d(un_d, np.array(gd))
                un_t = np.setdiff1d(un_t, np.array(gt))

        for ti in un_t:
            self.trackers[ti].update(None, None)
        for di in un_d:
            self.trackers.append(self._spawn(dets[di, :5], dets[di, 5], dets[di, 6]))

        rows = []
        i = len(self.trackers)
        for t in reversed(self.trackers):  # ocsort.py:315-331
            d = x_to_box(t.x)[0] if t.last_obs.sum() < 0 else t.last_obs[:4]
            if t.tsu < 1 and (t.streak >= self.min_hits or self.frame_count <= self.min_hits):
                rows.append(np.concatenate((d, [t.id + 1], [t.cls], [t.conf], [t.det_id])).reshape(1, -1))
            i -= 1
            if t.tsu > self.max_age:
                self.trackers.pop(i)
        return np.concatenate(rows) if rows else np.empty((0, 8))

    def run_video(self, dets, offsets):
        """Wrapper semantics of oc_sort_api.py:50-76 (see ByteTrackOracle.run_video)."""
        out, fr = [], []
        for f in range(len(offsets) - 1):
            d = dets[offsets[f]:offsets[f + 1]]
            if len(d) == 0:
                continue
            r = self.update(d[d[:, 4] > self.min_confidence])
            out.append(r)
            fr.append(np.full(len(r), f, dtype=np.int32))
        if not out:
            return np.zeros((0, 8)), np.zeros((0,), dtype=np.int32)
        return np.concatenate(out), np.concatenate(fr)
