"""ByteTrack per-frame association oracle (test infrastructure; never imported by tracklab_b200).

Restates /root/reference/plugins/track/byte_track/byte_tracker.py:167-320 (BYTETracker.update),
the STrack life cycle (:10-148), the list helpers (:323-361) and the wrapper filter
/root/reference/tracklab/wrappers/track/byte_track_api.py:50-56.

Tracks are plain records held in three Python lists exactly like the reference (order matters for
indices, ids and the duplicate-removal pass). Quirks kept on purpose:
  * the (cx, cy, w, h) box is stored in the field the filter treats as tlwh (byte_tracker.py:175,191-192),
    and converted back with xywh2xyxy on output (:311);
  * a freshly initiated mean is float32 (``_tlwh`` is float32, :15) until the first predict/update;
  * tracks removed for age stay in ``lost`` for one more frame (the ``removed`` list is extended
    after the subtraction, :296-298) and can be re-activated during that frame;
  * the id counter is global to the process in the reference (basetrack.py:13,35-37); here it is a
    constructor argument so tests can reproduce either behaviour.
"""
import numpy as np

from . import kalman_xyah_np as kf
from .assign_np import lapjv_extended
from .boxes_np import iou_plus1_f32

NEW, TRACKED, LOST, REMOVED = 0, 1, 2, 3


class _Rec:
    """One STrack (byte_tracker.py:10-148)."""

    __slots__ = ("box32", "mean", "cov", "activated", "state", "score", "cls", "det_id",
                 "track_id", "frame_id", "start_frame", "tracklet_len")

    def __init__(self, box, score, cls, det_id):
        self.box32 = np.asarray(box, dtype=np.float32)  # byte_tracker.py:15
        self.mean = None
        self.cov = None
        self.activated = False
        self.state = NEW
        self.score = score
        self.cls = cls
        self.det_id = det_id
        self.track_id = 0  # BaseTrack.track_id class default (basetrack.py:15)
        self.frame_id = 0
        self.start_frame = 0
        self.tracklet_len = 0

    def tlwh(self):  # byte_tracker.py:100-110
        if self.mean is None:
            return self.box32.copy()
        out = self.mean[:4].copy()
        out[2] *= out[3]
        out[:2] -= out[2:] / 2
        return out

    def tlbr(self):  # byte_tracker.py:114-120
        out = self.tlwh().copy()
        out[2:] += out[:2]
        return out


def _to_xyah(tlwh):  # byte_tracker.py:124-131
    out = np.asarray(tlwh).copy()
    out[:2] += out[2:] / 2
    out[2] /= out[3]
    return out


def _iou_dist(a, b):  # matching.py:72-89
    return 1 - iou_plus1_f32([t.tlbr() for t in a], [t.tlbr() for t in b])


def _fuse_score(cost, dets):  # matching.py:171-179
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


def _union(a, b):  # joint_stracks, byte_tracker.py:323-334
    seen, out = set(), []
    for t in a:
        seen.add(t.track_id)
        out.append(t)
    for t in b:
        if t.track_id not in seen:
            seen.add(t.track_id)
            out.append(t)
    return out


def _minus(a, b):  # sub_stracks, byte_tracker.py:337-345
    keep = {}
    for t in a:
        keep[t.track_id] = t
    for t in b:
        if keep.get(t.track_id, 0):
            del keep[t.track_id]
    return list(keep.values())


class ByteTrackOracle:
    def __init__(self, track_thresh=0.6, match_thresh=0.8, track_buffer=30, frame_rate=30,
                 min_confidence=0.4, first_id=1):
        self.track_thresh = track_thresh
        self.match_thresh = match_thresh
        self.det_thresh = track_thresh + 0.1
        self.max_time_lost = int(frame_rate / 30.0 * track_buffer)
        self.min_confidence = min_confidence
        self.tracked, self.lost, self.removed = [], [], []
        self.frame_id = 0
        self._next = first_id - 1

    def _new_id(self):
        self._next += 1
        return self._next

    # -- STrack transitions -------------------------------------------------------------
    def _activate(self, t):  # byte_tracker.py:43-56
        t.track_id = self._new_id()
        t.mean, t.cov = kf.bt_initiate(_to_xyah(t.box32))
        t.tracklet_len = 0
        t.state = TRACKED
        if self.frame_id == 1:
            t.activated = True
        t.frame_id = self.frame_id
        t.start_frame = self.frame_id

    def _correct(self, t, det, reactivate):  # byte_tracker.py:58-97
        t.mean, t.cov = kf.bt_update(t.mean, t.cov, _to_xyah(det.tlwh()))
        t.tracklet_len = 0 if reactivate else t.tracklet_len + 1
        t.state = TRACKED
        t.activated = True
        t.frame_id = self.frame_id
        t.score = det.score
        if reactivate:
            t.cls = det.cls
        t.det_id = det.det_id

    # -- one frame ----------------------------------------------------------------------
    def update(self, dets):
        """``dets`` float64[D,7] = [l,t,r,b,conf,cls,det_id] AFTER the wrapper's ``conf > min_confidence``
        filter; returns float64[M,8] = [x1,y1,x2,y2,track_id,cls,score,det_id]."""
        dets = np.asarray(dets, dtype=np.float64).reshape(-1, 7)
        self.frame_id += 1
        activated, refind, lost_now, removed_now = [], [], [], []
        xyxy = dets[:, :4]
        xywh = np.empty_like(xyxy)  # ultralytics xyxy2xywh
        xywh[..., 0] = (xyxy[..., 0] + xyxy[..., 2]) / 2
        xywh[..., 1] = (xyxy[..., 1] + xyxy[..., 3]) / 2
        xywh[..., 2] = xyxy[..., 2] - xyxy[..., 0]
        xywh[..., 3] = xyxy[..., 3] - xyxy[..., 1]
        conf, cls, ids = dets[:, 4], dets[:, 5], dets[:, 6]
        high = conf > self.track_thresh
        second = np.logical_and(conf > 0.1, conf < self.track_thresh)
        d_high = [_Rec(b, s, c, i) for b, s, c, i in zip(xywh[high], conf[high], cls[high], ids[high])]
        d_low = [_Rec(b, s, c, i) for b, s, c, i in zip(xywh[second], conf[second], cls[second], ids[second])]

        unconfirmed = [t for t in self.tracked if not t.activated]
        confirmed = [t for t in self.tracked if t.activated]

        # first association (byte_tracker.py:217-237)
        pool = _union(confirmed, self.lost)
        if pool:
            mm = np.asarray([t.mean.copy() for t in pool])
            cc = np.asarray([t.cov for t in pool])
            for i, t in enumerate(pool):
                if t.state != TRACKED:
                    mm[i][7] = 0
            mm, cc = kf.bt_multi_predict(mm, cc)
            for t, m_, c_ in zip(pool, mm, cc):
                t.mean, t.cov = m_, c_
        cost = _fuse_score(_iou_dist(pool, d_high), d_high)
        pairs, u_trk, u_det = _assign(cost, self.match_thresh)
        for it, idet in pairs:
            t = pool[it]
            if t.state == TRACKED:
                self._correct(t, d_high[idet], False)
                activated.append(t)
            else:
                self._correct(t, d_high[idet], True)
                refind.append(t)

        # second association with low-score boxes (byte_tracker.py:239-264)
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

        # unconfirmed tracks vs remaining high boxes (byte_tracker.py:266-278)
        left = [d_high[i] for i in u_det]
        cost = _fuse_score(_iou_dist(unconfirmed, left), left)
        pairs, u_unc, u_det = _assign(cost, 0.7)
        for it, idet in pairs:
            self._correct(unconfirmed[it], left[idet], False)
            activated.append(unconfirmed[it])
        for it in u_unc:
            unconfirmed[it].state = REMOVED
            removed_now.append(unconfirmed[it])

        # births (byte_tracker.py:280-286)
        for i in u_det:
            if left[i].score < self.det_thresh:
                continue
            self._activate(left[i])
            activated.append(left[i])

        # ageing and list maintenance (byte_tracker.py:288-299)
        for t in self.lost:
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
        # duplicate removal (byte_tracker.py:348-361)
        pd = _iou_dist(self.tracked, self.lost)
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
            x1, y1 = b[0] - b[2] / 2, b[1] - b[3] / 2  # xywh2xyxy (byte_tracker.py:311)
            x2, y2 = b[0] + b[2] / 2, b[1] + b[3] / 2
            rows.append([x1, y1, x2, y2, t.track_id, t.cls, t.score, t.det_id])
        return np.asarray(rows, dtype=np.float64).reshape(-1, 8)

    def run_video(self, dets, offsets):
        """Whole video through the wrapper semantics of byte_track_api.py:50-76: frames without any
        detection row skip ``update`` entirely; rows are filtered with ``conf > min_confidence``.
        Returns (rows float64[R,8], frame_of_row int32[R])."""
        out, fr = [], []
        for f in range(len(offsets) - 1):
            d = dets[offsets[f]:offsets[f + 1]]
            if len(d) == 0:
                continue
            d = d[d[:, 4] > self.min_confidence]
            r = self.update(d)
            out.append(r)
            fr.append(np.full(len(r), f, dtype=np.int32))
        if not out:
            return np.zeros((0, 8)), np.zeros((0,), dtype=np.int32)
        return np.concatenate(out), np.concatenate(fr)
