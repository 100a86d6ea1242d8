import ast
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return dict(rows=g["rows"], frames=g["frames"], gen=ast.literal_eval(str(g["gen"])),
                hyper=ast.literal_eval(str(g["hyper"])), min_conf=float(g["min_conf"]),
                tracker=str(g["tracker"]), dets_sha=bytes(g["dets_sha"].tobytes()) if "dets_sha" in g.files else None,
                **{k: g[k] for k in ("affines", "raw_emb") if k in g.files})


def load_bpbreid_golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return dict(rows=g["rows"], frames=g["frames"], gen=ast.literal_eval(str(g["gen"])), hyper=ast.literal_eval(str(g["hyper"])))


BPB_KEYS = ("ema_alpha", "mc_lambda", "max_dist", "max_iou_distance", "max_age", "n_init", "min_bbox_confidence",
            "max_kalman_prediction_without_update")
BPB_ORACLE_KEYS = BPB_KEYS + ("matching_strategy", "gating_thres_factor", "w_kfgd", "w_reid", "w_st")   # the oracle also restates bot_sort_matching


def assert_bpbreid_rows_match(rows, frames, ref_rows, ref_frames, box_tol=1e-6, dist_tol=1e-5, allow_relabel=False):
    """BPBReID rows [track_id, kf_ltwh(4), pred_kf_ltwh(4), matched code, matched dist, hits, age, det_id]: integer columns
    bit-exact (track ids up to a consistent relabelling when allow_relabel, see assert_rows_match), boxes / distances within
    tolerance, NaN pattern (births have no predicted box and no match) identical."""
    assert rows.shape == ref_rows.shape, (rows.shape, ref_rows.shape)
    ka, kb = np.lexsort((rows[:, 13], frames)), np.lexsort((ref_rows[:, 13], ref_frames))
    a, b = rows[ka], ref_rows[kb]
    assert np.array_equal(frames[ka], ref_frames[kb])
    for c, what in ((13, "det ids"), (9, "matched_with stage"), (11, "hits"), (12, "age")):
        assert np.array_equal(a[:, c], b[:, c]), what + " differ"
    if allow_relabel and not np.array_equal(a[:, 0], b[:, 0]):
        fwd, bwd = {}, {}
        for x, y in zip(a[:, 0], b[:, 0]):
            assert fwd.setdefault(x, y) == y and bwd.setdefault(y, x) == x, "tracks differ beyond a relabelling"
    else:
        assert np.array_equal(a[:, 0], b[:, 0]), "track ids differ"
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    err = np.nanmax(np.abs(a[:, 1:9] - b[:, 1:9])) if len(a) else 0.0
    derr = np.nanmax(np.abs(a[:, 10] - b[:, 10])) if np.isfinite(a[:, 10]).any() else 0.0
    assert err <= box_tol, f"box error {err}"
    assert derr <= dist_tol, f"matched distance error {derr}"
    return err, derr


def assert_rows_match(rows, frames, ref_rows, ref_frames, box_tol=1e-6, allow_relabel=False):
    """Integer outputs (frame, track id, det id, class) bit-exact; boxes/scores within tolerance.
    Rows are compared as sets per frame keyed by det id (the wrappers index results by det id).

    allow_relabel: OC-SORT only. When its solver is forced to pair leftover tracks with zero-overlap detections
    (association.py:272 on a rectangular matrix) WHICH zero-cost detections get those dummy pairs is a tie broken by
    the third-party solver's internals (lap 0.5.12 in the reference, scipy in the oracle's stand-in — unpinned, see
    DESIGN.md §6); the pairs are discarded (:286-292) but they reorder ``unmatched_detections`` and therefore the
    NUMBERING of the tracks born in that frame. The partition of detections into tracks is unaffected, so in that
    mode ids must agree up to one consistent bijection over the whole video."""
    assert rows.shape == ref_rows.shape, (rows.shape, ref_rows.shape)
    assert np.array_equal(np.bincount(frames, minlength=ref_frames.max() + 1 if len(ref_frames) else 0),
                          np.bincount(ref_frames, minlength=ref_frames.max() + 1 if len(ref_frames) else 0))
    ka = np.lexsort((rows[:, 7], frames))
    kb = np.lexsort((ref_rows[:, 7], ref_frames))
    a, b = rows[ka], ref_rows[kb]
    assert np.array_equal(frames[ka], ref_frames[kb])
    assert np.array_equal(a[:, 7], b[:, 7]), "det ids differ"
    if allow_relabel and not np.array_equal(a[:, 4], b[:, 4]):
        fwd, bwd = {}, {}
        for x, y in zip(a[:, 4], b[:, 4]):
            assert fwd.setdefault(x, y) == y and bwd.setdefault(y, x) == x, "tracks differ beyond a relabelling"
    else:
        assert np.array_equal(a[:, 4], b[:, 4]), "track ids differ"
    assert np.array_equal(a[:, 5], b[:, 5]), "classes differ"
    assert np.array_equal(a[:, 6], b[:, 6]), "scores differ"
    err = np.abs(a[:, :4] - b[:, :4]).max() if len(a) else 0.0
    assert err <= box_tol, f"box error {err}"
    return err


def blackout_video(video, period=12, span=3, keep_frac=0.15, seed=0):
    """Copy of ``video`` in which, for ``span`` frames of every ``period``, only ``keep_frac`` of the detections survive (a detector
    black-out): tracks die during the gap and the full set of detections returns at once, so detections outnumber the live
    trackers in the frames after it (exercises the unassigned-row handling of the association rounds)."""
    import dataclasses
    rng = np.random.default_rng(seed)
    keep = np.ones(len(video.dets), dtype=bool)
    for f in range(video.n_frames):
        if f % period >= period - span:
            sl = slice(video.offsets[f], video.offsets[f + 1])
            keep[sl] = rng.random(sl.stop - sl.start) < keep_frac
    off = np.concatenate([[0], np.cumsum([keep[video.offsets[f]:video.offsets[f + 1]].sum() for f in range(video.n_frames)])]).astype(video.offsets.dtype)
    rep = dict(dets=video.dets[keep].copy(), offsets=off, gt_identity=video.gt_identity[keep].copy())
    if video.embeddings is not None:
        rep["embeddings"] = video.embeddings[keep].copy()
    if video.visibility is not None:
        rep["visibility"] = video.visibility[keep].copy()
    return dataclasses.replace(video, **rep)
