import ast
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return dict(rows=g["rows"], frames=g["frames"], gen=ast.literal_eval(str(g["gen"])),
                hyper=ast.literal_eval(str(g["hyper"])), min_conf=float(g["min_conf"]),
                tracker=str(g["tracker"]), dets_sha=bytes(g["dets_sha"].tobytes()))


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
