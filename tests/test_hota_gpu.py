"""tk_hota_sequence vs the TrackEval fork vendored in the reference (goldens) and vs the oracle on fresh cases."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS = ("HOTA", "DetA", "AssA", "DetRe", "DetPr", "AssRe", "AssPr", "LocA", "HOTA_TP", "HOTA_FN", "HOTA_FP")


def _device_hota(gb, gi, go, tb, ti, to):
    from tracklab_b200.hota import HotaDevice
    ng, nt = np.diff(go), np.diff(to)
    n_g, n_t = (int(gi.max()) + 1 if len(gi) else 0), (int(ti.max()) + 1 if len(ti) else 0)
    h = HotaDevice(len(go) - 1, max(n_g, 1), max(n_t, 1), max(int((ng.astype(np.int64) * nt).sum()), 1), int(ng.max()), int(nt.max()))
    d = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()
    h.run(d(gb.reshape(-1, 4), np.float64), d(gi, np.int32), d(go, np.int32), d(tb.reshape(-1, 4), np.float64), d(ti, np.int32),
          d(to, np.int32), n_g, n_t)
    return h.result()


@pytest.mark.parametrize("name", ["hota_boxes_bytetrack_c2", "hota_boxes_random_s7", "hota_boxes_random_wide_s8"])
def test_device_hota_equals_vendored_trackeval(name):
    """Counts exact; float fields bit-equal (the kernel follows NumPy's summation orders and scipy's solver)."""
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    res = _device_hota(g["gt_boxes"], g["gt_ids"], g["gt_off"], g["tr_boxes"], g["tr_ids"], g["tr_off"])
    for k in ("HOTA_TP", "HOTA_FN", "HOTA_FP"):
        assert np.array_equal(res[k], g["ref_" + k]), k
    for k in FIELDS:
        assert np.array_equal(res[k], g["ref_" + k]), (k, np.abs(res[k] - g["ref_" + k]).max())


@pytest.mark.parametrize("seed,F,n_ids,n_false", [(11, 40, 12, 2), (12, 300, 44, 4), (13, 20, 200, 30)])
def test_device_hota_equals_oracle_on_fresh_cases(seed, F, n_ids, n_false):
    from oracle.hota_np import hota_from_boxes
    from tests.golden.make_hota_golden import case_random
    c = case_random(seed, F, n_ids, n_false, 5.0)
    ref, res = hota_from_boxes(*c), _device_hota(*c)
    for k in FIELDS:
        assert np.array_equal(res[k], ref[k]), (k, np.abs(res[k] - ref[k]).max())


def test_device_hota_empty_sides_and_tracker_rows_entry():
    """No tracker rows / no gt rows (hota.py:39-52) and the convenience entry on raw tracker rows of a synthetic video."""
    from oracle.hota_np import hota_from_boxes, hota_of_tracker_rows
    from tests.golden.make_hota_golden import case_random
    from tests.util import load_golden
    from tracklab_b200.hota import hota_of_rows
    from tracklab_b200.synth import make_video
    gb, gi, go, tb, ti, to = case_random(5, 10, 6, 1, 3.0)
    z = np.zeros_like(to)
    ref, res = hota_from_boxes(gb, gi, go, tb[:0], ti[:0], z), _device_hota(gb, gi, go, tb[:0], ti[:0], z)
    for k in FIELDS:
        assert np.array_equal(res[k], ref[k]), k
    z = np.zeros_like(go)
    ref, res = hota_from_boxes(gb[:0], gi[:0], z, tb, ti, to), _device_hota(gb[:0], gi[:0], z, tb, ti, to)
    for k in FIELDS:
        assert np.array_equal(res[k], ref[k]), k
    g = load_golden("bytetrack_small_s5")
    v = make_video(**g["gen"])
    ref = hota_of_tracker_rows(v, g["rows"], g["frames"])
    keep = v.gt_identity >= 0
    det_frame = np.repeat(np.arange(v.n_frames), np.diff(v.offsets))
    d = v.dets[keep]
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    r = g["rows"]
    res = hota_of_rows(cu(np.column_stack([d[:, 0], d[:, 1], d[:, 2] - d[:, 0], d[:, 3] - d[:, 1]])), cu(v.gt_identity[keep].astype(np.int64)),
                       cu(det_frame[keep].astype(np.int64)), cu(np.column_stack([r[:, 0], r[:, 1], r[:, 2] - r[:, 0], r[:, 3] - r[:, 1]])),
                       cu(r[:, 4].astype(np.int64)), cu(g["frames"].astype(np.int64)), v.n_frames)
    for k in ("HOTA_TP", "HOTA_FN", "HOTA_FP"):
        assert np.array_equal(res[k], ref[k]), k
    for k in ("HOTA", "DetA", "AssA", "LocA"):
        assert np.allclose(res[k], ref[k], rtol=1e-12, atol=0), k     # the row-level oracle uses iou_ltwh (same values, other op order)
