"""GPU parity: Deep OC-SORT whole-video kernel (C ABI tk_deepocsort_*) vs goldens of the UNMODIFIED plugin and the oracle."""
import warnings

import numpy as np
import pytest
import torch

from tests.test_oracle_cpu import DEEPOCSORT_GOLDENS, deepocsort_case

pytestmark = pytest.mark.gpu


def _run_device(video, embs, affines, hyper, min_conf, cap=128, n_copies=1):
    from tracklab_b200.device_trackers import DeepOCSortDevice, rows_to_frames
    trk = DeepOCSortDevice(embs.shape[1], **hyper, min_confidence=min_conf, cap_tracks=cap, cap_dets=cap, n_seq=n_copies)
    N = len(video.dets)
    dets = torch.from_numpy(np.concatenate([video.dets] * n_copies)).cuda()
    e = torch.from_numpy(np.concatenate([embs] * n_copies)).cuda()
    offs = torch.from_numpy(np.stack([video.offsets.astype(np.int32) + k * N for k in range(n_copies)])).cuda()
    aff = None if hyper.get("cmc_off") else torch.from_numpy(np.stack([affines] * n_copies)).cuda().contiguous()
    rows, fc, cnt = trk.run(dets, offs, e, aff, out_rows=torch.empty((n_copies * N, 8), dtype=torch.float64, device="cuda"))
    trk.check_status()
    return [rows_to_frames(rows, fc, offs[:, 0].contiguous(), seq=k) for k in range(n_copies)]


def _assert_same(rows, frames, ref_rows, ref_frames, box_tol):
    """Rows in emission order (reversed tracker list, ocsort.py:520-536): ids / det ids / classes / confidences exact, boxes to tolerance."""
    assert rows.shape == ref_rows.shape, (rows.shape, ref_rows.shape)
    assert np.array_equal(frames, ref_frames)
    bad = np.nonzero(~np.all(rows[:, 4:] == ref_rows[:, 4:], axis=1))[0]
    assert len(bad) == 0, f"first differing row {bad[0]} (frame {frames[bad[0]]}): {rows[bad[0]]} vs {ref_rows[bad[0]]}; {len(bad)} rows differ"
    err = np.abs(rows[:, :4] - ref_rows[:, :4]).max() if len(rows) else 0.0
    assert err <= box_tol, f"box error {err}"
    return err


@pytest.mark.parametrize("name", DEEPOCSORT_GOLDENS)
def test_deepocsort_matches_reference_golden(name):
    """Track ids, detection ids, classes, confidences and row order equal to the UNMODIFIED plugin; boxes (float64 Kalman / affine
    arithmetic through BLAS in the reference) within 1e-6 px."""
    g, v, e = deepocsort_case(name)
    (rows, frames), = _run_device(v, e, g["affines"], g["hyper"], g["min_conf"])
    err = _assert_same(rows, frames, g["rows"], g["frames"], box_tol=1e-6)
    print(name, "max box err", err)


def test_deepocsort_two_videos_in_one_launch_and_small_capacity():
    g, v, e = deepocsort_case("deepocsort_yaml_s7000")
    res = _run_device(v, e, g["affines"], g["hyper"], g["min_conf"], cap=96, n_copies=2)
    for rows, frames in res:
        _assert_same(rows, frames, g["rows"], g["frames"], box_tol=1e-6)


@pytest.mark.parametrize("seed,hyper", [
    (31, dict(det_thresh=0.3, max_age=15, min_hits=2, iou_threshold=0.25, delta_t=2, asso_func="giou", inertia=0.3)),
    (32, dict(det_thresh=0.0, max_age=5, min_hits=1, iou_threshold=0.3, delta_t=1, asso_func="iou", inertia=0.2, embedding_off=True)),
])
def test_deepocsort_matches_oracle_fresh_seed(seed, hyper):
    from oracle.deepocsort_np import DeepOCSortOracle
    from tests.golden.make_deepocsort_golden import make_affines
    from tracklab_b200.synth import make_video
    v = make_video(seed=seed, n_frames=120, n_ids=40, emb_dim=96, conf_range=(0.2, 1.0))
    e = np.ascontiguousarray(v.embeddings.astype(np.float32))
    aff = make_affines(seed, v.n_frames, 0.005)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_rows, ref_frames = DeepOCSortOracle(**hyper, min_confidence=0.4).run_video(v.dets, v.offsets, e, aff)
    (rows, frames), = _run_device(v, e, aff, hyper, 0.4)
    _assert_same(rows, frames, ref_rows, ref_frames, box_tol=1e-6)


def test_deepocsort_empty_and_all_filtered_frames():
    """Frames without detections are skipped by the wrapper (no predict, no CMC); frames whose detections are all below
    min_confidence still run update() with an empty array (predict + CMC + update(None) for every tracker)."""
    from oracle.deepocsort_np import DeepOCSortOracle
    from tests.golden.make_deepocsort_golden import make_affines
    from tracklab_b200.synth import make_video
    v = make_video(seed=41, n_frames=40, n_ids=10, emb_dim=16)
    offs, dets, e = v.offsets.copy(), v.dets.copy(), np.ascontiguousarray(v.embeddings.astype(np.float32))
    keep = np.ones(len(dets), dtype=bool)
    keep[offs[5]:offs[8]] = False
    dets[offs[12]:offs[14], 4] = 0.05
    new_off = np.concatenate([[0], np.cumsum([keep[offs[f]:offs[f + 1]].sum() for f in range(v.n_frames)])]).astype(np.int32)
    import dataclasses
    v2 = dataclasses.replace(v, dets=dets[keep].copy(), offsets=new_off, embeddings=e[keep].copy(), gt_identity=v.gt_identity[keep].copy())
    hyper = dict(det_thresh=0.0, max_age=6, min_hits=1, iou_threshold=0.25, delta_t=3, asso_func="giou", inertia=0.3)
    aff = make_affines(3, v.n_frames, 0.004)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_rows, ref_frames = DeepOCSortOracle(**hyper, min_confidence=0.4).run_video(v2.dets, v2.offsets, e[keep], aff)
    (rows, frames), = _run_device(v2, np.ascontiguousarray(e[keep]), aff, hyper, 0.4)
    _assert_same(rows, frames, ref_rows, ref_frames, box_tol=1e-6)
    assert not np.isin(frames, [5, 6, 7, 12, 13]).any()
