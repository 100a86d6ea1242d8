"""BPBReID-StrongSORT whole-video kernel vs goldens of the reference plugin and vs the oracle (GPU)."""
import numpy as np
import pytest
import torch

from tests.util import BPB_KEYS, BPB_ORACLE_KEYS, assert_bpbreid_rows_match, load_bpbreid_golden
from tracklab_b200.synth import make_video

pytestmark = pytest.mark.gpu


def _ltwh_rows(v):
    d = v.dets.copy()
    d[:, 2] = v.dets[:, 2] - v.dets[:, 0]
    d[:, 3] = v.dets[:, 3] - v.dets[:, 1]
    return d


def _run_device(v, hyper, ncta=8, chunks=1, cap=128):
    from tracklab_b200.device_trackers import BpbreidStrongSortDevice, rows_to_frames
    dev = torch.device("cuda:0")
    K, E = v.embeddings.shape[1:]
    trk = BpbreidStrongSortDevice(K, E, **{k: hyper[k] for k in BPB_ORACLE_KEYS if k in hyper}, ctas_per_video=ncta, cap_tracks=(8 * cap if cap >= 128 else cap), cap_dets=cap)
    dets = torch.from_numpy(_ltwh_rows(v)).to(dev)
    feats = torch.from_numpy(v.embeddings).to(dev)
    vis = torch.from_numpy(v.visibility.astype(np.float32)).to(dev)
    offs = torch.from_numpy(v.offsets).to(dev)
    out_rows = torch.empty((max(1, v.n_dets), 14), dtype=torch.float64, device=dev)
    out_start = torch.zeros(1, dtype=torch.int32, device=dev)
    out_count = torch.zeros(1, dtype=torch.int32, device=dev)
    F = v.n_frames
    fcs = []
    bounds = np.linspace(0, F, chunks + 1).astype(int)
    for a, b in zip(bounds[:-1], bounds[1:]):
        _, fc, _ = trk.run(dets, offs[a:b + 1].unsqueeze(0).contiguous(), feats, vis, out_rows=out_rows, out_start=out_start,
                           out_count=out_count)
        fcs.append(fc)
    trk.check_status()
    rows, fr = rows_to_frames(out_rows, torch.cat(fcs, dim=1), out_start)
    trk.close()
    return rows, fr


@pytest.mark.parametrize("name,ncta", [("bpbreid_yaml_s6000", 8), ("bpbreid_tight_s6001", 1), ("bpbreid_tight_s6001", 24)])
def test_bpbreid_matches_reference_golden(name, ncta):
    g = load_bpbreid_golden(name)
    v = make_video(**g["gen"])
    rows, fr = _run_device(v, g["hyper"], ncta=ncta)
    assert_bpbreid_rows_match(rows, fr, g["rows"], g["frames"], box_tol=1e-6, dist_tol=1e-5)


@pytest.mark.parametrize("ncta", [1, 8])
def test_bpbreid_bot_sort_matching_matches_reference_golden(ncta):
    """matching_strategy: bot_sort_matching (one stage over all tracks on the weighted sum of the Kalman position distance, the
    part-based appearance distance and 1 - IoU, sort/tracker.py:335-363,169-240, incl. the np.logical_or(.., .., out) quirk that
    leaves the spatio-temporal gate unapplied) vs the UNMODIFIED plugin (tests/golden/bpbreid_botsort_s6002.npz): exact ids."""
    g = load_bpbreid_golden("bpbreid_botsort_s6002")
    v = make_video(**g["gen"])
    rows, fr = _run_device(v, g["hyper"], ncta=ncta)
    assert_bpbreid_rows_match(rows, fr, g["rows"], g["frames"], box_tol=1e-6, dist_tol=1e-5)


def test_bpbreid_chunked_launches_equal_one_launch():
    g = load_bpbreid_golden("bpbreid_yaml_s6000")
    v = make_video(**g["gen"])
    a, fa = _run_device(v, g["hyper"], chunks=1)
    b, fb = _run_device(v, g["hyper"], chunks=7)
    assert np.array_equal(fa, fb) and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("seed", [6100, 6101])
def test_bpbreid_fresh_seed_vs_oracle(seed):
    from oracle.bpbreid_np import BpbreidStrongSortOracle
    hyper = dict(ema_alpha=0.85, mc_lambda=0.99, max_dist=0.4, max_iou_distance=0.75, max_age=20, n_init=1, min_bbox_confidence=0.25,
                 max_kalman_prediction_without_update=5)
    v = make_video(seed=seed, n_frames=140, n_ids=36, emb_dim=48, n_parts=5, conf_range=(0.1, 1.0), p_visible=0.7)
    ref_rows, ref_fr = BpbreidStrongSortOracle(**hyper).run_video(v.dets, v.offsets, v.embeddings, v.visibility)
    rows, fr = _run_device(v, hyper, ncta=6)
    assert_bpbreid_rows_match(rows, fr, ref_rows, ref_fr, box_tol=1e-6, dist_tol=1e-5)


def test_bpbreid_empty_frames_and_capacity_error():
    from tracklab_b200 import _lib
    hyper = dict(ema_alpha=0.9, mc_lambda=0.995, max_dist=0.5, max_iou_distance=0.8, max_age=300, n_init=0, min_bbox_confidence=0.0,
                 max_kalman_prediction_without_update=7)
    from oracle.bpbreid_np import BpbreidStrongSortOracle
    v = make_video(seed=6200, n_frames=60, n_ids=6, emb_dim=16, n_parts=4, p_detect=0.35, fp_rate=0.0)
    assert (np.diff(v.offsets) == 0).any()
    ref_rows, ref_fr = BpbreidStrongSortOracle(**hyper).run_video(v.dets, v.offsets, v.embeddings, v.visibility)
    rows, fr = _run_device(v, hyper, ncta=2)
    assert_bpbreid_rows_match(rows, fr, ref_rows, ref_fr)
    big = make_video(seed=6201, n_frames=10, n_ids=40, emb_dim=8, n_parts=2)
    with pytest.raises(_lib.TrackKernError):
        _run_device(big, hyper, cap=16)


def test_bpbreid_yaml_config_keeps_hundreds_of_stale_tracks():
    """n_init 0 / max_age 300 (the reference YAML): every false positive is a confirmed track for 300 frames, so the track
    table grows far beyond the detections per frame; the gate-first device path must still agree with the dense oracle."""
    from oracle.bpbreid_np import BpbreidStrongSortOracle
    hyper = dict(ema_alpha=0.9, mc_lambda=0.995, max_dist=0.5, max_iou_distance=0.8, max_age=300, n_init=0, min_bbox_confidence=0.0,
                 max_kalman_prediction_without_update=7)
    v = make_video(seed=6300, n_frames=330, n_ids=30, emb_dim=32, n_parts=6, fp_rate=0.05)
    orc = BpbreidStrongSortOracle(**hyper)
    ref_rows, ref_fr = orc.run_video(v.dets, v.offsets, v.embeddings, v.visibility)
    assert len(orc.tracks) > 300
    rows, fr = _run_device(v, hyper, ncta=8, chunks=3)
    assert_bpbreid_rows_match(rows, fr, ref_rows, ref_fr, box_tol=1e-6, dist_tol=1e-5)


def test_bpbreid_two_videos_in_one_launch():
    """n_seq = 2: two cooperative CTA groups in one launch, each equal to its own single-video oracle run."""
    from oracle.bpbreid_np import BpbreidStrongSortOracle
    from tracklab_b200.device_trackers import BpbreidStrongSortDevice, rows_to_frames
    hyper = dict(ema_alpha=0.9, mc_lambda=0.995, max_dist=0.5, max_iou_distance=0.8, max_age=30, n_init=1, min_bbox_confidence=0.2,
                 max_kalman_prediction_without_update=7)
    videos = [make_video(seed=6400 + i, n_frames=70, n_ids=18 + 6 * i, emb_dim=32, n_parts=4, conf_range=(0.1, 1.0)) for i in range(2)]
    F = videos[0].n_frames
    dets = torch.from_numpy(np.concatenate([_ltwh_rows(v) for v in videos])).cuda()
    feats = torch.from_numpy(np.concatenate([v.embeddings for v in videos])).cuda()
    vis = torch.from_numpy(np.concatenate([v.visibility.astype(np.float32) for v in videos])).cuda()
    base = np.cumsum([0] + [v.n_dets for v in videos])
    offs = torch.from_numpy(np.stack([v.offsets + base[i] for i, v in enumerate(videos)]).astype(np.int32)).cuda()
    trk = BpbreidStrongSortDevice(4, 32, **hyper, ctas_per_video=3, n_seq=2, cap_tracks=256, cap_dets=64)
    cap = max(v.n_dets for v in videos)
    out_rows = torch.empty((2 * cap, 14), dtype=torch.float64, device="cuda")
    out_start = torch.tensor([0, cap], dtype=torch.int32, device="cuda")
    rows, fc, cnt = trk.run(dets, offs, feats, vis, out_rows=out_rows, out_start=out_start)
    trk.check_status()
    for i, v in enumerate(videos):
        got, gf = rows_to_frames(rows, fc, out_start, seq=i)
        want, wf = BpbreidStrongSortOracle(**hyper).run_video(v.dets, v.offsets, v.embeddings, v.visibility)
        assert F == v.n_frames
        assert_bpbreid_rows_match(got, gf, want, wf, box_tol=1e-6, dist_tol=1e-5)


def test_bpbreid_frames_whose_detections_are_all_filtered_out():
    """A frame with rows, all below min_bbox_confidence: predict only, no tracker.update (strong_sort.py:86-89) — the group
    barriers of the cooperative kernel must still pair up on that path."""
    from oracle.bpbreid_np import BpbreidStrongSortOracle
    hyper = dict(ema_alpha=0.9, mc_lambda=0.995, max_dist=0.5, max_iou_distance=0.8, max_age=30, n_init=0, min_bbox_confidence=0.3,
                 max_kalman_prediction_without_update=3)
    v = make_video(seed=6600, n_frames=40, n_ids=10, emb_dim=16, n_parts=3, conf_range=(0.4, 1.0), fp_rate=0.0)
    for f in (5, 6, 17, 39):
        v.dets[v.offsets[f]:v.offsets[f + 1], 4] = 0.1
    ref_rows, ref_fr = BpbreidStrongSortOracle(**hyper).run_video(v.dets, v.offsets, v.embeddings, v.visibility)
    assert not np.isin(ref_fr, [5, 6, 17, 39]).any()
    for ncta in (1, 4):
        rows, fr = _run_device(v, hyper, ncta=ncta, chunks=2)
        assert_bpbreid_rows_match(rows, fr, ref_rows, ref_fr)
