"""GPU parity: BoT-SORT whole-video kernel (C ABI tk_botsort_*) vs goldens of the UNMODIFIED plugin and the oracle."""
import warnings

import numpy as np
import pytest
import torch

from tests.test_oracle_cpu import BOTSORT_GOLDENS, botsort_case
from tests.util import assert_rows_match

pytestmark = pytest.mark.gpu


def _run_device(video, embs, warps, hyper, min_conf, cap=128, n_copies=1):
    from tracklab_b200.device_trackers import BotSortDevice, rows_to_frames
    trk = BotSortDevice(embs.shape[1], **hyper, min_confidence=min_conf, cap_tracks=cap, cap_dets=cap, n_seq=n_copies)
    N = len(video.dets)
    dets = torch.from_numpy(np.concatenate([video.dets] * n_copies)).cuda()
    e = torch.from_numpy(np.concatenate([embs] * n_copies)).cuda()
    offs = torch.from_numpy(np.stack([video.offsets.astype(np.int32) + k * N for k in range(n_copies)])).cuda()
    w = torch.from_numpy(np.stack([warps] * n_copies)).cuda().contiguous()
    rows, fc, cnt = trk.run(dets, offs, e, w, out_rows=torch.empty((n_copies * N, 8), dtype=torch.float64, device="cuda"))
    trk.check_status()
    return [rows_to_frames(rows, fc, offs[:, 0].contiguous(), seq=k) for k in range(n_copies)]


@pytest.mark.parametrize("name", BOTSORT_GOLDENS)
def test_botsort_matches_reference_golden(name):
    """Track ids, detection ids, classes and scores equal to the UNMODIFIED plugin; boxes (float64 Kalman arithmetic through
    LAPACK / BLAS in the reference) within 1e-6 px."""
    g, v, e = botsort_case(name)
    (rows, frames), = _run_device(v, e, g["affines"], g["hyper"], g["min_conf"])
    err = assert_rows_match(rows, frames, g["rows"], g["frames"], box_tol=1e-6)
    print(name, "max box err", err)


def test_botsort_two_videos_in_one_launch():
    g, v, e = botsort_case("botsort_yaml_s8000")
    for rows, frames in _run_device(v, e, g["affines"], g["hyper"], g["min_conf"], cap=96, n_copies=2):
        assert_rows_match(rows, frames, g["rows"], g["frames"], box_tol=1e-6)


@pytest.mark.parametrize("seed,hyper", [
    (51, dict(track_high_thresh=0.5, new_track_thresh=0.55, track_buffer=10, match_thresh=0.7, lambda_=0.95)),
    (52, dict(track_high_thresh=0.35, new_track_thresh=0.4, track_buffer=40, match_thresh=0.5, proximity_thresh=0.6, appearance_thresh=0.4, lambda_=0.98)),
])
def test_botsort_matches_oracle_fresh_seed(seed, hyper):
    from oracle.botsort_np import BotSortOracle
    from tests.golden.make_deepocsort_golden import make_affines
    from tracklab_b200.synth import make_video
    v = make_video(seed=seed, n_frames=120, n_ids=40, emb_dim=96, conf_range=(0.1, 1.0))
    e = np.ascontiguousarray(v.embeddings.astype(np.float32)) * np.float32(2.0)
    warps = make_affines(seed, v.n_frames, 0.005)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_rows, ref_frames = BotSortOracle(**hyper, min_confidence=0.4).run_video(v.dets, v.offsets, e.copy(), warps)
    (rows, frames), = _run_device(v, e, warps, hyper, 0.4)
    assert_rows_match(rows, frames, ref_rows, ref_frames, box_tol=1e-6)
