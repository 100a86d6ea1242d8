"""GPU parity: StrongSORT whole-video kernel (C ABI) vs reference goldens and the oracle."""
import numpy as np
import pytest
import torch

from tests.util import assert_rows_match, load_golden
from tracklab_b200.synth import make_video

pytestmark = pytest.mark.gpu


def _run_device(video, hyper, min_conf, ncta=8):
    from tracklab_b200.device_trackers import StrongSortDevice, rows_to_frames
    trk = StrongSortDevice(video.embeddings.shape[1], **hyper, min_confidence=min_conf, image_size=(video.width, video.height),
                           ctas_per_video=ncta)
    dets = torch.from_numpy(video.dets).cuda()
    offs = torch.from_numpy(video.offsets.astype(np.int32))[None].cuda()
    rows, fc, cnt = trk.run(dets, offs, torch.from_numpy(video.embeddings).cuda())
    trk.check_status()
    return rows_to_frames(rows, fc, torch.zeros(1, dtype=torch.int32))


@pytest.mark.parametrize("name", ["strongsort_s4000", "strongsort_budget8_s4001"])
@pytest.mark.parametrize("ncta", [1, 8])
def test_strongsort_matches_reference_golden(name, ncta):
    g = load_golden(name)
    video = make_video(**g["gen"])
    rows, frames = _run_device(video, g["hyper"], g["min_conf"], ncta)
    # strict: track ids, detection ids and the int()-truncated boxes are bit-equal to the UNMODIFIED plugin's (the device solver
    # is scipy's algorithm including its tie-breaking, csrc/lsap_scipy.cuh, and unmatched detections keep the reference's order)
    assert_rows_match(rows, frames, g["rows"], g["frames"], box_tol=0.0)


def test_strongsort_matches_oracle_fresh_seed():
    from oracle.strongsort_np import StrongSortOracle
    video = make_video(seed=41, n_frames=120, n_ids=40, emb_dim=256)
    hyper = dict(max_dist=0.16, max_iou_dist=0.55, max_age=30, max_unmatched_preds=0, n_init=3, nn_budget=50, mc_lambda=0.995,
                 ema_alpha=0.9)
    ref_rows, ref_frames = StrongSortOracle(**hyper, min_confidence=0.4, image_size=(video.width, video.height)).run_video(
        video.dets, video.offsets, video.embeddings)
    rows, frames = _run_device(video, hyper, 0.4)
    assert_rows_match(rows, frames, ref_rows, ref_frames, box_tol=0.0)


@pytest.mark.parametrize("golden", ["strongsort_e2e_s5000", "strongsort_e2e_s5001"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_strongsort_end_to_end_with_reid_matches_reference_plugin(precision, golden):
    """frames -> crop kernel -> ResNet-50 -> StrongSORT kernel vs the UNMODIFIED plugin incl. its in-tracker ReID on the
    same synthetic frames (tests/golden/strongsort_e2e_s500{0,1}.npz, make_golden.py: run_strongsort_end_to_end; s5001 = 200
    frames / 44 identities / 8115 rows). Ids and boxes bit-equal for the fp32 AND the bf16 backbone; the cosine distances of the
    two differ by < 1e-4 (printed), inside north_star's tolerance, which is why bf16 is the default of the ReID stage."""
    from tracklab_b200.device_trackers import StrongSortDevice, rows_to_frames
    from tracklab_b200.reid import ReidStageDevice
    from tracklab_b200.synth import make_frames
    g = load_golden(golden)
    video = make_video(**g["gen"])
    frames = make_frames(video, 0, video.n_frames, device="cuda")
    dets = torch.from_numpy(video.dets).cuda()
    offs = torch.from_numpy(video.offsets.astype(np.int32))[None].cuda()
    det_frame = torch.from_numpy(np.repeat(np.arange(video.n_frames), np.diff(video.offsets)).astype(np.int32)).cuda()
    feats = ReidStageDevice(precision=precision).features(frames, dets, det_frame)
    trk = StrongSortDevice(feats.shape[1], **g["hyper"], min_confidence=g["min_conf"], image_size=(video.width, video.height))
    rows, fc, cnt = trk.run(dets, offs, feats)
    trk.check_status()
    got, gf = rows_to_frames(rows, fc, torch.zeros(1, dtype=torch.int32))
    assert_rows_match(got, gf, g["rows"], g["frames"], box_tol=0.0)
