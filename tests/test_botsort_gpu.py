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


def test_botsort_module_through_engine_equals_oracle_chain(tmp_path):
    """modules.BotSORT (drop-in for bot_sort_api.BotSORT) on PNG frames through the engine protocol: its ids must equal the oracle
    tracker fed with the module's own stage outputs (device ReID features of the StrongSORT-rule crops, device ECC warps)."""
    import types

    import cv2

    from oracle.botsort_np import BotSortOracle
    from tests.golden.make_botsort_golden import YAML
    from tests.test_engine_modules_gpu import _tracking_frames
    from tracklab_b200 import kernels, modules
    from tests.engine_mirror import OfflineEngineMirror
    from tracklab_b200.synth import make_frames, make_video
    F = 14
    video = make_video(seed=3200, n_frames=F, n_ids=16)
    frames = make_frames(video, 0, F, device="cpu").numpy()
    for f in range(F):
        cv2.imwrite(str(tmp_path / f"{f:06d}.png"), frames[f][..., ::-1])
    vmd, imd, det = _tracking_frames([video])
    imd["file_path"] = [str(tmp_path / f"{f:06d}.png") for f in range(F)]
    det = det[~det["image_id"].isin([6])]
    cfg = types.SimpleNamespace(min_confidence=0.4, hyperparams=dict(YAML), reid_arch="resnet50", reid_precision="fp32",
                                synthetic_weights=True, model_weights=None, cap_tracks=128, cap_dets=128)
    mod = modules.BotSORT(cfg, "cuda:0")
    assert mod.level == "image" and mod.name == "BotSORT"
    out = OfflineEngineMirror([mod], vmd, imd, det).track_dataset().sort_index()
    has = out["track_id"].notna().to_numpy()
    assert has.sum() > 0.5 * (det["bbox_conf"] > 0.4).sum()
    dsort = det.sort_index()
    ltwh = np.stack(dsort["bbox_ltwh"].to_numpy()).astype(np.float64)
    rows = np.zeros((len(dsort), 7))
    rows[:, :4] = ltwh; rows[:, 2] += rows[:, 0]; rows[:, 3] += rows[:, 1]
    rows[:, 4] = dsort["bbox_conf"].to_numpy(dtype=float); rows[:, 5] = dsort["category_id"].to_numpy(dtype=float); rows[:, 6] = dsort.index.to_numpy()
    img = dsort["image_id"].to_numpy().astype(int)
    offs = np.concatenate([[0], np.cumsum(np.bincount(img, minlength=F))])
    fr = torch.from_numpy(frames).cuda()
    feats = mod.reid.features(fr, torch.from_numpy(rows).cuda(), torch.from_numpy(img.astype(np.int32)).cuda(), ltwh_rows=kernels.CROP_RULE_STRONGSORT).cpu().numpy()
    w, _, _ = kernels.ecc_euclidean(kernels.ecc_gray_small(fr, 0.1), 100, 1e-5, 0.1)
    warps = modules.compose_skipped_affines(w.double().cpu().numpy().reshape(-1, 2, 3), np.diff(offs) > 0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want, wf = BotSortOracle(**YAML, min_confidence=0.4).run_video(rows, offs, feats.copy(), warps)
    got_ids = out["track_id"].to_numpy(dtype=float, na_value=np.nan)
    ref_ids = np.full(len(out), np.nan)
    pos = {int(i): k for k, i in enumerate(out.index.to_numpy())}
    for r in want:
        ref_ids[pos[int(r[7])]] = r[4]
    assert np.array_equal(np.isnan(got_ids), np.isnan(ref_ids)) and np.array_equal(got_ids[has], ref_ids[has])


def test_botsort_empty_and_all_filtered_frames():
    """Frames without detections are skipped by the wrapper; frames whose detections are all below min_confidence still run update()
    (predict + GMC + lost marking) - vs the oracle, with identity warps (cmc_method 'none')."""
    import dataclasses
    from oracle.botsort_np import BotSortOracle
    from tracklab_b200.synth import make_video
    v = make_video(seed=61, n_frames=40, n_ids=10, emb_dim=16)
    offs, dets, e = v.offsets.copy(), v.dets.copy(), np.ascontiguousarray(v.embeddings.astype(np.float32))
    keep = np.ones(len(dets), dtype=bool)
    keep[offs[5]:offs[8]] = False
    dets[offs[12]:offs[14], 4] = 0.05
    new_off = np.concatenate([[0], np.cumsum([keep[offs[f]:offs[f + 1]].sum() for f in range(v.n_frames)])]).astype(np.int32)
    v2 = dataclasses.replace(v, dets=dets[keep].copy(), offsets=new_off, embeddings=e[keep].copy(), gt_identity=v.gt_identity[keep].copy())
    hyper = dict(track_high_thresh=0.45, new_track_thresh=0.5, track_buffer=6, match_thresh=0.8, lambda_=0.97)
    warps = np.tile(np.eye(2, 3), (v.n_frames, 1, 1))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_rows, ref_frames = BotSortOracle(**hyper, min_confidence=0.4).run_video(v2.dets, v2.offsets, e[keep].copy(), warps)
    (rows, frames), = _run_device(v2, np.ascontiguousarray(e[keep]), warps, hyper, 0.4)
    assert_rows_match(rows, frames, ref_rows, ref_frames, box_tol=1e-6)
    assert not np.isin(frames, [5, 6, 7]).any()
