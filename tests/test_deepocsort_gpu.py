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


def test_xyxy_int_crop_rule_matches_pil():
    """TK_CROP_RULE_XYXY_INT (deep_oc_sort/ocsort.py:560-565: box.astype(int) + NumPy slice) vs PIL on the same crops: identical pixels."""
    from oracle.preprocess_np import reid_crops
    from tracklab_b200 import kernels
    rng = np.random.default_rng(4)
    frame = rng.integers(0, 256, size=(1, 1080, 1920, 3), dtype=np.uint8)
    boxes = np.array([[0, 0, 1919.9, 1079.9], [5.7, 9.2, 40.9, 80.1], [1800.3, 900.2, 1950.0, 1100.0], [100, 100, 228, 356],
                      [100.2, 50.7, 164.9, 178.9], [700.99, 300.01, 760.5, 480.99]], dtype=np.float64)
    dets = np.concatenate([boxes, np.ones((len(boxes), 3))], axis=1)
    out = kernels.crop_resize_norm(torch.from_numpy(frame).cuda(), torch.from_numpy(dets).cuda(),
                                   torch.zeros(len(boxes), dtype=torch.int32, device="cuda"), ltwh_rows=kernels.CROP_RULE_XYXY_INT)
    ref = reid_crops(frame[0], boxes, rule="xyxy_int")
    mean = np.asarray(kernels.REID_MEAN, np.float32)[None, :, None, None]
    std = np.asarray(kernels.REID_STD, np.float32)[None, :, None, None]
    assert np.array_equal(np.rint((out.cpu().numpy() * std + mean) * 255), np.rint((ref * std + mean) * 255))
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-6


def test_deepocsort_module_through_engine_equals_oracle_chain(tmp_path):
    """modules.DeepOCSORT (drop-in for deep_oc_sort_api.DeepOCSORT) on PNG frames through the engine protocol: its ids must equal
    the oracle tracker fed with the module's own stage outputs (device ReID features of the xyxy-int crops, device ECC affines
    composed over skipped frames)."""
    import types

    import cv2

    from oracle.deepocsort_np import DeepOCSortOracle
    from tests.golden.make_deepocsort_golden import YAML
    from tests.test_engine_modules_gpu import _tracking_frames
    from tracklab_b200 import kernels, modules
    from tests.engine_mirror import OfflineEngineMirror
    from tracklab_b200.synth import make_frames, make_video
    F = 14
    video = make_video(seed=3100, n_frames=F, n_ids=16)
    frames = make_frames(video, 0, F, device="cpu").numpy()
    for f in range(F):
        cv2.imwrite(str(tmp_path / f"{f:06d}.png"), frames[f][..., ::-1])
    vmd, imd, det = _tracking_frames([video])
    imd["file_path"] = [str(tmp_path / f"{f:06d}.png") for f in range(F)]
    det = det[~det["image_id"].isin([5])]                   # one frame without detections: skipped by the wrapper, affine composed
    cfg = types.SimpleNamespace(min_confidence=0.4, hyperparams=dict(YAML), reid_arch="resnet50", reid_precision="fp32",
                                synthetic_weights=True, model_weights=None, cap_tracks=128, cap_dets=128)
    mod = modules.DeepOCSORT(cfg, "cuda:0")
    assert mod.level == "image" and mod.name == "DeepOCSORT"
    out = OfflineEngineMirror([mod], vmd, imd, det).track_dataset().sort_index()
    has = out["track_id"].notna().to_numpy()
    assert has.sum() > 0.8 * (det["bbox_conf"] > 0.4).sum()
    # the same chain with the oracle tracker on the module's own stage outputs
    dsort = det.sort_index()
    ltwh = np.stack(dsort["bbox_ltwh"].to_numpy()).astype(np.float64)
    rows = np.zeros((len(dsort), 7))
    rows[:, :4] = ltwh; rows[:, 2] += rows[:, 0]; rows[:, 3] += rows[:, 1]
    rows[:, 4] = dsort["bbox_conf"].to_numpy(dtype=float); rows[:, 5] = dsort["category_id"].to_numpy(dtype=float); rows[:, 6] = dsort.index.to_numpy()
    img = dsort["image_id"].to_numpy().astype(int)
    offs = np.concatenate([[0], np.cumsum(np.bincount(img, minlength=F))])
    fr = torch.from_numpy(frames).cuda()
    feats = mod.reid.features(fr, torch.from_numpy(rows).cuda(), torch.from_numpy(img.astype(np.int32)).cuda(), ltwh_rows=kernels.CROP_RULE_XYXY_INT).cpu().numpy()
    warps, _, _ = kernels.ecc_euclidean(kernels.ecc_gray_small(fr, 0.1), 100, 1e-5, 0.1)
    aff = modules.compose_skipped_affines(warps.double().cpu().numpy().reshape(-1, 2, 3), np.diff(offs) > 0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want, wf = DeepOCSortOracle(**YAML, min_confidence=0.4).run_video(rows, offs, feats, aff)
    want = want[~pd_duplicated_first(want[:, 7], wf)]
    got_ids = out["track_id"].to_numpy(dtype=float, na_value=np.nan)
    ref_ids = np.full(len(out), np.nan)
    pos = {int(i): k for k, i in enumerate(out.index.to_numpy())}
    for r in want:
        ref_ids[pos[int(r[7])]] = r[4]
    assert np.array_equal(np.isnan(got_ids), np.isnan(ref_ids)) and np.array_equal(got_ids[has], ref_ids[has])


def pd_duplicated_first(det_ids, frames):
    """results[~results.index.duplicated(keep='first')] per frame (deep_oc_sort_api.py:88)."""
    seen, dup = set(), np.zeros(len(det_ids), dtype=bool)
    for k, (d, f) in enumerate(zip(det_ids, frames)):
        dup[k] = (f, d) in seen
        seen.add((f, d))
    return dup
