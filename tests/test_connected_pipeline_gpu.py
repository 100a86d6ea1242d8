"""GPU: the CONNECTED product pipeline (frames -> detector rows -> ReID features -> tracker, the tracker consuming the rows
tk_pack_detections_ex appended at the device cursor) against the oracle chain (oracle/pipeline_np.py) — exact ids.

The detector network runs in bf16 on the device and in fp32 in the CPU restatement, so the two detectors' rows agree only to
the precision of the network; the chain DOWNSTREAM of the rows (row order / id offsets / dtypes of the pack -> crop -> ReID ->
tracker hand-offs, which is what "connected" adds) is checked exactly: the oracle consumes the rows the device detector emitted
and must reproduce the device tracker's output bit for bit in its integer columns. The detector stage itself is compared with
its CPU restatement on the same frames (test_detector_rows_*)."""
import numpy as np
import pytest
import torch

from tests.util import assert_rows_match
from tracklab_b200.synth import make_frames, make_video

pytestmark = pytest.mark.gpu


def _pipe(config, B, F, **kw):
    from tracklab_b200.video_pipeline import build_pipeline
    return build_pipeline(config, device="cuda:0", batch=B, frames_cap=F, **kw)


def _frames(seed, F, n_ids=44):
    v = make_video(seed=seed, n_frames=F, n_ids=n_ids)
    return v, make_frames(v, 0, F, device="cpu").cuda()


def _prepare(pipe, frames):
    if not pipe.det.trained:
        pipe.det.calibrate(frames[:pipe.batch], target_per_image=60.0)


@pytest.mark.parametrize("config,oracle", [("config2", "bytetrack"), ("config2_ocsort", "ocsort")])
def test_config2_connected_chain_equals_oracle_on_the_detector_rows(config, oracle):
    from oracle.bytetrack_np import ByteTrackOracle
    from oracle.ocsort_np import OCSortOracle
    from tracklab_b200.video_pipeline import CONFIGS
    F, B = 48, 16
    v, frames = _frames(2000, F)
    pipe = _pipe(config, B, F)
    _prepare(pipe, frames)
    res = pipe.run_video(frames)
    torch.cuda.synchronize()
    pipe.check_status()
    h = pipe.results_to_host(res, with_detections=True)
    assert h.det_rows > F and h.det_offsets[-1] == h.det_rows
    assert np.array_equal(h.det_table[:, 6], np.arange(h.det_rows)), "running detection ids (rtmlib_api.py:42-45)"
    Orc = ByteTrackOracle if oracle == "bytetrack" else OCSortOracle
    want, wf = Orc(**CONFIGS[config]["hyper"], min_confidence=0.4).run_video(h.det_table, h.det_offsets)
    assert len(want) > 0
    # OC-SORT's lap.lapjv is un-vendored: ids up to the documented relabelling where the stand-in solver ties (tests/util.py)
    assert_rows_match(h.rows, h.frame, want, wf, box_tol=1e-9, allow_relabel=(oracle == "ocsort"))
    # pinned-host streaming gives the same rows (same graphs per batch)
    res2 = pipe.run_video(frames.cpu().pin_memory())
    torch.cuda.synchronize()
    h2 = pipe.results_to_host(res2)
    assert h2.det_rows == h.det_rows and np.array_equal(h2.rows, h.rows) and np.array_equal(h2.frame, h.frame)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_config3_connected_chain_equals_oracle_chain_on_the_detector_rows(precision):
    """YOLOX-m -> rows -> PIL-exact crops -> ResNet-50 -> StrongSORT. fp32 ReID: ids and boxes equal the oracle chain (PIL crops,
    fp32 CPU ResNet-50, NumPy StrongSORT) run on the device detector's rows. bf16 ReID (the throughput default): same check —
    the appearance distances differ by ~1e-3, the test reports whether any decision moved."""
    from oracle.pipeline_np import detect_reid_track_video
    from tracklab_b200.nets.resnet_reid import build_resnet50_reid
    from tracklab_b200.video_pipeline import CONFIGS
    F, B = 15, 6                                   # 3 batches (6/6/3): exercises the one-batch-late cursor read and a ragged tail
    v, frames = _frames(3000, F)
    pipe = _pipe("config3", B, F, reid_precision=precision)
    _prepare(pipe, frames)
    res = pipe.run_video(frames)
    torch.cuda.synchronize()
    pipe.check_status()
    h = pipe.results_to_host(res, with_detections=True)
    assert h.det_rows > F
    det_rows = [h.det_table[h.det_offsets[f]:h.det_offsets[f + 1]] for f in range(F)]
    want, wf, _, feats = detect_reid_track_video(None, build_resnet50_reid(1234).float().eval(), frames.cpu().numpy(),
                                                 CONFIGS["config3"]["hyper"], 0.4, detector_rows=det_rows)
    got_feats = res.features[:h.det_rows].cpu().numpy()
    ref_feats = np.concatenate(feats)
    a = got_feats / np.linalg.norm(got_feats, axis=1, keepdims=True)
    b = ref_feats / np.linalg.norm(ref_feats, axis=1, keepdims=True)
    dcos = np.abs((1 - a @ a.T) - (1 - b @ b.T)).max()
    print(f"connected config3 [{precision}]: {h.det_rows} detector rows, {len(h.rows)} track rows, max |d cosine distance| = {dcos:.2e}")
    if precision == "fp32":
        assert dcos < 1e-4, dcos                   # north_star: fp32 distances within 1e-4
        assert_rows_match(h.rows, h.frame, want, wf, box_tol=0.0)
    else:
        assert dcos < 2e-2, dcos
        assert h.rows.shape == want.shape and np.array_equal(np.sort(h.rows[:, 7]), np.sort(want[:, 7]))


def test_config3_bpbreid_connected_chain_equals_oracle():
    """Part-based flavour of configs[2]: ltwh rows (tk_pack_detections_ex TK_ROWS_LTWH), ReID-wrapper crop rule
    (TK_CROP_RULE_LTWH_ROUNDED), ResNet-50 feature as embeddings [D,1,2048] / visibility 1, tk_bpbreid_run."""
    from oracle.bpbreid_np import BpbreidStrongSortOracle
    from oracle.pipeline_np import kpreid_crop_box
    from tests.util import assert_bpbreid_rows_match
    from tracklab_b200.nets.resnet_reid import build_resnet50_reid
    from tracklab_b200.video_pipeline import CONFIGS
    F, B = 12, 6
    v, frames = _frames(3001, F)
    pipe = _pipe("config3_bpbreid", B, F, reid_precision="fp32")
    _prepare(pipe, frames)
    res = pipe.run_video(frames)
    torch.cuda.synchronize()
    pipe.check_status()
    h = pipe.results_to_host(res, with_detections=True)
    N = h.det_rows
    assert N > F
    # features of the oracle: ReID-wrapper crop rule + PIL resize + fp32 CPU ResNet-50 on the device detector's ltwh rows
    from PIL import Image
    from oracle.preprocess_np import REID_MEAN, REID_STD
    net = build_resnet50_reid(1234).float().eval()
    fr = frames.cpu().numpy()
    x = np.zeros((N, 3, 256, 128), np.float32)
    fo = np.repeat(np.arange(F), np.diff(h.det_offsets))
    for i in range(N):
        l, t, r, b = kpreid_crop_box(h.det_table[i, :4], v.width, v.height)
        small = np.asarray(Image.fromarray(fr[fo[i]][t:b, l:r]).resize((128, 256), Image.BILINEAR)).astype(np.float32) / np.float32(255)
        x[i] = ((small - np.asarray(REID_MEAN, np.float32)) / np.asarray(REID_STD, np.float32)).transpose(2, 0, 1)
    with torch.no_grad():
        ref = torch.cat([net(torch.from_numpy(x[i:i + 64])) for i in range(0, N, 64)]).numpy()
    got = res.features[:N].cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-4 * np.abs(ref).max()
    ltrb = h.det_table.copy()            # the oracle takes wrapper rows [l,t,r,b,..]; l + w is exact in float64 for float32 l, w
    ltrb[:, 2] += ltrb[:, 0]; ltrb[:, 3] += ltrb[:, 1]
    want, wf = BpbreidStrongSortOracle(**CONFIGS["config3_bpbreid"]["hyper"]).run_video(
        ltrb, h.det_offsets, ref[:, None, :], np.ones((N, 1), np.float32))
    assert_bpbreid_rows_match(h.rows, h.frame, want, wf, box_tol=1e-6, dist_tol=1e-4)


def test_detector_rows_match_the_cpu_restatement_and_localise_when_trained():
    """Detector stage alone: device (bf16, fused executor) vs the fp32 CPU restatement on the same frames — matched boxes; with
    the trained weights (weights/yolox_s_synth.pt) the rows are the generator's boxes."""
    from oracle.pipeline_np import detect_frame
    F = 6
    v, frames = _frames(2000, F)
    pipe = _pipe("config2", F, F)
    _prepare(pipe, frames)
    res = pipe.run_video(frames)
    torch.cuda.synchronize()
    h = pipe.results_to_host(res, with_detections=True)
    model = pipe.det.model.float().cpu().eval()
    ious, n_dev, n_cpu = [], 0, 0
    for f in range(F):
        cpu = detect_frame(model, frames[f].cpu().numpy())
        dev = h.det_table[h.det_offsets[f]:h.det_offsets[f + 1]]
        n_dev += len(dev); n_cpu += len(cpu)
        if len(cpu) and len(dev):
            a, b = torch.from_numpy(dev[:, :4]), torch.from_numpy(cpu[:, :4])
            lt = torch.maximum(a[:, None, :2], b[None, :, :2]); rb = torch.minimum(a[:, None, 2:], b[None, :, 2:])
            wh = (rb - lt).clamp(min=0); inter = wh[..., 0] * wh[..., 1]
            iou = inter / ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]))[:, None].add(((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))[None]).sub(inter)
            ious.extend(iou.max(0).values.tolist())
    ious = np.asarray(ious)
    print(f"detector rows: device {n_dev}, cpu {n_cpu}, matched IoU>0.9: {(ious > 0.9).mean():.3f}")
    assert abs(n_dev - n_cpu) <= 0.1 * n_cpu + 2 and (ious > 0.9).mean() > 0.85
    if pipe.det.trained:
        hit = 0
        for f in range(F):
            gt = torch.from_numpy(v.frame(f)[:, :4]); a = torch.from_numpy(h.det_table[h.det_offsets[f]:h.det_offsets[f + 1], :4])
            lt = torch.maximum(a[:, None, :2], gt[None, :, :2]); rb = torch.minimum(a[:, None, 2:], gt[None, :, 2:])
            wh = (rb - lt).clamp(min=0); inter = wh[..., 0] * wh[..., 1]
            iou = inter / ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]))[:, None].add(((gt[:, 2] - gt[:, 0]) * (gt[:, 3] - gt[:, 1]))[None]).sub(inter)
            hit += int((iou.max(0).values > 0.5).sum())
        assert hit >= 0.85 * v.offsets[F]
