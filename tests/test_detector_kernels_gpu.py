"""GPU parity of the detector-side kernels (letterbox, YOLOX decode+NMS, row packing) vs the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", [(1080, 1920), (2160, 3840), (720, 1280), (1000, 1777), (1280, 720), (540, 960)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_letterbox_bit_exact_vs_cv2(hw, dtype):
    from oracle.preprocess_np import letterbox_yolox
    from tracklab_b200 import kernels
    h, w = hw
    rng = np.random.default_rng(h * 7 + w)
    frames = rng.integers(0, 256, size=(2, h, w, 3), dtype=np.uint8)
    out, ratio = kernels.letterbox(torch.from_numpy(frames).cuda(), 640, dtype, swap_rb=False)
    for b in range(2):
        ref, r = letterbox_yolox(frames[b], 640)
        assert r == ratio
        got = out[b].float().cpu().numpy()
        assert np.array_equal(got, ref), (np.abs(got - ref).max(), (got != ref).mean())  # 0..255 are exact in bf16


def test_letterbox_swap_rb_and_strided_batch():
    from oracle.preprocess_np import letterbox_yolox
    from tracklab_b200 import kernels
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, size=(3, 1080, 1920, 3), dtype=np.uint8)
    out, _ = kernels.letterbox(torch.from_numpy(frames).cuda(), 640, torch.float32, swap_rb=True)
    for b in range(3):
        ref, _ = letterbox_yolox(frames[b][..., ::-1].copy(), 640)
        assert np.array_equal(out[b].cpu().numpy(), ref)


def _synthetic_pred(rng, B, nc, n_obj):
    A = 8400
    pred = np.zeros((B, A, 5 + nc), dtype=np.float32)
    pred[..., :2] = rng.uniform(-0.5, 1.5, size=(B, A, 2))
    pred[..., 2:4] = rng.uniform(0.5, 2.5, size=(B, A, 2))
    pred[..., 4] = rng.uniform(0.0, 0.6, size=(B, A))
    pred[..., 5:] = rng.uniform(0.0, 0.9, size=(B, A, nc))
    for b in range(B):
        hot = rng.choice(A, size=n_obj, replace=False)
        pred[b, hot, 4] = rng.uniform(0.85, 1.0, size=n_obj)
        pred[b, hot, 5 + rng.integers(0, nc, size=n_obj)] = rng.uniform(0.9, 1.0, size=n_obj)
        # clusters of near-duplicates around the hot anchors so NMS has work to do
        for a in hot[: n_obj // 2]:
            for d in (1, 2):
                if a + d < A:
                    pred[b, a + d] = pred[b, a]
                    pred[b, a + d, 4] *= (0.97 if d == 1 else 0.95)  # distinct scores: tie order is a NumPy sort detail
    return pred


@pytest.mark.parametrize("nc,n_obj", [(1, 40), (1, 300), (3, 120), (1, 0)])
def test_yolox_nms_matches_oracle(nc, n_obj):
    from oracle.yolox_post_np import yolox_postprocess
    from tracklab_b200 import kernels
    rng = np.random.default_rng(nc * 100 + n_obj)
    pred = _synthetic_pred(rng, 4, nc, n_obj)
    ratio = 1.0 / 3.0
    boxes, scores, cls, count, status = kernels.yolox_nms(torch.from_numpy(pred).cuda(), ratio, 640, logits=False,
                                                          max_out=1024)
    assert int(status.item()) == 0
    for b in range(4):
        rb, rs, rc = yolox_postprocess(pred[b], np.float32(ratio))
        k = int(count[b].item())
        assert k == len(rs)
        gs = scores[b, :k].cpu().numpy()
        gb = boxes[b, :k].cpu().numpy()
        gc = cls[b, :k].cpu().numpy()
        # same kept set: scores are exact float32 products; compare as sorted multisets
        o1 = np.lexsort((gb[:, 0], -gs)); o2 = np.lexsort((rb[:, 0], -rs))
        assert np.array_equal(gs[o1], rs[o2])
        assert np.array_equal(gc[o1], rc[o2])
        assert np.allclose(gb[o1], rb[o2], rtol=1e-5, atol=1e-3)  # expf vs np.exp: few ulp on ~1e3 px


def test_pack_detections_rows_match_wrapper():
    from oracle.yolox_post_np import wrapper_rows
    from tracklab_b200 import kernels
    rng = np.random.default_rng(0)
    B, K = 5, 64
    cnt = rng.integers(0, 50, size=B).astype(np.int32)
    xy = rng.uniform(-50, 1900, size=(B, K, 2)).astype(np.float32)
    wh = rng.uniform(5, 300, size=(B, K, 2)).astype(np.float32)
    boxes = np.concatenate([xy, xy + wh], axis=2).astype(np.float32)
    scores = rng.uniform(0.7, 1, size=(B, K)).astype(np.float32)
    cls = np.zeros((B, K), dtype=np.int32)
    dev = "cuda"
    dets = torch.zeros((int(cnt.sum()) + 10, 7), dtype=torch.float64, device=dev)
    offs = torch.zeros((B + 4,), dtype=torch.int32, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    first = torch.tensor([7, 2], dtype=torch.int32, device=dev)
    kernels.pack_detections(torch.from_numpy(boxes).to(dev), torch.from_numpy(scores).to(dev), torch.from_numpy(cls).to(dev),
                            torch.from_numpy(cnt).to(dev), 1920, 1080, first, dets, offs, status)
    assert int(status.item()) == 0
    o = offs.cpu().numpy()[2:2 + B + 1]
    assert first.cpu().tolist() == [7 + int(cnt.sum()), 2 + B]
    assert o[0] == 7 and np.array_equal(np.diff(o), cnt)
    for b in range(B):
        ref = wrapper_rows(boxes[b, :cnt[b]], 1920, 1080, first_id=o[b])
        assert np.array_equal(dets[o[b]:o[b + 1]].cpu().numpy(), ref)


def test_streamed_video_with_drain_schedule_equals_resident_video():
    """DetectTrackPipeline from pinned host memory (H2D overlapped, last batch drained as 25/13/12 frames on their own CUDA
    graphs) against the HBM-resident run: identical tracker rows on the same tracker input, and the same detector rows up to
    the bf16 algorithm choice cuDNN makes per batch size (row count within 2 %)."""
    from tracklab_b200.detector import YoloxDetectorDevice
    from tracklab_b200.device_trackers import ByteTrackDevice
    from tracklab_b200.synth import make_frames, make_video
    from tracklab_b200.video_pipeline import DetectTrackPipeline
    dev = torch.device("cuda:0")
    F, B = 150, 50
    video = make_video(seed=31, n_frames=F, n_ids=30)
    frames = make_frames(video, 0, F, device=dev)
    gen_dets = torch.from_numpy(video.dets).to(dev)
    gen_offs = torch.from_numpy(video.offsets.astype(np.int32)).to(dev)
    det = YoloxDetectorDevice("s", device=dev, batch=B, frames_cap=F, dets_cap=1 << 16)
    det.calibrate(frames[:B])
    trk = ByteTrackDevice(device=dev)
    pipe = DetectTrackPipeline(det, trk, B)
    assert pipe._schedule(F, True)[-3:] == [(100, 125), (125, 138), (138, 150)] and pipe._schedule(F, False)[-1] == (100, 150)
    rd = pipe.run_video(frames, tracker_dets=gen_dets, tracker_offsets=gen_offs)
    rows_d, fc_d, cnt_d = rd.out_rows, rd.out_fc, rd.out_count
    torch.cuda.synchronize()
    a = (rows_d[: int(cnt_d)].clone(), fc_d.clone(), int(det.cursor[0]), det.offsets[: F + 1].clone())
    assert a[2] > 0 and int(a[3][-1]) == a[2]
    host = frames.cpu().pin_memory()
    for _ in range(2):   # the second pass replays the tail graphs
        rh = pipe.run_video(host, tracker_dets=gen_dets, tracker_offsets=gen_offs)
        rows_h, fc_h, cnt_h = rh.out_rows, rh.out_fc, rh.out_count
        torch.cuda.synchronize()
        assert int(cnt_h) == int(cnt_d) and torch.equal(rows_h[: int(cnt_h)], a[0]) and torch.equal(fc_h, a[1])
        n = int(det.cursor[0])
        assert abs(n - a[2]) <= 0.02 * a[2] and int(det.offsets[F]) == n
        assert torch.equal(det.offsets[:101], a[3][:101])          # the full batches are the same graphs in both runs
    det.check_status(); trk.check_status()
