"""GPU parity of the ReID crop-gather kernel vs the PIL-based oracle (integer-exact pixels)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_crop_resize_norm_matches_pil():
    from oracle.preprocess_np import reid_crops
    from tracklab_b200 import kernels
    from tracklab_b200.synth import make_frames, make_video
    video = make_video(seed=7, n_frames=3, n_ids=30)
    frames = make_frames(video, 0, 3, device="cuda")
    dets = torch.from_numpy(video.dets).cuda()
    det_frame = torch.from_numpy(np.repeat(np.arange(3), np.diff(video.offsets)).astype(np.int32)).cuda()
    out = kernels.crop_resize_norm(frames, dets, det_frame).cpu().numpy()
    fr = frames.cpu().numpy()
    worst = 0.0
    for f in range(3):
        sl = slice(video.offsets[f], video.offsets[f + 1])
        ref = reid_crops(fr[f], video.dets[sl, :4])
        worst = max(worst, float(np.abs(out[sl] - ref).max()))
        # undo the normalisation: the underlying uint8 pixels must be identical
        mean = np.asarray(kernels.REID_MEAN, np.float32)[None, :, None, None]
        std = np.asarray(kernels.REID_STD, np.float32)[None, :, None, None]
        assert np.array_equal(np.rint((out[sl] * std + mean) * 255), np.rint((ref * std + mean) * 255))
    assert worst < 1e-6


def test_crop_resize_norm_extreme_boxes_and_bf16():
    from oracle.preprocess_np import reid_crops
    from tracklab_b200 import kernels
    rng = np.random.default_rng(3)
    frame = rng.integers(0, 256, size=(1, 1080, 1920, 3), dtype=np.uint8)
    boxes = np.array([[0, 0, 1919, 1079], [5.7, 9.2, 40.9, 80.1], [1800.3, 900.2, 1950.0, 1100.0], [-20.0, -30.0, 300.5, 700.9],
                      [100, 100, 228, 356], [100.2, 50.7, 164.9, 178.9]], dtype=np.float64)
    dets = np.concatenate([boxes, np.ones((len(boxes), 3))], axis=1)
    out = kernels.crop_resize_norm(torch.from_numpy(frame).cuda(), torch.from_numpy(dets).cuda(),
                                   torch.zeros(len(boxes), dtype=torch.int32, device="cuda"))
    ref = reid_crops(frame[0], boxes)
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-6
    out16 = kernels.crop_resize_norm(torch.from_numpy(frame).cuda(), torch.from_numpy(dets).cuda(),
                                     torch.zeros(len(boxes), dtype=torch.int32, device="cuda"), out_dtype=torch.bfloat16,
                                     channels_last=True)
    assert torch.allclose(out16.float().cpu(), torch.from_numpy(ref), atol=2e-2, rtol=1e-2)


def test_reid_stage_fused_matches_fp32_module():
    """bf16 fused executor vs the same ResNet-50 in fp32: cosine similarity of the features ~ 1."""
    from tracklab_b200 import kernels
    from tracklab_b200.nets.resnet_reid import build_resnet50_reid
    from tracklab_b200.reid import ReidStageDevice
    from tracklab_b200.synth import make_frames, make_video
    video = make_video(seed=8, n_frames=2, n_ids=12)
    frames = make_frames(video, 0, 2, device="cuda")
    dets = torch.from_numpy(video.dets).cuda()
    det_frame = torch.from_numpy(np.repeat(np.arange(2), np.diff(video.offsets)).astype(np.int32)).cuda()
    stage = ReidStageDevice()
    got = stage.features(frames, dets, det_frame)
    x32 = kernels.crop_resize_norm(frames, dets, det_frame)
    with torch.no_grad():
        ref = build_resnet50_reid().cuda().float()(x32)
    assert got.shape == ref.shape == (video.n_dets, 2048)
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=1)
    print("min cosine(fused bf16, fp32)", cos.min().item())
    assert cos.min().item() > 0.995
