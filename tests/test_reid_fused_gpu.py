"""GPU: the fused bf16 ReID executor (s2d stem, cuDNN fused epilogues, pooling kernels, CUDA graphs) against the plain
PyTorch fp32 module with the same weights, and the pooling kernels against torch."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pooling_kernels_match_torch():
    from tracklab_b200 import kernels
    torch.manual_seed(0)
    x = torch.randn(5, 64, 37, 22, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ref = torch.nn.functional.max_pool2d(x.float(), 3, 2, 1)
    got = kernels.maxpool3x3s2(x)
    assert got.shape == ref.shape and torch.equal(got.float(), ref)          # max of bf16 values is exact
    y = torch.randn(7, 2048, 8, 4, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    assert torch.allclose(kernels.avgpool(y), y.float().mean(dim=(2, 3)), atol=1e-5, rtol=1e-5)


def test_s2d16_crop_layout_is_a_permutation_of_the_plain_crop():
    from tracklab_b200 import kernels
    from tracklab_b200.synth import make_frames, make_video
    v = make_video(seed=77, n_frames=3, n_ids=12)
    frames = make_frames(v, 0, 3, device="cuda")
    dets = torch.from_numpy(v.dets).cuda()
    det_frame = torch.from_numpy(np.repeat(np.arange(3), np.diff(v.offsets)).astype(np.int32)).cuda()
    N = dets.shape[0]
    plain = kernels.crop_resize_norm(frames, dets, det_frame, out_dtype=torch.bfloat16, channels_last=True)      # [N,3,256,128]
    buf = torch.zeros((N + 3, 16, 131, 67), dtype=torch.bfloat16, device="cuda").contiguous(memory_format=torch.channels_last)
    kernels.crop_resize_norm(frames, dets, det_frame, s2d16_out=buf)
    inner = buf[:N, :12, 2:130, 2:66].reshape(N, 2, 2, 3, 128, 64)                   # [n, py, px, c, y2, x2]
    back = inner.permute(0, 3, 4, 1, 5, 2).reshape(N, 3, 256, 128)
    assert torch.equal(back, plain)
    assert float(buf[:N, 12:].abs().max()) == 0.0 and float(buf[N:].abs().max()) == 0.0
    assert float(buf[:N, :, :2].abs().max()) == 0.0 and float(buf[:N, :, 130:].abs().max()) == 0.0


@pytest.mark.parametrize("use_graphs", [False, True])
def test_fused_resnet_matches_fp32_module(use_graphs):
    from tracklab_b200 import kernels
    from tracklab_b200.nets.resnet_fused import ResNet50Fused
    from tracklab_b200.nets.resnet_reid import build_resnet50_reid
    from tracklab_b200.synth import make_frames, make_video
    v = make_video(seed=78, n_frames=4, n_ids=25)
    frames = make_frames(v, 0, 4, device="cuda")
    dets = torch.from_numpy(v.dets).cuda()
    det_frame = torch.from_numpy(np.repeat(np.arange(4), np.diff(v.offsets)).astype(np.int32)).cuda()
    N = dets.shape[0]
    model = build_resnet50_reid(1234).cuda().eval()
    fused = ResNet50Fused(model, "cuda:0", use_graphs=use_graphs)
    with torch.no_grad():
        x32 = kernels.crop_resize_norm(frames, dets, det_frame, out_dtype=torch.float32)
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            ref = model(x32)
        finally:
            torch.backends.cudnn.allow_tf32 = tf32
        for _ in range(2):   # second call replays the captured graph
            buf = fused.input_buffer(N)
            kernels.crop_resize_norm(frames, dets, det_frame, s2d16_out=buf)
            got = fused(buf, n_valid=N).clone()
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=1)
    assert got.shape == ref.shape and float(cos.min()) > 0.999, float(cos.min())
    rel = (got - ref).norm(dim=1) / ref.norm(dim=1)
    assert float(rel.max()) < 0.05, float(rel.max())


@pytest.mark.parametrize("arch", ["osnet_x1_0", "osnet_ibn_x1_0"])
def test_osnet_stage_bf16_matches_fp32_module(arch):
    """OSNet flavours (the StrongSORT YAML default is osnet_ibn_x1_0): bf16 channels-last under a CUDA graph vs the fp32 module."""
    from tracklab_b200.reid import ReidStageDevice
    from tracklab_b200.synth import make_frames, make_video
    v = make_video(seed=79, n_frames=4, n_ids=25)
    frames = make_frames(v, 0, 4, device="cuda")
    dets = torch.from_numpy(v.dets).cuda()
    det_frame = torch.from_numpy(np.repeat(np.arange(4), np.diff(v.offsets)).astype(np.int32)).cuda()
    ref = ReidStageDevice(precision="fp32", arch=arch).features(frames, dets, det_frame)
    stage = ReidStageDevice(precision="bf16", arch=arch)
    for _ in range(2):
        got = stage.features(frames, dets, det_frame)
    assert got.shape == ref.shape == (dets.shape[0], 512)
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=1)
    assert float(cos.min()) > 0.995, float(cos.min())
