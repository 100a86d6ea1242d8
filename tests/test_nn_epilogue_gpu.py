"""GPU: backbone epilogue kernels and the fused YOLOX executor vs plain PyTorch."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def test_bias_act_matches_torch():
    from tracklab_b200 import kernels
    torch.manual_seed(0)
    x = _cl(torch.randn(3, 64, 20, 28, device="cuda").bfloat16())
    b = torch.randn(64, device="cuda")
    res = _cl(torch.randn(3, 64, 20, 28, device="cuda").bfloat16())
    dst = _cl(torch.zeros(3, 160, 20, 28, device="cuda", dtype=torch.bfloat16))
    kernels.bias_act(x, b, dst, 32, 1, res)
    ref = (F.silu(x.float() + b[None, :, None, None]) + res.float()).bfloat16()
    assert torch.equal(dst[:, :32], torch.zeros_like(dst[:, :32])) and torch.equal(dst[:, 96:], torch.zeros_like(dst[:, 96:]))
    assert torch.allclose(dst[:, 32:96].float(), ref.float(), rtol=2e-2, atol=2e-2)
    y = x.clone()
    kernels.bias_act(y, b, y, 0, 2)   # in place, ReLU
    assert torch.allclose(y.float(), F.relu(x.float() + b[None, :, None, None]).bfloat16().float(), rtol=1e-2, atol=1e-2)


def test_spp_and_upsample_match_torch():
    from tracklab_b200 import kernels
    torch.manual_seed(1)
    x = _cl(torch.randn(4, 128, 20, 20, device="cuda").bfloat16())
    dst = _cl(torch.zeros(4, 512, 20, 20, device="cuda", dtype=torch.bfloat16))
    kernels.spp_pool(x, dst)
    ref = torch.cat([x] + [F.max_pool2d(x.float(), k, 1, k // 2).bfloat16() for k in (5, 9, 13)], 1)
    assert torch.equal(dst, ref)
    up = _cl(torch.zeros(4, 160, 40, 40, device="cuda", dtype=torch.bfloat16))
    kernels.upsample2x(dst, up, 32, src_offset=128, channels=128)
    assert torch.equal(up[:, 32:160], F.interpolate(dst[:, 128:256].float(), scale_factor=2, mode="nearest").bfloat16())


def test_letterbox_focus16_layout():
    from tracklab_b200 import kernels
    rng = np.random.default_rng(5)
    frames = torch.from_numpy(rng.integers(0, 256, size=(2, 1080, 1920, 3), dtype=np.uint8)).cuda()
    plain, _ = kernels.letterbox(frames, 640, torch.bfloat16, swap_rb=True)
    x16 = torch.zeros((2, 16, 320, 320), dtype=torch.bfloat16, device="cuda").contiguous(memory_format=torch.channels_last)
    kernels.letterbox(frames, 640, torch.bfloat16, swap_rb=True, out=x16, focus16=True)
    tl, bl = plain[..., ::2, ::2], plain[..., 1::2, ::2]
    tr, br = plain[..., ::2, 1::2], plain[..., 1::2, 1::2]
    assert torch.equal(x16[:, :12], torch.cat((tl, bl, tr, br), 1))
    assert torch.equal(x16[:, 12:], torch.zeros_like(x16[:, 12:]))
    # same content with a 32-channel pixel pitch (the stem layout the detector uses); channels 16..31 are never written
    x32 = torch.zeros((2, 32, 320, 320), dtype=torch.bfloat16, device="cuda").contiguous(memory_format=torch.channels_last)
    x32[:, 16:] = 7.0
    kernels.letterbox(frames, 640, torch.bfloat16, swap_rb=True, out=x32, focus16=True)
    assert torch.equal(x32[:, :16], x16) and bool((x32[:, 16:] == 7.0).all())


@pytest.mark.parametrize("variant", ["s", "m"])
def test_fused_yolox_matches_module(variant):
    from tracklab_b200 import kernels
    from tracklab_b200.nets.yolox import build_yolox
    from tracklab_b200.nets.yolox_fused import YoloxFused
    rng = np.random.default_rng(7)
    frames = torch.from_numpy(rng.integers(0, 256, size=(2, 1080, 1920, 3), dtype=np.uint8)).cuda()
    model = build_yolox(variant).cuda().bfloat16().to(memory_format=torch.channels_last)
    x, _ = kernels.letterbox(frames, 640, torch.bfloat16, swap_rb=True, channels_last=True)
    x16 = torch.zeros((2, YoloxFused.STEM_IN, 320, 320), dtype=torch.bfloat16, device="cuda").contiguous(memory_format=torch.channels_last)
    kernels.letterbox(frames, 640, torch.bfloat16, swap_rb=True, out=x16, focus16=True)
    with torch.no_grad():
        ref = model(x).float()
        ref32 = build_yolox(variant).cuda().float()(x.float())
        got = YoloxFused(model, "cuda")(x16).float()
    assert got.shape == ref.shape == (2, 8400, 6)
    scale = ref32.abs().max().item()
    e_fused, e_eager = (got - ref32).abs().max().item(), (ref - ref32).abs().max().item()
    print(variant, "max |fused - fp32|", e_fused, "max |eager bf16 - fp32|", e_eager, "scale", scale)
    assert e_fused <= max(2.0 * e_eager, 0.02 * scale) + 1e-3   # fused bf16 path is as close to fp32 as eager bf16 is
