"""GPU: the hand-written tcgen05 1x1-convolution GEMM with fused epilogue (csrc/conv1x1_tc.cu) vs a float32 reference of the same
op (fp32 reference: x.float() @ w.float().T + bias -> activation -> (+ residual) -> one rounding to bf16)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, w, b, act, res):
    y = x.float() @ w.float().T + (b if b is not None else 0.0)
    if act == 1:
        y = y * torch.sigmoid(y)
    elif act == 2:
        y = torch.relu(y)
    if res is not None:
        y = y + res.float()
    if act == 3:
        y = torch.relu(y)
    return y


@pytest.mark.parametrize("M,K,N", [(128, 64, 32), (1000, 64, 32), (128 * 148 * 2 + 77, 32, 16), (5000, 96, 48), (3001, 192, 96),
                                   (4096, 128, 256), (2500, 384, 384), (777, 1536, 768), (20000, 256, 64), (300, 48, 48), (1000, 64, 320), (513, 128, 160)])
@pytest.mark.parametrize("act", [1, 0])
def test_conv1x1_matches_fp32_reference(M, K, N, act):
    from tracklab_b200 import kernels
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    x = (torch.randn((M, K), device="cuda", generator=g)).to(torch.bfloat16)
    w = (torch.randn((N, K), device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn((N,), device="cuda", generator=g)
    out = kernels.conv1x1_bias_act(x, w, b, act=act)
    torch.cuda.synchronize()
    ref = _ref(x, w, b, act, None)
    err = (out.float() - ref).abs().max().item()
    tol = 2.0 ** -7 * max(1.0, ref.abs().max().item())
    assert err <= tol, (err, tol)
    # tighter: the result is the correctly rounded bf16 of an fp32-accumulated sum up to accumulation order
    assert (out.float() - ref).abs().mean().item() < 3e-3 * max(1.0, ref.abs().mean().item())


@pytest.mark.parametrize("act", [1, 2, 3])
def test_conv1x1_writes_concat_slice_with_residual_and_leaves_the_rest(act):
    from tracklab_b200 import kernels
    B, H, W, K, N, P = 3, 20, 24, 64, 32, 96
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((B, K, H, W), device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn((N, K), device="cuda", generator=g) / 8).to(torch.bfloat16)
    b = torch.randn((N,), device="cuda", generator=g)
    res = torch.randn((B, 2 * N, H, W), device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dst = torch.full((B, P, H, W), 7.0, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    kernels.conv1x1_bias_act(x, w, b, dst=dst, dst_offset=48, act=act, residual=res, res_offset=N)
    torch.cuda.synchronize()
    xm = x.permute(0, 2, 3, 1).reshape(-1, K)
    rm = res.permute(0, 2, 3, 1).reshape(-1, 2 * N)[:, N:]
    ref = _ref(xm, w, b, act, rm)
    got = dst.permute(0, 2, 3, 1).reshape(-1, P)
    assert (got[:, 48:80].float() - ref).abs().max().item() <= 2.0 ** -7 * max(1.0, ref.abs().max().item())
    assert torch.all(got[:, :48] == 7.0) and torch.all(got[:, 80:] == 7.0)


def test_yolox_fused_executor_with_tcgen05_1x1_layers_matches_module():
    """The fused YOLOX executor with its 1x1 layers on the tcgen05 GEMM vs the plain module in fp32 (same weights)."""
    from tracklab_b200.nets.yolox import build_yolox
    from tracklab_b200.nets.yolox_fused import YoloxFused
    torch.manual_seed(0)
    model = build_yolox("s", 1, 1234, prior_prob=0.01).cuda().eval()
    ex = YoloxFused(model.to(torch.bfloat16).to(memory_format=torch.channels_last), "cuda", use_tc3=True)   # 1x1 and 3x3 layers on tcgen05
    x = torch.rand((2, 3, 640, 640), device="cuda") * 255.0
    with torch.no_grad():
        ref = model.float()(x).float()
        tl, bl, tr, br = x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]
        x16 = torch.zeros((2, YoloxFused.STEM_IN, 320, 320), device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        x16[:, :12] = torch.cat((tl, bl, tr, br), 1).to(torch.bfloat16)
        got = ex(x16).float()
    assert ex.tc_layers > 20
    denom = ref.abs().max().item()
    assert (got - ref).abs().max().item() <= 0.05 * denom, ((got - ref).abs().max().item(), denom)


def _ref3(x, w, b, act, res):
    y = torch.nn.functional.conv2d(x.float(), w.float(), b, 1, 1)
    if act == 1:
        y = y * torch.sigmoid(y)
    elif act == 2:
        y = torch.relu(y)
    if res is not None:
        y = y + res.float()
    if act == 3:
        y = torch.relu(y)
    return y


@pytest.mark.parametrize("B,H,W,Cin,N", [(2, 16, 16, 64, 64), (3, 20, 20, 128, 128), (2, 40, 40, 96, 96), (1, 80, 80, 48, 48), (2, 33, 47, 32, 16),
                                         (1, 160, 160, 32, 64), (2, 20, 20, 384, 384), (1, 24, 24, 192, 320), (5, 8, 8, 16, 48)])
@pytest.mark.parametrize("act", [1, 0])
def test_conv3x3_matches_fp32_reference(B, H, W, Cin, N, act):
    """csrc/conv3x3_tc.cu (TMA im2col boxes, BK 64 / 32 / 16, clipped patches at the borders) vs F.conv2d in float32 on the same bf16 data."""
    from tracklab_b200 import kernels
    g = torch.Generator(device="cuda").manual_seed(B * H + W + Cin + N)
    x = torch.randn((B, Cin, H, W), device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn((N, Cin, 3, 3), device="cuda", generator=g) / (9 * Cin) ** 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = torch.randn((N,), device="cuda", generator=g)
    out = kernels.conv3x3_bias_act(x, w, b, act=act)
    torch.cuda.synchronize()
    ref = _ref3(x, w, b, act, None)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2.0 ** -7 * max(1.0, ref.abs().max().item()), err
    assert (out.float() - ref).abs().mean().item() < 3e-3 * max(1.0, ref.abs().mean().item())


@pytest.mark.parametrize("act", [1, 3])
def test_conv3x3_residual_and_concat_slice(act):
    from tracklab_b200 import kernels
    B, H, W, Cin, N, P = 2, 40, 40, 64, 64, 160
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn((B, Cin, H, W), device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn((N, Cin, 3, 3), device="cuda", generator=g) / 24).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = torch.randn((N,), device="cuda", generator=g)
    res = torch.randn((B, 2 * N, H, W), device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dst = torch.full((B, P, H, W), 7.0, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    kernels.conv3x3_bias_act(x, w, b, dst=dst, dst_offset=64, act=act, residual=res, res_offset=N)
    torch.cuda.synchronize()
    ref = _ref3(x, w, b, act, res[:, N:])
    got = dst[:, 64:128].float()
    assert (got - ref).abs().max().item() <= 2.0 ** -7 * max(1.0, ref.abs().max().item())
    assert torch.all(dst[:, :64] == 7.0) and torch.all(dst[:, 128:] == 7.0)
