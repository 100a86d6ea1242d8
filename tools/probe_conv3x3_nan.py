"""Where does conv3x3 produce NaN for block_n = 160? Prints NaN counts per output-channel range and pixel for a few N."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tracklab_b200 import kernels
for (B, H, W, Cin, N) in [(1, 24, 24, 192, 320), (1, 24, 24, 64, 160), (1, 16, 16, 64, 160), (1, 16, 16, 64, 144), (1, 16, 16, 64, 176), (1, 16, 16, 64, 224), (1, 16, 16, 64, 96), (1, 16, 16, 64, 80)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((B, Cin, H, W), device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn((N, Cin, 3, 3), device="cuda", generator=g) / (9 * Cin) ** 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    out = kernels.conv3x3_bias_act(x, w, None, act=0); torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.float(), w.float(), None, 1, 1)
    bad = ~torch.isfinite(out.float()) | ((out.float() - ref).abs() > 0.05 * ref.abs().max())
    per_c = bad.sum(dim=(0, 2, 3)).cpu().numpy()
    rng = [(i, int(per_c[i:i + 16].sum())) for i in range(0, N, 16)]
    print("conv3x3", (B, H, W, Cin, N), "bad", int(bad.sum()), "per 16 ch:", rng)
    x1 = x.permute(0, 2, 3, 1).reshape(-1, Cin).contiguous()
    w1 = w[:, :, 1, 1].contiguous()
    o1 = kernels.conv1x1_bias_act(x1, w1, None, act=0); torch.cuda.synchronize()
    r1 = x1.float() @ w1.float().T
    b1 = ~torch.isfinite(o1.float()) | ((o1.float() - r1).abs() > 0.05 * r1.abs().max())
    print("conv1x1", (x1.shape[0], Cin, N), "bad", int(b1.sum()), [(i, int(b1[:, i:i + 16].sum())) for i in range(0, N, 16)])
