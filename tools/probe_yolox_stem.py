"""Probe: cuDNN kernels for the YOLOX-s stem (3x3 on the 16-channel focus layout -> 32 channels at 320x320, batch 50)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
torch.backends.cudnn.benchmark = True
N = 50


def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for ci, co in [(16, 32), (32, 32), (16, 64), (32, 64), (64, 64)]:
    x = torch.randn(N, ci, 320, 320, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(co, ci, 3, 3, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    print(f"3x3 s1 320x320 Cin {ci} Cout {co}: {t(lambda: F.conv2d(x, w, None, 1, 1)):.1f} us", flush=True)
# the same stem on the 2x2 space-to-depth of the focus tensor: 64 channels at 160x160, 2x2-equivalent 3x3 taps -> 4 x 32 outputs
x = torch.randn(N, 64, 160, 160, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
for k, co in [(2, 128), (3, 128)]:
    w = torch.randn(co, 64, k, k, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    xp = F.pad(x, (1, 1, 1, 1)) if k == 2 else x
    print(f"s2d form {k}x{k} 160x160 Cin 64 Cout {co}: {t(lambda: F.conv2d(xp, w, None, 1, 0 if k == 2 else 1)):.1f} us", flush=True)
