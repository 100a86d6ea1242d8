"""Probe: does torch.cudnn_convolution_relu / _add_relu run bf16 channels-last on this GPU, and is it faster than
F.conv2d + tk_bias_act? Typical ResNet-50 ReID shapes (N crops of 256x128)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from tracklab_b200 import kernels
torch.backends.cudnn.benchmark = True
N = int(sys.argv[1]) if len(sys.argv) > 1 else 768
dev = "cuda"
shapes = [  # (Cin, Cout, k, stride, H, W)
    (8, 64, 7, 2, 256, 128), (64, 64, 1, 1, 64, 32), (64, 64, 3, 1, 64, 32), (64, 256, 1, 1, 64, 32), (256, 128, 1, 1, 64, 32),
    (128, 128, 3, 2, 64, 32), (128, 512, 1, 1, 32, 16), (256, 256, 3, 1, 16, 8), (1024, 256, 1, 1, 16, 8), (512, 512, 3, 1, 8, 4), (512, 2048, 1, 1, 8, 4)]


def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for (ci, co, k, st, H, W) in shapes:
    x = torch.randn(N, ci, H, W, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(co, ci, k, k, device=dev, dtype=torch.bfloat16) * 0.05).contiguous(memory_format=torch.channels_last)
    b32 = torch.randn(co, device=dev, dtype=torch.float32)
    b16 = b32.to(torch.bfloat16)
    pad = k // 2
    def ours():
        y = F.conv2d(x, w, None, st, pad)
        return kernels.bias_act(y, b32, y, 0, 2, None)
    ref = ours().float()
    line = f"Cin {ci:4d} Cout {co:4d} k{k} s{st} {H}x{W}: conv+tk_bias_act {t(ours):8.1f} us"
    try:
        def fused(): return torch.cudnn_convolution_relu(x, w, b16, (st, st), (pad, pad), (1, 1), 1)
        y = fused()
        err = (y.float() - ref).abs().max().item()
        line += f" | cudnn_convolution_relu {t(fused):8.1f} us (cl={y.is_contiguous(memory_format=torch.channels_last)}, max diff {err:.3f})"
    except Exception as e:
        line += f" | cudnn_convolution_relu FAILED {type(e).__name__}: {str(e)[:80]}"
    if st == 1 and ci != 8:
        try:
            z = torch.randn_like(ref).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            def fused2(): return torch.cudnn_convolution_add_relu(x, w, z, 1.0, b16, (st, st), (pad, pad), (1, 1), 1)
            fused2()
            line += f" | add_relu {t(fused2):8.1f} us"
        except Exception as e:
            line += f" | add_relu FAILED {str(e)[:60]}"
    def plain(): return F.conv2d(x, w, None, st, pad)
    line += f" | conv only {t(plain):8.1f} us"
    print(line, flush=True)
