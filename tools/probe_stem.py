"""Probe: ResNet stem (7x7/2 on RGB) as a 4x4/1 convolution on a 2x2 space-to-depth input (12 -> 16 channels), cuDNN bf16 NHWC."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
torch.backends.cudnn.benchmark = True
N = int(sys.argv[1]) if len(sys.argv) > 1 else 768
dev = "cuda"


def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


torch.manual_seed(0)
x = torch.randn(N, 3, 256, 128, device=dev)
w = torch.randn(64, 3, 7, 7, device=dev) * 0.05
b = torch.randn(64, device=dev)
ref = F.relu(F.conv2d(x, w, b, 2, 3))
# space-to-depth input, zero border: rows 2 before / 1 after
xs = x.view(N, 3, 128, 2, 64, 2).permute(0, 3, 5, 1, 2, 4).reshape(N, 12, 128, 64)           # channel = (py*2+px)*3 + c
xp = torch.zeros(N, 16, 131, 67, device=dev)
xp[:, :12, 2:130, 2:66] = xs
w2 = torch.zeros(64, 16, 4, 4, device=dev)
for ky in range(7):
    ay, py = divmod(ky - 3, 2)          # ky - 3 = 2*ay + py, py in {0,1}
    for kx in range(7):
        ax, px = divmod(kx - 3, 2)
        w2[:, (py * 2 + px) * 3:(py * 2 + px) * 3 + 3, ay + 2, ax + 2] = w[:, :, ky, kx]
y2 = F.relu(F.conv2d(xp, w2, b, 1, 0))
print("s2d formulation max abs diff vs 7x7/2 (fp32):", (y2 - ref).abs().max().item(), tuple(y2.shape))
xb = xp.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
wb = w2.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
bb = b.to(torch.bfloat16)
print(f"4x4/1 on 16ch: conv only {t(lambda: F.conv2d(xb, wb, None, 1, 0)):.1f} us, "
      f"cudnn_convolution_relu {t(lambda: torch.cudnn_convolution_relu(xb, wb, bb, (1, 1), (0, 0), (1, 1), 1)):.1f} us")
x8 = torch.zeros(N, 8, 256, 128, device=dev); x8[:, :3] = x
x8 = x8.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
w8 = torch.zeros(64, 8, 7, 7, device=dev); w8[:, :3] = w
w8 = w8.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
print(f"7x7/2 on 8ch (current): cudnn_convolution_relu {t(lambda: torch.cudnn_convolution_relu(x8, w8, bb, (2, 2), (3, 3), (1, 1), 1)):.1f} us")
# 2x2 taps on a 4x4 space-to-depth (48 -> 48 channels, stride 1 at quarter resolution would need stride handling) is not equivalent; skip.
y = torch.cudnn_convolution_relu(xb, wb, bb, (1, 1), (0, 0), (1, 1), 1)
print(f"max_pool2d torch NHWC: {t(lambda: F.max_pool2d(y, 3, 2, 1)):.1f} us; mean: {t(lambda: y.float().mean(dim=(2, 3))):.1f} us")
