"""Micro-benchmark of the tcgen05 1x1-convolution GEMM (csrc/conv1x1_tc.cu) on the layer shapes of the detector / ReID executors,
against the library path it replaces (cuDNN 1x1 convolution + tk_bias_act_nhwc, or cuDNN's fused conv+bias+ReLU for ResNet).
Shapes are recorded from one real forward pass; every shape is timed with CUDA events over buffers rotated to defeat L2 reuse.
    python tools/bench_conv1x1.py [--net yolox_s|yolox_m|resnet50] [--batch 50]     (env TK_C1_EPI_WARPS=4|8|12, TK_C1_ONE_CTA=1)"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tracklab_b200 import kernels

ap = argparse.ArgumentParser()
ap.add_argument("--net", default="yolox_s"); ap.add_argument("--batch", type=int, default=50); ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda:0")
shapes = []
orig = kernels.conv1x1_bias_act


def rec(x, w, bias, dst=None, dst_offset=0, act=1, residual=None, res_offset=0):
    M = x.numel() // x.shape[1]
    shapes.append((M, w.shape[1], w.shape[0], act, residual is not None, x.shape[2], x.shape[3]))
    return orig(x, w, bias, dst=dst, dst_offset=dst_offset, act=act, residual=residual, res_offset=res_offset)


orig3 = kernels.conv3x3_bias_act


def rec3(x, w, bias, dst=None, dst_offset=0, act=1, residual=None, res_offset=0):
    M = x.numel() // x.shape[1]
    shapes.append((M, w.shape[1], w.shape[0], act, residual is not None, x.shape[2], x.shape[3], 3))
    return orig3(x, w, bias, dst=dst, dst_offset=dst_offset, act=act, residual=residual, res_offset=res_offset)


kernels.conv1x1_bias_act = rec
kernels.conv3x3_bias_act = rec3
with torch.no_grad():
    if a.net.startswith("yolox"):
        from tracklab_b200.nets.yolox import build_yolox
        from tracklab_b200.nets.yolox_fused import YoloxFused
        m = build_yolox(a.net[-1], 1, 1234).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
        ex = YoloxFused(m, dev, use_tc=True)
        x = torch.zeros((a.batch, YoloxFused.STEM_IN, 320, 320), dtype=torch.bfloat16, device=dev).contiguous(memory_format=torch.channels_last)
        ex(x)
    else:
        from tracklab_b200.nets.resnet_fused import ResNet50Fused
        from tracklab_b200.nets.resnet_reid import build_resnet50_reid
        ex = ResNet50Fused(build_resnet50_reid(1234), dev, use_graphs=False, use_tc=True)
        ex(ex.input_buffer(a.batch))
kernels.conv1x1_bias_act = orig
kernels.conv3x3_bias_act = orig3
torch.cuda.synchronize()


def timeit(fn, reps):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


tot_tc = tot_lib = tot_bytes = 0.0
rows = []
uniq = {}
for s in shapes:
    uniq[s] = uniq.get(s, 0) + 1
for key, cnt in uniq.items():
    M, K, N, act, has_res, H, W = key[:7]
    ks = key[7] if len(key) > 7 else 1
    B = M // (H * W)
    nbuf = max(2, min(6, int(300e6 // max(1, M * (K + N) * 2)) + 1))
    xs = [torch.randn((B, K, H, W), device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for _ in range(nbuf)]
    ds = [torch.empty((B, N, H, W), dtype=torch.bfloat16, device=dev).contiguous(memory_format=torch.channels_last) for _ in range(nbuf)]
    rs = [torch.randn((B, N, H, W), device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for _ in range(nbuf)] if has_res else None
    if ks == 3:
        w4 = (torch.randn((N, K, 3, 3), device=dev) / (9 * K) ** 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = w4
    else:
        w = (torch.randn((N, K), device=dev) / K ** 0.5).to(torch.bfloat16)
        w4 = w.reshape(N, K, 1, 1).contiguous(memory_format=torch.channels_last)
    tcfn = kernels.conv3x3_bias_act if ks == 3 else kernels.conv1x1_bias_act
    bias = torch.randn((N,), device=dev); b16 = bias.to(torch.bfloat16)
    t_tc = timeit(lambda i: tcfn(xs[i % nbuf], w, bias, dst=ds[i % nbuf], act=act, residual=rs[i % nbuf] if rs else None), a.reps)
    if a.net.startswith("yolox"):
        def lib(i):
            y = F.conv2d(xs[i % nbuf], w4, None, 1, ks // 2)
            kernels.bias_act(y, bias, ds[i % nbuf], 0, act, rs[i % nbuf] if rs else None)
    else:
        def lib(i):
            if has_res:
                torch.cudnn_convolution_add_relu(xs[i % nbuf], w4, rs[i % nbuf], 1.0, b16, (1, 1), (0, 0), (1, 1), 1)
            else:
                torch.cudnn_convolution_relu(xs[i % nbuf], w4, b16, (1, 1), (0, 0), (1, 1), 1)
    t_lib = timeit(lib, a.reps)
    nbytes = M * (2 * K + 2 * N + (2 * N if has_res else 0))
    flops = 2.0 * M * K * N * ks * ks
    rows.append(dict(k=ks, HW=f"{H}x{W}", M=M, K=K, N=N, act=act, res=has_res, count=cnt, tc_us=round(t_tc, 1), lib_us=round(t_lib, 1), tc_GBps=round(nbytes / t_tc / 1e3, 0),
                     tc_TFLOPs=round(flops / t_tc / 1e6, 1), lib_GBps=round(nbytes / t_lib / 1e3, 0)))
    tot_tc += t_tc * cnt; tot_lib += t_lib * cnt; tot_bytes += nbytes * cnt
for r in sorted(rows, key=lambda r: -r["tc_us"] * r["count"]):
    print(json.dumps(r))
print(json.dumps({"net": a.net, "batch": a.batch, "layers": len(shapes), "tc_total_us": round(tot_tc, 1), "lib_total_us": round(tot_lib, 1),
                  "tc_GBps": round(tot_bytes / tot_tc / 1e3, 0), "lib_GBps": round(tot_bytes / tot_lib / 1e3, 0),
                  "epi_warps": os.environ.get("TK_C1_EPI_WARPS", "8"), "one_cta": os.environ.get("TK_C1_ONE_CTA", "0")}))
