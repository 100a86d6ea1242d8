"""Probe: cuDNN graph API (python frontend) conv + bias + SiLU (+ residual) in one kernel, bf16 NHWC, vs conv + tk_bias_act."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudnn, torch, torch.nn.functional as F
from tracklab_b200 import kernels
torch.backends.cudnn.benchmark = True
print("cudnn frontend", cudnn.__version__, "backend", cudnn.backend_version())
dev = "cuda"
handle = cudnn.create_handle()
cudnn.set_stream(handle=handle, stream=torch.cuda.current_stream().cuda_stream)


def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def build(x, w, b, y, stride, pad, res=None):
    g = cudnn.pygraph(handle=handle, io_data_type=cudnn.data_type.BFLOAT16, intermediate_data_type=cudnn.data_type.FLOAT,
                      compute_data_type=cudnn.data_type.FLOAT)
    X = g.tensor(name="X", dim=list(x.shape), stride=list(x.stride()), data_type=cudnn.data_type.BFLOAT16)
    W = g.tensor(name="W", dim=list(w.shape), stride=list(w.stride()), data_type=cudnn.data_type.BFLOAT16)
    B = g.tensor(name="B", dim=[1, b.shape[0], 1, 1], stride=[b.shape[0], 1, b.shape[0], b.shape[0]], data_type=cudnn.data_type.FLOAT)
    C = g.conv_fprop(image=X, weight=W, padding=[pad, pad], stride=[stride, stride], dilation=[1, 1], compute_data_type=cudnn.data_type.FLOAT)
    Cb = g.bias(name="bias", input=C, bias=B)
    Y = g.mul(a=Cb, b=g.sigmoid(input=Cb))   # SiLU (the swish binding of frontend 1.18 has swapped argument types)
    ins = {X: x, W: w, B: b}
    if res is not None:
        R = g.tensor(name="R", dim=list(res.shape), stride=list(res.stride()), data_type=cudnn.data_type.BFLOAT16)
        Y = g.add(a=Y, b=R)
        ins[R] = res
    Y.set_output(True).set_data_type(cudnn.data_type.BFLOAT16).set_dim(list(y.shape)).set_stride(list(y.stride()))
    g.validate(); g.build_operation_graph(); g.create_execution_plans([cudnn.heur_mode.A, cudnn.heur_mode.FALLBACK]); g.check_support(); g.build_plans()
    ws = torch.empty(max(1, g.get_workspace_size()), dtype=torch.uint8, device=dev)
    ins[Y] = y
    return lambda: g.execute(ins, ws, handle=handle)


N = 50
for (ci, co, k, st, H, W_) in [(16, 32, 3, 1, 320, 320), (32, 64, 3, 2, 320, 320), (64, 64, 1, 1, 160, 160), (64, 64, 3, 1, 160, 160), (128, 128, 3, 1, 80, 80),
                               (256, 256, 1, 1, 40, 40), (256, 256, 3, 1, 40, 40), (512, 512, 1, 1, 20, 20)]:
    x = torch.randn(N, ci, H, W_, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(co, ci, k, k, device=dev) * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = torch.randn(co, device=dev)
    pad = k // 2
    def ours():
        y = F.conv2d(x, w, None, st, pad)
        return kernels.bias_act(y, b, y, 0, 1, None)
    ref = ours().float()
    line = f"Cin {ci:3d} Cout {co:3d} k{k} s{st} {H}x{W_}: conv+tk_bias_act {t(ours):7.1f} us | conv only {t(lambda: F.conv2d(x, w, None, st, pad)):7.1f} us"
    try:
        y = torch.empty_like(ref, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        run = build(x, w, b, y, st, pad)
        run(); torch.cuda.synchronize()
        line += f" | cudnn graph conv+bias+swish {t(run):7.1f} us (max diff {float((y.float() - ref).abs().max()):.3f})"
        # into a channel slice of a wider (concat) buffer
        wide = torch.zeros(N, 2 * co, ref.shape[2], ref.shape[3], device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ys = wide[:, co:]
        run2 = build(x, w, b, ys, st, pad)
        run2(); torch.cuda.synchronize()
        line += f" | into concat slice {t(run2):7.1f} us (diff {float((ys.float() - ref).abs().max()):.3f})"
    except Exception as e:
        line += f" | cudnn graph FAILED {type(e).__name__}: {str(e)[:160]}"
    print(line, flush=True)
