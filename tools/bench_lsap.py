"""Latency of the one-warp scipy-exact solver (tk_lsap_scipy_batched) vs the CTA solver (tk_lap_batched) on single problems."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tracklab_b200 import kernels

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n

rng = np.random.default_rng(0)
for (n, m, kind) in [(36, 40, "random"), (80, 80, "random"), (80, 80, "extended"), (150, 150, "random")]:
    if kind == "random":
        c = rng.random((1, n, m))
    else:                      # lap extension of a 36 x 44 tracking-like matrix: mostly zeros (no overlap), one strong entry per row
        a, b = 36, 44
        real = np.zeros((a, b)); idx = rng.permutation(b)[:a]; real[np.arange(a), idx] = -rng.uniform(0.5, 1.5, a)
        c = np.full((1, a + b, a + b), real.max() + 1.0); c[0, a:, b:] = 0.0; c[0, :a, :b] = real
    ct = torch.from_numpy(np.ascontiguousarray(c)).cuda()
    t1 = timeit(lambda: kernels.lsap_scipy_batched(ct))
    t2 = timeit(lambda: kernels.lap_batched(ct))
    print(f"{kind} {c.shape[1]}x{c.shape[2]}: lsap_scipy {t1:.1f} us, lap_cta {t2:.1f} us (includes ~10 us launch + allocation overhead)")
