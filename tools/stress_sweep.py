"""BASELINE configs[4] stress sweep: 4K-scale association problems (D = T = 150, E = 256), cost-matrix + assignment kernels
alone, B in {1, 8, 64, 512} problems per launch. One JSON line per (kernel, B): CUDA-event time, algorithmic bytes, GB/s
(or problems/s) and the fraction of the measured HBM peak. Inputs are rewritten between timed launches only through their
size: at B = 512 every operand set exceeds nothing close to L2, so an explicit 256 MB L2 flush precedes each timed launch.
Usage: python tools/stress_sweep.py [--reps 20]   (torchrun: every rank runs the same sweep, rank 0 prints aggregate)"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tracklab_b200 import kernels



def sweep(dev, rank=0, world=1, reps=20, batches=(1, 8, 64, 512), emit=None):
    """Runs the sweep on this rank; times are the max over ranks (one problem set per rank: weak scaling, no data-path collective).
    Returns the list of records; ``emit(rec)`` is called on rank 0 for every record."""
    from tracklab_b200 import dist as tdist
    peak = 6579.6
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    D = T = 150; E = 256
    rng = np.random.default_rng(rank)

    def boxes(B, N):
        xy = rng.uniform(0, [3600, 1900], size=(B, N, 2)); wh = rng.uniform([40, 80], [240, 520], size=(B, N, 2))
        return torch.from_numpy(np.concatenate([xy, xy + wh], 2)).to(dev)

    def timeit(fn):
        for _ in range(3): fn()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        ms = tot / reps
        return tdist.max_over_ranks(ms, dev) if world > 1 else ms

    out = []
    for B in batches:
        A, Bx = boxes(B, D), boxes(B, T)
        Bx[:, :D] = A + torch.randn_like(A) * 3
        fa = torch.randn(B, D, E, device=dev); fb = fa + 0.15 * torch.randn(B, T, E, device=dev)
        cost = torch.rand(B, D, T, dtype=torch.float64, device=dev)
        idx = torch.stack([torch.randperm(T, device=dev) for _ in range(B)])
        cost.scatter_(2, idx[:, :D, None], torch.rand(B, D, 1, dtype=torch.float64, device=dev) * 0.3)
        for name, fn, nbytes, flops in [
            ("tk_iou_matrix[giou]", lambda: kernels.iou_matrix(A, Bx, "giou"), B * ((D + T) * 32 + D * T * 8), 0),
            ("tk_iou_p1_f32", lambda: kernels.iou_p1_dist(A.float(), Bx.float()), B * ((D + T) * 16 + D * T * 4), 0),
            ("tk_cosine_dist", lambda: kernels.cosine_dist(fa, fb), B * ((D + T) * E * 4 + D * T * 8), 2.0 * B * D * T * E),
            ("tk_lap_batched[limit 0.8]", lambda: kernels.lap_batched(cost, 0.8), B * (D * T * 8 + (D + T) * 4), 0)]:
            ms = timeit(fn)
            rec = {"kernel": name, "B": B, "D": D, "T": T, "E": E, "ms": ms, "us_per_problem": 1e3 * ms / B, "algorithmic_bytes": nbytes,
                   "GBps": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak, "n_gpus": world,
                   "problems_per_s_all_gpus": world * B / (ms * 1e-3)}
            if flops: rec["GFLOPs_per_gpu"] = flops / (ms * 1e-3) / 1e9
            out.append(rec)
            if rank == 0 and emit is not None: emit(rec)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser(); ap.add_argument("--reps", type=int, default=20); a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); dist.init_process_group("nccl", device_id=dev)
    sweep(dev, rank, world, a.reps, emit=lambda r: print(json.dumps(r)))
    if world > 1: dist.destroy_process_group()
