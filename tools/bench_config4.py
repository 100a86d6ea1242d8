"""Secondary measurement (BASELINE configs[3] shape): RT-DETR-r50vd + KPR-shaped part embeddings (K x 512, synthetic) + the
part-based StrongSORT association (tk_bpbreid_*), one synthetic 1080p video per GPU, one all_gather of per-video metrics.
Usage: python tools/bench_config4.py [--frames 96] [--batch 16] [--steps 2]   (torchrun for several GPUs); bench.py calls run()."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch


def run(dev, rank=0, world=1, frames=96, batch=16, steps=2, warmup=1, parts=6, dim=512, ctas=24, precision="bf16", graphs=True):
    """One video per rank; device-timed, max over ranks; returns the JSON-able record (same on every rank)."""
    import torch.distributed as dist
    from tracklab_b200 import dist as tdist
    from tracklab_b200.device_trackers import BpbreidStrongSortDevice
    from tracklab_b200.rtdetr_detector import RTDetrDetectorDevice
    from tracklab_b200.synth import make_frames, make_video
    F, B = frames, batch
    video = make_video(seed=4000 + rank, n_frames=F, n_ids=44, emb_dim=dim, n_parts=parts)
    fr = torch.empty((F, video.height, video.width, 3), dtype=torch.uint8, device=dev)
    for f0 in range(0, F, 48):
        fr[f0:min(F, f0 + 48)] = make_frames(video, f0, min(F, f0 + 48), device=dev)
    d = video.dets.copy(); d[:, 2] -= d[:, 0]; d[:, 3] -= d[:, 1]
    dets = torch.from_numpy(d).to(dev); offs = torch.from_numpy(video.offsets.astype(np.int32)).to(dev)
    feats = torch.from_numpy(video.embeddings).to(dev); vis = torch.from_numpy(video.visibility.astype(np.float32)).to(dev)
    det = RTDetrDetectorDevice(dev, 0.4, precision, use_graphs=graphs); det.calibrate(fr[:1])
    trk = BpbreidStrongSortDevice(parts, dim, ctas_per_video=ctas, device=dev)
    out_rows = torch.empty((video.n_dets, 14), dtype=torch.float64, device=dev)
    s_trk = torch.cuda.Stream(device=dev)

    def one_video():
        trk.reset()
        out_start = torch.zeros(1, dtype=torch.int32, device=dev); out_count = torch.zeros(1, dtype=torch.int32, device=dev)
        cur = torch.cuda.current_stream()
        counts = None
        for f0 in range(0, F, B):
            f1 = min(F, f0 + B)
            rows, counts = det.detect_batch(fr[f0:f1])          # detector rows are computed in full (seeded weights) and discarded
            ev = torch.cuda.Event(); ev.record(cur); s_trk.wait_event(ev)
            with torch.cuda.stream(s_trk):
                trk.run(dets, offs[f0:f1 + 1].unsqueeze(0).contiguous(), feats, vis, out_rows=out_rows, out_start=out_start, out_count=out_count)
        cur.wait_stream(s_trk)
        return out_count, counts

    def timed(n):
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): cnt, dc = one_video()
        e1.record(); torch.cuda.synchronize()
        return tdist.max_over_ranks(e0.elapsed_time(e1), dev), cnt, dc

    timed(warmup)
    ms, cnt, dc = timed(steps)
    trk.check_status()
    m = tdist.gather_video_metrics(torch.tensor([[F, video.n_dets, int(cnt.item()), float(dc.float().mean().item()), ms / steps]],
                                                dtype=torch.float64, device=dev))
    rec = {"metric": "tracking_fps_1080p_40det", "value": world * steps * F / (ms / 1e3), "unit": "frames/s", "n_gpus": world,
           "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
           "config": {"workload": f"config4 shape: RT-DETR-r50vd ({precision}, transformers module, seeded weights) + synthetic part embeddings "
                                  f"{parts}x{dim} + part-based StrongSORT ({ctas} CTAs/video), {F} frames 1080p per video, one video per GPU",
                      "detector_batch": B, "tracker_input": "generator detections + generator part embeddings (the detector's rows are computed and discarded)"},
           "per_video_rows": m[:, 0, 2].tolist(), "detector_rows_per_frame_last_batch": m[:, 0, 3].tolist(),
           "detector_cuda_graph": bool(det.use_graphs), "graph_error": getattr(det, "graph_error", None),
           "dtype": f"{precision} detector / f32 appearance / f64 association"}
    del det, trk, fr
    torch.cuda.empty_cache()
    return rec


if __name__ == "__main__":
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=96); ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=2); ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--parts", type=int, default=6); ap.add_argument("--dim", type=int, default=512); ap.add_argument("--ctas", type=int, default=24)
    ap.add_argument("--precision", default="bf16"); ap.add_argument("--no-graphs", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); dist.init_process_group("nccl", device_id=dev)
    rec = run(dev, rank, world, a.frames, a.batch, a.steps, a.warmup, a.parts, a.dim, a.ctas, a.precision, not a.no_graphs)
    if rank == 0:
        print(json.dumps(rec))
    if world > 1: dist.destroy_process_group()
