"""Secondary measurement (BASELINE configs[2]/[3]): YOLOX-m + ResNet-50 ReID (256x128 crops) + StrongSORT on one synthetic
1080p video per GPU. Prints one JSON line. Usage: python tools/bench_config3.py [--frames 200] [--batch 20] [--steps 3]
(torchrun for several GPUs: one video per rank, single all_gather of per-video metrics at the end)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from tracklab_b200 import dist as tdist
from tracklab_b200.detector import YoloxDetectorDevice
from tracklab_b200.device_trackers import StrongSortDevice
from tracklab_b200.reid import ReidStageDevice
from tracklab_b200.synth import make_frames, make_video

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=200); ap.add_argument("--batch", type=int, default=20)
ap.add_argument("--steps", type=int, default=3); ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--variant", default="m"); ap.add_argument("--ctas", type=int, default=32)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); dist.init_process_group("nccl", device_id=dev)
F, B = a.frames, a.batch
video = make_video(seed=3000 + rank, n_frames=F, n_ids=44)
frames = torch.empty((F, video.height, video.width, 3), dtype=torch.uint8, device=dev)
for f0 in range(0, F, 50):
    frames[f0:min(F, f0 + 50)] = make_frames(video, f0, min(F, f0 + 50), device=dev)
dets = torch.from_numpy(video.dets).to(dev); offs = torch.from_numpy(video.offsets.astype(np.int32)).to(dev)
det_frame = torch.from_numpy(np.repeat(np.arange(F), np.diff(video.offsets)).astype(np.int32)).to(dev)
det = YoloxDetectorDevice(a.variant, device=dev, batch=B, frames_cap=F, dets_cap=max(1 << 16, 300 * F)); det.calibrate(frames[:B])
reid = ReidStageDevice(device=dev)
trk = StrongSortDevice(reid.feature_dim, image_size=(video.width, video.height), ctas_per_video=a.ctas, device=dev)
feats = torch.empty((video.n_dets, reid.feature_dim), dtype=torch.float32, device=dev)
out_rows = torch.empty((2 * video.n_dets, 8), dtype=torch.float64, device=dev)
s_trk = torch.cuda.Stream(device=dev)

def one_video():
    det.reset(); trk.reset()
    out_start = torch.zeros(1, dtype=torch.int32, device=dev); out_count = torch.zeros(1, dtype=torch.int32, device=dev)
    cur = torch.cuda.current_stream()
    for f0 in range(0, F, B):
        f1 = min(F, f0 + B)
        det.detect_batch(frames[f0:f1])
        r0, r1 = int(video.offsets[f0]), int(video.offsets[f1])
        # crops index frames relative to the batch view
        feats[r0:r1] = reid.features(frames[f0:f1], dets[r0:r1], det_frame[r0:r1] - f0)
        ev = torch.cuda.Event(); ev.record(cur)
        s_trk.wait_event(ev)
        with torch.cuda.stream(s_trk):
            trk.run(dets, offs[f0:f1 + 1].unsqueeze(0), feats, out_rows=out_rows, out_start=out_start, out_count=out_count)
    cur.wait_stream(s_trk)
    return out_count

def timed(n):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): cnt = one_video()
    e1.record(); torch.cuda.synchronize()
    return tdist.max_over_ranks(e0.elapsed_time(e1), dev), cnt

timed(a.warmup)
ms, cnt = timed(a.steps)
det.check_status(); trk.check_status()
m = tdist.gather_video_metrics(torch.tensor([[F, video.n_dets, int(cnt.item()), 0, ms / a.steps]], dtype=torch.float64, device=dev))
if rank == 0:
    print(json.dumps({"metric": "tracking_fps_1080p_40det", "value": world * a.steps * F / (ms / 1e3), "unit": "frames/s", "n_gpus": world,
                      "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
                      "config": {"workload": f"config3/4: YOLOX-{a.variant} + ResNet-50 ReID (256x128, ~38 crops/frame) + StrongSORT "
                                             f"(budget 100, {a.ctas} CTAs/video), {F} frames 1080p per video, one video per GPU",
                                 "detector_batch": B, "tracker_input": "generator detections; ReID features from real crops"},
                      "per_video_rows": m[:, 0, 2].tolist(), "dtype": "bf16 backbones / f32 appearance / f64 association"}))
if world > 1: dist.destroy_process_group()
