"""Stage breakdown of the config-2/3 shape (YOLOX-m + ResNet-50 ReID + StrongSORT), each stage timed alone on one GPU.
Usage: python tools/profile_config3.py [--frames 200] [--batch 20] [--precision bf16]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tracklab_b200.detector import YoloxDetectorDevice
from tracklab_b200.device_trackers import StrongSortDevice
from tracklab_b200.reid import ReidStageDevice
from tracklab_b200.synth import make_frames, make_video

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=200); ap.add_argument("--batch", type=int, default=20)
ap.add_argument("--variant", default="m"); ap.add_argument("--ctas", type=int, default=32); ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
dev = torch.device("cuda:0")
F, B = a.frames, a.batch
video = make_video(seed=3000, n_frames=F, n_ids=44)
frames = torch.empty((F, video.height, video.width, 3), dtype=torch.uint8, device=dev)
for f0 in range(0, F, 50):
    frames[f0:min(F, f0 + 50)] = make_frames(video, f0, min(F, f0 + 50), device=dev)
dets = torch.from_numpy(video.dets).to(dev); offs = torch.from_numpy(video.offsets.astype(np.int32)).to(dev)
det_frame = torch.from_numpy(np.repeat(np.arange(F), np.diff(video.offsets)).astype(np.int32)).to(dev)
det = YoloxDetectorDevice(a.variant, device=dev, batch=B, frames_cap=F, dets_cap=max(1 << 16, 300 * F)); det.calibrate(frames[:B])
reid = ReidStageDevice(device=dev, precision=a.precision)
trk = StrongSortDevice(reid.feature_dim, image_size=(video.width, video.height), ctas_per_video=a.ctas, device=dev)
feats = torch.empty((video.n_dets, reid.feature_dim), dtype=torch.float32, device=dev)


def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run_det():
    det.reset()
    for f0 in range(0, F, B): det.detect_batch(frames[f0:min(F, f0 + B)])


def run_reid():
    for f0 in range(0, F, B):
        f1 = min(F, f0 + B); r0, r1 = int(video.offsets[f0]), int(video.offsets[f1])
        feats[r0:r1] = reid.features(frames[f0:f1], dets[r0:r1], det_frame[r0:r1] - f0)


def run_crops():
    for f0 in range(0, F, B):
        f1 = min(F, f0 + B); r0, r1 = int(video.offsets[f0]), int(video.offsets[f1])
        reid.crops(frames[f0:f1], dets[r0:r1], det_frame[r0:r1] - f0)


def run_trk():
    trk.reset(); trk.run(dets, offs.unsqueeze(0), feats)


t_det, t_reid, t_trk = timeit(run_det), timeit(run_reid), timeit(run_trk)
t_crop = timeit(run_crops) if hasattr(reid, "crops") else float("nan")
print(f"config3 stages, {F} frames, batch {B}, {video.n_dets / F:.1f} crops/frame, ReID {a.precision}: "
      f"detector {1e3 * t_det / F:.1f} us/frame, reid {1e3 * t_reid / F:.1f} us/frame (crop kernel {1e3 * t_crop / F:.1f}), "
      f"tracker {1e3 * t_trk / F:.1f} us/frame -> serial sum {F / (t_det + t_reid + t_trk) * 1e3:.0f} FPS")
