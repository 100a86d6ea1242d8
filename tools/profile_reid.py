"""Kernel-level breakdown of the ReID stage (crop gather + ResNet-50) with torch.profiler (CUPTI).
Usage: python tools/profile_reid.py [crops] [precision]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from tracklab_b200.reid import ReidStageDevice
from tracklab_b200.synth import make_frames, make_video

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 20
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
legacy = len(sys.argv) > 3 and sys.argv[3] == "legacy"
arch = sys.argv[4] if len(sys.argv) > 4 else "resnet50"
video = make_video(seed=3000, n_frames=n_frames, n_ids=44)
frames = make_frames(video, 0, n_frames, device="cuda")
dets = torch.from_numpy(video.dets).cuda()
det_frame = torch.from_numpy(np.repeat(np.arange(n_frames), np.diff(video.offsets)).astype(np.int32)).cuda()
reid = ReidStageDevice(precision=precision, legacy=legacy, arch=arch)
for _ in range(3): reid.features(frames, dets, det_frame)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); reid.features(frames, dets, det_frame); e1.record(); torch.cuda.synchronize()
print(f"{dets.shape[0]} crops: {e0.elapsed_time(e1):.2f} ms -> {1e3 * e0.elapsed_time(e1) / dets.shape[0]:.1f} us/crop")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    reid.features(frames, dets, det_frame); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
