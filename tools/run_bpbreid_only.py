"""Time the BPBReID-StrongSORT whole-video kernel: python tools/run_bpbreid_only.py [frames] [E] [K] [ctas]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tracklab_b200.synth import make_video
from tracklab_b200.device_trackers import BpbreidStrongSortDevice
F = int(sys.argv[1]) if len(sys.argv) > 1 else 200
E = int(sys.argv[2]) if len(sys.argv) > 2 else 512
K = int(sys.argv[3]) if len(sys.argv) > 3 else 6
ncta = int(sys.argv[4]) if len(sys.argv) > 4 else 8
video = make_video(seed=2000, n_frames=F, n_ids=44, emb_dim=E, n_parts=K)
d = video.dets.copy(); d[:, 2] -= d[:, 0]; d[:, 3] -= d[:, 1]
dets = torch.from_numpy(d).cuda(); offs = torch.from_numpy(video.offsets.astype(np.int32))[None].cuda()
feats = torch.from_numpy(video.embeddings).cuda(); vis = torch.from_numpy(video.visibility.astype(np.float32)).cuda()
trk = BpbreidStrongSortDevice(K, E, ctas_per_video=ncta)
for _ in range(3):
    trk.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rows, fc, cnt = trk.run(dets, offs, feats, vis); e1.record(); torch.cuda.synchronize()
    print(f"bpbreid F={F} E={E} K={K} ctas={ncta}: {e0.elapsed_time(e1) * 1e3 / F:.1f} us/frame, rows {int(cnt.item())}")
trk.check_status()
import ctypes
lib = trk.lib
if hasattr(lib, "tk_debug_bpbreid_phases"):
    buf = (ctypes.c_ulonglong * 64)()
    lib.tk_debug_bpbreid_phases(buf, 1)
    trk.reset(); trk.run(dets, offs, feats, vis); torch.cuda.synchronize()
    lib.tk_debug_bpbreid_phases(buf, 0)
    names = ["compact", "predict+xyah", "cache", "rect", "B1", "P1", "B2", "fuse+live", "costA", "lapA", "pairsA+cand", "costB", "lapB",
             "kf_update", "miss+birth", "prune+out", "B3"]
    tot = sum(buf[:17])
    print("master phases (cycles/frame):", ", ".join(f"{n} {buf[i] / F:.0f}" for i, n in enumerate(names)), f"| total {tot / F:.0f}")
