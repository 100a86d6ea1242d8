"""Time the StrongSORT whole-video kernel: python tools/run_strongsort_only.py [frames] [E] [ctas] [budget]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tracklab_b200.synth import make_video
from tracklab_b200.device_trackers import StrongSortDevice
F = int(sys.argv[1]) if len(sys.argv) > 1 else 200
E = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ncta = int(sys.argv[3]) if len(sys.argv) > 3 else 8
budget = int(sys.argv[4]) if len(sys.argv) > 4 else 100
video = make_video(seed=2000, n_frames=F, n_ids=44, emb_dim=E)
dets = torch.from_numpy(video.dets).cuda(); offs = torch.from_numpy(video.offsets.astype(np.int32))[None].cuda()
feats = torch.from_numpy(video.embeddings).cuda()
trk = StrongSortDevice(E, nn_budget=budget, ctas_per_video=ncta)
for _ in range(3):
    trk.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rows, fc, cnt = trk.run(dets, offs, feats); e1.record(); torch.cuda.synchronize()
    print(f"strongsort F={F} E={E} ctas={ncta} budget={budget}: {e0.elapsed_time(e1) * 1e3 / F:.1f} us/frame, rows {int(cnt.item())}")
trk.check_status()
