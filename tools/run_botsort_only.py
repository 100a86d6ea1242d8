"""BoT-SORT whole-video launch timed alone (us/frame): python tools/run_botsort_only.py [frames] [emb_dim]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tracklab_b200.synth import make_video
from tracklab_b200.device_trackers import BotSortDevice
from tests.golden.make_botsort_golden import YAML
from tests.golden.make_deepocsort_golden import make_affines

F = int(sys.argv[1]) if len(sys.argv) > 1 else 500
E = int(sys.argv[2]) if len(sys.argv) > 2 else 512
video = make_video(seed=2000, n_frames=F, n_ids=44, emb_dim=E)
dets = torch.from_numpy(video.dets).cuda()
embs = torch.from_numpy(np.ascontiguousarray(video.embeddings.astype(np.float32))).cuda()
offs = torch.from_numpy(video.offsets.astype(np.int32))[None].cuda()
warps = torch.from_numpy(make_affines(1, F, 0.004))[None].cuda().contiguous()
trk = BotSortDevice(E, **YAML)
for _ in range(3):
    trk.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rows, fc, cnt = trk.run(dets, offs, embs, warps)
    e1.record()
    torch.cuda.synchronize()
    print("botsort", F, "frames:", e0.elapsed_time(e1) * 1e3 / F, "us/frame", int(cnt.item()), "rows,", len(video.dets) / F, "det/frame, E", E)
trk.check_status()
