"""One pass over the kernels that are not on the headline bench path, for `ncu -k regex:...` captures (see profiles/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tracklab_b200 import kernels
from tracklab_b200.device_trackers import BpbreidStrongSortDevice, OCSortDevice, StrongSortDevice
from tracklab_b200.reid import ReidStageDevice
from tracklab_b200.synth import make_frames, make_video

F = 20
v = make_video(seed=3000, n_frames=F, n_ids=44)
frames = make_frames(v, 0, F, device="cuda")
dets = torch.from_numpy(v.dets).cuda()
det_frame = torch.from_numpy(np.repeat(np.arange(F), np.diff(v.offsets)).astype(np.int32)).cuda()
reid = ReidStageDevice(use_graphs=False)           # crop (s2d16) + stem + maxpool + ... + avgpool, eager launches
for _ in range(2):
    reid.features(frames, dets, det_frame)
x = kernels.resize_frames(frames, (640, 640), torch.float32, 1.0 / 255.0)                 # RT-DETR pre-processing, 20 frames
lg = torch.randn(F, 300, 80, device="cuda"); bx = torch.rand(F, 300, 4, device="cuda") * 0.5 + 0.25
kernels.rtdetr_decode(lg, bx, (1920, 1080), 0.4, 0)
vv = make_video(seed=2000, n_frames=100, n_ids=44, emb_dim=512)
d = torch.from_numpy(vv.dets).cuda(); o = torch.from_numpy(vv.offsets.astype(np.int32))[None].cuda(); f = torch.from_numpy(vv.embeddings).cuda()
oc = OCSortDevice(); oc.run(d, o)
ss = StrongSortDevice(512, ctas_per_video=32); ss.run(d, o, f)
vp = make_video(seed=2000, n_frames=100, n_ids=44, emb_dim=512, n_parts=6)
d2 = vp.dets.copy(); d2[:, 2] -= d2[:, 0]; d2[:, 3] -= d2[:, 1]
bp = BpbreidStrongSortDevice(6, 512, ctas_per_video=24)
bp.run(torch.from_numpy(d2).cuda(), torch.from_numpy(vp.offsets.astype(np.int32))[None].cuda(), torch.from_numpy(vp.embeddings).cuda(),
       torch.from_numpy(vp.visibility.astype(np.float32)).cuda())
torch.cuda.synchronize()
for t in (oc, ss, bp): t.check_status()
print("done")
