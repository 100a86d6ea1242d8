"""Deep OC-SORT whole-video launch timed alone (us/frame): python tools/run_deepocsort_only.py [frames] [emb_dim]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tracklab_b200.synth import make_video
from tracklab_b200.device_trackers import DeepOCSortDevice
from tests.golden.make_deepocsort_golden import YAML, make_affines

F = int(sys.argv[1]) if len(sys.argv) > 1 else 500
E = int(sys.argv[2]) if len(sys.argv) > 2 else 512
video = make_video(seed=2000, n_frames=F, n_ids=44, emb_dim=E)
dets = torch.from_numpy(video.dets).cuda()
embs = torch.from_numpy(np.ascontiguousarray(video.embeddings.astype(np.float32))).cuda()
offs = torch.from_numpy(video.offsets.astype(np.int32))[None].cuda()
aff = torch.from_numpy(make_affines(1, F, 0.004))[None].cuda().contiguous()
trk = DeepOCSortDevice(E, **YAML)
for _ in range(3):
    trk.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rows, fc, cnt = trk.run(dets, offs, embs, aff)
    e1.record()
    torch.cuda.synchronize()
    print("deepocsort", F, "frames:", e0.elapsed_time(e1) * 1e3 / F, "us/frame", int(cnt.item()), "rows,", len(video.dets) / F, "det/frame, E", E)
trk.check_status()

import ctypes
from tracklab_b200 import _lib
lib = _lib.load()
if hasattr(lib, "tk_debug_deepocsort_phases"):      # library built with EXTRA=-DTK_PHASE_PROF
    buf = (ctypes.c_ulonglong * 32)()
    lib.tk_debug_deepocsort_phases(buf, 1)
    trk.reset(); trk.run(dets, offs, embs, aff); torch.cuda.synchronize()
    lib.tk_debug_deepocsort_phases(buf, 0)
    names = {0: "filter", 1: "cmc+alpha", 2: "predict", 3: "snapshot", 4: "round1 lists", 5: "update1", 6: "ocr", 7: "miss", 8: "birth", 9: "out+death",
             10: "round1 iou/emb/cost (inside 4)", 11: "round1 solver (inside 4)"}
    print("phase cycles/frame:", ", ".join(f"{names.get(k, k)} {buf[k] / F:.0f}" for k in range(12)), "| total", int(sum(buf[:10]) / F + (buf[10] + buf[11]) / F))
