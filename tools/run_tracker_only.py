"""Run one whole-video tracker launch (for ncu captures): python tools/run_tracker_only.py [bytetrack|ocsort] [frames]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tracklab_b200.synth import make_video
from tracklab_b200.device_trackers import ByteTrackDevice, OCSortDevice

kind = sys.argv[1] if len(sys.argv) > 1 else "bytetrack"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 100
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 128
video = make_video(seed=2000, n_frames=F, n_ids=44)
dets = torch.from_numpy(video.dets).cuda()
offs = torch.from_numpy(video.offsets.astype(np.int32))[None].cuda()
trk = (ByteTrackDevice if kind == "bytetrack" else OCSortDevice)(cap_tracks=cap, cap_dets=cap)
for _ in range(3):
    trk.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rows, fc, cnt = trk.run(dets, offs)
    e1.record()
    torch.cuda.synchronize()
    print(kind, F, "frames:", e0.elapsed_time(e1) * 1e3 / F, "us/frame", int(cnt.item()), "rows")
trk.check_status()

# optional per-phase cycle breakdown (library built with -DTK_PHASE_PROF)
import ctypes
from tracklab_b200 import _lib
lib = _lib.load()
if hasattr(lib, "tk_debug_bytetrack_phases") and kind == "bytetrack":
    buf = (ctypes.c_ulonglong * 64)()
    lib.tk_debug_bytetrack_phases(buf, 1)
    trk.reset(); trk.run(dets, offs); torch.cuda.synchronize()
    lib.tk_debug_bytetrack_phases(buf, 0)
    tot = sum(buf)
    print("phase cycles/frame:", {k: int(v / F) for k, v in enumerate(buf) if v}, "total", int(tot / F))
if hasattr(lib, "tk_debug_ocsort_phases") and kind == "ocsort":
    buf = (ctypes.c_ulonglong * 64)()
    lib.tk_debug_ocsort_phases(buf, 1)
    trk.reset(); trk.run(dets, offs); torch.cuda.synchronize()
    lib.tk_debug_ocsort_phases(buf, 0)
    names = ["split", "predict", "drop", "kobs", "cost1", "lap1", "rnd1_lists(4->6)", "byte", "ocr", "miss+birth", "update1", "out+death"]
    order = [0, 1, 2, 3, 4, 5, 10, 6, 7, 8, 9, 11]
    print("ocsort phase cycles/frame:", ", ".join(f"{names[k]} {buf[k] / F:.0f}" for k in order), "| total", int(sum(buf) / F),
          "| cost1 split: init+precompute", int(buf[12] / F), "main loop (thread 0)", int(buf[13] / F), "wait", int(buf[4] / F))
