"""Print the first frame where a device tracker and its oracle disagree (debug aid, GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tracklab_b200.synth import make_video
from tracklab_b200.device_trackers import OCSortDevice, rows_to_frames
from oracle.ocsort_np import OCSortOracle

np.set_printoptions(linewidth=220, precision=3, suppress=True)
asso = sys.argv[1] if len(sys.argv) > 1 else "diou"
video = make_video(seed=21, n_frames=150, n_ids=50, conf_range=(0.2, 1.0))
hyper = dict(det_thresh=0.5, max_age=20, min_hits=2, iou_threshold=0.25, delta_t=2, asso_func=asso, inertia=0.3, use_byte=True)
for k, v in [("base", {}), ("nobyte", dict(use_byte=False))]:
    h = dict(hyper, **v)
    ref, rf = OCSortOracle(**h, min_confidence=0.4).run_video(video.dets, video.offsets)
    trk = OCSortDevice(**h, min_confidence=0.4)
    rows, fc, cnt = trk.run(torch.from_numpy(video.dets).cuda(), torch.from_numpy(video.offsets.astype(np.int32))[None].cuda())
    got, gf = rows_to_frames(rows, fc, torch.zeros(1, dtype=torch.int32))
    fwd = {}
    for f in range(video.n_frames):
        a = got[gf == f]; b = ref[rf == f]
        a = a[np.argsort(a[:, 7])]; b = b[np.argsort(b[:, 7])]
        bad = a.shape != b.shape or not np.array_equal(a[:, 7], b[:, 7])
        if not bad:
            for x, y in zip(a[:, 4], b[:, 4]):
                if fwd.setdefault(x, y) != y:
                    bad = True
        if bad:
            print(asso, k, "first structurally differing frame:", f, "rows", a.shape, b.shape)
            sa, sb = set(a[:, 7]), set(b[:, 7])
            print(" det ids only on device:", sorted(sa - sb), " only in oracle:", sorted(sb - sa))
            common = sorted(sa & sb)
            ia = {d: t for d, t in zip(a[:, 7], a[:, 4])}; ib = {d: t for d, t in zip(b[:, 7], b[:, 4])}
            print(" id mismatches (det: device id -> oracle id, mapped expectation):", [(d, ia[d], ib[d], fwd.get(ia[d])) for d in common if fwd.get(ia[d]) != ib[d]][:12])
            break
    else:
        print(asso, k, "equal up to relabelling")
if asso == "ct_dist":
    from tests.util import load_golden
    g = load_golden("ocsort_ctdist_byte_s1003")
    video = make_video(**g["gen"])
    ref, rf = g["rows"], g["frames"]
    trk = OCSortDevice(**g["hyper"], min_confidence=g["min_conf"])
    rows, fc, cnt = trk.run(torch.from_numpy(video.dets).cuda(), torch.from_numpy(video.offsets.astype(np.int32))[None].cuda())
    print("status", trk.status())
    got, gf = rows_to_frames(rows, fc, torch.zeros(1, dtype=torch.int32))
    print("rows", got.shape, ref.shape)
    for f in range(video.n_frames):
        a = got[gf == f]; b = ref[rf == f]
        a = a[np.argsort(a[:, 7])]; b = b[np.argsort(b[:, 7])]
        if a.shape != b.shape or not np.array_equal(a[:, [4, 7]], b[:, [4, 7]]):
            print("golden: first differing frame", f); print(a[:, [4, 6, 7]]); print(b[:, [4, 6, 7]]); break
