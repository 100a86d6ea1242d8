"""Print the first frame where a device tracker and its oracle disagree (debug aid, GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tracklab_b200.synth import make_video
from tracklab_b200.device_trackers import OCSortDevice, rows_to_frames
from oracle.ocsort_np import OCSortOracle

np.set_printoptions(linewidth=200, precision=3, suppress=True)
for asso in ["iou", "giou", "diou", "ciou"]:
    video = make_video(seed=21, n_frames=150, n_ids=50, conf_range=(0.2, 1.0))
    hyper = dict(det_thresh=0.5, max_age=20, min_hits=2, iou_threshold=0.25, delta_t=2, asso_func=asso, inertia=0.3, use_byte=True)
    for k, v in [("base", {}), ("nobyte", dict(use_byte=False)), ("dt1", dict(delta_t=1)), ("minhits1", dict(min_hits=1))]:
        h = dict(hyper, **v)
        ref, rf = OCSortOracle(**h, min_confidence=0.4).run_video(video.dets, video.offsets)
        trk = OCSortDevice(**h, min_confidence=0.4)
        rows, fc, cnt = trk.run(torch.from_numpy(video.dets).cuda(), torch.from_numpy(video.offsets.astype(np.int32))[None].cuda())
        got, gf = rows_to_frames(rows, fc, torch.zeros(1, dtype=torch.int32))
        first = None
        for f in range(video.n_frames):
            a = got[gf == f]; b = ref[rf == f]
            a = a[np.argsort(a[:, 7])]; b = b[np.argsort(b[:, 7])]
            if a.shape != b.shape or not np.array_equal(a[:, [4, 7]], b[:, [4, 7]]):
                first = f
                break
        print(asso, k, "first differing frame:", first, "status", trk.status())
        if first is not None and k == "base" and asso == "iou":
            print("device:\n", a[:, [4, 7, 6]].T, "\noracle:\n", b[:, [4, 7, 6]].T)
            d = video.frame(first)
            print("dets conf of frame", np.round(d[:, 4], 3))
