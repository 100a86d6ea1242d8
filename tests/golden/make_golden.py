"""Generate golden vectors by running the UNMODIFIED reference plugins (build container only).

    python tests/golden/make_golden.py [name ...]

Imports the tracker plugins from /root/reference/plugins/track with the third-party stand-ins of
``oracle/ref_shims`` (see its README), feeds them the synthetic videos of ``tracklab_b200.synth``
through the same per-frame filter the reference wrappers apply
(/root/reference/tracklab/wrappers/track/byte_track_api.py:50-56), and stores inputs' generator
parameters + the reference outputs as ``tests/golden/<name>.npz``. The GPU box has no
/root/reference, so the fixtures (small) are committed and this script is the provenance.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_shims"), os.path.join(REF, "plugins", "track"), REF, ROOT]

from tracklab_b200.synth import make_video  # noqa: E402

# name -> (tracker, generator kwargs, hyper-parameters)   (hyper-parameters: SURVEY.md Appendix A)
CASES = {
    "bytetrack_c2_s2000": ("bytetrack", dict(seed=2000, n_frames=160, n_ids=44),
                           dict(track_thresh=0.6, track_buffer=30, match_thresh=0.8, frame_rate=30)),
    "bytetrack_small_s5": ("bytetrack", dict(seed=5, n_frames=64, n_ids=20, conf_range=(0.05, 1.0)),
                           dict(track_thresh=0.6, track_buffer=30, match_thresh=0.8, frame_rate=30)),
    "bytetrack_buffer5_s9": ("bytetrack", dict(seed=9, n_frames=120, n_ids=30, conf_range=(0.3, 1.0)),
                             dict(track_thresh=0.5, track_buffer=5, match_thresh=0.8, frame_rate=30)),
    "ocsort_c1_iou_s1000": ("ocsort", dict(seed=1000, n_frames=64, n_ids=20, conf_range=(0.5, 1.0)),
                            dict(det_thresh=0, max_age=50, min_hits=1, iou_threshold=0.22136877277096445,
                                 delta_t=1, asso_func="iou", inertia=0.3941737016672115, use_byte=False)),
    "ocsort_giou_s1001": ("ocsort", dict(seed=1001, n_frames=200, n_ids=44),
                          dict(det_thresh=0, max_age=50, min_hits=1, iou_threshold=0.22136877277096445,
                               delta_t=1, asso_func="giou", inertia=0.3941737016672115, use_byte=False)),
    "ocsort_byte_dt3_s1002": ("ocsort", dict(seed=1002, n_frames=160, n_ids=30, conf_range=(0.05, 1.0)),
                              dict(det_thresh=0.5, max_age=8, min_hits=3, iou_threshold=0.3,
                                   delta_t=3, asso_func="iou", inertia=0.2, use_byte=True)),
}
MIN_CONF = 0.4


def run_reference(tracker, video, hyper):
    if tracker == "bytetrack":
        from byte_track import basetrack, byte_tracker
        basetrack.BaseTrack._count = 0  # process-global in the reference (basetrack.py:13)
        model = byte_tracker.BYTETracker(**hyper)
    elif tracker == "ocsort":
        from oc_sort import ocsort
        model = ocsort.OCSort(**hyper)
    else:
        raise KeyError(tracker)
    rows, frames = [], []
    for f in range(video.n_frames):
        d = video.frame(f)
        if len(d) == 0:
            continue
        d = d[d[:, 4] > MIN_CONF]
        with torch.no_grad():
            res = np.asarray(model.update(torch.from_numpy(d.copy()), None), dtype=np.float64)
        if res.size:
            rows.append(res.reshape(-1, 8))
            frames.append(np.full(res.reshape(-1, 8).shape[0], f, dtype=np.int32))
    if not rows:
        return np.zeros((0, 8)), np.zeros((0,), dtype=np.int32)
    return np.concatenate(rows), np.concatenate(frames)


def main(names):
    for name in names:
        tracker, gen, hyper = CASES[name]
        video = make_video(**gen)
        rows, frames = run_reference(tracker, video, hyper)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), rows=rows, frames=frames,
            dets_sha=np.frombuffer(__import__("hashlib").sha256(video.dets.tobytes()).digest(), dtype=np.uint8),
            tracker=tracker, gen=repr(gen), hyper=repr(hyper), min_conf=MIN_CONF)
        print(f"{name}: {video.n_dets} dets -> {rows.shape[0]} rows, {len(np.unique(rows[:, 4]))} ids")


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
