"""Goldens for BoT-SORT: the UNMODIFIED plugin /root/reference/plugins/track/bot_sort/bot_sort.py run in the build
container on the synthetic videos, with its two learned / image-based inputs supplied where the plugin obtains them: `_get_features`
(the in-tracker ReID forward) returns the generator's per-detection embeddings and `gmc.apply` a seeded small similarity transform.

    python tests/golden/make_botsort_golden.py [name ...]
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_env  # noqa: E402
from tests.golden.make_deepocsort_golden import make_affines  # noqa: E402

YAML = dict(appearance_thresh=0.4818211117541298, cmc_method="sparseOptFlow", frame_rate=30, lambda_=0.9896143462366406,
            match_thresh=0.22734550911325851, new_track_thresh=0.21144301345190655, proximity_thresh=0.5945380911899254,
            track_buffer=60, track_high_thresh=0.33824964456239337)           # configs/modules/track/bot_sort.yaml
DEFAULTS = dict(track_high_thresh=0.45, new_track_thresh=0.6, track_buffer=30, match_thresh=0.8, proximity_thresh=0.5,
                appearance_thresh=0.25, cmc_method="sparseOptFlow", frame_rate=30, lambda_=0.985)
CASES = {
    "botsort_yaml_s8000": (dict(seed=8000, n_frames=160, n_ids=30, emb_dim=64), YAML, 0.004),
    "botsort_defaults_s8001": (dict(seed=8001, n_frames=120, n_ids=44, emb_dim=128, conf_range=(0.05, 1.0)), DEFAULTS, 0.008),
    "botsort_buffer5_nogmc_s8002": (dict(seed=8002, n_frames=100, n_ids=24, emb_dim=32, conf_range=(0.2, 1.0)),
                                    dict(DEFAULTS, track_buffer=5, match_thresh=0.6, lambda_=0.9, new_track_thresh=0.5), 0.0),
}
MIN_CONF = 0.4


def run_reference(video, hyper, embs, warps):
    ref_env.install()
    import bot_sort.bot_sort as bs
    from bot_sort.basetrack import BaseTrack
    from bot_sort.kalman_filter import KalmanFilter
    m = object.__new__(bs.BoTSORT)              # __init__ builds the ReID network and the GMC estimator; set the rest by hand
    m.tracked_stracks, m.lost_stracks, m.removed_stracks = [], [], []
    BaseTrack.clear_count()
    m.frame_id = 0
    m.lambda_ = hyper["lambda_"]
    m.track_high_thresh, m.new_track_thresh = hyper["track_high_thresh"], hyper["new_track_thresh"]
    m.buffer_size = int(hyper["frame_rate"] / 30.0 * hyper["track_buffer"])
    m.max_time_lost = m.buffer_size
    m.kalman_filter = KalmanFilter()
    m.proximity_thresh, m.appearance_thresh, m.match_thresh = hyper["proximity_thresh"], hyper["appearance_thresh"], hyper["match_thresh"]

    class _Gmc:
        def apply(self, img, dets=None):
            return warps[_Gmc.f].copy()
    m.gmc = _Gmc()
    rows, frames = [], []
    img = np.zeros((video.height, video.width, 3), dtype=np.uint8)
    for f in range(video.n_frames):
        sl = slice(video.offsets[f], video.offsets[f + 1])
        d = video.dets[sl]
        if len(d) == 0:
            continue
        keep = d[:, 4] > MIN_CONF
        d, e = d[keep], embs[sl][keep]
        _Gmc.f = f
        m._get_features = lambda xywh, im, _e=e, _d=d, _m=m: torch.from_numpy(_e[_d[:, 4] > _m.track_high_thresh].copy())
        with torch.no_grad():
            res = np.asarray(m.update(torch.from_numpy(d.copy()), img), dtype=np.float64)
        if res.size:
            rows.append(res.reshape(-1, 8))
            frames.append(np.full(len(rows[-1]), f, dtype=np.int32))
    if not rows:
        return np.zeros((0, 8)), np.zeros((0,), dtype=np.int32)
    return np.concatenate(rows), np.concatenate(frames)


def main(names):
    from tracklab_b200.synth import make_video
    for name in names or CASES:
        gen, hyper, mag = CASES[name]
        v = make_video(**gen)
        embs = np.ascontiguousarray(v.embeddings.astype(np.float32)) * np.float32(3.0)     # un-normalised, the plugin normalises
        warps = make_affines(gen["seed"], v.n_frames, mag)
        rows, frames = run_reference(v, hyper, embs, warps)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rows=rows, frames=frames, affines=warps, tracker="botsort",
                            gen=repr(gen), hyper=repr(hyper), min_conf=MIN_CONF)
        print(name, rows.shape, "ids", len(np.unique(rows[:, 4])) if len(rows) else 0)


if __name__ == "__main__":
    main(sys.argv[1:])
