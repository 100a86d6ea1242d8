"""Goldens for Deep OC-SORT: the UNMODIFIED plugin /root/reference/plugins/track/deep_oc_sort/ocsort.py run in the
build container on the synthetic videos, with its two learned / image-based inputs supplied from outside exactly where the plugin
obtains them: `_get_features` (the in-tracker ReID forward) returns the generator's per-detection embeddings, and
`cmc.compute_affine` returns a seeded small similarity transform per frame (camera jitter).

    python tests/golden/make_deepocsort_golden.py [name ...]
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_env  # noqa: E402

YAML = dict(det_thresh=0, max_age=50, min_hits=1, iou_threshold=0.22136877277096445, delta_t=1, asso_func="giou",
            inertia=0.3941737016672115, w_association_emb=0.75, alpha_fixed_emb=0.95, aw_param=0.5, embedding_off=False,
            cmc_off=False, aw_off=False, new_kf_off=False)           # configs/modules/track/deep_oc_sort.yaml
CASES = {
    "deepocsort_yaml_s7000": (dict(seed=7000, n_frames=160, n_ids=30, emb_dim=64), YAML, 0.004),
    "deepocsort_dt3_s7001": (dict(seed=7001, n_frames=120, n_ids=44, emb_dim=128, conf_range=(0.3, 1.0)),
                             dict(YAML, max_age=12, min_hits=3, iou_threshold=0.3, delta_t=3, asso_func="iou", inertia=0.2,
                                  w_association_emb=0.5, alpha_fixed_emb=0.9, aw_param=0.4), 0.01),
    "deepocsort_awoff_nocmc_s7002": (dict(seed=7002, n_frames=100, n_ids=24, emb_dim=32, conf_range=(0.05, 1.0)),
                                     dict(YAML, det_thresh=0.3, max_age=8, delta_t=2, asso_func="diou", aw_off=True, cmc_off=True), 0.0),
    # tracks die after one missed frame and the IoU gate is tight: detections outnumber trackers in most frames, so the
    # `[y[-1], -1]` pairs of linear_assignment (oracle q1) fire in both association rounds
    "deepocsort_births_s7004": (dict(seed=7004, n_frames=120, n_ids=40, emb_dim=32, conf_range=(0.3, 1.0)),
                                dict(YAML, max_age=1, iou_threshold=0.08, delta_t=2, asso_func="iou", aw_param=0.6), 0.012),
    "deepocsort_raw_emb_s7003": (dict(seed=7003, n_frames=80, n_ids=20, emb_dim=48), dict(YAML, asso_func="ciou"), 0.006),
}
MIN_CONF = 0.4


def make_affines(seed, n_frames, mag):
    """Small similarity transforms (rotation, scale, translation in pixels) as float64 [F,2,3]; frame 0 = identity like the estimator."""
    rng = np.random.default_rng(seed + 99)
    A = np.zeros((n_frames, 2, 3))
    for f in range(n_frames):
        th, sc = rng.normal(0, mag * 0.5), 1.0 + rng.normal(0, mag * 0.5)
        tx, ty = rng.normal(0, mag * 800, 2)
        A[f] = [[sc * np.cos(th), -sc * np.sin(th), tx], [sc * np.sin(th), sc * np.cos(th), ty]]
    A[0] = np.eye(2, 3)
    return A


def embeddings_of(video, name):
    e = video.embeddings.astype(np.float32)
    if "raw_emb" in name:      # un-normalised features like a backbone without a normalising head (the plugin does not normalise them)
        rng = np.random.default_rng(5)
        e = e * rng.uniform(2.0, 9.0, (len(e), 1)).astype(np.float32)
    return np.ascontiguousarray(e)


def run_reference(video, hyper, embs, affines):
    ref_env.install()
    from deep_oc_sort import ocsort
    model = object.__new__(ocsort.OCSort)           # __init__ builds the ReID network and the CMC estimator; set the rest by hand
    for k, v in hyper.items():
        setattr(model, k, v)
    model.asso_func = ocsort.ASSO_FUNCS[hyper["asso_func"]]
    model.trackers, model.frame_count = [], 0
    ocsort.KalmanBoxTracker.count = 0

    class _Cmc:
        def compute_affine(self, img, bbox, tag):
            return affines[_Cmc.f].copy()
    model.cmc = _Cmc()
    rows, frames = [], []
    img = np.zeros((video.height, video.width, 3), dtype=np.uint8)
    for f in range(video.n_frames):
        sl = slice(video.offsets[f], video.offsets[f + 1])
        d = video.dets[sl]
        if len(d) == 0:
            continue
        keep = d[:, 4] > MIN_CONF
        d = d[keep]
        e = embs[sl][keep]
        _Cmc.f = f
        model._get_features = lambda xyxy, im, _e=e, _d=d, _m=model: torch.from_numpy(_e[_d[:, 4] > _m.det_thresh].copy())
        with torch.no_grad():
            res = np.asarray(model.update(torch.from_numpy(d.copy()), img), dtype=np.float64)
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
        if "births" in name:
            from tests.util import blackout_video
            v = blackout_video(v, seed=gen["seed"])
        embs = embeddings_of(v, name)
        aff = make_affines(gen["seed"], v.n_frames, mag)
        rows, frames = run_reference(v, hyper, embs, aff)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rows=rows, frames=frames, affines=aff, tracker="deepocsort",
                            gen=repr(gen), hyper=repr(hyper), min_conf=MIN_CONF, raw_emb=("raw_emb" in name))
        print(name, rows.shape, "ids", len(np.unique(rows[:, 4])) if len(rows) else 0)


if __name__ == "__main__":
    main(sys.argv[1:])
