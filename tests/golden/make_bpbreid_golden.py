"""Golden vectors for the BPBReID-StrongSORT association (build container only).

    python tests/golden/make_bpbreid_golden.py

Runs the UNMODIFIED /root/reference/plugins/track/bpbreid_strong_sort plugin on synthetic part-based videos. Its two
torchreid imports (un-vendored git dependency) are served by ``oracle/ref_shims_torchreid`` — the appearance distance is
therefore PARITY UNPINNED (restated, see that README); everything else (Kalman model, two-stage matching, EMA, life cycle,
output rule) is the reference's own code. Output rows: see oracle/bpbreid_np.py::BpbreidStrongSortOracle.update.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_shims_torchreid"), os.path.join(ROOT, "oracle", "ref_shims"),
                os.path.join(REF, "plugins", "track"), ROOT]

from tracklab_b200.synth import make_video  # noqa: E402

YAML = dict(ema_alpha=0.9, mc_lambda=0.995, max_dist=0.5, motion_criterium="iou", max_iou_distance=0.8, max_age=300, n_init=0,
            nn_budget=100, min_bbox_confidence=0.0, only_position_for_kf_gating=False, max_kalman_prediction_without_update=7,
            matching_strategy="strong_sort_matching", gating_thres_factor=1, w_kfgd=1, w_reid=1, w_st=1)
CASES = {
    # hyper-parameters of /root/reference/tracklab/configs/modules/track/bpbreid_strong_sort.yaml
    "bpbreid_yaml_s6000": (dict(seed=6000, n_frames=160, n_ids=30, emb_dim=32, n_parts=6), YAML),
    "bpbreid_tight_s6001": (dict(seed=6001, n_frames=120, n_ids=44, emb_dim=64, n_parts=9, conf_range=(0.2, 1.0), p_visible=0.6),
                            dict(YAML, max_dist=0.3, max_iou_distance=0.7, max_age=10, n_init=2, ema_alpha=0.8, mc_lambda=0.98,
                                 min_bbox_confidence=0.3, max_kalman_prediction_without_update=3)),
    # single-stage weighted-sum matching (tracker.py:335-363): oracle only, the device kernel implements strong_sort_matching
    "bpbreid_botsort_s6002": (dict(seed=6002, n_frames=100, n_ids=28, emb_dim=32, n_parts=6, conf_range=(0.3, 1.0)),
                              dict(YAML, matching_strategy="bot_sort_matching", gating_thres_factor=1.5, w_kfgd=1, w_reid=2, w_st=1, max_age=20,
                                   n_init=1)),
}


def run_reference(video, hyper):
    from bpbreid_strong_sort.strong_sort import StrongSORT
    model = StrongSORT(**hyper)
    rows, frames = [], []
    for f in range(video.n_frames):
        sl = slice(video.offsets[f], video.offsets[f + 1])
        d = video.dets[sl]
        if len(d) == 0:   # bpbreid_strong_sort_api.py:75-84,105-106
            continue
        ltwh = np.column_stack([d[:, 0], d[:, 1], d[:, 2] - d[:, 0], d[:, 3] - d[:, 1]])
        with torch.no_grad():
            res = model.update(torch.from_numpy(d[:, 6].copy()), torch.from_numpy(ltwh), torch.from_numpy(video.embeddings[sl].copy()),
                               torch.from_numpy(video.visibility[sl].copy()), torch.from_numpy(d[:, 4].copy()),
                               torch.zeros(len(d), dtype=torch.float64), torch.full((len(d),), float(f)))
        for det_id, r in res.iterrows():
            assert r.time_since_update == 0 and r.state == "c"
            mw = r.matched_with
            code, dist = (0, np.nan) if mw is None else ((1 if mw[0] == "R" else 2), float(mw[1]))
            pred = r.track_bbox_pred_kf_ltwh if r.track_bbox_pred_kf_ltwh is not None else np.full(4, np.nan)
            rows.append([r.track_id, *r.track_bbox_kf_ltwh, *pred, code, dist, r.hits, r.age, float(det_id)])
            frames.append(f)
    return np.asarray(rows, dtype=np.float64).reshape(-1, 14), np.asarray(frames, dtype=np.int32)


if __name__ == "__main__":
    for name, (gen, hyper) in CASES.items():
        video = make_video(**gen)
        rows, frames = run_reference(video, hyper)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rows=rows, frames=frames, tracker="bpbreid", gen=repr(gen),
                            hyper=repr(hyper))
        print(f"{name}: {video.n_dets} dets -> {rows.shape[0]} rows, {len(np.unique(rows[:, 0]))} ids, "
              f"R={int((rows[:, 9] == 1).sum())} S={int((rows[:, 9] == 2).sum())} births={int((rows[:, 9] == 0).sum())}")
