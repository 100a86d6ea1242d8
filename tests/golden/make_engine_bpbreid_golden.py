"""Engine-level golden for the BPBReID-StrongSORT module: the UNMODIFIED reference engine + reference wrapper
(/root/reference/tracklab/wrappers/track/bpbreid_strong_sort_api.py) + plugin on a synthetic set whose detections carry
part-based embeddings / visibility scores (torchreid's distance function served by oracle/ref_shims_torchreid: parity
unpinned for that term). Build container only.

    python tests/golden/make_engine_bpbreid_golden.py
"""
import os
import sys
import tempfile
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_shims_torchreid"), ROOT]

from oracle import ref_env  # noqa: E402

ref_env.install()

import cv2  # noqa: E402
from tracklab.datastruct import TrackerState, TrackingSet  # noqa: E402
from tracklab.engine import OfflineTrackingEngine  # noqa: E402
from tracklab.pipeline import Pipeline  # noqa: E402
from tracklab.wrappers.track.bpbreid_strong_sort_api import BPBReIDStrongSORT  # noqa: E402

from tracklab_b200.synth import make_video  # noqa: E402

GEN = dict(seed=6500, n_frames=60, n_ids=18, emb_dim=32, n_parts=5, conf_range=(0.3, 1.0))
CFG = dict(ecc=False, ema_alpha=0.9, mc_lambda=0.995, max_dist=0.5, motion_criterium="iou", max_iou_distance=0.8, max_oks_distance=0.7,
           max_age=300, n_init=0, nn_budget=100, min_bbox_confidence=0.0, only_position_for_kf_gating=False,
           max_kalman_prediction_without_update=7, matching_strategy="strong_sort_matching", gating_thres_factor=1, w_kfgd=1, w_reid=1, w_st=1)


def main():
    tmp = tempfile.mkdtemp()
    img_path = os.path.join(tmp, "blank.jpg")
    cv2.imwrite(img_path, np.zeros((1080, 1920, 3), dtype=np.uint8))
    video = make_video(**GEN)
    imgs, dets = [], []
    k = 0
    for f in range(video.n_frames):
        imgs.append(dict(id=f, video_id=0, frame=f, file_path=img_path, nframes=video.n_frames, is_labeled=True))
        for row in video.frame(f):
            l, t, r, b, conf, cls, _ = row
            dets.append(dict(image_id=f, video_id=0, category_id=int(cls), bbox_ltwh=np.array([l, t, r - l, b - t]), bbox_conf=conf,
                             embeddings=video.embeddings[k], visibility_scores=video.visibility[k].astype(bool)))
            k += 1
    tset = TrackingSet(pd.DataFrame([dict(id=0, name="synthetic_0")]).set_index("id", drop=False),
                       pd.DataFrame(imgs).set_index("id", drop=False), pd.DataFrame(dets))
    pipeline = Pipeline([BPBReIDStrongSORT(types.SimpleNamespace(**CFG), "cpu", batch_size=1)])
    state = TrackerState(tset, load_from_groundtruth=True, pipeline=pipeline)
    engine = OfflineTrackingEngine(modules=pipeline, tracker_state=state, num_workers=0, callbacks={})
    engine.track_dataset()
    df = state.detections_pred.sort_index()
    has = df["track_id"].notna().to_numpy()
    kf = np.stack([np.asarray(x, dtype=np.float64) if h else np.full(4, np.nan) for x, h in zip(df["track_bbox_kf_ltwh"], has)])
    code = np.array([(1 if m[0] == "R" else 2) if isinstance(m, tuple) else 0 for m in df["matched_with"]])
    np.savez_compressed(os.path.join(HERE, "engine_bpbreid.npz"), det_index=df.index.to_numpy().astype(np.int64),
                        track_id=np.where(has, df["track_id"].to_numpy(dtype=float, na_value=np.nan), np.nan), kf_ltwh=kf, matched_code=code,
                        hits=np.where(has, df["hits"].to_numpy(dtype=float, na_value=np.nan), np.nan),
                        age=np.where(has, df["age"].to_numpy(dtype=float, na_value=np.nan), np.nan), gen=repr(GEN), cfg=repr(CFG))
    print(len(df), "detections,", int(has.sum()), "with a track id,", len(np.unique(df["track_id"].dropna())), "ids; stages", np.bincount(code))


if __name__ == "__main__":
    main()
