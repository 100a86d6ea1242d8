"""Engine-level goldens: the UNMODIFIED reference engine + reference tracker wrappers on a synthetic set.

    python tests/golden/make_engine_golden.py

Runs /root/reference/tracklab/engine/offline.py:OfflineTrackingEngine with the reference wrappers
(/root/reference/tracklab/wrappers/track/{byte_track,oc_sort}_api.py) over a synthetic ``TrackingSet`` whose
ground-truth detections are injected with ``TrackerState(load_from_groundtruth=True)``,
``num_workers=0``. Stores the resulting per-detection columns (track_id, track_bbox_ltwh, track_bbox_conf) as
``tests/golden/engine_<name>.npz`` — the contract the drop-in modules of tracklab_b200 must reproduce through
the engine (tests/test_engine_*). Build container only (needs /root/reference).
"""
import os
import sys
import tempfile
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_env  # noqa: E402

ref_env.install()

import cv2  # noqa: E402
from tracklab.datastruct import TrackerState, TrackingSet  # noqa: E402
from tracklab.engine import OfflineTrackingEngine  # noqa: E402
from tracklab.pipeline import Pipeline  # noqa: E402

from tracklab_b200.synth import make_video  # noqa: E402

CASES = {
    "bytetrack_2videos": ("bytetrack", [dict(seed=3000, n_frames=48, n_ids=20), dict(seed=3001, n_frames=40, n_ids=16)],
                          dict(track_thresh=0.6, track_buffer=30, match_thresh=0.8, frame_rate=30)),
    "ocsort_c1": ("ocsort", [dict(seed=1000, n_frames=64, n_ids=20, conf_range=(0.5, 1.0))],
                  dict(det_thresh=0, max_age=50, min_hits=1, iou_threshold=0.22136877277096445, delta_t=1,
                       asso_func="iou", inertia=0.3941737016672115, use_byte=False)),
}


def tracking_set(videos, img_path):
    vids, imgs, dets = [], [], []
    image_id = 0
    for v, video in enumerate(videos):
        vids.append(dict(id=v, name=f"synthetic_{v}"))
        for f in range(video.n_frames):
            imgs.append(dict(id=image_id, video_id=v, frame=f, file_path=img_path, nframes=video.n_frames,
                             is_labeled=True))
            for row in video.frame(f):
                l, t, r, b, conf, cls, _ = row
                dets.append(dict(image_id=image_id, video_id=v, category_id=int(cls),
                                 bbox_ltwh=np.array([l, t, r - l, b - t]), bbox_conf=conf))
            image_id += 1
    video_md = pd.DataFrame(vids).set_index("id", drop=False)
    image_md = pd.DataFrame(imgs).set_index("id", drop=False)
    det_gt = pd.DataFrame(dets)
    return TrackingSet(video_md, image_md, det_gt)


def make_module(kind, hyper):
    cfg = types.SimpleNamespace(min_confidence=0.4, hyperparams=hyper)
    if kind == "bytetrack":
        from byte_track import basetrack
        from tracklab.wrappers.track.byte_track_api import ByteTrack
        basetrack.BaseTrack._count = 0
        return ByteTrack(cfg, "cpu")
    from tracklab.wrappers.track.oc_sort_api import OCSORT
    return OCSORT(cfg, "cpu")


def main():
    tmp = tempfile.mkdtemp()
    img_path = os.path.join(tmp, "blank.jpg")
    cv2.imwrite(img_path, np.zeros((1080, 1920, 3), dtype=np.uint8))
    for name, (kind, gens, hyper) in CASES.items():
        videos = [make_video(**g) for g in gens]
        tset = tracking_set(videos, img_path)
        pipeline = Pipeline([make_module(kind, hyper)])
        state = TrackerState(tset, load_from_groundtruth=True, pipeline=pipeline)
        engine = OfflineTrackingEngine(modules=pipeline, tracker_state=state, num_workers=0, callbacks={})
        engine.track_dataset()
        df = state.detections_pred.sort_index()
        has = df["track_id"].notna().to_numpy()
        ltwh = np.stack([np.asarray(x, dtype=np.float64) if h else np.full(4, np.nan)
                         for x, h in zip(df["track_bbox_ltwh"], has)])
        np.savez_compressed(os.path.join(HERE, f"engine_{name}.npz"),
                            det_index=df.index.to_numpy().astype(np.int64), image_id=df["image_id"].to_numpy().astype(np.int64),
                            video_id=df["video_id"].to_numpy().astype(np.int64),
                            track_id=np.where(has, df["track_id"].to_numpy(dtype=float, na_value=np.nan), np.nan),
                            track_bbox_ltwh=ltwh,
                            track_bbox_conf=np.where(has, df["track_bbox_conf"].to_numpy(dtype=float, na_value=np.nan), np.nan),
                            kind=kind, gens=repr(gens), hyper=repr(hyper))
        print(name, len(df), "detections,", int(has.sum()), "with a track id,", len(np.unique(df['track_id'].dropna())), "ids")


if __name__ == "__main__":
    main()
