"""GPU: the drop-in modules driven through the engine protocol reproduce the REAL reference engine's output
(tests/golden/engine_*.npz were produced by /root/reference's OfflineTrackingEngine + reference wrappers)."""
import ast
import os
import types

import numpy as np
import pandas as pd
import pytest

from tracklab_b200.synth import make_video

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _tracking_frames(videos):
    vids, imgs, dets = [], [], []
    image_id = 0
    for v, video in enumerate(videos):
        vids.append(dict(id=v, name=f"synthetic_{v}"))
        for f in range(video.n_frames):
            imgs.append(dict(id=image_id, video_id=v, frame=f, file_path="unused.jpg"))
            for row in video.frame(f):
                l, t, r, b, conf, cls, _ = row
                dets.append(dict(image_id=image_id, video_id=v, category_id=int(cls),
                                 bbox_ltwh=np.array([l, t, r - l, b - t]), bbox_conf=conf))
            image_id += 1
    det = pd.DataFrame(dets).reset_index(drop=True)
    det["id"] = det.index   # TrackerState.load_groundtruth (tracker_state.py:146-147)
    return (pd.DataFrame(vids).set_index("id", drop=False), pd.DataFrame(imgs).set_index("id", drop=False), det)


@pytest.mark.parametrize("name,frames_per_batch", [("bytetrack_2videos", None), ("bytetrack_2videos", 7), ("ocsort_c1", None)])
def test_modules_through_engine_match_reference_engine(name, frames_per_batch):
    from tracklab_b200 import modules
    from tracklab_b200.engine_mirror import OfflineEngineMirror
    g = np.load(os.path.join(HERE, "golden", f"engine_{name}.npz"))
    gens, hyper, kind = ast.literal_eval(str(g["gens"])), ast.literal_eval(str(g["hyper"])), str(g["kind"])
    videos = [make_video(**k) for k in gens]
    vmd, imd, det = _tracking_frames(videos)
    cfg = types.SimpleNamespace(min_confidence=0.4, hyperparams=hyper, frames_per_batch=frames_per_batch,
                                cap_tracks=128, cap_dets=128)
    mod = (modules.ByteTrack if kind == "bytetrack" else modules.OCSORT)(cfg, "cuda:0")
    assert mod.level == "image" and mod.name == ("ByteTrack" if kind == "bytetrack" else "OCSORT")
    out = OfflineEngineMirror([mod], vmd, imd, det).track_dataset().sort_index()
    assert np.array_equal(out.index.to_numpy(), g["det_index"])
    has = out["track_id"].notna().to_numpy()
    ref_has = ~np.isnan(g["track_id"])
    assert np.array_equal(has, ref_has)
    assert np.array_equal(out["track_id"].to_numpy(dtype=float, na_value=np.nan)[has], g["track_id"][ref_has])
    got = np.stack([np.asarray(x, dtype=np.float64) for x in out["track_bbox_ltwh"][has]])
    assert np.abs(got - g["track_bbox_ltwh"][ref_has]).max() < 1e-6
    assert np.array_equal(out["track_bbox_conf"].to_numpy(dtype=float, na_value=np.nan)[has], g["track_bbox_conf"][ref_has])
