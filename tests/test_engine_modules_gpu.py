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
    from tests.engine_mirror import OfflineEngineMirror
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


def test_bpbreid_module_through_engine_matches_reference_plugin():
    """BPBReIDStrongSORT drop-in (visibility as booleans, like the BPBReID ReID module emits) vs the golden of the
    unmodified reference plugin: same detection -> track partition, stages, hits and ages; float64 boxes within 1e-6."""
    from tests.util import load_bpbreid_golden
    from tracklab_b200 import modules
    from tests.engine_mirror import OfflineEngineMirror
    g = load_bpbreid_golden("bpbreid_yaml_s6000")
    video = make_video(**g["gen"])
    vmd, imd, det = _tracking_frames([video])
    det["embeddings"] = list(video.embeddings)
    det["visibility_scores"] = list(video.visibility.astype(bool))
    cfg = types.SimpleNamespace(ecc=False, **g["hyper"])
    mod = modules.BPBReIDStrongSORT(cfg, "cuda:0")
    assert mod.level == "image" and mod.name == "BPBReIDStrongSORT"
    out = OfflineEngineMirror([mod], vmd, imd, det).track_dataset().sort_index()
    ref = g["rows"]
    ref = ref[np.argsort(ref[:, 13])]
    has = out["track_id"].notna().to_numpy()
    assert np.array_equal(out.index.to_numpy()[has], ref[:, 13].astype(int))
    tid = out["track_id"].to_numpy(dtype=float, na_value=np.nan)[has]
    assert np.array_equal(tid, ref[:, 0]), "track ids differ from the reference plugin's"   # exact: scipy's tie-breaking is reproduced on device
    sel = out[has]
    assert np.array_equal(sel["hits"].to_numpy(dtype=float), ref[:, 11]) and np.array_equal(sel["age"].to_numpy(dtype=float), ref[:, 12])
    assert (sel["state"] == "c").all() and (sel["time_since_update"] == 0).all()
    code = np.array([(1 if m[0] == "R" else 2) if isinstance(m, tuple) else 0 for m in sel["matched_with"]])   # None -> NaN after the merge
    assert np.array_equal(code, ref[:, 9].astype(int))
    box = np.stack([np.asarray(b, dtype=np.float64) for b in sel["track_bbox_kf_ltwh"]])
    assert np.abs(box - ref[:, 1:5]).max() < 1e-6
    born = np.isnan(ref[:, 5])
    assert all((not isinstance(p, np.ndarray)) == b for p, b in zip(sel["track_bbox_pred_kf_ltwh"], born))


def test_bpbreid_module_matches_the_real_engine_golden():
    """tests/golden/engine_bpbreid.npz was produced by the REAL OfflineTrackingEngine + the reference BPBReIDStrongSORT wrapper
    (make_engine_bpbreid_golden.py); the drop-in through the mirrored engine protocol must give the same per-detection columns."""
    from tracklab_b200 import modules
    from tests.engine_mirror import OfflineEngineMirror
    g = np.load(os.path.join(HERE, "golden", "engine_bpbreid.npz"))
    video = make_video(**ast.literal_eval(str(g["gen"])))
    cfgd = ast.literal_eval(str(g["cfg"]))
    vmd, imd, det = _tracking_frames([video])
    det["embeddings"] = list(video.embeddings)
    det["visibility_scores"] = list(video.visibility.astype(bool))
    out = OfflineEngineMirror([modules.BPBReIDStrongSORT(types.SimpleNamespace(**cfgd), "cuda:0")], vmd, imd, det).track_dataset().sort_index()
    assert np.array_equal(out.index.to_numpy(), g["det_index"])
    has, ref_has = out["track_id"].notna().to_numpy(), ~np.isnan(g["track_id"])
    assert np.array_equal(has, ref_has)
    fwd, bwd = {}, {}
    for x, y in zip(out["track_id"].to_numpy(dtype=float, na_value=np.nan)[has], g["track_id"][ref_has]):
        assert fwd.setdefault(x, y) == y and bwd.setdefault(y, x) == x
    sel = out[has]
    assert np.array_equal(sel["hits"].to_numpy(dtype=float), g["hits"][ref_has]) and np.array_equal(sel["age"].to_numpy(dtype=float), g["age"][ref_has])
    code = np.array([(1 if m[0] == "R" else 2) if isinstance(m, tuple) else 0 for m in sel["matched_with"]])
    assert np.array_equal(code, g["matched_code"][ref_has])
    box = np.stack([np.asarray(b, dtype=np.float64) for b in sel["track_bbox_kf_ltwh"]])
    assert np.abs(box - g["kf_ltwh"][ref_has]).max() < 1e-6
