"""Drive the REAL reference engine (OfflineTrackingEngine + TrackerState + Pipeline, imported from /root/reference or from the
staged copy oracle/_ref/) over tracklab_b200's drop-in modules, in a process of its own (oracle/ref_env installs an import hook
for the reference's optional third-party imports, which must not leak into the pytest process). Prints one JSON line.

    python tests/real_engine_driver.py bytetrack [--cpu-standin]
    python tests/real_engine_driver.py chain
"""
import json
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_env  # noqa: E402

ref_env.install()           # BEFORE tracklab_b200.modules: the modules then subclass the real tracklab.pipeline classes

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402
from tracklab.datastruct import TrackerState, TrackingSet  # noqa: E402
from tracklab.engine import OfflineTrackingEngine  # noqa: E402
from tracklab.pipeline import ImageLevelModule, Pipeline  # noqa: E402

from tracklab_b200.synth import make_frames, make_video  # noqa: E402


def tracking_set(videos, paths_of, with_dets=True):
    vids, imgs, dets = [], [], []
    image_id = 0
    for v, video in enumerate(videos):
        vids.append(dict(id=v, name=f"synthetic_{v}"))
        for f in range(video.n_frames):
            imgs.append(dict(id=image_id, video_id=v, frame=f, file_path=paths_of(v, f), nframes=video.n_frames, is_labeled=True))
            if with_dets:
                for row in video.frame(f):
                    l, t, r, b, conf, cls, _ = row
                    dets.append(dict(image_id=image_id, video_id=v, category_id=int(cls), bbox_ltwh=np.array([l, t, r - l, b - t]), bbox_conf=conf))
            image_id += 1
    video_md = pd.DataFrame(vids).set_index("id", drop=False)
    image_md = pd.DataFrame(imgs).set_index("id", drop=False)
    det_gt = pd.DataFrame(dets) if dets else pd.DataFrame(columns=["image_id", "video_id"])
    return TrackingSet(video_md, image_md, det_gt)


def case_bytetrack(cpu_standin):
    """Golden engine_bytetrack_2videos.npz (REAL engine + reference wrapper) vs the REAL engine + the drop-in module."""
    import ast

    from tracklab_b200 import modules
    g = np.load(os.path.join(HERE, "golden", "engine_bytetrack_2videos.npz"))
    gens, hyper = ast.literal_eval(str(g["gens"])), ast.literal_eval(str(g["hyper"]))
    videos = [make_video(**k) for k in gens]
    tmp = tempfile.mkdtemp()
    img = os.path.join(tmp, "blank.jpg")
    cv2.imwrite(img, np.zeros((1080, 1920, 3), dtype=np.uint8))
    cfg = types.SimpleNamespace(min_confidence=0.4, hyperparams=hyper)
    if cpu_standin:
        # protocol check without a GPU: the module's device tracker is replaced by the NumPy oracle behind the same interface
        import torch

        from oracle.bytetrack_np import ByteTrackOracle

        class _Standin:
            def __init__(self, **kw):
                self.kw = {k: v for k, v in kw.items() if k in ("track_thresh", "match_thresh", "track_buffer", "frame_rate", "min_confidence")}
                self.next_id = 1

            def reset(self, keep_id_counter=False):
                self.keep = keep_id_counter

            def run(self, dets, offs):
                o = ByteTrackOracle(**self.kw, first_id=self.next_id if getattr(self, "keep", False) else 1)
                rows, fr = o.run_video(dets.numpy(), offs[0].numpy())
                self.next_id = o._next + 1
                fc = np.bincount(fr, minlength=offs.shape[1] - 1).astype(np.int32)
                return torch.from_numpy(rows), torch.from_numpy(fc)[None], torch.tensor([len(rows)])

            def check_status(self):
                pass
        modules.torch.cuda.is_available = lambda: True
        modules.ByteTrack._device_cls = _Standin
        _to = torch.Tensor.to
        torch.Tensor.to = lambda self, *a, **k: self if (a and isinstance(a[0], torch.device)) else _to(self, *a, **k)
    mod = modules.ByteTrack(cfg, "cuda:0" if not cpu_standin else "cpu")
    assert isinstance(mod, ImageLevelModule) and mod.level == "image" and mod.name == "ByteTrack"
    tset = tracking_set(videos, lambda v, f: img)
    pipeline = Pipeline([mod])
    state = TrackerState(tset, load_from_groundtruth=True, pipeline=pipeline)
    engine = OfflineTrackingEngine(modules=pipeline, tracker_state=state, num_workers=0, callbacks={})
    engine.track_dataset()
    df = state.detections_pred.sort_index()
    has = df["track_id"].notna().to_numpy()
    tid = np.where(has, df["track_id"].to_numpy(dtype=float, na_value=np.nan), np.nan)
    ltwh = np.stack([np.asarray(x, dtype=np.float64) if h else np.full(4, np.nan) for x, h in zip(df["track_bbox_ltwh"], has)])
    ok_index = bool(np.array_equal(df.index.to_numpy(), g["det_index"]))
    ok_ids = bool(np.array_equal(np.isnan(tid), np.isnan(g["track_id"])) and np.array_equal(tid[has], g["track_id"][has]))
    err = float(np.nanmax(np.abs(ltwh - g["track_bbox_ltwh"]))) if has.any() else 0.0
    return {"case": "bytetrack", "real_base_class": True, "rows": int(len(df)), "with_track": int(has.sum()), "index_equal": ok_index,
            "ids_equal": ok_ids, "max_box_err": err}


def case_chain(n_frames=10):
    """[RTMLibDetector, KPReId, BPBReIDStrongSORT] through the REAL engine on synthetic PNG frames (no ground-truth detections)."""
    from tracklab_b200 import modules
    video = make_video(seed=3000, n_frames=n_frames, n_ids=30)
    tmp = tempfile.mkdtemp()
    frames = make_frames(video, 0, n_frames, device="cpu").numpy()
    for f in range(n_frames):
        cv2.imwrite(os.path.join(tmp, f"{f:06d}.png"), frames[f][..., ::-1])
    det = modules.RTMLibDetector("cuda:0", model=dict(onnx_model="https://x/yolox_m_8xb8-300e_humanart-c2c7a14a.zip", model_input_size=[640, 640]),
                                 batch_size=4)
    reid = modules.KPReId(dict(reid_arch="resnet50", reid_precision="fp32"), "cuda:0", batch_size=256)
    trk = modules.BPBReIDStrongSORT(types.SimpleNamespace(ecc=False), "cuda:0")
    pipeline = Pipeline([det, reid, trk])
    levels = [m.level for m in pipeline.models]
    tset = tracking_set([video], lambda v, f: os.path.join(tmp, f"{f:06d}.png"), with_dets=False)
    state = TrackerState(tset, pipeline=pipeline)          # runs Pipeline.validate on the real classes
    engine = OfflineTrackingEngine(modules=pipeline, tracker_state=state, num_workers=0, callbacks={})
    engine.track_dataset()
    df = state.detections_pred
    out = {"case": "chain", "levels": levels, "detections": int(len(df)), "columns": sorted(df.columns),
           "with_track": int(df["track_id"].notna().sum()), "ids": int(df["track_id"].dropna().nunique()),
           "embedding_shape": list(np.asarray(df["embeddings"].iloc[0]).shape), "bbox_dtype": str(np.asarray(df["bbox_ltwh"].iloc[0]).dtype)}
    # the same detections through the oracle chain (ReID-wrapper crop rule, fp32 CPU ResNet-50, NumPy BPBReID tracker)
    from PIL import Image

    import torch

    from oracle.bpbreid_np import BpbreidStrongSortOracle
    from oracle.pipeline_np import kpreid_crop_box
    from oracle.preprocess_np import REID_MEAN, REID_STD
    from tracklab_b200.nets.resnet_reid import build_resnet50_reid
    df = df.sort_index()
    ltwh = np.stack(df["bbox_ltwh"].to_numpy()).astype(np.float32)
    imgs = df["image_id"].to_numpy().astype(int)
    x = np.zeros((len(df), 3, 256, 128), np.float32)
    for i in range(len(df)):
        l, t, r, b = kpreid_crop_box(ltwh[i], video.width, video.height)
        small = np.asarray(Image.fromarray(frames[imgs[i]][t:b, l:r]).resize((128, 256), Image.BILINEAR)).astype(np.float32) / np.float32(255)
        x[i] = ((small - np.asarray(REID_MEAN, np.float32)) / np.asarray(REID_STD, np.float32)).transpose(2, 0, 1)
    net = build_resnet50_reid(1234).float().eval()
    with torch.no_grad():
        ref = torch.cat([net(torch.from_numpy(x[i:i + 64])) for i in range(0, len(x), 64)]).numpy()
    got = np.stack(df["embeddings"].to_numpy())[:, 0, :]
    out["max_embedding_rel_err"] = float(np.abs(got - ref).max() / np.abs(ref).max())
    rows = np.zeros((len(df), 7))
    rows[:, :4] = ltwh.astype(np.float64); rows[:, 2] += rows[:, 0]; rows[:, 3] += rows[:, 1]
    rows[:, 4] = df["bbox_conf"].to_numpy(dtype=float); rows[:, 6] = df.index.to_numpy()
    offs = np.concatenate([[0], np.cumsum(np.bincount(imgs, minlength=n_frames))])
    want, wf = BpbreidStrongSortOracle().run_video(rows, offs, ref[:, None, :], np.ones((len(df), 1), np.float32))
    want = want[np.argsort(want[:, 13])]
    has = df["track_id"].notna().to_numpy()
    out["oracle_ids_equal"] = bool(np.array_equal(df.index.to_numpy()[has], want[:, 13].astype(int))
                                   and np.array_equal(df["track_id"].to_numpy(dtype=float, na_value=np.nan)[has], want[:, 0]))
    return out


if __name__ == "__main__":
    case = sys.argv[1]
    res = case_bytetrack("--cpu-standin" in sys.argv) if case == "bytetrack" else case_chain()
    print("RESULT " + json.dumps(res))
