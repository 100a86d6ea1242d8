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
    "ocsort_ctdist_byte_s1003": ("ocsort", dict(seed=1003, n_frames=120, n_ids=30, conf_range=(0.05, 1.0)),
                                 dict(det_thresh=0.5, max_age=10, min_hits=2, iou_threshold=0.3,
                                      delta_t=2, asso_func="ct_dist", inertia=0.2, use_byte=True)),
    "ocsort_diou_s1004": ("ocsort", dict(seed=1004, n_frames=80, n_ids=24, conf_range=(0.05, 1.0)),
                          dict(det_thresh=0.5, max_age=12, min_hits=2, iou_threshold=0.3, delta_t=3, asso_func="diou", inertia=0.2, use_byte=True)),
    "ocsort_ciou_s1005": ("ocsort", dict(seed=1005, n_frames=80, n_ids=24, conf_range=(0.05, 1.0)),
                          dict(det_thresh=0.5, max_age=12, min_hits=2, iou_threshold=0.3, delta_t=1, asso_func="ciou", inertia=0.3, use_byte=False)),
    "strongsort_s4000": ("strongsort", dict(seed=4000, n_frames=160, n_ids=30, emb_dim=64),
                         dict(max_dist=0.1594374041012136, max_iou_dist=0.5431835667667874, max_age=40, max_unmatched_preds=0,
                              n_init=3, nn_budget=100, mc_lambda=0.995, ema_alpha=0.8962157769329083)),
    "strongsort_budget8_s4001": ("strongsort", dict(seed=4001, n_frames=120, n_ids=44, emb_dim=128, conf_range=(0.3, 1.0)),
                                 dict(max_dist=0.2, max_iou_dist=0.7, max_age=12, max_unmatched_preds=0, n_init=2, nn_budget=8,
                                      mc_lambda=0.98, ema_alpha=0.9)),
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
    elif tracker == "strongsort":
        # StrongSORT.__init__ builds the in-tracker ReID network (strong_sort.py:33); the association is exercised with
        # externally supplied features, so the object is assembled without it and _get_features is replaced.
        from strong_sort.sort.nn_matching import NearestNeighborDistanceMetric
        from strong_sort.sort.tracker import Tracker
        from strong_sort.strong_sort import StrongSORT
        model = object.__new__(StrongSORT)
        model.max_dist = hyper["max_dist"]
        metric = NearestNeighborDistanceMetric("cosine", hyper["max_dist"], hyper["nn_budget"])
        model.tracker = Tracker(metric, max_iou_dist=hyper["max_iou_dist"], max_age=hyper["max_age"], n_init=hyper["n_init"],
                                max_unmatched_preds=hyper["max_unmatched_preds"], mc_lambda=hyper["mc_lambda"],
                                ema_alpha=hyper["ema_alpha"])
    else:
        raise KeyError(tracker)
    rows, frames = [], []
    fake_img = np.zeros((video.height, video.width, 3), dtype=np.uint8)
    for f in range(video.n_frames):
        d = video.frame(f)
        if len(d) == 0:
            continue
        keep = d[:, 4] > MIN_CONF
        d = d[keep]
        if tracker == "strongsort":
            feats = video.embeddings[video.offsets[f]:video.offsets[f + 1]][keep]
            model._get_features = lambda xywhs, img, _f=feats: torch.from_numpy(_f.copy())
            with torch.no_grad():
                res = model.update(torch.from_numpy(d.copy()), fake_img)
            res = np.asarray(res)
            res = res[:, [0, 1, 2, 3, 4, 5, 6, 8]].astype(np.float64) if res.size else np.zeros((0, 8))
        else:
            with torch.no_grad():
                res = np.asarray(model.update(torch.from_numpy(d.copy()), None), dtype=np.float64)
        if res.size:
            rows.append(res.reshape(-1, 8))
            frames.append(np.full(res.reshape(-1, 8).shape[0], f, dtype=np.int32))
    if not rows:
        return np.zeros((0, 8)), np.zeros((0,), dtype=np.int32)
    return np.concatenate(rows), np.concatenate(frames)


def run_strongsort_end_to_end(name="strongsort_e2e_s5000", gen=None):
    """The UNMODIFIED StrongSORT plugin incl. its in-tracker ReID (vendored ResNet-50, PIL crops) on synthetic frames.
    Weights: tracklab_b200.nets.resnet_reid.build_resnet50_reid(seed) exported into the reference's conv+BN format
    (identity BatchNorm statistics), so both sides hold the same function without committing 94 MB of weights."""
    import tempfile
    from pathlib import Path

    from strong_sort.strong_sort import StrongSORT

    from tracklab_b200.nets.resnet_reid import build_resnet50_reid
    from tracklab_b200.synth import make_frames
    gen = gen or dict(seed=5000, n_frames=30, n_ids=12)
    hyper = dict(max_dist=0.1594374041012136, max_iou_dist=0.5431835667667874, max_age=40, max_unmatched_preds=0,
                 n_init=3, nn_budget=100, mc_lambda=0.995, ema_alpha=0.8962157769329083)
    video = make_video(**gen)
    mine = build_resnet50_reid(1234)
    sd = {}
    eps = 1e-5

    def put(conv_key, bn_key, m):
        sd[conv_key + ".weight"] = m.conv.weight.detach() * (1.0 + eps) ** 0.5
        c = m.conv.weight.shape[0]
        sd[bn_key + ".weight"], sd[bn_key + ".bias"] = torch.ones(c), m.conv.bias.detach().clone()
        sd[bn_key + ".running_mean"], sd[bn_key + ".running_var"] = torch.zeros(c), torch.ones(c)
        sd[bn_key + ".num_batches_tracked"] = torch.tensor(0)

    put("conv1", "bn1", mine.conv1)
    for li, layer in enumerate((mine.layer1, mine.layer2, mine.layer3, mine.layer4), start=1):
        for bi, blk in enumerate(layer):
            q = f"layer{li}.{bi}"
            put(q + ".conv1", q + ".bn1", blk.conv1); put(q + ".conv2", q + ".bn2", blk.conv2); put(q + ".conv3", q + ".bn3", blk.conv3)
            if blk.down is not None:
                put(q + ".downsample.0", q + ".downsample.1", blk.down)
    tmp = Path(tempfile.mkdtemp()) / "resnet50_synth.pt"
    torch.save(sd, tmp)
    model = StrongSORT(tmp, torch.device("cpu"), False, **hyper)
    rows, frames = [], []
    for f in range(video.n_frames):
        d = video.frame(f)
        d = d[d[:, 4] > MIN_CONF]
        img = make_frames(video, f, f + 1, device="cpu")[0].numpy()
        with torch.no_grad():
            res = np.asarray(model.update(torch.from_numpy(d.copy()), img))
        if res.size:
            res = res[:, [0, 1, 2, 3, 4, 5, 6, 8]].astype(np.float64)
            rows.append(res); frames.append(np.full(len(res), f, dtype=np.int32))
    rows, frames = np.concatenate(rows), np.concatenate(frames)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), rows=rows, frames=frames,
                        dets_sha=np.frombuffer(__import__("hashlib").sha256(video.dets.tobytes()).digest(), dtype=np.uint8),
                        tracker="strongsort_e2e", gen=repr(gen), hyper=repr(hyper), min_conf=MIN_CONF)
    print(f"{name}: {video.n_dets} dets -> {rows.shape[0]} rows, {len(np.unique(rows[:, 4]))} ids")


def run_strongsort_ecc(name="strongsort_ecc_s4002"):
    """The UNMODIFIED StrongSORT tracker with camera compensation (strong_sort_api.py:59-65: tracker.camera_update(prev, img) before
    update) on a camera_drift video with externally supplied features. Every track runs cv2.findTransformECC itself (track.py:129-214);
    the matrix of each frame pair is recorded too (it is the same for all tracks) so that the tests can feed it to the device tracker."""
    import cv2

    from strong_sort.sort.nn_matching import NearestNeighborDistanceMetric
    from strong_sort.sort.track import Track
    from strong_sort.sort.tracker import Tracker
    from strong_sort.strong_sort import StrongSORT

    from tracklab_b200.synth import make_frames
    gen = dict(seed=4002, n_frames=90, n_ids=24, emb_dim=64, camera_drift=True)
    hyper = dict(max_dist=0.1594374041012136, max_iou_dist=0.5431835667667874, max_age=40, max_unmatched_preds=0,
                 n_init=3, nn_budget=100, mc_lambda=0.995, ema_alpha=0.8962157769329083)
    video = make_video(**gen)
    model = object.__new__(StrongSORT)
    model.max_dist = hyper["max_dist"]
    metric = NearestNeighborDistanceMetric("cosine", hyper["max_dist"], hyper["nn_budget"])
    model.tracker = Tracker(metric, max_iou_dist=hyper["max_iou_dist"], max_age=hyper["max_age"], n_init=hyper["n_init"],
                            max_unmatched_preds=hyper["max_unmatched_preds"], mc_lambda=hyper["mc_lambda"], ema_alpha=hyper["ema_alpha"])
    rows, frames, warps = [], [], np.full((video.n_frames, 6), np.nan, dtype=np.float32)
    prev = None
    probe = object.__new__(Track)
    for f in range(video.n_frames):
        img = make_frames(video, f, f + 1, device="cpu")[0].numpy()
        if prev is not None:
            wm, _ = probe.ECC(prev, img)
            if wm is not None:
                warps[f] = np.asarray(wm, dtype=np.float32).reshape(-1)
            model.tracker.camera_update(prev, img)
        prev = img
        d = video.frame(f)
        if len(d) == 0:
            continue
        keep = d[:, 4] > MIN_CONF
        d = d[keep]
        feats = video.embeddings[video.offsets[f]:video.offsets[f + 1]][keep]
        model._get_features = lambda xywhs, im, _f=feats: torch.from_numpy(_f.copy())
        with torch.no_grad():
            res = np.asarray(model.update(torch.from_numpy(d.copy()), img))
        if res.size:
            res = res[:, [0, 1, 2, 3, 4, 5, 6, 8]].astype(np.float64)
            rows.append(res); frames.append(np.full(len(res), f, dtype=np.int32))
    rows, frames = np.concatenate(rows), np.concatenate(frames)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), rows=rows, frames=frames, warps=warps,
                        dets_sha=np.frombuffer(__import__("hashlib").sha256(video.dets.tobytes()).digest(), dtype=np.uint8),
                        tracker="strongsort_ecc", gen=repr(gen), hyper=repr(hyper), min_conf=MIN_CONF)
    print(f"{name}: {video.n_dets} dets -> {rows.shape[0]} rows, {len(np.unique(rows[:, 4]))} ids; warps finite in {int(np.isfinite(warps[:, 0]).sum())} frames")


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
    args = sys.argv[1:]
    if "strongsort_e2e" in args:
        run_strongsort_end_to_end()
        args.remove("strongsort_e2e")
        if not args:
            sys.exit(0)
    if "strongsort_ecc" in args:
        run_strongsort_ecc()
        args.remove("strongsort_ecc")
        if not args:
            sys.exit(0)
    if "strongsort_e2e_long" in args:   # 200 frames / 44 identities (~7.5k crops through the plugin's CPU ResNet-50: minutes)
        run_strongsort_end_to_end("strongsort_e2e_s5001", dict(seed=5001, n_frames=200, n_ids=44))
        args.remove("strongsort_e2e_long")
        if not args:
            sys.exit(0)
    main(args or list(CASES))
