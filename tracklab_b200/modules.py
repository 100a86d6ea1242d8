"""B200-native drop-in modules for the reference engine (``tracklab.pipeline`` API).

``ByteTrack`` and ``OCSORT`` keep the reference wrappers' class names, column contracts, constructor signature
``(cfg, device, **kwargs)``, ``reset()`` and the result layout
(/root/reference/tracklab/wrappers/track/byte_track_api.py:14-76, oc_sort_api.py:14-76), but:

  * the per-frame Python tracker is replaced by ONE whole-video kernel launch (libtrackkern, tk_*_run),
    triggered when the engine hands the video's detections to the module's datapipe
    (/root/reference/tracklab/engine/offline.py:24) — explicitly allowed by the module API
    ("datapipe (optional) ... dataloader (optional)", imagelevel_module.py:18-23);
  * the module's dataloader yields whole-video batches (image ids only): no worker processes, no per-sample
    image decode (datapipe.py:27-48 decodes the frame for every tracker sample although trackers ignore it),
    and ``process`` returns the rows of the batch's images from the precomputed device result, so the engine's
    DataFrame merge (engine.py:180-181) runs once per batch instead of once per frame.

There is no CPU fallback: constructing these modules without CUDA raises.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd
import torch

from . import _lib
from .device_trackers import ByteTrackDevice, OCSortDevice
from .pipeline import DetectionLevelModule, ImageLevelModule


def _cfg_get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


class _VideoBatches:
    """Iterable the engine treats as the module's DataLoader: yields ``(image_ids, batch)``."""

    def __init__(self, pipe):
        self.pipe = pipe

    def __iter__(self):
        ids = self.pipe.image_ids
        n = self.pipe.frames_per_batch or len(ids)
        for i in range(0, len(ids), max(1, n)):
            chunk = ids[i:i + n]
            yield chunk, {"image_ids": chunk}

    def __len__(self):
        n = self.pipe.frames_per_batch or max(1, len(self.pipe.image_ids))
        return (len(self.pipe.image_ids) + n - 1) // max(1, n)


class _TrackerDatapipe:
    """Stands in for EngineDatapipe (datastruct/datapipe.py:5-48): ``update`` receives the video's detections."""

    def __init__(self, module, frames_per_batch):
        self.module = module
        self.frames_per_batch = frames_per_batch
        self.image_ids = []

    def update(self, image_filepaths, img_metadatas, detections):
        self.image_ids = list(img_metadatas.index)
        self.module._track_video(img_metadatas, detections)

    def __len__(self):
        return len(self.image_ids)


# ``Module.level`` is derived from the name of the FIRST base class (pipeline/module.py:34-37), so every concrete module
# below inherits ImageLevelModule directly and takes its behaviour from the functions of this section.
_IN_COLS = ["bbox_ltwh", "bbox_conf", "category_id"]
_OUT_COLS = ["track_id", "track_bbox_ltwh", "track_bbox_conf"]


# Constructor defaults of the REFERENCE plugin classes: a partial `hyperparams` dict must run the tracker the reference would run
# (the device classes default to the tuned values of tracklab/configs/modules/track/*.yaml instead).
REF_DEFAULTS = {
    "ByteTrack": dict(track_thresh=0.45, track_buffer=25, match_thresh=0.8, frame_rate=30),                 # byte_tracker.py:152
    "OCSORT": dict(max_age=30, min_hits=3, iou_threshold=0.3, delta_t=3, asso_func="iou", inertia=0.2, use_byte=False),   # ocsort.py:183-184 (det_thresh is required)
    "StrongSORT": dict(max_dist=0.2, max_iou_dist=0.7, max_age=70, max_unmatched_preds=7, n_init=3, nn_budget=100, mc_lambda=0.995,
                       ema_alpha=0.9),                                                                      # strong_sort.py:20-31
}


def _with_reference_defaults(name, hyper):
    out = dict(REF_DEFAULTS.get(name, {}))
    out.update(hyper)
    if name == "OCSORT" and "det_thresh" not in out:
        raise _lib.TrackKernError("OCSORT: hyperparams.det_thresh is required (OCSort.__init__ has no default for it, oc_sort/ocsort.py:183)")
    return out


class _Impl:
    def __init__(self, cfg, device, **kwargs):
        ImageLevelModule.__init__(self, batch_size=1)
        if not torch.cuda.is_available():
            raise _lib.TrackKernError(f"{type(self).__name__} needs a CUDA device: tracklab_b200 has no CPU path")
        self.cfg = cfg
        self.device = torch.device(device if str(device) != "cuda" else "cuda:0")
        self.cap_tracks = int(_cfg_get(cfg, "cap_tracks", 256))
        self.cap_dets = int(_cfg_get(cfg, "cap_dets", 256))
        self.frames_per_batch = _cfg_get(cfg, "frames_per_batch", None)
        hyper = _with_reference_defaults(type(self).__name__, dict(_cfg_get(cfg, "hyperparams", {}) or {}))
        self.tracker = self._device_cls(**hyper, min_confidence=float(_cfg_get(cfg, "min_confidence", 0.4)),
                                        cap_tracks=self.cap_tracks, cap_dets=self.cap_dets, device=self.device)
        self._pipe = _TrackerDatapipe(self, self.frames_per_batch)
        self._result = None
        self._first_reset = True

    # -- module API --------------------------------------------------------------------------------
    def reset(self):
        """New video: drop the tracker state (the reference re-creates the tracker, byte_track_api.py:28-30)."""
        self.tracker.reset(keep_id_counter=self._continue_ids and not self._first_reset)
        self._first_reset = False
        self._result = None

    def datapipe(self):
        return self._pipe

    def dataloader(self, engine=None):
        return _VideoBatches(self._pipe)

    def preprocess(self, image, detections, metadata):   # never called: the datapipe is overridden
        return {"input": []}

    def _track_video(self, img_metadatas: pd.DataFrame, detections: pd.DataFrame):
        """Rows ``[l,t,r,b,conf,cls,det_id]`` of the whole video -> device -> one kernel launch -> cached frame."""
        image_ids = np.asarray(img_metadatas.index)
        if detections is None or len(detections) == 0:
            self._result = pd.DataFrame(columns=["image_id"] + self.output_columns)
            return
        pos = {int(i): k for k, i in enumerate(image_ids)}
        img = detections["image_id"].to_numpy()
        keep = np.fromiter((int(i) in pos for i in img), dtype=bool, count=len(img))
        det = detections[keep]
        frame = np.fromiter((pos[int(i)] for i in det["image_id"].to_numpy()), dtype=np.int64, count=len(det))
        order = np.argsort(frame, kind="stable")          # rows grouped by frame, detection order kept inside a frame
        ltwh = np.stack(det["bbox_ltwh"].to_numpy()).astype(np.float64).reshape(-1, 4)[order]
        rows = np.empty((len(det), 7), dtype=np.float64)
        rows[:, 0] = ltwh[:, 0]
        rows[:, 1] = ltwh[:, 1]
        rows[:, 2] = ltwh[:, 0] + ltwh[:, 2]                # ltwh_to_ltrb (utils/coordinates.py:257-267)
        rows[:, 3] = ltwh[:, 1] + ltwh[:, 3]
        rows[:, 4] = det["bbox_conf"].to_numpy(dtype=np.float64)[order]
        rows[:, 5] = det["category_id"].to_numpy(dtype=np.float64)[order]
        rows[:, 6] = np.asarray(det.index, dtype=np.float64)[order]
        counts = np.bincount(frame, minlength=len(image_ids))
        offsets = np.zeros(len(image_ids) + 1, dtype=np.int32)
        np.cumsum(counts, out=offsets[1:])
        if counts.max() > self.cap_dets:
            raise _lib.TrackKernError(f"{counts.max()} detections in one frame exceed cap_dets={self.cap_dets}")
        d_dev = torch.from_numpy(rows).to(self.device)
        o_dev = torch.from_numpy(offsets)[None].to(self.device)
        out_rows, out_fc, out_cnt = self.tracker.run(d_dev, o_dev)
        self.tracker.check_status()                          # one synchronisation per video
        n = int(out_cnt[0].item())
        res = out_rows[:n].cpu().numpy()
        fc = out_fc[0].cpu().numpy()
        frame_of_row = np.repeat(np.arange(len(fc)), fc)
        ltrb = res[:, :4]
        self._result = pd.DataFrame({
            "track_bbox_ltwh": list(np.column_stack([ltrb[:, 0], ltrb[:, 1], ltrb[:, 2] - ltrb[:, 0], ltrb[:, 3] - ltrb[:, 1]])),
            "track_bbox_conf": res[:, 6],
            "track_id": res[:, 4],
            "image_id": image_ids[frame_of_row],
        }, index=pd.Index(res[:, 7].astype(int), name="idxs"))

    def process(self, batch, detections: pd.DataFrame, metadatas: pd.DataFrame):
        if len(detections) == 0 or self._result is None or len(self._result) == 0:
            return []
        sel = self._result[self._result["image_id"].isin(list(metadatas.index))]
        if len(sel) == 0:
            return []
        assert set(sel.index).issubset(detections.index), \
            "Mismatch of indexes during the tracking. The results should match the detections."
        return sel[["track_bbox_ltwh", "track_bbox_conf", "track_id"]]


def _bind(cls):
    """Copy the shared implementation into a class whose first base is ImageLevelModule."""
    for name in ("__init__", "reset", "dataloader", "preprocess", "_track_video", "process"):
        setattr(cls, name, _Impl.__dict__[name])
    cls.datapipe = property(_Impl.__dict__["datapipe"])
    cls.__abstractmethods__ = frozenset()
    return cls


@_bind
class ByteTrack(ImageLevelModule):
    """Drop-in for tracklab.wrappers.track.byte_track_api.ByteTrack (same name => same ``module.name``)."""
    input_columns = list(_IN_COLS)
    output_columns = list(_OUT_COLS)
    collate_fn = None
    _device_cls = ByteTrackDevice
    _continue_ids = True   # BaseTrack._count is process-global in the reference (byte_track/basetrack.py:13)


@_bind
class OCSORT(ImageLevelModule):
    """Drop-in for tracklab.wrappers.track.oc_sort_api.OCSORT."""
    input_columns = list(_IN_COLS)
    output_columns = list(_OUT_COLS)
    collate_fn = None
    _device_cls = OCSortDevice
    _continue_ids = False


# ---- StrongSORT: detect -> ReID -> associate inside one module, like the reference wrapper ---------------------------
class _StrongSortImpl:
    """Shared implementation bound into ``StrongSORT`` (see _bind). Replaces
    /root/reference/tracklab/wrappers/track/strong_sort_api.py:16-93: the reference decodes the frame in ``process``, crops
    and runs the ReID network inside the tracker for every frame; here the video's frames are decoded once per batch of
    images, all crops of the batch go through the crop-gather kernel + backbone together, and the whole video is
    associated by one tk_strongsort_run_cmc launch. ``cfg.ecc`` (the YAML default, strong_sort.yaml:13): the camera motion of every
    consecutive frame pair is estimated on the device (tk_ecc_gray_small + tk_ecc_euclidean = cv2.findTransformECC, once per pair
    instead of once per track and pair) and applied to every track's box before the frame is processed (sort/track.py:224-239)."""

    def __init__(self, cfg, device, **kwargs):
        ImageLevelModule.__init__(self, batch_size=1)
        if not torch.cuda.is_available():
            raise _lib.TrackKernError(f"{type(self).__name__} needs a CUDA device: tracklab_b200 has no CPU path")
        self.ecc = bool(_cfg_get(cfg, "ecc", False))
        from .device_trackers import StrongSortDevice
        from .reid import ReidStageDevice
        self.cfg = cfg
        self.device = torch.device(device if str(device) != "cuda" else "cuda:0")
        self.cap_tracks = int(_cfg_get(cfg, "cap_tracks", 128))
        self.cap_dets = int(_cfg_get(cfg, "cap_dets", 128))
        self.frames_per_batch = _cfg_get(cfg, "frames_per_batch", None)
        self.decode_batch = int(_cfg_get(cfg, "decode_batch", 16))
        self.hyper = _with_reference_defaults("StrongSORT", dict(_cfg_get(cfg, "hyperparams", {}) or {}))
        if int(self.hyper.get("max_unmatched_preds", 0)) != 0:
            raise _lib.TrackKernError("StrongSORT: only max_unmatched_preds = 0 (tracklab/configs/modules/track/strong_sort.yaml) is built on the "
                                      "device; the plugin's constructor default 7 applies when hyperparams omits it - set it explicitly")
        self.min_confidence = float(_cfg_get(cfg, "min_confidence", 0.4))
        # the plugin's factory reads the architecture from the weights file name (deep/reid_model_factory.py:122-127)
        weights = _cfg_get(cfg, "model_weights", None)
        name = os.path.basename(str(weights)) if weights is not None else ""
        arch = _cfg_get(cfg, "reid_arch", None) or next((a for a in ("osnet_ibn_x1_0", "osnet_x1_0", "resnet50") if a in name), "resnet50")
        model = None
        if weights is not None and not os.path.isfile(str(weights)) and not bool(_cfg_get(cfg, "synthetic_weights", False)):
            # the reference loads or downloads real weights or fails; a typo must not produce plausible tracks from random weights
            raise _lib.TrackKernError(f"ReID weights {weights!r} not found (pass synthetic_weights=True to run on seeded random weights)")
        if weights is not None and os.path.isfile(str(weights)):   # a reference-format state_dict (conv + BN), folded on load
            from .reid import build_reid_model
            sd = torch.load(str(weights), map_location="cpu")
            model = build_reid_model(arch).from_reference_state_dict(sd.get("state_dict", sd))
        self.reid = ReidStageDevice(device=self.device, model=model, precision=_cfg_get(cfg, "reid_precision", "bf16"), arch=arch)
        self._trk_cls = StrongSortDevice
        self.tracker = None
        self._pipe = _TrackerDatapipe(self, self.frames_per_batch)
        self._result = None

    def reset(self):
        self._result = None
        if self.tracker is not None:
            self.tracker.reset()

    def datapipe(self):
        return self._pipe

    def dataloader(self, engine=None):
        return _VideoBatches(self._pipe)

    def preprocess(self, image, detections, metadata):   # never called: the datapipe is overridden
        return {"input": []}

    def _track_video(self, img_metadatas, detections):
        import cv2
        image_ids = np.asarray(img_metadatas.index)
        if detections is None or len(detections) == 0:
            self._result = pd.DataFrame(columns=["image_id"] + self.output_columns)
            return
        rows, offsets, _ = _rows_from_detections(image_ids, detections, self.cap_dets)
        paths = list(img_metadatas["file_path"])
        first = cv2.imread(paths[0])
        H, W = first.shape[:2]
        if self.tracker is None or self.tracker.params.image_width != W or self.tracker.params.image_height != H:
            self.tracker = self._trk_cls(self.reid.feature_dim, **self.hyper, min_confidence=self.min_confidence,
                                         image_size=(W, H), cap_tracks=self.cap_tracks, cap_dets=self.cap_dets, device=self.device)
        d_dev = torch.from_numpy(rows).to(self.device)
        feats = torch.empty((len(rows), self.reid.feature_dim), dtype=torch.float32, device=self.device)
        small = []
        for f0 in range(0, len(paths), self.decode_batch):   # decode once per image (cv2_load_image: BGR -> RGB, utils/cv2.py:54-66)
            f1 = min(len(paths), f0 + self.decode_batch)
            from .ingest import load_frames
            fr = load_frames(paths[f0:f1], self.device, str(_cfg_get(self.cfg, "decode", "cv2")))     # JPEG: nvJPEG on the device; else cv2_load_image
            if self.ecc:
                from . import kernels
                small.append(kernels.ecc_gray_small(fr, 0.1))
            r0, r1 = int(offsets[f0]), int(offsets[f1])
            if r1 > r0:
                det_frame = torch.from_numpy(np.repeat(np.arange(f1 - f0), np.diff(offsets[f0:f1 + 1])).astype(np.int32)).to(self.device)
                feats[r0:r1] = self.reid.features(fr, d_dev[r0:r1], det_frame)
        o_dev = torch.from_numpy(offsets)[None].to(self.device)
        warps = None
        if self.ecc:
            from . import kernels
            warps, _, _ = kernels.ecc_euclidean(torch.cat(small), 100, 1e-5, 0.1)   # row f: frame f-1 -> f, row 0 NaN (no previous frame)
            warps = warps[None].contiguous()
        out_rows, out_fc, out_cnt = self.tracker.run(d_dev, o_dev, feats, warps=warps)
        self.tracker.check_status()
        n = int(out_cnt[0].item())
        res = out_rows[:n].cpu().numpy()
        fc = out_fc[0].cpu().numpy()
        frame_of_row = np.repeat(np.arange(len(fc)), fc)
        ltrb = res[:, :4]
        self._result = pd.DataFrame({
            "track_bbox_ltwh": list(np.column_stack([ltrb[:, 0], ltrb[:, 1], ltrb[:, 2] - ltrb[:, 0], ltrb[:, 3] - ltrb[:, 1]])),
            "track_bbox_conf": res[:, 6], "track_id": res[:, 4], "image_id": image_ids[frame_of_row],
        }, index=pd.Index(res[:, 7].astype(int), name="idxs"))

    def process(self, batch, detections, metadatas):
        if len(detections) == 0 or self._result is None or len(self._result) == 0:
            return []
        sel = self._result[self._result["image_id"].isin(list(metadatas.index))]
        if len(sel) == 0:
            return []
        # a track is reported up to one frame after its last update with the detection id of that update, so a row may
        # point at a detection of the previous image — the reference wrapper lets the override happen (strong_sort_api.py:78-82)
        sel = sel[~sel.index.duplicated(keep="last")]
        return sel[["track_bbox_ltwh", "track_bbox_conf", "track_id"]]


class _DeepOCSortImpl(_StrongSortImpl):
    """Shared implementation bound into ``DeepOCSORT``. Replaces /root/reference/tracklab/wrappers/track/deep_oc_sort_api.py:16-91:
    the reference decodes the frame in ``process``, crops (``box.astype(int)`` + NumPy slice, deep_oc_sort/ocsort.py:560-565), runs
    the ReID network and the camera-motion estimator inside the tracker for every frame; here all crops of a batch of frames go through
    the crop-gather kernel (TK_CROP_RULE_XYXY_INT) + backbone together and the whole video is associated by one tk_deepocsort_run
    launch. Camera motion: the plugin's estimator is sparse optical flow + RANSAC (cmc.py:138-166, OpenCV); here the 2x3 transform of
    every consecutive frame pair comes from the device ECC (tk_ecc_gray_small + tk_ecc_euclidean, the estimator StrongSORT's wrapper
    uses) and the transforms of frames the wrapper skips (no detections) are composed into the next processed frame - a different
    estimator of the same camera motion; the association itself is the plugin's, operation by operation, given the matrices."""

    def __init__(self, cfg, device, **kwargs):
        ImageLevelModule.__init__(self, batch_size=1)
        if not torch.cuda.is_available():
            raise _lib.TrackKernError(f"{type(self).__name__} needs a CUDA device: tracklab_b200 has no CPU path")
        from .device_trackers import DeepOCSortDevice
        from .reid import ReidStageDevice
        self.cfg = cfg
        self.device = torch.device(device if str(device) != "cuda" else "cuda:0")
        self.cap_tracks = int(_cfg_get(cfg, "cap_tracks", 128))
        self.cap_dets = int(_cfg_get(cfg, "cap_dets", 128))
        self.frames_per_batch = _cfg_get(cfg, "frames_per_batch", None)
        self.decode_batch = int(_cfg_get(cfg, "decode_batch", 16))
        self.hyper = dict(_cfg_get(cfg, "hyperparams", {}) or {})
        self.min_confidence = float(_cfg_get(cfg, "min_confidence", 0.4))
        self.ecc = not bool(self.hyper.get("cmc_off", False))
        weights = _cfg_get(cfg, "model_weights", None)
        name = os.path.basename(str(weights)) if weights is not None else ""
        arch = _cfg_get(cfg, "reid_arch", None) or next((a for a in ("osnet_ibn_x1_0", "osnet_x1_0", "resnet50") if a in name), "osnet_x1_0")
        model = None
        if weights is not None and not os.path.isfile(str(weights)) and not bool(_cfg_get(cfg, "synthetic_weights", False)):
            raise _lib.TrackKernError(f"ReID weights {weights!r} not found (pass synthetic_weights=True to run on seeded random weights)")
        if weights is not None and os.path.isfile(str(weights)):
            from .reid import build_reid_model
            sd = torch.load(str(weights), map_location="cpu")
            model = build_reid_model(arch).from_reference_state_dict(sd.get("state_dict", sd))
        self.reid = None
        if not bool(self.hyper.get("embedding_off", False)):
            self.reid = ReidStageDevice(device=self.device, model=model, precision=_cfg_get(cfg, "reid_precision", "bf16"), arch=arch)
        self._trk_cls = DeepOCSortDevice
        self.tracker = None
        self._pipe = _TrackerDatapipe(self, self.frames_per_batch)
        self._result = None

    def _track_video(self, img_metadatas, detections):
        import cv2

        from . import kernels
        image_ids = np.asarray(img_metadatas.index)
        if detections is None or len(detections) == 0:
            self._result = pd.DataFrame(columns=["image_id"] + self.output_columns)
            return
        rows, offsets, _ = _rows_from_detections(image_ids, detections, self.cap_dets)
        paths = list(img_metadatas["file_path"])
        E = self.reid.feature_dim if self.reid is not None else 1
        if self.tracker is None:
            self.tracker = self._trk_cls(E, **self.hyper, min_confidence=self.min_confidence, cap_tracks=self.cap_tracks,
                                         cap_dets=self.cap_dets, device=self.device)
        d_dev = torch.from_numpy(rows).to(self.device)
        feats = torch.empty((len(rows), E), dtype=torch.float32, device=self.device) if self.reid is not None else None
        small = []
        for f0 in range(0, len(paths), self.decode_batch):
            f1 = min(len(paths), f0 + self.decode_batch)
            from .ingest import load_frames
            fr = load_frames(paths[f0:f1], self.device, str(_cfg_get(self.cfg, "decode", "cv2")))     # JPEG: nvJPEG on the device; else cv2_load_image
            if self.ecc:
                small.append(kernels.ecc_gray_small(fr, 0.1))
            r0, r1 = int(offsets[f0]), int(offsets[f1])
            if r1 > r0 and self.reid is not None:
                det_frame = torch.from_numpy(np.repeat(np.arange(f1 - f0), np.diff(offsets[f0:f1 + 1])).astype(np.int32)).to(self.device)
                feats[r0:r1] = self.reid.features(fr, d_dev[r0:r1], det_frame, ltwh_rows=getattr(self, "_crop_rule", kernels.CROP_RULE_XYXY_INT))
        o_dev = torch.from_numpy(offsets)[None].to(self.device)
        affines = None
        if self.ecc:
            warps, _, _ = kernels.ecc_euclidean(torch.cat(small), 100, 1e-5, 0.1)   # row f: frame f-1 -> f (NaN: first frame / failed)
            affines = torch.from_numpy(compose_skipped_affines(warps.double().cpu().numpy().reshape(-1, 2, 3), np.diff(offsets) > 0))
            affines = affines[None].contiguous().to(self.device)
        out_rows, out_fc, out_cnt = self.tracker.run(d_dev, o_dev, feats, affines)
        self.tracker.check_status()
        n = int(out_cnt[0].item())
        res = out_rows[:n].cpu().numpy()
        fc = out_fc[0].cpu().numpy()
        frame_of_row = np.repeat(np.arange(len(fc)), fc)
        ltrb = res[:, :4]
        self._result = pd.DataFrame({
            "track_bbox_ltwh": list(np.column_stack([ltrb[:, 0], ltrb[:, 1], ltrb[:, 2] - ltrb[:, 0], ltrb[:, 3] - ltrb[:, 1]])),
            "track_bbox_conf": res[:, 6], "track_id": res[:, 4], "image_id": image_ids[frame_of_row],
        }, index=pd.Index(res[:, 7].astype(int), name="idxs"))

    def process(self, batch, detections, metadatas):
        if len(detections) == 0 or self._result is None or len(self._result) == 0:
            return []
        sel = self._result[self._result["image_id"].isin(list(metadatas.index))]
        if len(sel) == 0:
            return []
        return sel[~sel.index.duplicated(keep="first")][["track_bbox_ltwh", "track_bbox_conf", "track_id"]]   # deep_oc_sort_api.py:88


class _BotSortImpl(_DeepOCSortImpl):
    """Shared implementation bound into ``BotSORT``. Replaces /root/reference/tracklab/wrappers/track/bot_sort_api.py:16-87 like
    ``_DeepOCSortImpl`` replaces the Deep OC-SORT wrapper: ReID forward of all crops of a batch of frames (the plugin's crop is the
    StrongSORT rule: centre box, int(), clip - bot_sort.py:487-495 = TK_CROP_RULE_STRONGSORT), camera motion from the device ECC for every
    ``cmc_method`` except 'none' (the plugin's default estimator is OpenCV sparse optical flow, gmc.py:239-303: a different estimator of the
    same motion), one tk_botsort_run launch per video."""

    def __init__(self, cfg, device, **kwargs):
        ImageLevelModule.__init__(self, batch_size=1)
        if not torch.cuda.is_available():
            raise _lib.TrackKernError(f"{type(self).__name__} needs a CUDA device: tracklab_b200 has no CPU path")
        from .device_trackers import BotSortDevice
        from .reid import ReidStageDevice
        self.cfg = cfg
        self.device = torch.device(device if str(device) != "cuda" else "cuda:0")
        self.cap_tracks = int(_cfg_get(cfg, "cap_tracks", 128))
        self.cap_dets = int(_cfg_get(cfg, "cap_dets", 128))
        self.frames_per_batch = _cfg_get(cfg, "frames_per_batch", None)
        self.decode_batch = int(_cfg_get(cfg, "decode_batch", 16))
        self.hyper = dict(_cfg_get(cfg, "hyperparams", {}) or {})
        self.min_confidence = float(_cfg_get(cfg, "min_confidence", 0.4))
        self.ecc = str(self.hyper.get("cmc_method", "sparseOptFlow")).lower() != "none"
        weights = _cfg_get(cfg, "model_weights", None)
        name = os.path.basename(str(weights)) if weights is not None else ""
        arch = _cfg_get(cfg, "reid_arch", None) or next((a for a in ("osnet_ibn_x1_0", "osnet_x1_0", "resnet50") if a in name), "osnet_ibn_x1_0")
        model = None
        if weights is not None and not os.path.isfile(str(weights)) and not bool(_cfg_get(cfg, "synthetic_weights", False)):
            raise _lib.TrackKernError(f"ReID weights {weights!r} not found (pass synthetic_weights=True to run on seeded random weights)")
        if weights is not None and os.path.isfile(str(weights)):
            from .reid import build_reid_model
            sd = torch.load(str(weights), map_location="cpu")
            model = build_reid_model(arch).from_reference_state_dict(sd.get("state_dict", sd))
        self.reid = ReidStageDevice(device=self.device, model=model, precision=_cfg_get(cfg, "reid_precision", "bf16"), arch=arch)
        self._trk_cls = BotSortDevice
        self._crop_rule = 0          # kernels.CROP_RULE_STRONGSORT
        self.tracker = None
        self._pipe = _TrackerDatapipe(self, self.frames_per_batch)
        self._result = None

    def process(self, batch, detections, metadatas):
        if len(detections) == 0 or self._result is None or len(self._result) == 0:
            return []
        sel = self._result[self._result["image_id"].isin(list(metadatas.index))]
        if len(sel) == 0:
            return []
        return sel[["track_bbox_ltwh", "track_bbox_conf", "track_id"]]


def compose_skipped_affines(pair_warps, processed):
    """pair_warps float64 [F,2,3]: frame f-1 -> f (row 0 / failed pairs NaN -> identity). The plugin's estimator is only called on
    frames the wrapper processes (deep_oc_sort_api.py:61-62) and relates each call to the previous CALL, so the transforms of skipped
    frames are folded into the next processed one; the first processed frame gets the identity (cmc.py:144-148)."""
    F = len(pair_warps)
    out = np.tile(np.eye(2, 3), (F, 1, 1))
    acc, seen = np.eye(3), False
    for f in range(F):
        A = np.eye(3)
        if f > 0 and np.isfinite(pair_warps[f]).all():
            A[:2] = pair_warps[f]
        acc = A @ acc
        if processed[f]:
            if seen:
                out[f] = acc[:2]
            seen = True
            acc = np.eye(3)
    return out


def _rows_from_detections(image_ids, detections, cap_dets):
    """DataFrame rows of one video -> float64 [N,7] = [l,t,r,b,conf,cls,det_id] grouped by frame + int32 offsets."""
    pos = {int(i): k for k, i in enumerate(image_ids)}
    img = detections["image_id"].to_numpy()
    keep = np.fromiter((int(i) in pos for i in img), dtype=bool, count=len(img))
    det = detections[keep]
    frame = np.fromiter((pos[int(i)] for i in det["image_id"].to_numpy()), dtype=np.int64, count=len(det))
    order = np.argsort(frame, kind="stable")
    ltwh = np.stack(det["bbox_ltwh"].to_numpy()).astype(np.float64).reshape(-1, 4)[order]
    rows = np.empty((len(det), 7), dtype=np.float64)
    rows[:, 0], rows[:, 1] = ltwh[:, 0], ltwh[:, 1]
    rows[:, 2], rows[:, 3] = ltwh[:, 0] + ltwh[:, 2], ltwh[:, 1] + ltwh[:, 3]
    rows[:, 4] = det["bbox_conf"].to_numpy(dtype=np.float64)[order]
    rows[:, 5] = det["category_id"].to_numpy(dtype=np.float64)[order]
    rows[:, 6] = np.asarray(det.index, dtype=np.float64)[order]
    counts = np.bincount(frame, minlength=len(image_ids))
    offsets = np.zeros(len(image_ids) + 1, dtype=np.int32)
    np.cumsum(counts, out=offsets[1:])
    if len(counts) and counts.max() > cap_dets:
        raise _lib.TrackKernError(f"{counts.max()} detections in one frame exceed cap_dets={cap_dets}")
    return rows, offsets, frame[order]


def _bind_from(cls, impl):
    for name in ("__init__", "reset", "dataloader", "preprocess", "_track_video", "process"):
        setattr(cls, name, impl.__dict__[name])
    cls.datapipe = property(impl.__dict__["datapipe"])
    cls.__abstractmethods__ = frozenset()
    return cls


class StrongSORT(ImageLevelModule):
    """Drop-in for tracklab.wrappers.track.strong_sort_api.StrongSORT (ReID + ECC camera compensation + association on device)."""
    input_columns = list(_IN_COLS)
    output_columns = list(_OUT_COLS)
    collate_fn = None


_bind_from(StrongSORT, _StrongSortImpl)


class DeepOCSORT(ImageLevelModule):
    """Drop-in for tracklab.wrappers.track.deep_oc_sort_api.DeepOCSORT (ReID + camera-motion affines + association on device)."""
    input_columns = list(_IN_COLS)
    output_columns = list(_OUT_COLS)
    collate_fn = None


_bind_from(DeepOCSORT, _StrongSortImpl)
for _name in ("__init__", "_track_video", "process"):
    setattr(DeepOCSORT, _name, _DeepOCSortImpl.__dict__[_name])


class BotSORT(ImageLevelModule):
    """Drop-in for tracklab.wrappers.track.bot_sort_api.BotSORT (ReID + camera-motion warps + association on device)."""
    input_columns = list(_IN_COLS)
    output_columns = list(_OUT_COLS)
    collate_fn = None


_bind_from(BotSORT, _StrongSortImpl)
BotSORT._track_video = _DeepOCSortImpl.__dict__["_track_video"]
for _name in ("__init__", "process"):
    setattr(BotSORT, _name, _BotSortImpl.__dict__[_name])


# ---- BPBReID-StrongSORT: part-based embeddings + visibility scores come from the upstream ReID module ---------------
class _BpbreidImpl:
    """Shared implementation bound into ``BPBReIDStrongSORT``. Replaces
    /root/reference/tracklab/wrappers/track/bpbreid_strong_sort_api.py:13-118: the video's ``bbox_ltwh`` / ``embeddings``
    [K,E] / ``visibility_scores`` [K] columns go to the device once and tk_bpbreid_run associates the whole video.
    Only ``matching_strategy: strong_sort_matching`` + ``motion_criterium: iou`` with ``ecc: false`` (the reference YAML) are
    implemented; the ``costs`` column (per-detection visualisation dictionaries, sort/tracker.py:365-407) is left empty."""

    def __init__(self, cfg, device, batch_size=None, **kwargs):
        ImageLevelModule.__init__(self, batch_size=1)
        if not torch.cuda.is_available():
            raise _lib.TrackKernError(f"{type(self).__name__} needs a CUDA device: tracklab_b200 has no CPU path")
        if bool(_cfg_get(cfg, "ecc", False)):
            raise _lib.TrackKernError("ecc=True (cv2.findTransformECC camera compensation) is not implemented on device")
        if _cfg_get(cfg, "motion_criterium", "iou") != "iou":
            raise _lib.TrackKernError("only motion_criterium=iou is implemented on device (oks needs the pose module's keypoints)")
        self.cfg = cfg
        self.device = torch.device(device if str(device) != "cuda" else "cuda:0")
        self.cap_tracks = int(_cfg_get(cfg, "cap_tracks", 1024))
        self.cap_dets = int(_cfg_get(cfg, "cap_dets", 128))
        self.ctas_per_video = int(_cfg_get(cfg, "ctas_per_video", 24))
        self.frames_per_batch = _cfg_get(cfg, "frames_per_batch", None)
        self.hyper = dict(ema_alpha=float(_cfg_get(cfg, "ema_alpha", 0.9)), mc_lambda=float(_cfg_get(cfg, "mc_lambda", 0.995)),
                          max_dist=float(_cfg_get(cfg, "max_dist", 0.5)), max_iou_distance=float(_cfg_get(cfg, "max_iou_distance", 0.8)),
                          max_age=int(_cfg_get(cfg, "max_age", 300)), n_init=int(_cfg_get(cfg, "n_init", 0)),
                          min_bbox_confidence=float(_cfg_get(cfg, "min_bbox_confidence", 0.0)),
                          max_kalman_prediction_without_update=int(_cfg_get(cfg, "max_kalman_prediction_without_update", 7)),
                          matching_strategy=str(_cfg_get(cfg, "matching_strategy", "strong_sort_matching")),
                          gating_thres_factor=float(_cfg_get(cfg, "gating_thres_factor", 1.0)), w_kfgd=float(_cfg_get(cfg, "w_kfgd", 1.0)),
                          w_reid=float(_cfg_get(cfg, "w_reid", 1.0)), w_st=float(_cfg_get(cfg, "w_st", 1.0)))
        self.tracker = None
        self._pipe = _TrackerDatapipe(self, self.frames_per_batch)
        self._result = None

    def reset(self):
        self._result = None
        if self.tracker is not None:
            self.tracker.reset()

    def datapipe(self):
        return self._pipe

    def dataloader(self, engine=None):
        return _VideoBatches(self._pipe)

    def preprocess(self, image, detections, metadata):   # never called: the datapipe is overridden
        return {"input": []}

    def _track_video(self, img_metadatas, detections):
        from .device_trackers import BpbreidStrongSortDevice
        image_ids = np.asarray(img_metadatas.index)
        if detections is None or len(detections) == 0:
            self._result = pd.DataFrame(columns=["image_id"] + self.output_columns)
            return
        pos = {int(i): k for k, i in enumerate(image_ids)}
        img = detections["image_id"].to_numpy()
        keep = np.fromiter((int(i) in pos for i in img), dtype=bool, count=len(img))
        det = detections[keep]
        frame = np.fromiter((pos[int(i)] for i in det["image_id"].to_numpy()), dtype=np.int64, count=len(det))
        order = np.argsort(frame, kind="stable")
        rows = np.zeros((len(det), 7), dtype=np.float64)
        rows[:, :4] = np.stack(det["bbox_ltwh"].to_numpy()).astype(np.float64).reshape(-1, 4)[order]   # np.asarray(..., dtype=float) strong_sort.py:69
        conf_col = "bbox_conf" if "bbox_conf" in det.columns else "keypoints_conf"                     # bpbreid_strong_sort_api.py:85-88
        rows[:, 4] = det[conf_col].to_numpy(dtype=np.float64)[order]
        rows[:, 6] = np.asarray(det.index, dtype=np.float64)[order]
        feats = np.ascontiguousarray(np.stack(det["embeddings"].to_numpy()).astype(np.float32)[order])
        vis = np.ascontiguousarray(np.stack(det["visibility_scores"].to_numpy()).astype(np.float32)[order])
        counts = np.bincount(frame, minlength=len(image_ids))
        offsets = np.zeros(len(image_ids) + 1, dtype=np.int32)
        np.cumsum(counts, out=offsets[1:])
        if counts.max() > self.cap_dets:
            raise _lib.TrackKernError(f"{counts.max()} detections in one frame exceed cap_dets={self.cap_dets}")
        K, E = feats.shape[1:]
        if self.tracker is None or (self.tracker.n_parts, self.tracker.feature_dim) != (K, E):
            self.tracker = BpbreidStrongSortDevice(K, E, **self.hyper, ctas_per_video=self.ctas_per_video, cap_tracks=self.cap_tracks,
                                                   cap_dets=self.cap_dets, device=self.device)
        out_rows, out_fc, out_cnt = self.tracker.run(torch.from_numpy(rows).to(self.device), torch.from_numpy(offsets)[None].to(self.device),
                                                     torch.from_numpy(feats).to(self.device), torch.from_numpy(vis).to(self.device))
        self.tracker.check_status()
        n = int(out_cnt[0].item())
        res = out_rows[:n].cpu().numpy()
        fc = out_fc[0].cpu().numpy()
        frame_of_row = np.repeat(np.arange(len(fc)), fc)
        code = res[:, 9].astype(int)
        self._result = pd.DataFrame({
            "track_id": res[:, 0].astype(int),
            "track_bbox_kf_ltwh": list(res[:, 1:5]),
            "track_bbox_pred_kf_ltwh": [None if np.isnan(r[0]) else r for r in res[:, 5:9]],
            "matched_with": [None if c == 0 else ("R" if c == 1 else "S", float(d)) for c, d in zip(code, res[:, 10])],
            "costs": [{} for _ in range(n)],
            "hits": res[:, 11].astype(int), "age": res[:, 12].astype(int),
            "time_since_update": np.zeros(n, dtype=int), "state": ["c"] * n,
            "image_id": image_ids[frame_of_row],
        }, index=pd.Index(res[:, 13].astype(int)))

    def process(self, batch, detections, metadatas):
        if len(detections) == 0 or self._result is None or len(self._result) == 0:
            return []
        sel = self._result[self._result["image_id"].isin(list(metadatas.index))]
        if len(sel) == 0:
            return []
        assert set(sel.index).issubset(detections.index), \
            "Mismatch of indexes during the tracking. The results should match the detections."
        return sel[list(self.output_columns)]


class BPBReIDStrongSORT(ImageLevelModule):
    """Drop-in for tracklab.wrappers.track.bpbreid_strong_sort_api.BPBReIDStrongSORT."""
    input_columns = ["bbox_ltwh", "embeddings", "visibility_scores"]
    output_columns = ["track_id", "track_bbox_kf_ltwh", "track_bbox_pred_kf_ltwh", "matched_with", "costs", "hits", "age",
                      "time_since_update", "state"]
    collate_fn = None


_bind_from(BPBReIDStrongSORT, _BpbreidImpl)


# ---- RT-DETR detector (transformers flavour of the bbox_detector step) -------------------------------------------------
class RTDetr(ImageLevelModule):
    """Drop-in for tracklab.wrappers.bbox_detector.transformers_api.RTDetr: same constructor, columns, running detection id and
    row contents; resize, model and post-processing run on the device for the whole batch (tk_resize_frames_u8 ->
    RTDetrForObjectDetection -> tk_rtdetr_decode). ``model_name`` selects the architecture; weights come from ``weights``
    (a state_dict file) or are seeded (there is no network for ``from_pretrained``)."""
    input_columns = []
    output_columns = ["image_id", "video_id", "category_id", "bbox_ltwh", "bbox_conf"]

    def __init__(self, device, batch_size, model_name="rtdetr_r50vd_coco_o365", min_confidence=0.4, weights=None, precision="bf16",
                 seed=1234, **kwargs):
        super().__init__(batch_size)
        if not torch.cuda.is_available():
            raise _lib.TrackKernError("RTDetr needs a CUDA device: tracklab_b200 has no CPU path")
        if "r50vd" not in model_name or "v2" in model_name:
            raise _lib.TrackKernError(f"{model_name}: only the r50vd RT-DETR architecture is built (configs/modules/bbox_detector/rtdetr_transformers.yaml)")
        from .nets.rtdetr import build_rtdetr
        from .rtdetr_detector import RTDetrDetectorDevice
        self.device = torch.device(device if str(device) != "cuda" else "cuda:0")
        if weights is not None and not os.path.isfile(str(weights)):
            raise _lib.TrackKernError(f"RT-DETR weights {weights!r} not found (omit `weights` to run on seeded random weights)")
        sd = torch.load(str(weights), map_location="cpu") if weights is not None else None
        self.detector = RTDetrDetectorDevice(self.device, min_confidence, precision, model=build_rtdetr(seed, state_dict=sd))
        self._calibrate = sd is None     # seeded weights: the person bias is set on the first batch (nets/rtdetr.py)
        self.min_confidence = min_confidence
        self.id = 0

    @torch.no_grad()
    def preprocess(self, image, detections: pd.DataFrame, metadata: pd.Series):
        return image

    @torch.no_grad()
    def process(self, batch, detections: pd.DataFrame, metadatas: pd.DataFrame):
        frames = torch.as_tensor(batch)
        if frames.dim() == 3:
            frames = frames[None]
        frames = frames.to(self.device, non_blocking=True).contiguous()
        if self._calibrate:
            self.detector.calibrate(frames[:1])
            self._calibrate = False
        rows, counts = self.detector.detect_batch(frames)
        rows, counts = rows.cpu().numpy(), counts.cpu().numpy()        # one read per batch
        out = []
        for i in range(len(counts)):
            for r in rows[i, :counts[i]]:
                out.append(pd.Series(dict(image_id=metadatas["id"].values[i], bbox_ltwh=r[:4].astype(np.float32), bbox_conf=float(r[4]),
                                          video_id=metadatas["video_id"].values[i], category_id=1), name=self.id))
                self.id += 1
        return out


# ---- YOLOX detector (rtmlib flavour of the bbox_detector step) -----------------------------------------------------------
class _ImageBatches:
    """Iterable the engine treats as a DataLoader: yields ``(image_ids, {"image_ids": ids})`` chunks of ``batch`` images."""

    def __init__(self, pipe, batch):
        self.pipe, self.batch = pipe, max(1, int(batch))

    def __iter__(self):
        ids = self.pipe.image_ids
        for i in range(0, len(ids), self.batch):
            yield ids[i:i + self.batch], {"image_ids": ids[i:i + self.batch]}

    def __len__(self):
        return (len(self.pipe.image_ids) + self.batch - 1) // self.batch


class _PathsDatapipe:
    """Stands in for EngineDatapipe (datastruct/datapipe.py:5-48) without per-sample decoding in worker processes: ``update``
    just remembers the video's image paths; ``process`` decodes the images of its batch once."""

    def __init__(self):
        self.image_ids, self.paths = [], {}

    def update(self, image_filepaths, img_metadatas, detections):
        self.image_ids = list(img_metadatas.index)
        self.paths = dict(image_filepaths) if image_filepaths else {i: p for i, p in img_metadatas["file_path"].items()}

    def __len__(self):
        return len(self.image_ids)


def _decode_rgb(paths):
    import cv2
    return np.stack([cv2.cvtColor(cv2.imread(str(p)), cv2.COLOR_BGR2RGB) for p in paths])     # cv2_load_image (utils/cv2.py:54-66)


class RTMLibDetector(ImageLevelModule):
    """Drop-in for tracklab.wrappers.bbox_detector.rtmlib_api.RTMLibDetector (same name, columns, running detection id,
    ``bbox_conf = 1.0``, float32 clipped ``bbox_ltwh``, /root/reference/tracklab/wrappers/bbox_detector/rtmlib_api.py:12-46).
    The reference instantiates ``rtmlib.YOLOX(onnx_model=<url>, model_input_size=[640, 640])`` on onnxruntime and calls it once
    per image; here the YOLOX variant is read from the same config node (``yolox_s`` / ``yolox_m`` in ``model.onnx_model``) and a
    batch of images goes through letterbox -> YOLOX (bf16, CUDA graph) -> decode + NMS -> tk_pack_detections_ex on the device.
    Weights: ``weights`` (a tracklab_b200 YOLOX state_dict; "auto" = weights/yolox_<v>_synth.pt); a missing file is an error."""
    input_columns = []
    output_columns = ["image_id", "video_id", "category_id", "bbox_ltwh", "bbox_conf"]
    collate_fn = None

    def __init__(self, device, model=None, batch_size=16, weights="auto", score_thr=0.7, nms_thr=0.45, synthetic_weights=False, **kwargs):
        super().__init__(batch_size=1)
        if not torch.cuda.is_available():
            raise _lib.TrackKernError("RTMLibDetector needs a CUDA device: tracklab_b200 has no CPU path")
        import re

        from .detector import YoloxDetectorDevice, synth_weights_path
        self.device = torch.device(device if str(device) != "cuda" else "cuda:0")
        url = str(_cfg_get(model, "onnx_model", "") or "") if model is not None else ""
        mt = re.search(r"yolox[_-](tiny|s|m|l)(?![a-z])", url)
        self.variant = str(_cfg_get(model, "variant", None) or (mt.group(1) if mt else "s"))
        size = _cfg_get(model, "model_input_size", [640, 640]) if model is not None else [640, 640]
        if weights == "auto":
            weights = synth_weights_path(self.variant)
            if weights is None and not synthetic_weights:
                raise _lib.TrackKernError(f"no weights for YOLOX-{self.variant} under weights/ (pass weights=<state_dict file>, or "
                                          "synthetic_weights=True to run on seeded random weights)")
        self.frames_per_batch = int(batch_size)
        self.detector = YoloxDetectorDevice(self.variant, device=self.device, batch=self.frames_per_batch, input_size=int(size[0]),
                                            score_thr=score_thr, nms_thr=nms_thr, frames_cap=self.frames_per_batch,
                                            dets_cap=self.frames_per_batch * 256, weights=weights, rows_ltwh=True)
        self._needs_calibration = weights is None
        self._pipe = _PathsDatapipe()
        self.id = 0

    @property
    def datapipe(self):
        return self._pipe

    def dataloader(self, engine=None):
        return _ImageBatches(self._pipe, self.frames_per_batch)

    def preprocess(self, image, detections: pd.DataFrame, metadata: pd.Series):
        return {}

    @torch.no_grad()
    def process(self, batch, detections: pd.DataFrame, metadatas: pd.DataFrame):
        ids = list(metadatas.index)
        paths = [self._pipe.paths[i] if i in self._pipe.paths else metadatas.loc[i, "file_path"] for i in ids]
        from .ingest import load_frames
        frames = load_frames(paths, self.device, getattr(self, "decode", "cv2"))      # JPEG: nvJPEG on the device; else cv2_load_image
        det = self.detector
        if self._needs_calibration:
            det.calibrate(frames, target_per_image=60.0)
            self._needs_calibration = False
        n = det.detect_into(frames)                             # one synchronisation per batch of images
        rows = det.dets[:n].cpu().numpy()
        offs = det.offsets[:len(ids) + 1].cpu().numpy()
        frame_of_row = np.repeat(np.arange(len(ids)), np.diff(offs))
        image_ids = np.asarray(ids)[frame_of_row]
        video_ids = metadatas["video_id"].to_numpy()[frame_of_row]
        out = pd.DataFrame({"image_id": image_ids, "bbox_ltwh": list(rows[:, :4].astype(np.float32)), "bbox_conf": rows[:, 4],
                            "video_id": video_ids, "category_id": np.ones(n, dtype=int)},
                           index=pd.RangeIndex(self.id, self.id + n))
        self.id += n
        return out


# ---- ReID embeddings of the detections (DetectionLevelModule) ---------------------------------------------------------
class _DetectionBatches:
    def __init__(self, pipe, batch):
        self.pipe, self.batch = pipe, max(1, int(batch))

    def __iter__(self):
        ids = self.pipe.det_ids
        for i in range(0, len(ids), self.batch):
            yield ids[i:i + self.batch], {"detection_ids": ids[i:i + self.batch]}

    def __len__(self):
        return (len(self.pipe.det_ids) + self.batch - 1) // self.batch


class _ReidDatapipe:
    def __init__(self, module):
        self.module, self.det_ids = module, []

    def update(self, image_filepaths, img_metadatas, detections):
        self.det_ids = list(detections.index) if detections is not None else []
        self.module._embed_video(image_filepaths, img_metadatas, detections)

    def __len__(self):
        return len(self.det_ids)


class KPReId(DetectionLevelModule):
    """Drop-in for the ReID step of the pipeline (tracklab.wrappers.reid.kpreid_api.KPReId: DetectionLevelModule,
    input ``bbox_ltwh``, output ``embeddings`` [K, E] / ``visibility_scores`` [K], /root/reference/tracklab/wrappers/reid/
    kpreid_api.py:20-24,115-182). The crop rule is the wrapper's (``detection.bbox.ltrb(image_shape, rounded=True)``: sanitise in
    float32, round half to even, ``image[t:b, l:r]`` — tk_crop_resize_norm_ex / TK_CROP_RULE_LTWH_ROUNDED); the backbone is one of
    the architectures vendored in the reference (ResNet-50 -> K = 1, E = 2048, the configs[2] shape of SURVEY.md 8d; OSNet -> E =
    512), because KPR itself lives in the un-vendored torchreid fork. All detections of a video are embedded when the engine
    hands them to the datapipe (images decoded once, crops batched through the CUDA-graphed backbone); ``process`` returns the
    rows of its batch of detection ids."""
    input_columns = ["bbox_ltwh"]
    output_columns = ["embeddings", "visibility_scores"]
    collate_fn = None

    def __init__(self, cfg=None, device="cuda:0", save_path=None, training_enabled=False, batch_size=4096, job_id=0, *args, **kwargs):
        super().__init__(batch_size)
        if not torch.cuda.is_available():
            raise _lib.TrackKernError("KPReId needs a CUDA device: tracklab_b200 has no CPU path")
        from .reid import ReidStageDevice, build_reid_model
        self.cfg = cfg
        self.device = torch.device(device if str(device) != "cuda" else "cuda:0")
        arch = _cfg_get(cfg, "reid_arch", "resnet50") if cfg is not None else "resnet50"
        weights = _cfg_get(cfg, "model_weights", None) if cfg is not None else None
        model = None
        if weights is not None:
            if not os.path.isfile(str(weights)):
                raise _lib.TrackKernError(f"ReID weights {weights!r} not found")
            sd = torch.load(str(weights), map_location="cpu")
            model = build_reid_model(arch).from_reference_state_dict(sd.get("state_dict", sd))
        self.reid = ReidStageDevice(device=self.device, model=model, arch=arch,
                                    precision=(_cfg_get(cfg, "reid_precision", "bf16") if cfg is not None else "bf16"))
        self.decode_batch = int(_cfg_get(cfg, "decode_batch", 16)) if cfg is not None else 16
        self.frames_per_batch = int(batch_size)
        self._pipe = _ReidDatapipe(self)
        self._result = None

    @property
    def datapipe(self):
        return self._pipe

    def dataloader(self, engine=None):
        return _DetectionBatches(self._pipe, self.frames_per_batch)

    def preprocess(self, image, detection: pd.Series, metadata: pd.Series):   # never called: the datapipe is overridden
        return {}

    @torch.no_grad()
    def _embed_video(self, image_filepaths, img_metadatas, detections):
        if detections is None or len(detections) == 0:
            self._result = pd.DataFrame(columns=self.output_columns)
            return
        image_ids = np.asarray(img_metadatas.index)
        pos = {int(i): k for k, i in enumerate(image_ids)}
        frame = np.fromiter((pos[int(i)] for i in detections["image_id"].to_numpy()), dtype=np.int64, count=len(detections))
        order = np.argsort(frame, kind="stable")
        ltwh = np.stack(detections["bbox_ltwh"].to_numpy()).astype(np.float32).reshape(-1, 4)[order]
        rows = np.zeros((len(order), 7), dtype=np.float64)
        rows[:, :4] = ltwh
        counts = np.bincount(frame, minlength=len(image_ids))
        offsets = np.zeros(len(image_ids) + 1, dtype=np.int64)
        np.cumsum(counts, out=offsets[1:])
        paths = [image_filepaths[i] if image_filepaths and i in image_filepaths else img_metadatas.loc[i, "file_path"] for i in image_ids]
        d_dev = torch.from_numpy(rows).to(self.device)
        feats = torch.empty((len(rows), self.reid.feature_dim), dtype=torch.float32, device=self.device)
        for f0 in range(0, len(paths), self.decode_batch):
            f1 = min(len(paths), f0 + self.decode_batch)
            r0, r1 = int(offsets[f0]), int(offsets[f1])
            if r1 == r0:
                continue
            from .ingest import load_frames
            fr = load_frames(paths[f0:f1], self.device, getattr(self, "decode", "cv2"))
            det_frame = torch.from_numpy(np.repeat(np.arange(f1 - f0), counts[f0:f1]).astype(np.int32)).to(self.device)
            feats[r0:r1] = self.reid.features(fr, d_dev[r0:r1], det_frame, ltwh_rows=True)
        emb = feats.cpu().numpy()
        idx = np.asarray(detections.index)[order]
        self._result = pd.DataFrame({"embeddings": list(emb[:, None, :]), "visibility_scores": list(np.ones((len(emb), 1), dtype=np.float32))},
                                    index=pd.Index(idx))

    def process(self, batch, detections: pd.DataFrame, metadatas: pd.DataFrame):
        if self._result is None or len(detections) == 0:
            return []
        return self._result.loc[detections.index, ["embeddings", "visibility_scores"]]
