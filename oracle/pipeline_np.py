"""CPU restatement of the detect -> associate loop for one video (test infrastructure / CPU baseline).

Per frame, like the reference's per-image module calls
(/root/reference/tracklab/engine/offline.py:20-35 -> /root/reference/tracklab/wrappers/bbox_detector/rtmlib_api.py:27-46
 -> /root/reference/tracklab/wrappers/track/byte_track_api.py:50-76): letterbox (cv2), detector forward at
batch 1 (rtmlib_api.py:19: batch_size=1) on CPU threads in float32, decode + NMS (NumPy), wrapper rows, then the
tracker oracle. The detector network is the same PyTorch module the GPU path uses (the reference's ONNX
graphs/onnxruntime are un-vendored — BASELINE.md §3), run here in fp32 on the host.
"""
import numpy as np
import torch

from .bytetrack_np import ByteTrackOracle
from .preprocess_np import letterbox_yolox
from .yolox_post_np import wrapper_rows, yolox_postprocess


@torch.no_grad()
def detect_frame(model_cpu, frame_rgb, size=640, score_thr=0.7, nms_thr=0.45, first_id=0):
    h, w = frame_rgb.shape[:2]
    bgr = np.ascontiguousarray(frame_rgb[..., ::-1])       # cv2.imread order (rtmlib_api.py:28)
    x, ratio = letterbox_yolox(bgr, size)
    raw = model_cpu(torch.from_numpy(x)[None]).numpy()[0].astype(np.float32)
    raw[:, 4:] = 1.0 / (1.0 + np.exp(-raw[:, 4:]))        # YOLOX head sigmoid on obj/cls
    boxes, scores, cls = yolox_postprocess(raw, np.float32(ratio), size, score_thr, nms_thr)
    boxes = boxes[cls == 0]
    return wrapper_rows(boxes, w, h, first_id)


def detect_track_video(model_cpu, frames_rgb, tracker_dets=None, tracker_offsets=None, hyper=None, min_conf=0.4):
    """frames uint8 [F,H,W,3]; returns (tracker rows [R,8], frame index [R], detector rows per frame)."""
    hyper = hyper or dict(track_thresh=0.6, match_thresh=0.8, track_buffer=30, frame_rate=30)
    trk = ByteTrackOracle(**hyper, min_confidence=min_conf)
    out, fr, det_rows = [], [], []
    next_id = 0
    for f in range(len(frames_rgb)):
        rows = detect_frame(model_cpu, frames_rgb[f], first_id=next_id)
        next_id += len(rows)
        det_rows.append(rows)
        d = rows if tracker_dets is None else tracker_dets[tracker_offsets[f]:tracker_offsets[f + 1]]
        if len(d) == 0:
            continue
        r = trk.update(d[d[:, 4] > min_conf])
        out.append(r)
        fr.append(np.full(len(r), f, dtype=np.int32))
    rows = np.concatenate(out) if out else np.zeros((0, 8))
    frames = np.concatenate(fr) if fr else np.zeros((0,), dtype=np.int32)
    return rows, frames, det_rows


def kpreid_crop_box(ltwh, width, height):
    """ReID wrapper crop rule (/root/reference/tracklab/wrappers/reid/kpreid_api.py:118-121 -> utils/__init__.py:47-48 ->
    utils/coordinates.py:216-267): float32 bbox_ltwh, sanitize_bbox_ltwh in place, ltrb, round half-to-even."""
    b = np.asarray(ltwh, dtype=np.float32).copy()
    b[0] = max(0, min(b[0], width - 2))
    b[1] = max(0, min(b[1], height - 2))
    b[2] = max(1, min(b[2], width - 1 - b[0]))
    b[3] = max(1, min(b[3], height - 1 - b[1]))
    l, t, r, bb = np.array([b[0], b[1], b[0] + b[2], b[1] + b[3]]).round().astype(int)
    return int(l), int(t), int(r), int(bb)


@torch.no_grad()
def reid_features_frame(reid_model_cpu, frame_rgb, rows, batch=64):
    """Per-frame in-tracker ReID of the StrongSORT plugin (/root/reference/plugins/track/strong_sort/strong_sort.py:135-145,
    reid_multibackend.py:184-237): PIL crops -> float32 network -> float32 features [D,E]."""
    from .preprocess_np import reid_crops
    if len(rows) == 0:
        return np.zeros((0, reid_model_cpu.feature_dim), dtype=np.float32)
    x = torch.from_numpy(reid_crops(frame_rgb, rows[:, :4]))
    return torch.cat([reid_model_cpu(x[i:i + batch]) for i in range(0, len(x), batch)]).numpy().astype(np.float32)


def detect_reid_track_video(det_model_cpu, reid_model_cpu, frames_rgb, hyper, min_conf=0.4, detector_rows=None, score_thr=0.7):
    """CPU restatement of the connected detect -> ReID -> associate loop for one video, frame by frame like the reference
    (rtmlib_api.py:27-46 -> strong_sort_api.py:43-93 -> strong_sort.py:41-85). ``detector_rows`` (list of [D,7] per frame)
    replaces the CPU detector forward — used to check the stages downstream of given rows.
    Returns (tracker rows [R,8], frame [R], list of detector rows per frame, list of features per frame)."""
    from .strongsort_np import StrongSortOracle
    H, W = frames_rgb.shape[1:3]
    trk = StrongSortOracle(**hyper, min_confidence=min_conf, image_size=(W, H))
    out, fr, det_rows, feats = [], [], [], []
    next_id = 0
    for f in range(len(frames_rgb)):
        if detector_rows is None:
            rows = detect_frame(det_model_cpu, frames_rgb[f], first_id=next_id, score_thr=score_thr)
        else:
            rows = detector_rows[f]
        next_id += len(rows)
        det_rows.append(rows)
        if len(rows) == 0:
            feats.append(np.zeros((0, reid_model_cpu.feature_dim), np.float32))
            continue
        keep = rows[:, 4] > min_conf               # strong_sort_api.py:54-56
        e = reid_features_frame(reid_model_cpu, frames_rgb[f], rows[keep])
        feats.append(e)
        r = trk.update(rows[keep], e)
        out.append(r)
        fr.append(np.full(len(r), f, dtype=np.int32))
    rows = np.concatenate(out) if out else np.zeros((0, 8))
    frames = np.concatenate(fr) if fr else np.zeros((0,), dtype=np.int32)
    return rows, frames, det_rows, feats
