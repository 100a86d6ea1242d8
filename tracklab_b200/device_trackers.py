"""Whole-video device trackers: thin Python objects over the stateful C-ABI handles.

PyTorch is used only for device memory and streams; all association arithmetic runs inside
libtrackkern's kernels.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.TrackKernError(f"{name} must be a CUDA tensor (tracklab_b200 has no CPU path)")


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _VideoTrackerDevice:
    """Shared driver of the stateful tk_<name>_{create,reset,run,status,destroy} entry points."""

    _prefix = None

    def _create(self, params, n_seq, cap_tracks, cap_dets, device):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.TrackKernError("CUDA device required (tracklab_b200 has no CPU path)")
        self.device = torch.device(device)
        self.n_seq, self.cap_tracks, self.cap_dets = n_seq, cap_tracks, cap_dets
        self.params = params
        self.handle = ctypes.c_void_p()
        self._fn = {k: getattr(self.lib, f"tk_{self._prefix}_{k}") for k in ("create", "reset", "run", "status", "destroy")}
        with torch.cuda.device(self.device):
            _lib.check(self._fn["create"](ctypes.byref(params), n_seq, cap_tracks, cap_dets, ctypes.byref(self.handle)),
                       f"tk_{self._prefix}_create")

    def reset(self, keep_id_counter: bool = False):
        with torch.cuda.device(self.device):
            _lib.check(self._fn["reset"](self.handle, int(keep_id_counter), _stream_ptr()), f"tk_{self._prefix}_reset")

    def run(self, dets: torch.Tensor, offsets: torch.Tensor, out_rows: torch.Tensor | None = None,
            out_start: torch.Tensor | None = None, out_count: torch.Tensor | None = None):
        """dets float64[N,7] (device), offsets int32[n_seq, F+1] (device, absolute row indices).
        Returns (out_rows float64[.,8], out_frame_count int32[n_seq,F], out_count int32[n_seq]) — all
        device tensors, nothing is synchronised."""
        _require_cuda(dets, "dets"); _require_cuda(offsets, "offsets")
        assert dets.dtype == torch.float64 and dets.is_contiguous()
        assert offsets.dtype == torch.int32 and offsets.is_contiguous() and offsets.shape[0] == self.n_seq
        n_frames = offsets.shape[1] - 1
        if out_rows is None:
            out_rows = torch.empty((max(1, dets.shape[0]), 8), dtype=torch.float64, device=dets.device)
        if out_start is None:
            out_start = offsets[:, 0].contiguous()
        if out_count is None:
            out_count = torch.zeros(self.n_seq, dtype=torch.int32, device=dets.device)
        out_fc = torch.empty((self.n_seq, n_frames), dtype=torch.int32, device=dets.device)
        with torch.cuda.device(self.device):
            _lib.check(self._fn["run"](self.handle, dets.data_ptr(), offsets.data_ptr(), n_frames, out_rows.data_ptr(),
                                       out_start.data_ptr(), out_fc.data_ptr(), out_count.data_ptr(), _stream_ptr()),
                       f"tk_{self._prefix}_run")
        return out_rows, out_fc, out_count

    def status(self) -> np.ndarray:
        st = (ctypes.c_int * self.n_seq)()
        with torch.cuda.device(self.device):
            _lib.check(self._fn["status"](self.handle, st, _stream_ptr()), f"tk_{self._prefix}_status")
        return np.asarray(list(st), dtype=np.int32)

    def check_status(self):
        st = self.status()
        if (st != 0).any():
            raise _lib.TrackKernError("device tracker error: " + "; ".join(
                f"video {i}: {_lib.status_text(int(s))}" for i, s in enumerate(st) if s))

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self._fn["destroy"](self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ByteTrackDevice(_VideoTrackerDevice):
    """ByteTrack for ``n_seq`` independent videos (C ABI: tk_bytetrack_*).

    Mirrors BYTETracker(**hyperparams) + the wrapper's ``min_confidence`` filter
    (/root/reference/plugins/track/byte_track/byte_tracker.py:151-165,
    /root/reference/tracklab/wrappers/track/byte_track_api.py:50-56)."""

    _prefix = "bytetrack"

    def __init__(self, track_thresh=0.6, match_thresh=0.8, track_buffer=30, frame_rate=30,
                 min_confidence=0.4, first_id=1, n_seq=1, cap_tracks=128, cap_dets=128, device="cuda:0"):
        self._create(_lib.BytetrackParams(track_thresh, match_thresh, min_confidence, track_buffer, frame_rate, first_id),
                     n_seq, cap_tracks, cap_dets, device)


class OCSortDevice(_VideoTrackerDevice):
    """OC-SORT for ``n_seq`` independent videos (C ABI: tk_ocsort_*).

    Mirrors OCSort(**hyperparams) (/root/reference/plugins/track/oc_sort/ocsort.py:184-201) + the wrapper filter
    (/root/reference/tracklab/wrappers/track/oc_sort_api.py:50-56)."""

    _prefix = "ocsort"

    def __init__(self, det_thresh=0.0, max_age=50, min_hits=1, iou_threshold=0.22136877277096445, delta_t=1,
                 asso_func="giou", inertia=0.3941737016672115, use_byte=False, min_confidence=0.4,
                 n_seq=1, cap_tracks=128, cap_dets=128, device="cuda:0"):
        if asso_func not in _lib.ASSO_CODES:
            raise _lib.TrackKernError(f"asso_func {asso_func!r} not supported on device (iou/giou/diou/ciou/ct_dist)")
        self._create(_lib.OcsortParams(det_thresh, iou_threshold, inertia, min_confidence, max_age, min_hits, delta_t,
                                       _lib.ASSO_CODES[asso_func], int(bool(use_byte))), n_seq, cap_tracks, cap_dets, device)


class DeepOCSortDevice(_VideoTrackerDevice):
    """Deep OC-SORT for ``n_seq`` independent videos (C ABI: tk_deepocsort_*).

    Mirrors OCSort(model_weights, device, fp16, **hyperparams) of the deep_oc_sort plugin + the wrapper's ``min_confidence`` filter
    (/root/reference/plugins/track/deep_oc_sort/ocsort.py:324-372, /root/reference/tracklab/wrappers/track/deep_oc_sort_api.py:57-67)
    with the plugin's constructor defaults; the in-tracker ReID forward and the camera-motion estimator are separate stages whose
    outputs (embeddings float32 [N,E], affines float64 [n_seq,F,2,3]) are passed to ``run``."""

    _prefix = "deepocsort"

    def __init__(self, feature_dim, det_thresh=0.0, max_age=30, min_hits=3, iou_threshold=0.3, delta_t=3, asso_func="iou", inertia=0.2,
                 w_association_emb=0.75, alpha_fixed_emb=0.95, aw_param=0.5, embedding_off=False, cmc_off=False, aw_off=False,
                 new_kf_off=False, min_confidence=0.4, n_seq=1, cap_tracks=128, cap_dets=128, device="cuda:0"):
        if new_kf_off:
            raise _lib.TrackKernError("new_kf_off=True (the 7-d SORT filter inside Deep OC-SORT) is not built; the reference default is False")
        if asso_func not in _lib.ASSO_CODES:
            raise _lib.TrackKernError(f"asso_func {asso_func!r} not supported on device (iou/giou/diou/ciou/ct_dist)")
        self.feature_dim = int(feature_dim)
        self.embedding_off, self.cmc_off = bool(embedding_off), bool(cmc_off)
        self._create(_lib.DeepocsortParams(det_thresh, iou_threshold, inertia, min_confidence, w_association_emb, alpha_fixed_emb, aw_param,
                                           max_age, min_hits, delta_t, _lib.ASSO_CODES[asso_func], int(self.embedding_off),
                                           int(self.cmc_off), int(bool(aw_off)), self.feature_dim), n_seq, cap_tracks, cap_dets, device)

    def reset(self, keep_id_counter: bool = False):
        with torch.cuda.device(self.device):
            _lib.check(self._fn["reset"](self.handle, _stream_ptr()), "tk_deepocsort_reset")

    def run(self, dets: torch.Tensor, offsets: torch.Tensor, embeddings: torch.Tensor | None = None, affines: torch.Tensor | None = None,
            out_rows: torch.Tensor | None = None, out_start: torch.Tensor | None = None, out_count: torch.Tensor | None = None):
        """dets float64 [N,7], offsets int32 [n_seq,F+1], embeddings float32 [N,E], affines float64 [n_seq,F,2,3] (device).
        Returns (out_rows, out_frame_count, out_count), nothing synchronised."""
        _require_cuda(dets, "dets"); _require_cuda(offsets, "offsets")
        assert dets.dtype == torch.float64 and dets.is_contiguous()
        assert offsets.dtype == torch.int32 and offsets.is_contiguous() and offsets.shape[0] == self.n_seq
        n_frames = offsets.shape[1] - 1
        if not self.embedding_off:
            if embeddings is None:
                raise _lib.TrackKernError("DeepOCSortDevice.run: embeddings are required unless embedding_off")
            _require_cuda(embeddings, "embeddings")
            assert embeddings.dtype == torch.float32 and embeddings.is_contiguous() and embeddings.shape == (dets.shape[0], self.feature_dim)
        if not self.cmc_off:
            if affines is None:
                raise _lib.TrackKernError("DeepOCSortDevice.run: one 2x3 affine per frame is required unless cmc_off")
            _require_cuda(affines, "affines")
            assert affines.dtype == torch.float64 and affines.is_contiguous() and tuple(affines.shape) == (self.n_seq, n_frames, 2, 3)
        if out_rows is None:
            out_rows = torch.empty((max(1, dets.shape[0]), 8), dtype=torch.float64, device=dets.device)
        if out_start is None:
            out_start = offsets[:, 0].contiguous()
        if out_count is None:
            out_count = torch.zeros(self.n_seq, dtype=torch.int32, device=dets.device)
        out_fc = torch.empty((self.n_seq, n_frames), dtype=torch.int32, device=dets.device)
        with torch.cuda.device(self.device):
            _lib.check(self._fn["run"](self.handle, dets.data_ptr(), embeddings.data_ptr() if embeddings is not None else None,
                                       affines.data_ptr() if affines is not None else None, offsets.data_ptr(), n_frames,
                                       out_rows.data_ptr(), out_start.data_ptr(), out_fc.data_ptr(), out_count.data_ptr(),
                                       int(out_rows.shape[0]), _stream_ptr()), "tk_deepocsort_run")
        return out_rows, out_fc, out_count


class BotSortDevice(_VideoTrackerDevice):
    """BoT-SORT for ``n_seq`` independent videos (C ABI: tk_botsort_*).

    Mirrors BoTSORT(model_weights, device, fp16, **hyperparams) of the bot_sort plugin with its constructor defaults + the wrapper's
    ``min_confidence`` filter (/root/reference/plugins/track/bot_sort/bot_sort.py:243-273, /root/reference/tracklab/wrappers/track/
    bot_sort_api.py:57-66); the in-tracker ReID forward and the camera-motion estimator (``cmc_method``) are separate stages whose
    outputs (embeddings float32 [N,E], warps float64 [n_seq,F,2,3]; identity rows for cmc_method 'none') are passed to ``run``."""

    _prefix = "botsort"

    def __init__(self, feature_dim, track_high_thresh=0.45, new_track_thresh=0.6, track_buffer=30, match_thresh=0.8, proximity_thresh=0.5,
                 appearance_thresh=0.25, cmc_method="sparseOptFlow", frame_rate=30, lambda_=0.985, min_confidence=0.4, n_seq=1,
                 cap_tracks=128, cap_dets=128, device="cuda:0"):
        self.feature_dim = int(feature_dim)
        self.cmc_method = cmc_method
        self._create(_lib.BotsortParams(track_high_thresh, new_track_thresh, match_thresh, proximity_thresh, appearance_thresh, lambda_,
                                        min_confidence, int(track_buffer), int(frame_rate), self.feature_dim), n_seq, cap_tracks, cap_dets, device)

    def run(self, dets: torch.Tensor, offsets: torch.Tensor, embeddings: torch.Tensor, warps: torch.Tensor | None = None,
            out_rows: torch.Tensor | None = None, out_start: torch.Tensor | None = None, out_count: torch.Tensor | None = None):
        """dets float64 [N,7], offsets int32 [n_seq,F+1], embeddings float32 [N,E], warps float64 [n_seq,F,2,3] (None = identity)."""
        _require_cuda(dets, "dets"); _require_cuda(offsets, "offsets"); _require_cuda(embeddings, "embeddings")
        assert dets.dtype == torch.float64 and dets.is_contiguous()
        assert offsets.dtype == torch.int32 and offsets.is_contiguous() and offsets.shape[0] == self.n_seq
        assert embeddings.dtype == torch.float32 and embeddings.is_contiguous() and tuple(embeddings.shape) == (dets.shape[0], self.feature_dim)
        n_frames = offsets.shape[1] - 1
        if warps is None:
            warps = torch.eye(2, 3, dtype=torch.float64, device=dets.device).repeat(self.n_seq, n_frames, 1, 1).contiguous()
        _require_cuda(warps, "warps")
        assert warps.dtype == torch.float64 and warps.is_contiguous() and tuple(warps.shape) == (self.n_seq, n_frames, 2, 3)
        if out_rows is None:
            out_rows = torch.empty((max(1, dets.shape[0]), 8), dtype=torch.float64, device=dets.device)
        if out_start is None:
            out_start = offsets[:, 0].contiguous()
        if out_count is None:
            out_count = torch.zeros(self.n_seq, dtype=torch.int32, device=dets.device)
        out_fc = torch.empty((self.n_seq, n_frames), dtype=torch.int32, device=dets.device)
        with torch.cuda.device(self.device):
            _lib.check(self._fn["run"](self.handle, dets.data_ptr(), embeddings.data_ptr(), warps.data_ptr(), offsets.data_ptr(), n_frames,
                                       out_rows.data_ptr(), out_start.data_ptr(), out_fc.data_ptr(), out_count.data_ptr(), _stream_ptr()),
                       "tk_botsort_run")
        return out_rows, out_fc, out_count


class StrongSortDevice(_VideoTrackerDevice):
    """StrongSORT association for ``n_seq`` videos with externally supplied ReID features (C ABI: tk_strongsort_*).

    Mirrors StrongSORT(**hyperparams).update minus the in-tracker ReID forward and ECC
    (/root/reference/plugins/track/strong_sort/strong_sort.py:23-85, /root/reference/tracklab/wrappers/track/strong_sort_api.py:66-93)."""

    _prefix = "strongsort"

    def __init__(self, feature_dim, max_dist=0.1594374041012136, max_iou_dist=0.5431835667667874, max_age=40,
                 max_unmatched_preds=0, n_init=3, nn_budget=100, mc_lambda=0.995, ema_alpha=0.8962157769329083,
                 min_confidence=0.4, image_size=(1920, 1080), ctas_per_video=8, n_seq=1, cap_tracks=128, cap_dets=128,
                 device="cuda:0"):
        self.feature_dim = feature_dim
        self._create(_lib.StrongsortParams(max_dist, max_iou_dist, mc_lambda, ema_alpha, min_confidence, max_age, n_init,
                                           nn_budget, max_unmatched_preds, feature_dim, image_size[0], image_size[1],
                                           ctas_per_video), n_seq, cap_tracks, cap_dets, device)

    def run(self, dets: torch.Tensor, offsets: torch.Tensor, features: torch.Tensor, out_rows: torch.Tensor | None = None,
            out_start: torch.Tensor | None = None, out_count: torch.Tensor | None = None, warps: torch.Tensor | None = None):
        """features float32 [N, E] aligned with dets rows. Output capacity: 2 rows per detection per video by default.
        warps float32 [n_seq, F, 6] (optional): per-frame camera motion (kernels.ecc_euclidean), applied to every track before
        the frame is processed (cfg.ecc of the reference wrapper, C ABI: tk_strongsort_run_cmc)."""
        _require_cuda(dets, "dets"); _require_cuda(offsets, "offsets"); _require_cuda(features, "features")
        assert dets.dtype == torch.float64 and dets.is_contiguous() and features.dtype == torch.float32 and features.is_contiguous()
        assert features.shape == (dets.shape[0], self.feature_dim)
        assert offsets.dtype == torch.int32 and offsets.is_contiguous() and offsets.shape[0] == self.n_seq
        n_frames = offsets.shape[1] - 1
        cap_rows = 2 * max(1, dets.shape[0])
        if out_rows is None:
            assert self.n_seq == 1, "pass out_rows/out_start for several videos"
            out_rows = torch.empty((cap_rows, 8), dtype=torch.float64, device=dets.device)
        else:
            cap_rows = out_rows.shape[0] // self.n_seq
        if out_start is None:
            out_start = torch.zeros(self.n_seq, dtype=torch.int32, device=dets.device)
        if out_count is None:
            out_count = torch.zeros(self.n_seq, dtype=torch.int32, device=dets.device)
        out_fc = torch.empty((self.n_seq, n_frames), dtype=torch.int32, device=dets.device)
        wp = None
        if warps is not None:
            _require_cuda(warps, "warps")
            assert warps.dtype == torch.float32 and warps.is_contiguous() and warps.numel() == self.n_seq * n_frames * 6
            wp = warps.data_ptr()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tk_strongsort_run_cmc(self.handle, dets.data_ptr(), features.data_ptr(), offsets.data_ptr(), n_frames, wp,
                                                      out_rows.data_ptr(), out_start.data_ptr(), out_fc.data_ptr(), out_count.data_ptr(),
                                                      cap_rows, _stream_ptr()), "tk_strongsort_run_cmc")
        return out_rows, out_fc, out_count


class BpbreidStrongSortDevice(_VideoTrackerDevice):
    """BPBReID-StrongSORT association for ``n_seq`` videos (C ABI: tk_bpbreid_*): part-based features + visibility scores.

    Mirrors bpbreid_strong_sort.StrongSORT(**cfg).update for ``matching_strategy="strong_sort_matching"``,
    ``motion_criterium="iou"`` (/root/reference/plugins/track/bpbreid_strong_sort/strong_sort.py:11-141,
    /root/reference/tracklab/configs/modules/track/bpbreid_strong_sort.yaml)."""

    _prefix = "bpbreid"
    COLS = 14

    def __init__(self, n_parts, feature_dim, ema_alpha=0.9, mc_lambda=0.995, max_dist=0.5, max_iou_distance=0.8, max_age=300,
                 n_init=0, min_bbox_confidence=0.0, max_kalman_prediction_without_update=7, ctas_per_video=8, n_seq=1,
                 cap_tracks=1024, cap_dets=128, device="cuda:0", matching_strategy="strong_sort_matching", gating_thres_factor=1.0,
                 w_kfgd=1.0, w_reid=1.0, w_st=1.0):
        self.n_parts, self.feature_dim = n_parts, feature_dim
        if matching_strategy not in ("strong_sort_matching", "bot_sort_matching"):
            raise _lib.TrackKernError(f"matching_strategy {matching_strategy!r} unknown (strong_sort_matching / bot_sort_matching)")
        if matching_strategy == "bot_sort_matching" and not w_kfgd > 0:
            raise _lib.TrackKernError("bot_sort_matching on device needs w_kfgd > 0 (the Kalman position gate prunes the pair list)")
        self._create(_lib.BpbreidParams(max_dist, max_iou_distance, mc_lambda, ema_alpha, min_bbox_confidence, max_age, n_init,
                                        max_kalman_prediction_without_update, n_parts, feature_dim, ctas_per_video,
                                        int(matching_strategy == "bot_sort_matching"), gating_thres_factor, w_kfgd, w_reid, w_st),
                     n_seq, cap_tracks, cap_dets, device)

    def run(self, dets: torch.Tensor, offsets: torch.Tensor, features: torch.Tensor, visibility: torch.Tensor,
            out_rows: torch.Tensor | None = None, out_start: torch.Tensor | None = None, out_count: torch.Tensor | None = None):
        """dets float64 [N,7] = [l,t,w,h,conf,cls,det id]; features float32 [N,K,E]; visibility float32 [N,K].
        At most one row per detection is produced."""
        for t, n in ((dets, "dets"), (offsets, "offsets"), (features, "features"), (visibility, "visibility")):
            _require_cuda(t, n)
            assert t.is_contiguous(), n
        assert dets.dtype == torch.float64 and features.dtype == torch.float32 and visibility.dtype == torch.float32
        assert features.shape == (dets.shape[0], self.n_parts, self.feature_dim) and visibility.shape == (dets.shape[0], self.n_parts)
        assert offsets.dtype == torch.int32 and offsets.shape[0] == self.n_seq
        n_frames = offsets.shape[1] - 1
        cap_rows = max(1, dets.shape[0])
        if out_rows is None:
            assert self.n_seq == 1, "pass out_rows/out_start for several videos"
            out_rows = torch.empty((cap_rows, self.COLS), dtype=torch.float64, device=dets.device)
        else:
            cap_rows = out_rows.shape[0] // self.n_seq
        if out_start is None:
            out_start = torch.zeros(self.n_seq, dtype=torch.int32, device=dets.device)
        if out_count is None:
            out_count = torch.zeros(self.n_seq, dtype=torch.int32, device=dets.device)
        out_fc = torch.empty((self.n_seq, n_frames), dtype=torch.int32, device=dets.device)
        with torch.cuda.device(self.device):
            _lib.check(self._fn["run"](self.handle, dets.data_ptr(), features.data_ptr(), visibility.data_ptr(), offsets.data_ptr(),
                                       n_frames, out_rows.data_ptr(), out_start.data_ptr(), out_fc.data_ptr(), out_count.data_ptr(),
                                       cap_rows, _stream_ptr()), "tk_bpbreid_run")
        return out_rows, out_fc, out_count


def rows_to_frames(out_rows: torch.Tensor, out_fc: torch.Tensor, out_start: torch.Tensor, seq: int = 0):
    """Host helper: split the rows of video ``seq`` per frame -> (rows float64[R,8], frame int32[R])."""
    fc = out_fc[seq].cpu().numpy()
    start = int(out_start[seq].item())
    total = int(fc.sum())
    rows = out_rows[start:start + total].cpu().numpy()
    return rows, np.repeat(np.arange(len(fc), dtype=np.int32), fc)
