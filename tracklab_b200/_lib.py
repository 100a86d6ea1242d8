"""ctypes binding of libtrackkern.so (the C ABI declared in include/trackkern.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C tracklab_b200/csrc``.
A missing library is a hard error — there is no Python or CPU fallback for the product path.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtrackkern.so")

TK_ERRORS = {-1: "TK_ERR_ARG", -2: "TK_ERR_CUDA", -3: "TK_ERR_CAPACITY", -4: "TK_ERR_INFEASIBLE"}
DEV_STATUS = {1: "track capacity overflow", 2: "detections-per-frame capacity overflow",
              4: "assignment infeasible", 8: "output capacity overflow", 16: "non-PD innovation covariance",
              32: "assignment-side capacity overflow", 64: "undefined appearance cost (no commonly visible part)"}


class TrackKernError(RuntimeError):
    pass


class BytetrackParams(ctypes.Structure):
    _fields_ = [("track_thresh", ctypes.c_double), ("match_thresh", ctypes.c_double),
                ("min_confidence", ctypes.c_double), ("track_buffer", ctypes.c_int),
                ("frame_rate", ctypes.c_int), ("first_id", ctypes.c_int)]


class OcsortParams(ctypes.Structure):
    _fields_ = [("det_thresh", ctypes.c_double), ("iou_threshold", ctypes.c_double), ("inertia", ctypes.c_double),
                ("min_confidence", ctypes.c_double), ("max_age", ctypes.c_int), ("min_hits", ctypes.c_int),
                ("delta_t", ctypes.c_int), ("asso_func", ctypes.c_int), ("use_byte", ctypes.c_int)]


class StrongsortParams(ctypes.Structure):
    _fields_ = [("max_dist", ctypes.c_double), ("max_iou_dist", ctypes.c_double), ("mc_lambda", ctypes.c_double),
                ("ema_alpha", ctypes.c_double), ("min_confidence", ctypes.c_double), ("max_age", ctypes.c_int),
                ("n_init", ctypes.c_int), ("nn_budget", ctypes.c_int), ("max_unmatched_preds", ctypes.c_int),
                ("feature_dim", ctypes.c_int), ("image_width", ctypes.c_int), ("image_height", ctypes.c_int),
                ("ctas_per_video", ctypes.c_int)]


class BpbreidParams(ctypes.Structure):
    _fields_ = [("max_dist", ctypes.c_double), ("max_iou_distance", ctypes.c_double), ("mc_lambda", ctypes.c_double),
                ("ema_alpha", ctypes.c_double), ("min_bbox_confidence", ctypes.c_double), ("max_age", ctypes.c_int),
                ("n_init", ctypes.c_int), ("max_kalman_prediction_without_update", ctypes.c_int), ("n_parts", ctypes.c_int),
                ("feature_dim", ctypes.c_int), ("ctas_per_video", ctypes.c_int), ("matching_strategy", ctypes.c_int),
                ("gating_thres_factor", ctypes.c_double), ("w_kfgd", ctypes.c_double), ("w_reid", ctypes.c_double), ("w_st", ctypes.c_double)]


class DeepocsortParams(ctypes.Structure):
    _fields_ = [("det_thresh", ctypes.c_double), ("iou_threshold", ctypes.c_double), ("inertia", ctypes.c_double),
                ("min_confidence", ctypes.c_double), ("w_association_emb", ctypes.c_double), ("alpha_fixed_emb", ctypes.c_double),
                ("aw_param", ctypes.c_double), ("max_age", ctypes.c_int), ("min_hits", ctypes.c_int), ("delta_t", ctypes.c_int),
                ("asso_func", ctypes.c_int), ("embedding_off", ctypes.c_int), ("cmc_off", ctypes.c_int), ("aw_off", ctypes.c_int),
                ("feature_dim", ctypes.c_int)]


class BotsortParams(ctypes.Structure):
    _fields_ = [("track_high_thresh", ctypes.c_double), ("new_track_thresh", ctypes.c_double), ("match_thresh", ctypes.c_double),
                ("proximity_thresh", ctypes.c_double), ("appearance_thresh", ctypes.c_double), ("lambda_", ctypes.c_double),
                ("min_confidence", ctypes.c_double), ("track_buffer", ctypes.c_int), ("frame_rate", ctypes.c_int),
                ("feature_dim", ctypes.c_int)]


ASSO_CODES = {"iou": 0, "giou": 1, "diou": 2, "ciou": 3, "ct_dist": 4}

_lib = None


def _declare(lib):
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    P = ctypes.POINTER
    sig = {
        "tk_abi_version": ([], ci),
        "tk_last_cuda_error": ([], ci),
        "tk_letterbox_u8": ([vp, ci, ci, ci, ctypes.c_longlong, vp, ci, ci, ci, ci, ci, P(cd), vp], ci),
        "tk_crop_resize_norm": ([vp, ci, ci, ctypes.c_longlong, vp, vp, ci, vp, ci, ci, ci, ci, P(ctypes.c_float), P(ctypes.c_float), vp], ci),
        "tk_crop_resize_norm_ex": ([vp, ci, ci, ctypes.c_longlong, vp, vp, ci, vp, ci, ci, ci, ci, P(ctypes.c_float), P(ctypes.c_float), ci, vp], ci),
        "tk_yolox_nms": ([vp, ci, ci, ci, ci, ci, ci, ctypes.c_float, ctypes.c_float, ctypes.c_float, ci, vp, vp, vp, vp, vp, vp], ci),
        "tk_pack_detections": ([vp, vp, vp, vp, ci, ci, ci, ci, ci, cd, cd, vp, vp, vp, ci, ci, vp, vp], ci),
        "tk_pack_detections_ex": ([vp, vp, vp, vp, ci, ci, ci, ci, ci, cd, cd, vp, vp, vp, ci, ci, vp, ci, vp, vp], ci),
        "tk_bias_act_nhwc": ([vp, vp, vp, vp, ctypes.c_longlong, ci, ci, ci, ci, ci, ci, vp], ci),
        "tk_conv1x1_bias_act_bf16": ([vp, ctypes.c_longlong, ci, ci, vp, ci, vp, vp, ci, ci, vp, ci, ci, ci, vp], ci),
        "tk_conv3x3_bias_act_bf16": ([vp, ci, ci, ci, ci, vp, ci, vp, vp, ci, ci, vp, ci, ci, ci, vp], ci),
        "tk_spp_nhwc": ([vp, vp, ci, ci, ci, ci, ci, ci, vp], ci),
        "tk_upsample2x_nhwc": ([vp, ci, ci, vp, ci, ci, ci, ci, ci, ci, vp], ci),
        "tk_resize_frames_u8": ([vp, ci, ci, ci, ctypes.c_longlong, vp, ci, ci, ci, ctypes.c_float, vp], ci),
        "tk_rtdetr_decode": ([vp, vp, ci, ci, ci, ci, ci, ctypes.c_float, ci, vp, vp, vp], ci),
        "tk_maxpool3x3s2_nhwc": ([vp, ci, ci, ci, ci, vp, vp], ci),
        "tk_avgpool_nhwc": ([vp, ci, ci, ci, vp, vp], ci),
        "tk_kf_predict": ([vp, vp, ci, ci, vp], ci),
        "tk_kf_update": ([vp, vp, vp, ci, ci, vp, vp], ci),
        "tk_vdc_cost": ([vp, vp, vp, vp, ci, ci, cd, ci, vp], ci),
        "tk_iou_matrix": ([vp, vp, vp, ci, ci, ci, ci, vp], ci),
        "tk_iou_p1_f32": ([vp, vp, vp, ci, ci, ci, vp], ci),
        "tk_cosine_dist": ([vp, vp, vp, vp, ci, ci, ci, ci, vp], ci),
        "tk_lap_batched": ([vp, ci, ci, ci, cd, ci, vp, vp, vp, vp], ci),
        "tk_lsap_scipy_batched": ([vp, ci, ci, ci, vp, vp, vp, vp], ci),
        "tk_hota_workspace_bytes": ([ci, ci, ci, ci, ctypes.c_longlong, P(ctypes.c_longlong)], ci),
        "tk_hota_sequence": ([vp, vp, vp, ctypes.c_longlong, vp, vp, vp, ctypes.c_longlong, ci, ci, ci, ci, ci, P(cd), ci,
                             ctypes.c_longlong, vp, ctypes.c_longlong, vp, vp, vp], ci),
        "tk_part_dist": ([vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp], ci),
        "tk_kf_gate": ([vp, vp, vp, vp, ci, ci, ci, vp, vp], ci),
        "tk_bytetrack_create": ([P(BytetrackParams), ci, ci, ci, P(vp)], ci),
        "tk_bytetrack_reset": ([vp, ci, vp], ci),
        "tk_bytetrack_run": ([vp, vp, vp, ci, vp, vp, vp, vp, vp], ci),
        "tk_bytetrack_status": ([vp, P(ci), vp], ci),
        "tk_bytetrack_destroy": ([vp], ci),
        "tk_ocsort_create": ([P(OcsortParams), ci, ci, ci, P(vp)], ci),
        "tk_ocsort_reset": ([vp, ci, vp], ci),
        "tk_ocsort_run": ([vp, vp, vp, ci, vp, vp, vp, vp, vp], ci),
        "tk_ocsort_status": ([vp, P(ci), vp], ci),
        "tk_ocsort_destroy": ([vp], ci),
        "tk_botsort_create": ([P(BotsortParams), ci, ci, ci, P(vp)], ci),
        "tk_botsort_reset": ([vp, ci, vp], ci),
        "tk_botsort_run": ([vp, vp, vp, vp, vp, ci, vp, vp, vp, vp, vp], ci),
        "tk_botsort_status": ([vp, P(ci), vp], ci),
        "tk_botsort_destroy": ([vp], ci),
        "tk_deepocsort_create": ([P(DeepocsortParams), ci, ci, ci, P(vp)], ci),
        "tk_deepocsort_reset": ([vp, vp], ci),
        "tk_deepocsort_run": ([vp, vp, vp, vp, vp, ci, vp, vp, vp, vp, ci, vp], ci),
        "tk_deepocsort_status": ([vp, P(ci), vp], ci),
        "tk_deepocsort_destroy": ([vp], ci),
        "tk_strongsort_create": ([P(StrongsortParams), ci, ci, ci, P(vp)], ci),
        "tk_strongsort_reset": ([vp, ci, vp], ci),
        "tk_strongsort_run": ([vp, vp, vp, vp, ci, vp, vp, vp, vp, ci, vp], ci),
        "tk_strongsort_run_cmc": ([vp, vp, vp, vp, ci, vp, vp, vp, vp, vp, ci, vp], ci),
        "tk_ecc_small_size": ([ci, ci, cd, P(ci), P(ci)], ci),
        "tk_ecc_gray_small": ([vp, ci, ci, ci, ctypes.c_longlong, cd, vp, vp], ci),
        "tk_ecc_euclidean": ([vp, ci, ci, ci, ci, cd, cd, vp, vp, vp, vp], ci),
        "tk_strongsort_status": ([vp, P(ci), vp], ci),
        "tk_strongsort_destroy": ([vp], ci),
        "tk_bpbreid_create": ([P(BpbreidParams), ci, ci, ci, P(vp)], ci),
        "tk_bpbreid_reset": ([vp, ci, vp], ci),
        "tk_bpbreid_run": ([vp, vp, vp, vp, vp, ci, vp, vp, vp, vp, ci, vp], ci),
        "tk_bpbreid_status": ([vp, P(ci), vp], ci),
        "tk_bpbreid_destroy": ([vp], ci),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    return sig


def load():
    """Load (once) and return the ctypes handle. Raises TrackKernError when the library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TrackKernError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). tracklab_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        _declare(lib)
        _lib = lib
    return _lib


def exported_symbols():
    """Names include/trackkern.h declares (used by the not-gpu ABI test)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "trackkern.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"^\s*int\s+(tk_[a-z0-9_]+)\s*\(", txt, flags=re.M)))


def check(code: int, what: str):
    if code != 0:
        lib = load()
        raise TrackKernError(f"{what} failed: {TK_ERRORS.get(code, code)} (last CUDA error {lib.tk_last_cuda_error()})")


def status_text(bits: int) -> str:
    return ", ".join(v for k, v in DEV_STATUS.items() if bits & k) or "ok"
