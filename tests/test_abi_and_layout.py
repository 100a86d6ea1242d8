"""CPU suite: the C-ABI library loads and exports every symbol include/trackkern.h declares (no compute calls
without a GPU); the product package never imports the oracle; the product path fails loudly without CUDA."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from tracklab_b200 import _lib
    lib = _lib.load()
    names = _lib.exported_symbols()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.tk_abi_version() == 1


def test_argument_validation_without_gpu():
    """Entry points reject bad arguments before touching the device."""
    import ctypes

    from tracklab_b200 import _lib
    lib = _lib.load()
    assert lib.tk_bytetrack_create(None, 1, 64, 64, None) == -1
    h = ctypes.c_void_p()
    p = _lib.BytetrackParams(0.6, 0.8, 0.4, 30, 30, 1)
    assert lib.tk_bytetrack_create(ctypes.byref(p), 1, 4096, 64, ctypes.byref(h)) == -3   # capacity
    assert lib.tk_letterbox_u8(None, 1, 10, 10, 300, None, 0, 0, 640, 114, 0, None, None) == -1
    assert lib.tk_yolox_nms(None, 0, 1, 8400, 1, 640, 1, 1.0, 0.7, 0.45, 10, None, None, None, None, None, None) == -1


def test_product_package_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tracklab_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, f"tracklab_b200 must not import the oracle: {bad}"


def test_product_path_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from tracklab_b200 import _lib
    from tracklab_b200.device_trackers import ByteTrackDevice
    with pytest.raises(_lib.TrackKernError):
        ByteTrackDevice()
    import types

    from tracklab_b200 import modules
    with pytest.raises(_lib.TrackKernError):
        modules.ByteTrack(types.SimpleNamespace(min_confidence=0.4, hyperparams={}), "cuda")
    with pytest.raises(_lib.TrackKernError):
        from tracklab_b200 import kernels
        kernels.letterbox(torch.zeros((1, 8, 8, 3), dtype=torch.uint8))


def test_module_api_mirror_contract():
    """Level inference from the first base-class name and the column protocol (pipeline/module.py:34-61,69-84)."""
    from tracklab_b200 import modules
    from tracklab_b200.pipeline import Pipeline
    assert modules.ByteTrack.__bases__[0].__name__ == "ImageLevelModule"     # => level == "image" (q14)
    assert modules.OCSORT.__bases__[0].__name__ == "ImageLevelModule"
    assert modules.ByteTrack.input_columns == ["bbox_ltwh", "bbox_conf", "category_id"]
    assert modules.OCSORT.output_columns == ["track_id", "track_bbox_ltwh", "track_bbox_conf"]
    assert callable(Pipeline.validate)


def test_every_device_stage_and_module_refuses_to_run_without_cuda():
    import types

    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from tracklab_b200 import _lib, modules
    from tracklab_b200.device_trackers import BotSortDevice, BpbreidStrongSortDevice, DeepOCSortDevice, OCSortDevice, StrongSortDevice
    from tracklab_b200.detector import YoloxDetectorDevice
    from tracklab_b200.reid import ReidStageDevice
    from tracklab_b200.rtdetr_detector import RTDetrDetectorDevice
    for ctor in (lambda: OCSortDevice(), lambda: StrongSortDevice(64), lambda: DeepOCSortDevice(64), lambda: BotSortDevice(64),
                 lambda: modules.BotSORT(types.SimpleNamespace(hyperparams={}), "cuda"),
                 lambda: modules.DeepOCSORT(types.SimpleNamespace(hyperparams={}), "cuda"), lambda: BpbreidStrongSortDevice(4, 32), lambda: YoloxDetectorDevice("s"),
                 lambda: ReidStageDevice(), lambda: ReidStageDevice(arch="osnet_x1_0"), lambda: RTDetrDetectorDevice(model=object()),
                 lambda: modules.StrongSORT(types.SimpleNamespace(hyperparams={}, ecc=False), "cuda"),
                 lambda: modules.BPBReIDStrongSORT(types.SimpleNamespace(ecc=False), "cuda"),
                 lambda: modules.RTDetr("cuda", 8, "rtdetr_r50vd_coco_o365", 0.4)):
        with pytest.raises(_lib.TrackKernError):
            ctor()


def test_new_modules_follow_the_reference_column_contracts():
    from tracklab_b200 import modules
    assert modules.DeepOCSORT.input_columns == ["bbox_ltwh", "bbox_conf", "category_id"]                        # deep_oc_sort_api.py:17-22
    assert modules.DeepOCSORT.output_columns == ["track_id", "track_bbox_ltwh", "track_bbox_conf"]
    assert modules.BotSORT.input_columns == ["bbox_ltwh", "bbox_conf", "category_id"]                           # bot_sort_api.py:19-24
    assert modules.BotSORT.output_columns == ["track_id", "track_bbox_ltwh", "track_bbox_conf"]
    for cls in (modules.StrongSORT, modules.BPBReIDStrongSORT, modules.RTDetr, modules.DeepOCSORT, modules.BotSORT):
        assert cls.__bases__[0].__name__ == "ImageLevelModule"
    assert modules.BPBReIDStrongSORT.input_columns == ["bbox_ltwh", "embeddings", "visibility_scores"]          # bpbreid_strong_sort_api.py:15-19
    assert modules.BPBReIDStrongSORT.output_columns == ["track_id", "track_bbox_kf_ltwh", "track_bbox_pred_kf_ltwh", "matched_with", "costs",
                                                        "hits", "age", "time_since_update", "state"]                 # :20-30
    assert modules.RTDetr.input_columns == [] and modules.RTDetr.output_columns == ["image_id", "video_id", "category_id", "bbox_ltwh", "bbox_conf"]


def test_jpeg_ingest_library_loads_and_exports_the_declared_symbols():
    """libtkjpeg.so (include/tkjpeg.h, nvJPEG frame ingest): loads in the build container and exports every declared entry point."""
    import ctypes
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = sorted(set(re.findall(r"^\s*int\s+(tk_jpeg_[a-z0-9_]+)\s*\(", open(os.path.join(root, "include", "tkjpeg.h")).read(), flags=re.M)))
    assert len(names) == 6
    from tracklab_b200 import ingest
    lib = ingest._load()
    for n in names:
        assert hasattr(lib, n), n


def test_skipped_frame_affines_are_folded_into_the_next_processed_frame():
    """modules.compose_skipped_affines: the plugins' camera-motion estimators are only called on frames the wrapper processes and relate
    each call to the previous call (deep_oc_sort/cmc.py:138-166, bot_sort/gmc.py:239-303), so pair-wise device transforms of skipped
    frames are composed; the first processed frame gets the identity; failed pairs (NaN) count as identity."""
    import numpy as np
    from tracklab_b200.modules import compose_skipped_affines
    def T(tx, ty, th=0.0):
        return np.array([[np.cos(th), -np.sin(th), tx], [np.sin(th), np.cos(th), ty]])
    w = np.full((6, 2, 3), np.nan)
    w[1], w[2], w[3], w[5] = T(2, 0), T(3, 1, 0.01), T(1, 0), T(5, 5)            # w[4] failed
    out = compose_skipped_affines(w, np.array([True, True, False, True, True, True]))
    assert np.array_equal(out[0], np.eye(2, 3)) and np.array_equal(out[1], T(2, 0)) and np.array_equal(out[2], np.eye(2, 3))
    H = lambda a: np.vstack([a, [0, 0, 1]])
    assert np.allclose(out[3], (H(T(1, 0)) @ H(T(3, 1, 0.01)))[:2])             # frame 2 was skipped: 1->2 then 2->3
    assert np.array_equal(out[4], np.eye(2, 3)) and np.array_equal(out[5], T(5, 5))
    out = compose_skipped_affines(w, np.array([False, False, True, True, True, True]))   # the first PROCESSED frame gets the identity
    assert np.array_equal(out[2], np.eye(2, 3)) and np.array_equal(out[3], T(1, 0))


def test_partial_hyperparams_are_completed_with_the_reference_constructor_defaults():
    """A hyperparams dict that omits keys must run what the reference would run (plugin constructor defaults), not the tuned YAML values
    the device classes default to; OC-SORT's det_thresh has no default in the reference."""
    import pytest
    from tracklab_b200 import _lib
    from tracklab_b200.modules import _with_reference_defaults
    h = _with_reference_defaults("ByteTrack", {"track_buffer": 60})
    assert h == dict(track_thresh=0.45, track_buffer=60, match_thresh=0.8, frame_rate=30)          # byte_tracker.py:152
    h = _with_reference_defaults("OCSORT", {"det_thresh": 0.1, "asso_func": "giou"})
    assert h["max_age"] == 30 and h["min_hits"] == 3 and h["delta_t"] == 3 and h["asso_func"] == "giou"   # ocsort.py:183-184
    with pytest.raises(_lib.TrackKernError):
        _with_reference_defaults("OCSORT", {"max_age": 5})
    assert _with_reference_defaults("StrongSORT", {})["max_unmatched_preds"] == 7                   # strong_sort.py:20-31 (rejected loudly by the module)
    assert _with_reference_defaults("DeepOCSORT", {"a": 1}) == {"a": 1}
