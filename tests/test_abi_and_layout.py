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
