"""GPU: the REAL reference engine (staged into oracle/_ref/ by __graft_entry__.build(), or /root/reference) over the CUDA modules.

  * ByteTrack drop-in under the real OfflineTrackingEngine + TrackerState == the golden the real engine produced with the
    reference wrapper (ids exact, boxes < 1e-9);
  * Pipeline([RTMLibDetector, KPReId, BPBReIDStrongSORT]) validated by the real Pipeline / TrackerState and run by the real engine
    on synthetic PNG frames (no ground-truth detections: every column comes from the modules); the tracker's ids equal the oracle
    chain on the emitted detections."""
import pytest

from tests.test_real_engine_cpu import _have_reference, run_driver

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not _have_reference(), reason="neither /root/reference nor the staged oracle/_ref is present")
def test_real_engine_over_cuda_bytetrack_module_equals_real_engine_golden():
    r = run_driver("bytetrack")
    assert r["index_equal"] and r["ids_equal"] and r["max_box_err"] < 1e-9 and r["with_track"] > 1000


@pytest.mark.skipif(not _have_reference(), reason="neither /root/reference nor the staged oracle/_ref is present")
def test_real_engine_runs_detector_reid_tracker_modules_from_png_frames():
    r = run_driver("chain")
    assert r["levels"] == ["image", "detection", "image"]
    for c in ("bbox_ltwh", "bbox_conf", "category_id", "embeddings", "visibility_scores", "track_id", "track_bbox_kf_ltwh", "hits", "age"):
        assert c in r["columns"], c
    assert r["detections"] > 100 and r["with_track"] == r["detections"] and r["embedding_shape"] == [1, 2048] and r["bbox_dtype"] == "float32"
    assert r["max_embedding_rel_err"] < 2e-4
    assert r["oracle_ids_equal"]
