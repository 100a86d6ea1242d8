"""GPU parity: ByteTrack whole-video kernel (C ABI) vs the committed reference goldens and the oracle."""
import numpy as np
import pytest
import torch

from tests.util import assert_rows_match, load_golden
from tracklab_b200.synth import make_video

pytestmark = pytest.mark.gpu

GOLDENS = ["bytetrack_c2_s2000", "bytetrack_small_s5", "bytetrack_buffer5_s9"]


def _run_device(video, hyper, min_conf, chunk=None, n_seq=1):
    from tracklab_b200.device_trackers import ByteTrackDevice, rows_to_frames
    trk = ByteTrackDevice(**hyper, min_confidence=min_conf, n_seq=n_seq, cap_tracks=128, cap_dets=128)
    dets = torch.from_numpy(video.dets).cuda()
    offs = torch.from_numpy(video.offsets.astype(np.int32))[None].cuda()
    if chunk is None:
        rows, fc, cnt = trk.run(dets, offs)
        trk.check_status()
        return rows_to_frames(rows, fc, offs[:, 0].contiguous())
    out_rows = torch.empty((video.n_dets, 8), dtype=torch.float64, device="cuda")
    out_start = torch.zeros(1, dtype=torch.int32, device="cuda")
    out_count = torch.zeros(1, dtype=torch.int32, device="cuda")
    fcs = []
    for f0 in range(0, video.n_frames, chunk):
        f1 = min(video.n_frames, f0 + chunk)
        o = offs[:, f0:f1 + 1].contiguous()
        _, fc, _ = trk.run(dets, o, out_rows=out_rows, out_start=out_start, out_count=out_count)
        fcs.append(fc)
    trk.check_status()
    return rows_to_frames(out_rows, torch.cat(fcs, dim=1), out_start)


@pytest.mark.parametrize("name", GOLDENS)
def test_bytetrack_matches_reference_golden(name):
    g = load_golden(name)
    video = make_video(**g["gen"])
    rows, frames = _run_device(video, g["hyper"], g["min_conf"])
    err = assert_rows_match(rows, frames, g["rows"], g["frames"], box_tol=1e-6)
    print(name, "max box err", err)


def test_bytetrack_chunked_equals_whole():
    g = load_golden("bytetrack_c2_s2000")
    video = make_video(**g["gen"])
    rows, frames = _run_device(video, g["hyper"], g["min_conf"], chunk=17)
    assert_rows_match(rows, frames, g["rows"], g["frames"], box_tol=1e-6)


@pytest.mark.parametrize("seed", [11, 12])
def test_bytetrack_matches_oracle_fresh_seed(seed):
    from oracle.bytetrack_np import ByteTrackOracle
    video = make_video(seed=seed, n_frames=200, n_ids=60, conf_range=(0.2, 1.0))
    hyper = dict(track_thresh=0.6, match_thresh=0.8, track_buffer=30, frame_rate=30)
    ref_rows, ref_frames = ByteTrackOracle(**hyper, min_confidence=0.4).run_video(video.dets, video.offsets)
    rows, frames = _run_device(video, hyper, 0.4)
    assert_rows_match(rows, frames, ref_rows, ref_frames, box_tol=1e-6)


def test_hota_of_device_rows_equals_hota_of_reference_rows():
    """BASELINE metric 'HOTA vs ref': the device tracker's rows score exactly the reference plugin's HOTA on the same video
    (oracle/hota_np.py restates the TrackEval HOTA vendored in the reference, tests/test_oracle_cpu.py)."""
    from oracle.hota_np import hota_of_tracker_rows
    g = load_golden("bytetrack_c2_s2000")
    video = make_video(**g["gen"])
    rows, frames = _run_device(video, g["hyper"], g["min_conf"])
    dev, ref = hota_of_tracker_rows(video, rows, frames), hota_of_tracker_rows(video, g["rows"], g["frames"])
    for k in ("HOTA", "DetA", "AssA", "LocA"):
        assert np.array_equal(dev[k], ref[k]), k
    print("HOTA", dev["HOTA"].mean(), "DetA", dev["DetA"].mean(), "AssA", dev["AssA"].mean())
