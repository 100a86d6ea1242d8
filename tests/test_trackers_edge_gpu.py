"""GPU: batched videos per launch, chunked continuation, empty / ragged frames, capacity errors."""
import numpy as np
import pytest
import torch

from tests.util import assert_rows_match
from tracklab_b200.synth import make_video

pytestmark = pytest.mark.gpu


def _stack(videos, with_feats=False):
    dets = np.concatenate([v.dets for v in videos])
    F = videos[0].n_frames
    offs, base = [], 0
    for v in videos:
        offs.append(v.offsets + base)
        base += v.n_dets
    out = [torch.from_numpy(dets).cuda(), torch.from_numpy(np.stack(offs).astype(np.int32)).cuda()]
    if with_feats:
        out.append(torch.from_numpy(np.concatenate([v.embeddings for v in videos])).cuda())
    return out


@pytest.mark.parametrize("kind", ["bytetrack", "ocsort"])
def test_three_videos_in_one_launch_equal_single_runs(kind):
    from oracle.bytetrack_np import ByteTrackOracle
    from oracle.ocsort_np import OCSortOracle
    from tracklab_b200.device_trackers import ByteTrackDevice, OCSortDevice, rows_to_frames
    videos = [make_video(seed=70 + i, n_frames=80, n_ids=25 + 5 * i) for i in range(3)]
    dets, offs = _stack(videos)
    Dev, Orc = (ByteTrackDevice, ByteTrackOracle) if kind == "bytetrack" else (OCSortDevice, OCSortOracle)
    trk = Dev(n_seq=3, cap_tracks=128, cap_dets=128)
    rows, fc, cnt = trk.run(dets, offs)
    trk.check_status()
    start = offs[:, 0].contiguous()
    for i, v in enumerate(videos):
        got, gf = rows_to_frames(rows, fc, start, seq=i)
        want, wf = Orc().run_video(v.dets, v.offsets)
        assert_rows_match(got, gf, want, wf, allow_relabel=(kind == "ocsort"))


def test_strongsort_two_videos_in_one_launch():
    from oracle.strongsort_np import StrongSortOracle
    from tracklab_b200.device_trackers import StrongSortDevice, rows_to_frames
    videos = [make_video(seed=80 + i, n_frames=60, n_ids=20, emb_dim=96) for i in range(2)]
    dets, offs, feats = _stack(videos, with_feats=True)
    trk = StrongSortDevice(96, nn_budget=30, n_seq=2, ctas_per_video=4)
    cap = 2 * max(v.n_dets for v in videos)
    out_rows = torch.empty((2 * cap, 8), dtype=torch.float64, device="cuda")
    out_start = torch.tensor([0, cap], dtype=torch.int32, device="cuda")
    rows, fc, cnt = trk.run(dets, offs, feats, out_rows=out_rows, out_start=out_start)
    trk.check_status()
    for i, v in enumerate(videos):
        got, gf = rows_to_frames(rows, fc, out_start, seq=i)
        want, wf = StrongSortOracle(nn_budget=30).run_video(v.dets, v.offsets, v.embeddings)
        assert_rows_match(got, gf, want, wf, box_tol=0.0)


@pytest.mark.parametrize("kind", ["bytetrack", "ocsort", "strongsort"])
def test_empty_low_confidence_and_single_detection_frames(kind):
    from oracle.bytetrack_np import ByteTrackOracle
    from oracle.ocsort_np import OCSortOracle
    from oracle.strongsort_np import StrongSortOracle
    from tracklab_b200.device_trackers import ByteTrackDevice, OCSortDevice, StrongSortDevice, rows_to_frames
    v = make_video(seed=90, n_frames=50, n_ids=8, emb_dim=32)
    dets, offs, emb = v.dets.copy(), v.offsets.copy(), v.embeddings
    keep = np.ones(len(dets), dtype=bool)
    for f in (3, 4, 5, 20):                      # frames with no detection rows at all
        keep[offs[f]:offs[f + 1]] = False
    dets[offs[10]:offs[11], 4] = 0.05            # a frame whose rows are all below the wrapper threshold
    keep[offs[30] + 1:offs[31]] = False          # a frame with a single detection
    new_off = np.concatenate([[0], np.cumsum([keep[offs[f]:offs[f + 1]].sum() for f in range(v.n_frames)])]).astype(np.int32)
    dets, emb = dets[keep], emb[keep]
    d = torch.from_numpy(dets).cuda(); o = torch.from_numpy(new_off)[None].cuda()
    if kind == "strongsort":
        trk = StrongSortDevice(32, nn_budget=10, ctas_per_video=2)
        rows, fc, cnt = trk.run(d, o, torch.from_numpy(emb).cuda())
        want, wf = StrongSortOracle(nn_budget=10).run_video(dets, new_off, emb)
        start = torch.zeros(1, dtype=torch.int32)
    else:
        trk = (ByteTrackDevice if kind == "bytetrack" else OCSortDevice)()
        rows, fc, cnt = trk.run(d, o)
        want, wf = (ByteTrackOracle if kind == "bytetrack" else OCSortOracle)().run_video(dets, new_off)
        start = o[:, 0].contiguous()
    trk.check_status()
    got, gf = rows_to_frames(rows, fc, start)
    assert_rows_match(got, gf, want, wf, box_tol=0.0 if kind == "strongsort" else 1e-6, allow_relabel=kind == "ocsort")


def test_ocsort_chunked_equals_whole():
    from tracklab_b200.device_trackers import OCSortDevice, rows_to_frames
    v = make_video(seed=95, n_frames=90, n_ids=30)
    d = torch.from_numpy(v.dets).cuda(); o = torch.from_numpy(v.offsets.astype(np.int32))[None].cuda()
    whole = OCSortDevice()
    r1, f1, _ = whole.run(d, o)
    a, af = rows_to_frames(r1, f1, o[:, 0].contiguous())
    trk = OCSortDevice()
    out_rows = torch.empty((v.n_dets, 8), dtype=torch.float64, device="cuda")
    start = torch.zeros(1, dtype=torch.int32, device="cuda"); count = torch.zeros(1, dtype=torch.int32, device="cuda")
    fcs = []
    for f0 in range(0, v.n_frames, 13):
        fe = min(v.n_frames, f0 + 13)
        _, fc, _ = trk.run(d, o[:, f0:fe + 1].contiguous(), out_rows=out_rows, out_start=start, out_count=count)
        fcs.append(fc)
    b, bf = rows_to_frames(out_rows, torch.cat(fcs, 1), start)
    assert np.array_equal(a, b) and np.array_equal(af, bf)


def test_capacity_overflow_is_reported_not_silent():
    from tracklab_b200 import _lib
    from tracklab_b200.device_trackers import ByteTrackDevice
    v = make_video(seed=96, n_frames=10, n_ids=40)
    trk = ByteTrackDevice(cap_tracks=16, cap_dets=64)
    trk.run(torch.from_numpy(v.dets).cuda(), torch.from_numpy(v.offsets.astype(np.int32))[None].cuda())
    with pytest.raises(_lib.TrackKernError, match="capacity"):
        trk.check_status()
    trk2 = ByteTrackDevice(cap_tracks=128, cap_dets=16)
    trk2.run(torch.from_numpy(v.dets).cuda(), torch.from_numpy(v.offsets.astype(np.int32))[None].cuda())
    with pytest.raises(_lib.TrackKernError, match="capacity"):
        trk2.check_status()
