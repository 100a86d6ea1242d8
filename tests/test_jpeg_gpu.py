"""Frame ingest on the device: nvJPEG batched decode (libtkjpeg.so) vs cv2.imread (libjpeg-turbo), and the detector
rows on both. JPEG decoders are not pixel-identical by design; the tests pin HOW different they are."""
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _write_jpegs(tmp_path, n, sampling):
    import cv2
    from tracklab_b200.synth import make_frames, make_video
    video = make_video(seed=3300, n_frames=n, n_ids=20)
    frames = make_frames(video, 0, n, device="cpu").numpy()
    paths = []
    for f in range(n):
        p = str(tmp_path / f"{sampling}_{f:06d}.jpg")
        cv2.imwrite(p, frames[f][..., ::-1], [cv2.IMWRITE_JPEG_QUALITY, 92, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sampling])
        paths.append(p)
    return video, paths


def test_nvjpeg_batch_decode_vs_libjpeg_turbo(tmp_path):
    import cv2
    from tracklab_b200.ingest import load_frames
    for sampling, mean_tol, p999_tol in ((cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, 0.8, 6), (cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, 2.5, 80)):
        video, paths = _write_jpegs(tmp_path, 8, sampling)
        load_frames(paths, "cuda:0", "nvjpeg")
        t0 = time.perf_counter(); dev_frames = load_frames(paths, "cuda:0", "nvjpeg"); torch.cuda.synchronize(); t_nv = time.perf_counter() - t0
        t0 = time.perf_counter(); same = load_frames(paths, "cuda:0", "cv2"); torch.cuda.synchronize(); t_cv = time.perf_counter() - t0
        ref = np.stack([cv2.cvtColor(cv2.imread(p), cv2.COLOR_BGR2RGB) for p in paths])
        assert np.array_equal(same.cpu().numpy(), ref)                         # decode="cv2" keeps the reference's pixels exactly
        assert tuple(dev_frames.shape) == ref.shape and dev_frames.dtype == torch.uint8
        diff = np.abs(dev_frames.cpu().numpy().astype(np.int16) - ref.astype(np.int16))
        print(f"sampling {sampling:#x}: nvjpeg vs libjpeg-turbo mean {diff.mean():.3f} p99.9 {np.percentile(diff, 99.9):.0f} max {diff.max()}; "
              f"8 x 1080p: nvjpeg {t_nv * 1e3:.1f} ms, cv2 + H2D {t_cv * 1e3:.1f} ms")
        assert diff.mean() < mean_tol and np.percentile(diff, 99.9) <= p999_tol


def test_detector_rows_on_nvjpeg_frames_match_cv2_frames(tmp_path):
    """4:4:4 files (decoders differ by IDCT rounding only): the trained YOLOX-s gives the same detections up to sub-pixel differences."""
    import cv2
    from tracklab_b200.detector import YoloxDetectorDevice, synth_weights_path
    from tracklab_b200.ingest import load_frames
    video, paths = _write_jpegs(tmp_path, 4, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444)
    det = YoloxDetectorDevice("s", device="cuda:0", batch=4, frames_cap=8, dets_cap=4096, weights=synth_weights_path("s"))
    out = []
    for mode in ("nvjpeg", "cv2"):
        det.reset()
        det.detect_batch(load_frames(paths, "cuda:0", mode))
        torch.cuda.synchronize()
        det.check_status()
        n = int(det.cursor[0].item())
        out.append((det.dets[:n].cpu().numpy().copy(), det.offsets[:5].cpu().numpy().copy()))
    (a, oa), (b, ob) = out
    assert abs(len(a) - len(b)) <= max(2, 0.05 * len(b))
    matched = 0
    for f in range(4):
        fa, fb = a[oa[f]:oa[f + 1]], b[ob[f]:ob[f + 1]]
        for r in fb:
            if len(fa) and np.abs(fa[:, :4] - r[:4]).max(axis=1).min() < 2.0:
                matched += 1
    assert matched >= 0.9 * len(b)
