"""Pre-processing oracles (test infrastructure): letterbox for the YOLOX detectors and the ReID crop rule.

``letterbox_yolox`` follows rtmlib 0.0.13 ``YOLOX.preprocess`` (un-vendored third party behind
/root/reference/tracklab/wrappers/bbox_detector/rtmlib_api.py:19-30; SURVEY.md §3.2 [3P-memory]):
ratio = min(S/h, S/w); cv2.resize(INTER_LINEAR) to (int(w*ratio), int(h*ratio)); paste top-left into
a 114-filled SxS canvas; HWC uint8 -> CHW float32, no mean/std, channel order untouched.
The resize itself is OpenCV's (present on both boxes) — ``resize_linear_u8`` restates its 11-bit
fixed-point arithmetic so the CUDA kernel has an integer-exact model; tests pin it to cv2.resize.
"""
import numpy as np

COEF_BITS = 11
COEF_ONE = 1 << COEF_BITS


def _linear_taps(src_n, dst_n):
    """Source index + 11-bit weights per destination index (OpenCV resize.cpp, INTER_LINEAR, 8U)."""
    scale = 1.0 / (dst_n / src_n)  # scale = 1/inv_scale with inv_scale = dst/src (double)
    d = np.arange(dst_n, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    f[lo] = 0.0
    s[lo] = 0
    hi = s >= src_n - 1
    f[hi] = 0.0
    s[hi] = src_n - 1
    w1 = np.rint(f * np.float32(COEF_ONE)).astype(np.int32)  # saturate_cast<short>: round half to even
    w0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_ONE)).astype(np.int32)
    s1 = np.minimum(s + 1, src_n - 1)
    return s, s1, w0, w1


def resize_linear_u8(img, dst_w, dst_h):
    """Integer-exact restatement of cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for uint8 HWC."""
    h, w = img.shape[:2]
    if w == 2 * dst_w and h == 2 * dst_h:
        # OpenCV switches exact 2x down-scaling to the INTER_AREA fast path: rounded 2x2 mean
        a = img.astype(np.int32)
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    x0, x1, a0, a1 = _linear_taps(w, dst_w)
    y0, y1, b0, b1 = _linear_taps(h, dst_h)
    src = img.astype(np.int32)
    r0 = src[y0][:, x0] * a0[None, :, None] + src[y0][:, x1] * a1[None, :, None]
    r1 = src[y1][:, x0] * a0[None, :, None] + src[y1][:, x1] * a1[None, :, None]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox_geometry(h, w, size):
    ratio = min(size / h, size / w)
    return ratio, int(w * ratio), int(h * ratio)


def letterbox_yolox(img, size=640, pad=114, use_cv2=True):
    """uint8 HWC -> (float32 CHW [3,size,size], ratio)."""
    h, w = img.shape[:2]
    ratio, rw, rh = letterbox_geometry(h, w, size)
    if use_cv2:
        import cv2
        small = cv2.resize(img, (rw, rh), interpolation=cv2.INTER_LINEAR).astype(np.uint8)
    else:
        small = resize_linear_u8(img, rw, rh)
    canvas = np.ones((size, size, 3), dtype=np.uint8) * pad
    canvas[:rh, :rw] = small
    return np.ascontiguousarray(canvas.transpose(2, 0, 1), dtype=np.float32), ratio
