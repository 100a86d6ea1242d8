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


# ---- ReID crop pre-processing (StrongSORT in-tracker ReID) -------------------------------------------------------
# /root/reference/plugins/track/strong_sort/strong_sort.py:102-108,135-145 (crop rule: int() truncation, clipping) and
# /root/reference/plugins/track/strong_sort/reid_multibackend.py:45-52,184-195 (PIL resize to 256x128 bilinear, ToTensor,
# Normalize). ``pil_resize_bilinear_u8`` restates Pillow's two-pass 8-bit resampler (antialiased: support = scale when
# down-scaling, 22-bit fixed-point coefficients, uint8 intermediate) so the CUDA kernel has an integer-exact model;
# tests pin it to PIL itself.
PIL_PRECISION_BITS = 32 - 8 - 2


def _pil_coeffs(in_size, out_size):
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(ksize)
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
        ww = w.sum()
        if ww != 0.0:
            w = w / ww
        kk[xx] = [int(-0.5 + v * (1 << PIL_PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PIL_PRECISION_BITS)) for v in w]
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def pil_resize_bilinear_u8(img, out_w, out_h):
    """Integer-exact restatement of PIL.Image.fromarray(img).resize((out_w, out_h), BILINEAR) for uint8 HWC."""
    h, w = img.shape[:2]
    src = img.astype(np.int64)
    if w != out_w:
        bounds, kk = _pil_coeffs(w, out_w)
        tmp = np.empty((h, out_w, src.shape[2]), dtype=np.int64)
        for xx in range(out_w):
            x0, n = bounds[xx]
            acc = (1 << (PIL_PRECISION_BITS - 1)) + (src[:, x0:x0 + n, :] * kk[xx, :n][None, :, None]).sum(axis=1)
            tmp[:, xx, :] = np.clip(acc >> PIL_PRECISION_BITS, 0, 255)
        src = tmp
    if h != out_h:
        bounds, kk = _pil_coeffs(h, out_h)
        out = np.empty((out_h, src.shape[1], src.shape[2]), dtype=np.int64)
        for yy in range(out_h):
            y0, n = bounds[yy]
            acc = (1 << (PIL_PRECISION_BITS - 1)) + (src[y0:y0 + n, :, :] * kk[yy, :n][:, None, None]).sum(axis=0)
            out[yy] = np.clip(acc >> PIL_PRECISION_BITS, 0, 255)
        src = out
    return src.astype(np.uint8)


def strongsort_crop_box(xyxy, width, height):
    """strong_sort.py:102-108 applied to the wrapper row: xyxy -> xywh (centre) -> int()-truncated, clipped corners."""
    cx, cy, w, h = (xyxy[0] + xyxy[2]) / 2, (xyxy[1] + xyxy[3]) / 2, xyxy[2] - xyxy[0], xyxy[3] - xyxy[1]
    x1 = max(int(cx - w / 2), 0)
    x2 = min(int(cx + w / 2), width - 1)
    y1 = max(int(cy - h / 2), 0)
    y2 = min(int(cy + h / 2), height - 1)
    return x1, y1, x2, y2


REID_MEAN = (0.485, 0.456, 0.406)
REID_STD = (0.229, 0.224, 0.225)


def xyxy_int_crop_box(bb, W, H):
    """Deep OC-SORT / BoT-SORT crop (deep_oc_sort/ocsort.py:560-565, bot_sort/bot_sort.py `_get_features`): `box.astype(int)` on
    x1,y1,x2,y2 and the NumPy slice `img[y1:y2, x1:x2]` (stops clamp at the image size). Negative starts would wrap around in
    NumPy; boxes are clipped to the image by the detector wrappers, so they are clamped to 0 here (documented deviation)."""
    x1, y1, x2, y2 = (int(v) for v in np.asarray(bb, dtype=np.float64)[:4].astype(int))
    return max(x1, 0), max(y1, 0), min(max(x2, 0), W), min(max(y2, 0), H)


def reid_crops(frame_rgb, dets_xyxy, out_hw=(256, 128), use_pil=True, rule="strongsort"):
    """float32 [D,3,256,128] network input of the in-tracker ReID (reid_multibackend.py:184-195) for one frame."""
    H, W = frame_rgb.shape[:2]
    out = np.zeros((len(dets_xyxy), 3, out_hw[0], out_hw[1]), dtype=np.float32)
    for i, bb in enumerate(dets_xyxy):
        x1, y1, x2, y2 = strongsort_crop_box(bb, W, H) if rule == "strongsort" else xyxy_int_crop_box(bb, W, H)
        crop = frame_rgb[y1:y2, x1:x2]
        if use_pil:
            from PIL import Image
            small = np.asarray(Image.fromarray(crop).resize((out_hw[1], out_hw[0]), Image.BILINEAR))
        else:
            small = pil_resize_bilinear_u8(crop, out_hw[1], out_hw[0])
        t = small.astype(np.float32).transpose(2, 0, 1) / np.float32(255)      # ToTensor
        for c in range(3):
            out[i, c] = (t[c] - np.float32(REID_MEAN[c])) / np.float32(REID_STD[c])   # Normalize
    return out
