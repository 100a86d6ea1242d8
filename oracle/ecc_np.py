"""ECC camera-motion compensation of the StrongSORT plugin, restated (test infrastructure).

Reference: /root/reference/plugins/track/strong_sort/sort/track.py:129-243 (Track.ECC + camera_update, called once per track and
frame from /root/reference/tracklab/wrappers/track/strong_sort_api.py:62-65 when cfg.ecc): gray conversion with cv2.COLOR_BGR2GRAY
applied to the RGB frame, cv2.resize(fx = fy = 0.1, INTER_LINEAR), cv2.findTransformECC(MOTION_EUCLIDEAN, 100 iterations,
eps 1e-5, gaussFiltSize 1), translation rescaled by 1 / 0.1, box corners warped.

cv2.findTransformECC is OpenCV (third party, present here): `find_transform_ecc_euclidean` restates its algorithm
(modules/video/src/ecc.cpp: forward-additive ECC of Evangelidis & Psarakis) INCLUDING warpAffine's fixed-point source
coordinates (1/32-pixel interpolation grid) so that the device kernel (csrc/ecc.cu), which follows this restatement, can be held
to 1e-3 of OpenCV itself; tests/test_oracle_cpu.py pins the restatement to cv2.findTransformECC."""
import numpy as np

AB_BITS, INTER_BITS = 10, 5
AB_SCALE, INTER_TAB = 1 << AB_BITS, 1 << INTER_BITS


def gray_bgr2gray_u8(img):
    """cv2.cvtColor(img, COLOR_BGR2GRAY) for uint8 (OpenCV 4.x, 15-bit coefficients): (ch0 * 3735 + ch1 * 19235 + ch2 * 9798 + 16384) >> 15.
    The plugin applies it to the RGB frame (strong_sort_api.py:61 loads RGB), so ch0 is the red channel there."""
    i = img.astype(np.int64)
    return ((i[..., 0] * 3735 + i[..., 1] * 19235 + i[..., 2] * 9798 + 16384) >> 15).astype(np.uint8)


def _warp_coords(M, h, w):
    """Source coordinates of cv2.warpAffine(..., WARP_INVERSE_MAP): integer part + 1/32 fraction per destination pixel."""
    M = np.asarray(M, dtype=np.float64)
    xs = np.arange(w)
    adelta = np.rint(M[0, 0] * xs * AB_SCALE).astype(np.int64)
    bdelta = np.rint(M[1, 0] * xs * AB_SCALE).astype(np.int64)
    rd = AB_SCALE // INTER_TAB // 2
    ys = np.arange(h)
    X0 = np.rint((M[0, 1] * ys + M[0, 2]) * AB_SCALE).astype(np.int64) + rd
    Y0 = np.rint((M[1, 1] * ys + M[1, 2]) * AB_SCALE).astype(np.int64) + rd
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    return X >> INTER_BITS, Y >> INTER_BITS, X & (INTER_TAB - 1), Y & (INTER_TAB - 1)


def warp_affine_linear(src, M):
    """cv2.warpAffine(src float32, M, flags = INTER_LINEAR | WARP_INVERSE_MAP), constant border 0."""
    h, w = src.shape
    sx, sy, fx, fy = _warp_coords(M, h, w)
    ax, ay = (fx / np.float32(INTER_TAB)).astype(np.float32), (fy / np.float32(INTER_TAB)).astype(np.float32)

    def at(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        return np.where(ok, src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], np.float32(0))

    w00, w01 = (1 - ay) * (1 - ax), (1 - ay) * ax
    w10, w11 = ay * (1 - ax), ay * ax
    return (at(sy, sx) * w00 + at(sy, sx + 1) * w01 + at(sy + 1, sx) * w10 + at(sy + 1, sx + 1) * w11).astype(np.float32)


def warp_mask_nearest(M, h, w):
    """cv2.warpAffine(ones uint8, M, flags = INTER_NEAREST | WARP_INVERSE_MAP): 1 where the rounded source pixel is inside."""
    Md = np.asarray(M, dtype=np.float64)
    xs, ys = np.arange(w), np.arange(h)
    adelta = np.rint(Md[0, 0] * xs * AB_SCALE).astype(np.int64)
    bdelta = np.rint(Md[1, 0] * xs * AB_SCALE).astype(np.int64)
    rd = AB_SCALE // 2
    X0 = np.rint((Md[0, 1] * ys + Md[0, 2]) * AB_SCALE).astype(np.int64) + rd
    Y0 = np.rint((Md[1, 1] * ys + Md[1, 2]) * AB_SCALE).astype(np.int64) + rd
    X = (X0[:, None] + adelta[None, :]) >> AB_BITS
    Y = (Y0[:, None] + bdelta[None, :]) >> AB_BITS
    return ((X >= 0) & (X < w) & (Y >= 0) & (Y < h))


def gradients(img):
    """filter2D with [-0.5, 0, 0.5] (and its transpose), BORDER_REFLECT_101."""
    p = np.pad(img, 1, mode="reflect")
    gx = (p[1:-1, 2:] - p[1:-1, :-2]) * np.float32(0.5)
    gy = (p[2:, 1:-1] - p[:-2, 1:-1]) * np.float32(0.5)
    return gx.astype(np.float32), gy.astype(np.float32)


def find_transform_ecc_euclidean(template_u8, image_u8, max_iter=100, eps=1e-5):
    """cv2.findTransformECC(template, image, eye(2,3), MOTION_EUCLIDEAN, (COUNT|EPS, max_iter, eps), None, 1) -> (rho, 2x3 float32)."""
    tmpl = template_u8.astype(np.float32)
    img = image_u8.astype(np.float32)
    h, w = tmpl.shape
    gx, gy = gradients(img)
    Xg, Yg = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    M = np.eye(2, 3, dtype=np.float32)
    rho, last_rho = -1.0, -eps
    it = 1
    while it <= max_iter and abs(rho - last_rho) >= eps:
        iw = warp_affine_linear(img, M)
        gxw, gyw = warp_affine_linear(gx, M), warp_affine_linear(gy, M)
        mask = warp_mask_nearest(M, h, w)
        n = int(mask.sum())
        im, tm = iw[mask].astype(np.float64), tmpl[mask].astype(np.float64)
        i_mean, t_mean = im.mean(), tm.mean()
        i_std = np.sqrt(max((im * im).mean() - i_mean * i_mean, 0.0))
        t_std = np.sqrt(max((tm * tm).mean() - t_mean * t_mean, 0.0))
        izm = np.where(mask, iw - np.float32(i_mean), iw).astype(np.float32)        # subtract(..., mask): unmasked pixels keep their value
        tzm = np.where(mask, tmpl - np.float32(t_mean), np.float32(0)).astype(np.float32)   # templateZM starts as zeros
        tnorm, inorm = np.sqrt(n * t_std * t_std), np.sqrt(n * i_std * i_std)
        h0, h1 = M[0, 0], M[1, 0]
        hatx = -(Xg * h1) - (Yg * h0)
        haty = (Xg * h0) - (Yg * h1)
        J = np.stack([gxw * hatx + gyw * haty, gxw, gyw]).astype(np.float32)        # [3, h, w]
        Jd = J.reshape(3, -1).astype(np.float64)
        H = Jd @ Jd.T
        Hinv = np.linalg.inv(H)
        corr = float(tzm.astype(np.float64).ravel() @ izm.astype(np.float64).ravel())
        last_rho, rho = rho, corr / (inorm * tnorm)
        ip = Jd @ izm.astype(np.float64).ravel()
        tp = Jd @ tzm.astype(np.float64).ravel()
        iph = Hinv @ ip
        lam_n = inorm * inorm - ip @ iph
        lam_d = corr - tp @ iph
        if lam_d <= 0.0:
            raise RuntimeError("ECC did not converge")
        lam = lam_n / lam_d
        err = (np.float32(lam) * tzm - izm).astype(np.float32)
        ep = Jd @ err.astype(np.float64).ravel()
        dp = Hinv @ ep
        theta = np.arcsin(np.float64(M[1, 0])) + dp[0]
        M = M.copy()
        M[0, 2] += np.float32(dp[1]); M[1, 2] += np.float32(dp[2])
        M[0, 0] = M[1, 1] = np.float32(np.cos(theta)); M[1, 0] = np.float32(np.sin(theta)); M[0, 1] = -M[1, 0]
        it += 1
    return rho, M
