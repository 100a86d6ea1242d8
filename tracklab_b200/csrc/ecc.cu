// ECC camera-motion compensation of the StrongSORT plugin on device.
//
// Reference: /root/reference/plugins/track/strong_sort/sort/track.py:129-243 (Track.ECC + camera_update), run for every track and
// frame by /root/reference/tracklab/wrappers/track/strong_sort_api.py:62-65 when cfg.ecc (the default of
// configs/modules/track/strong_sort.yaml:13): cv2.cvtColor(COLOR_BGR2GRAY) applied to the RGB frame, cv2.resize(fx = fy = 0.1,
// INTER_LINEAR), cv2.findTransformECC(MOTION_EUCLIDEAN, 100 iterations, eps 1e-5, gaussFiltSize 1), translation / 0.1.
// The reference recomputes the same transform once PER TRACK; it only depends on the frame pair, so it is computed once per pair
// here (all pairs of a batch of frames in parallel, one CTA each) and handed to tk_strongsort_run_cmc as a [frames, 6] array.
//
//   tk_ecc_gray_small   frames uint8 [n,H,W,3] -> uint8 [n, round(0.1 H), round(0.1 W)]: OpenCV's 15-bit gray conversion and its
//                       11-bit fixed-point bilinear resize, evaluated exactly (integer arithmetic, bit-equal to cv2)
//   tk_ecc_euclidean    consecutive pairs of small gray images -> 2x3 warp (translation already divided by the scale), rho, ok:
//                       the forward-additive ECC iteration of OpenCV (modules/video/src/ecc.cpp) incl. warpAffine's 1/32-pixel
//                       fixed-point source coordinates, restated in oracle/ecc_np.py (pinned to cv2.findTransformECC).
// Both images of a pair (2 x 20 KB at 1080p) live in shared memory; per iteration two passes over the pixels (masked mean / std,
// then correlation + 3x3 Hessian + projections) with block reductions in float64, the 3x3 solve on one thread.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "tk_common.cuh"
#include "trackkern.h"

namespace {

constexpr int ECC_THREADS = 512;
constexpr int AB_BITS = 10, INTER_BITS = 5, AB_SCALE = 1 << AB_BITS, INTER_TAB = 1 << INTER_BITS;

__device__ __forceinline__ int gray_px(const unsigned char* p) {   // COLOR_BGR2GRAY on the given channel order
    return (p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15;
}

// cv2.resize(gray, (0,0), fx=s, fy=s, INTER_LINEAR) on uint8, one channel: source coordinate (d + 0.5) / s - 0.5, 11-bit weights,
// horizontal pass in int, vertical pass (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2 (resize.cpp, VResizeLinear)
__global__ void ecc_gray_small_kernel(const unsigned char* __restrict__ frames, long long frame_stride, int H, int W, int h, int w,
                                      double inv_scale_x, double inv_scale_y, unsigned char* __restrict__ out) {
    const int n = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const unsigned char* img = frames + (size_t)n * frame_stride;
    auto coef = [](int d, double scale, int n_src, int& s0, int& a0, int& a1) {
        float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
        int s = (int)floorf(f);
        f = __fsub_rn(f, (float)s);
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= n_src - 1) { s = n_src - 1; f = 0.f; }   // cv: sx >= ssize.width-1 -> fx = 0, sx = width-1 (second tap clamped)
        s0 = s;
        a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
        a1 = __float2int_rn(__fmul_rn(f, 2048.0f));
    };
    int sx, ax0, ax1, sy, ay0, ay1;
    coef(x, inv_scale_x, W, sx, ax0, ax1);
    coef(y, inv_scale_y, H, sy, ay0, ay1);
    const int sx1 = min(sx + 1, W - 1), sy1 = min(sy + 1, H - 1);
    const unsigned char* r0 = img + (size_t)sy * W * 3;
    const unsigned char* r1 = img + (size_t)sy1 * W * 3;
    const int h0 = gray_px(r0 + 3 * sx) * ax0 + gray_px(r0 + 3 * sx1) * ax1;
    const int h1 = gray_px(r1 + 3 * sx) * ax0 + gray_px(r1 + 3 * sx1) * ax1;
    const int v = (((ay0 * (h0 >> 4)) >> 16) + ((ay1 * (h1 >> 4)) >> 16) + 2) >> 2;
    out[((size_t)n * h + y) * w + x] = (unsigned char)min(max(v, 0), 255);
}

struct Sums13 { double v[13]; };

__device__ __forceinline__ double warp_sum_d(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// block reduction of N doubles per thread; result valid in every thread (red: shared scratch [N][32])
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_sum_d(v[i]);
    __syncthreads();
    if (lane == 0)
        for (int i = 0; i < N; ++i) red[i * 32 + warp] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double t = lane < nw ? red[i * 32 + lane] : 0.0;
        v[i] = warp_sum_d(t);
    }
}

struct Warped { float iw, gx, gy; bool m; };

// one destination pixel of warpAffine(INTER_LINEAR | WARP_INVERSE_MAP) of the image and of its two gradient images, + the
// INTER_NEAREST mask. Md: the float32 map converted to double (cv2 does the same). img: uint8 [h][w] in shared memory.
__device__ __forceinline__ Warped warp_px(const unsigned char* img, int h, int w, const double* Md, int x, int y) {
    const int adelta = (int)llrint(Md[0] * x * AB_SCALE), bdelta = (int)llrint(Md[3] * x * AB_SCALE);
    const long long xb = llrint((Md[1] * y + Md[2]) * AB_SCALE), yb = llrint((Md[4] * y + Md[5]) * AB_SCALE);
    const int rdl = AB_SCALE / INTER_TAB / 2;
    const int X = (int)((xb + rdl + adelta) >> (AB_BITS - INTER_BITS)), Y = (int)((yb + rdl + bdelta) >> (AB_BITS - INTER_BITS));
    const int sx = X >> INTER_BITS, sy = Y >> INTER_BITS;
    const float ax = (float)(X & (INTER_TAB - 1)) / (float)INTER_TAB, ay = (float)(Y & (INTER_TAB - 1)) / (float)INTER_TAB;
    const float w00 = (1.f - ay) * (1.f - ax), w01 = (1.f - ay) * ax, w10 = ay * (1.f - ax), w11 = ay * ax;
    auto I = [&](int yy, int xx) -> float {   // image value, 0 outside (BORDER_CONSTANT)
        return (yy >= 0 && yy < h && xx >= 0 && xx < w) ? (float)img[yy * w + xx] : 0.f;
    };
    auto R = [&](int v, int n) { return v < 0 ? -v : (v >= n ? 2 * n - 2 - v : v); };   // BORDER_REFLECT_101
    auto GX = [&](int yy, int xx) -> float {  // filter2D [-0.5 0 0.5] of the (unwarped) image, 0 outside
        if (!(yy >= 0 && yy < h && xx >= 0 && xx < w)) return 0.f;
        return ((float)img[yy * w + R(xx + 1, w)] - (float)img[yy * w + R(xx - 1, w)]) * 0.5f;
    };
    auto GY = [&](int yy, int xx) -> float {
        if (!(yy >= 0 && yy < h && xx >= 0 && xx < w)) return 0.f;
        return ((float)img[R(yy + 1, h) * w + xx] - (float)img[R(yy - 1, h) * w + xx]) * 0.5f;
    };
    Warped o;
    o.iw = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(I(sy, sx), w00), __fmul_rn(I(sy, sx + 1), w01)), __fmul_rn(I(sy + 1, sx), w10)),
                     __fmul_rn(I(sy + 1, sx + 1), w11));
    o.gx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(GX(sy, sx), w00), __fmul_rn(GX(sy, sx + 1), w01)), __fmul_rn(GX(sy + 1, sx), w10)),
                     __fmul_rn(GX(sy + 1, sx + 1), w11));
    o.gy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(GY(sy, sx), w00), __fmul_rn(GY(sy, sx + 1), w01)), __fmul_rn(GY(sy + 1, sx), w10)),
                     __fmul_rn(GY(sy + 1, sx + 1), w11));
    const int rdn = AB_SCALE / 2;
    const int Xn = (int)((xb + rdn + adelta) >> AB_BITS), Yn = (int)((yb + rdn + bdelta) >> AB_BITS);
    o.m = Xn >= 0 && Xn < w && Yn >= 0 && Yn < h;
    return o;
}

__device__ bool inv3(const double* H, double* Hi) {   // 3x3 inverse by cofactors (cv::Mat::inv(DECOMP_LU) on a 3x3)
    const double a = H[0], b = H[1], c = H[2], d = H[3], e = H[4], f = H[5], g = H[6], hh = H[7], i = H[8];
    const double A = e * i - f * hh, B = -(d * i - f * g), C = d * hh - e * g;
    const double det = a * A + b * B + c * C;
    if (det == 0.0 || !(det == det)) return false;
    const double r = 1.0 / det;
    Hi[0] = A * r; Hi[1] = -(b * i - c * hh) * r; Hi[2] = (b * f - c * e) * r;
    Hi[3] = B * r; Hi[4] = (a * i - c * g) * r;  Hi[5] = -(a * f - c * d) * r;
    Hi[6] = C * r; Hi[7] = -(a * hh - b * g) * r; Hi[8] = (a * e - b * d) * r;
    return true;
}

// pair p: template = small[p], image = small[p + 1]; writes warps[(p + 1) * 6 ..], rho[p + 1], ok[p + 1]
__global__ void __launch_bounds__(ECC_THREADS)
ecc_euclidean_kernel(const unsigned char* __restrict__ small, int h, int w, int max_iter, double eps, float inv_scale_div,
                     float* __restrict__ warps, double* __restrict__ rho_out, int* __restrict__ ok_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int p = blockIdx.x, np = h * w;
    unsigned char* tmpl = smem_raw;
    unsigned char* img = smem_raw + ((np + 15) & ~15);
    double* red = (double*)(img + ((np + 15) & ~15));             // [13][32]
    __shared__ float Mf[6];
    __shared__ double ctl[8];    // 0 rho, 1 last_rho, 2 stop flag, 3 i_mean, 4 t_mean
    __shared__ int s_ok;
    for (int i = threadIdx.x; i < np; i += blockDim.x) { tmpl[i] = small[(size_t)p * np + i]; img[i] = small[(size_t)(p + 1) * np + i]; }
    if (threadIdx.x == 0) {
        Mf[0] = 1.f; Mf[1] = 0.f; Mf[2] = 0.f; Mf[3] = 0.f; Mf[4] = 1.f; Mf[5] = 0.f;
        ctl[0] = -1.0; ctl[1] = -eps; ctl[2] = 0.0; s_ok = 1;
    }
    __syncthreads();
    for (int it = 1; it <= max_iter; ++it) {
        if (!(fabs(ctl[0] - ctl[1]) >= eps) || ctl[2] != 0.0) break;      // uniform: ctl is shared
        double Md[6];
        for (int i = 0; i < 6; ++i) Md[i] = (double)Mf[i];
        // ---- pass 1: masked sums of the warped image and of the template
        double s5[5] = {0, 0, 0, 0, 0};
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            const int y = i / w, x = i - y * w;
            const Warped wv = warp_px(img, h, w, Md, x, y);
            if (wv.m) {
                const double a = (double)wv.iw, t = (double)tmpl[i];
                s5[0] += 1.0; s5[1] += a; s5[2] += a * a; s5[3] += t; s5[4] += t * t;
            }
        }
        block_sum<5>(s5, red);
        const double n = s5[0];
        const double i_mean = s5[1] / n, t_mean = s5[3] / n;
        const double i_var = fmax(s5[2] / n - i_mean * i_mean, 0.0), t_var = fmax(s5[4] / n - t_mean * t_mean, 0.0);
        const double inorm = sqrt(n * i_var), tnorm = sqrt(n * t_var);   // sqrt(countNonZero * std^2)
        const float i_mean_f = (float)i_mean, t_mean_f = (float)t_mean;
        // ---- pass 2: correlation, Hessian (upper triangle), projections of the zero-mean images onto the Jacobian
        const float h0 = Mf[0], h1 = Mf[3];
        double s13[13];
        for (int k = 0; k < 13; ++k) s13[k] = 0.0;
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            const int y = i / w, x = i - y * w;
            const Warped wv = warp_px(img, h, w, Md, x, y);
            // subtract(imageWarped, mean, imageWarped, mask): pixels outside the mask keep their warped value; templateZM is 0 there
            const float izm = wv.m ? __fsub_rn(wv.iw, i_mean_f) : wv.iw;
            const float tzm = wv.m ? __fsub_rn((float)tmpl[i], t_mean_f) : 0.f;
            const float xf = (float)x, yf = (float)y;
            const float hatx = __fsub_rn(-__fmul_rn(xf, h1), __fmul_rn(yf, h0));
            const float haty = __fsub_rn(__fmul_rn(xf, h0), __fmul_rn(yf, h1));
            const double j0 = (double)__fadd_rn(__fmul_rn(wv.gx, hatx), __fmul_rn(wv.gy, haty)), j1 = (double)wv.gx, j2 = (double)wv.gy;
            const double iz = (double)izm, tz = (double)tzm;
            s13[0] += tz * iz;
            s13[1] += j0 * j0; s13[2] += j0 * j1; s13[3] += j0 * j2; s13[4] += j1 * j1; s13[5] += j1 * j2; s13[6] += j2 * j2;
            s13[7] += j0 * iz; s13[8] += j1 * iz; s13[9] += j2 * iz;
            s13[10] += j0 * tz; s13[11] += j1 * tz; s13[12] += j2 * tz;
        }
        block_sum<13>(s13, red);
        if (threadIdx.x == 0) {
            const double corr = s13[0];
            const double H[9] = {s13[1], s13[2], s13[3], s13[2], s13[4], s13[5], s13[3], s13[5], s13[6]};
            double Hi[9];
            const double rho = corr / (inorm * tnorm);
            ctl[1] = ctl[0]; ctl[0] = rho;
            bool good = inv3(H, Hi) && rho == rho;
            if (good) {
                const double ip[3] = {s13[7], s13[8], s13[9]}, tp[3] = {s13[10], s13[11], s13[12]};
                double iph[3];
                for (int r = 0; r < 3; ++r) iph[r] = Hi[r * 3] * ip[0] + Hi[r * 3 + 1] * ip[1] + Hi[r * 3 + 2] * ip[2];
                const double lam_n = inorm * inorm - (ip[0] * iph[0] + ip[1] * iph[1] + ip[2] * iph[2]);
                const double lam_d = corr - (tp[0] * iph[0] + tp[1] * iph[1] + tp[2] * iph[2]);
                if (lam_d <= 0.0) good = false;
                else {
                    const double lam = lam_n / lam_d;
                    // error = lambda * templateZM - imageWarped is linear in the two images: its projection needs no third pass
                    // (OpenCV rounds lambda * tz - iz to float per pixel; the difference is far below the 1e-3 tolerance)
                    double ep[3], dp[3];
                    for (int r = 0; r < 3; ++r) ep[r] = lam * tp[r] - ip[r];
                    for (int r = 0; r < 3; ++r) dp[r] = Hi[r * 3] * ep[0] + Hi[r * 3 + 1] * ep[1] + Hi[r * 3 + 2] * ep[2];
                    const double theta = asin((double)Mf[3]) + dp[0];
                    Mf[2] = (float)((double)Mf[2] + dp[1]);      // map(0,2) += deltaP(1): float += double, stored as float
                    Mf[5] = (float)((double)Mf[5] + dp[2]);
                    Mf[0] = Mf[4] = (float)cos(theta);
                    Mf[3] = (float)sin(theta);
                    Mf[1] = -Mf[3];
                }
            }
            if (!good) { s_ok = 0; ctl[2] = 1.0; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float* o = warps + (size_t)(p + 1) * 6;
        o[0] = Mf[0]; o[1] = Mf[1]; o[2] = __fdiv_rn(Mf[2], inv_scale_div);   // warp_matrix[0, 2] / scale[0] (float32 / 0.1)
        o[3] = Mf[3]; o[4] = Mf[4]; o[5] = __fdiv_rn(Mf[5], inv_scale_div);
        rho_out[p + 1] = ctl[0];
        ok_out[p + 1] = s_ok;
    }
}

}  // namespace

extern "C" {

int tk_ecc_small_size(int H, int W, double scale, int* h_out, int* w_out) {
    if (!h_out || !w_out || H <= 0 || W <= 0 || !(scale > 0)) return TK_ERR_ARG;
    *w_out = (int)llrint(W * scale);      // cvRound(ssize.width * inv_scale_x)
    *h_out = (int)llrint(H * scale);
    return (*w_out > 0 && *h_out > 0) ? TK_OK : TK_ERR_ARG;
}

int tk_ecc_gray_small(const unsigned char* frames, int n_frames, int H, int W, long long frame_stride_bytes, double scale,
                      unsigned char* out, void* stream) {
    if (!frames || !out || n_frames <= 0) return TK_ERR_ARG;
    int h, w;
    if (tk_ecc_small_size(H, W, scale, &h, &w) != TK_OK) return TK_ERR_ARG;
    dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8, n_frames);
    ecc_gray_small_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(frames, frame_stride_bytes, H, W, h, w, 1.0 / scale, 1.0 / scale, out);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_ecc_euclidean(const unsigned char* small, int n_images, int h, int w, int max_iter, double eps, double scale,
                     float* warps_out, double* rho_out, int* ok_out, void* stream) {
    if (!small || !warps_out || !rho_out || !ok_out || n_images <= 0 || h <= 0 || w <= 0 || max_iter <= 0) return TK_ERR_ARG;
    if (n_images == 1) return TK_OK;
    const size_t np = (size_t)h * w;
    const size_t smem = 2 * ((np + 15) & ~(size_t)15) + 13 * 32 * sizeof(double);
    if (smem > 200 * 1024) return TK_ERR_CAPACITY;
    TK_CUDA_TRY(cudaFuncSetAttribute(ecc_euclidean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ecc_euclidean_kernel<<<n_images - 1, ECC_THREADS, smem, (cudaStream_t)stream>>>(small, h, w, max_iter, eps, (float)scale, warps_out,
                                                                                   rho_out, ok_out);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

}  // extern "C"
