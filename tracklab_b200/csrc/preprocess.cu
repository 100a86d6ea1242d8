// Frame pre-processing kernels: letterbox (detector input) — HBM-bound byte work.
//
// tk_letterbox_u8 replaces rtmlib 0.0.13 YOLOX.preprocess + the HWC->CHW float conversion that runs
// behind /root/reference/tracklab/wrappers/bbox_detector/rtmlib_api.py:19-30 (un-vendored; restated in
// oracle/preprocess_np.py): ratio = min(S/h, S/w), cv2.resize(INTER_LINEAR) to (int(w*ratio), int(h*ratio)),
// paste top-left into a 114-filled SxS canvas, no mean/std. The resize is OpenCV's 8-bit path: 11-bit
// fixed-point taps, horizontal pass in int32, vertical pass ((b*(S>>4))>>16 summed, +2 >> 2) — restated
// here integer-exactly so the detector sees the same bytes as the CPU path.
//
// B200 shape: one CTA per output row. The one or two source rows an output row needs are contiguous
// byte spans in HBM, so one elected thread moves them into shared memory with 1-D bulk async copies
// (cp.async.bulk, the TMA engine, completion on an mbarrier) — fully coalesced, no register staging —
// and every thread then blends its pixels out of shared memory and writes the three channel planes
// with coalesced stores. Rows whose vertical weight is zero (exact integer ratios such as 1080p -> 360)
// are not fetched at all.
#include <cuda_bf16.h>
#include "tk_common.cuh"
#include "trackkern.h"

namespace {

constexpr int LB_THREADS = 256;

struct Tap { int s0, s1, w0, w1; };

// OpenCV resize.cpp (INTER_LINEAR, 8U): fx = (float)((d+0.5)*scale-0.5), floor, clamp, 11-bit weights
__device__ __forceinline__ Tap linear_tap(int d, int src_n, double scale) {
    float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (s < 0) { f = 0.0f; s = 0; }
    if (s >= src_n - 1) { f = 0.0f; s = src_n - 1; }
    Tap t;
    t.s0 = s;
    t.s1 = min(s + 1, src_n - 1);
    t.w1 = __float2int_rn(__fmul_rn(f, 2048.0f));
    t.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
    return t;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(phase));
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

template <typename OutT> __device__ __forceinline__ OutT cvt_out(float v);
template <> __device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// src u8 [B, H, W, 3] (row pitch = W*3, frame pitch = frame_stride bytes), dst OutT [B, 3, S, S]
template <typename OutT>
__global__ void __launch_bounds__(LB_THREADS)
letterbox_kernel(const unsigned char* __restrict__ src, size_t frame_stride, int H, int W, OutT* __restrict__ dst,
                 int S, int rw, int rh, double scale_x, double scale_y, int area2x, int pad, int swap_rb, int nhwc,
                 const unsigned char* __restrict__ src_end) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const int y = blockIdx.x, b = blockIdx.y;
    // layouts: 0 planar [B,3,S,S]; 1 channels-last [B,S,S,3]; 2 "focus16" = YOLOX Focus space-to-depth already applied,
    // channels-last [B,S/2,S/2,16]: channel = patch*3 + c with patch order (tl, bl, tr, br) as in Focus.forward, 12..15 stay zero
    // (the stem convolution's input channels are zero-padded to 16 so cuDNN needs no NHWC padding pass).
    size_t xs = 1;
    OutT *out0, *out1, *out2;
    if (nhwc == 2) {
        const int S2 = S >> 1;
        out0 = dst + (((size_t)b * S2 + (y >> 1)) * S2) * 16 + (y & 1) * 3;
        out1 = out0 + 1; out2 = out0 + 2;
    } else if (nhwc == 1) {
        xs = 3;
        out0 = dst + (((size_t)b * S + y) * S) * 3; out1 = out0 + 1; out2 = out0 + 2;
    } else {
        out0 = dst + ((size_t)b * 3 + 0) * S * S + (size_t)y * S; out1 = out0 + (size_t)S * S; out2 = out1 + (size_t)S * S;
    }
    // element offset of output column x inside the row pointers above
    auto xo = [&](int x) -> size_t { return nhwc == 2 ? (size_t)(x >> 1) * 16 + (size_t)(x & 1) * 6 : (size_t)x * xs; };
    const OutT padv = cvt_out<OutT>((float)pad);
    if (y >= rh) {  // pure padding row
        for (int x = threadIdx.x; x < S; x += LB_THREADS) { const size_t o = xo(x); out0[o] = padv; out1[o] = padv; out2[o] = padv; }
        return;
    }
    Tap ty;
    if (area2x) { ty.s0 = 2 * y; ty.s1 = 2 * y + 1; ty.w0 = 1; ty.w1 = 1; }
    else ty = linear_tap(y, H, scale_y);
    const bool need1 = area2x || ty.w1 != 0;
    const size_t row_bytes = (size_t)W * 3;
    const unsigned char* base = src + (size_t)b * frame_stride;
    const unsigned char* g0 = base + (size_t)ty.s0 * row_bytes;
    const unsigned char* g1 = base + (size_t)ty.s1 * row_bytes;
    // 16-byte aligned spans for the bulk copies
    const size_t a0 = (size_t)g0 & 15, a1 = (size_t)g1 & 15;
    const unsigned n0 = (unsigned)((a0 + row_bytes + 15) & ~(size_t)15);
    const unsigned n1 = (unsigned)((a1 + row_bytes + 15) & ~(size_t)15);
    const size_t pitch = (row_bytes + 31 + 15) & ~(size_t)15;
    unsigned char* s0 = smem;
    unsigned char* s1 = smem + pitch;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // the 16-byte rounding may reach past the end of the frame buffer on the very last rows: plain loads there
    const bool tail = (g0 - a0 + n0 > src_end) || (need1 && (g1 - a1 + n1 > src_end));
    if (tail) {
        for (size_t i = threadIdx.x; i < row_bytes; i += LB_THREADS) {
            s0[a0 + i] = g0[i];
            if (need1) s1[a1 + i] = g1[i];
        }
        __syncthreads();
    } else {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, n0 + (need1 ? n1 : 0u));
            bulk_g2s(s0, g0 - a0, n0, &bar);
            if (need1) bulk_g2s(s1, g1 - a1, n1, &bar);
        }
        mbar_wait(&bar, 0);
    }
    const unsigned char* r0 = s0 + a0;
    const unsigned char* r1 = need1 ? s1 + a1 : r0;
    const int c0 = swap_rb ? 2 : 0, c2 = swap_rb ? 0 : 2;
    for (int x = threadIdx.x; x < S; x += LB_THREADS) {
        const size_t o = xo(x);
        if (x >= rw) { out0[o] = padv; out1[o] = padv; out2[o] = padv; continue; }
        int v[3];
        if (area2x) {
            const unsigned char* p0 = r0 + 6 * x;
            const unsigned char* p1 = r1 + 6 * x;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = (p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2;
        } else {
            const Tap tx = linear_tap(x, W, scale_x);
            const unsigned char* p00 = r0 + 3 * tx.s0;
            const unsigned char* p01 = r0 + 3 * tx.s1;
            const unsigned char* p10 = r1 + 3 * tx.s0;
            const unsigned char* p11 = r1 + 3 * tx.s1;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int h0 = p00[c] * tx.w0 + p01[c] * tx.w1;
                const int h1 = p10[c] * tx.w0 + p11[c] * tx.w1;
                int o = (((ty.w0 * (h0 >> 4)) >> 16) + ((ty.w1 * (h1 >> 4)) >> 16) + 2) >> 2;
                v[c] = min(max(o, 0), 255);
            }
        }
        out0[o] = cvt_out<OutT>((float)v[c0]);
        out1[o] = cvt_out<OutT>((float)v[1]);
        out2[o] = cvt_out<OutT>((float)v[c2]);
    }
}


// Focus-unfolded layout (out_layout 2): one CTA per PAIR of output rows, so every thread owns whole 16-channel pixels of
// the [S/2, S/2, 16] tensor and writes them with full 32-byte stores (the per-row version wrote 2-byte pieces of those
// pixels; the partial-sector writes made L2 read each sector back from HBM: 2.4x the algorithmic traffic in ncu).
template <typename OutT>
__global__ void __launch_bounds__(LB_THREADS)
letterbox_focus16_kernel(const unsigned char* __restrict__ src, size_t frame_stride, int H, int W, OutT* __restrict__ dst,
                         int S, int rw, int rh, double scale_x, double scale_y, int area2x, int pad, int swap_rb,
                         const unsigned char* __restrict__ src_end, int pix_pitch) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const int fy = blockIdx.x, b = blockIdx.y, S2 = S >> 1;
    OutT* orow = dst + (((size_t)b * S2 + fy) * S2) * pix_pitch;   // pix_pitch 16, or 32 with caller-zeroed channels 16..31
    const float padf = (float)pad;
    const size_t row_bytes = (size_t)W * 3;
    const size_t pitch = (row_bytes + 31 + 15) & ~(size_t)15;
    const unsigned char* base = src + (size_t)b * frame_stride;
    Tap ty[2];
    const unsigned char* g[4];
    size_t al[4];
    unsigned nb[4];
    bool need[4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int y = 2 * fy + r;
        if (area2x) { ty[r].s0 = 2 * y; ty[r].s1 = 2 * y + 1; ty[r].w0 = 1; ty[r].w1 = 1; }
        else ty[r] = linear_tap(min(y, rh - 1), H, scale_y);
        const bool live = y < rh;
        g[2 * r] = base + (size_t)ty[r].s0 * row_bytes;
        g[2 * r + 1] = base + (size_t)ty[r].s1 * row_bytes;
        need[2 * r] = live;
        need[2 * r + 1] = live && (area2x || ty[r].w1 != 0);
    }
    bool tail = false;
    unsigned total = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        al[k] = (size_t)g[k] & 15;
        nb[k] = (unsigned)((al[k] + row_bytes + 15) & ~(size_t)15);
        if (need[k]) { total += nb[k]; tail = tail || (g[k] - al[k] + nb[k] > src_end); }
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (total > 0) {
        if (tail) {
            for (int k = 0; k < 4; ++k)
                if (need[k]) for (size_t i = threadIdx.x; i < row_bytes; i += LB_THREADS) smem[k * pitch + al[k] + i] = g[k][i];
            __syncthreads();
        } else {
            if (threadIdx.x == 0) {
                mbar_expect_tx(&bar, total);
#pragma unroll
                for (int k = 0; k < 4; ++k) if (need[k]) bulk_g2s(smem + k * pitch, g[k] - al[k], nb[k], &bar);
            }
            mbar_wait(&bar, 0);
        }
    }
    const int c0 = swap_rb ? 2 : 0, c2 = swap_rb ? 0 : 2;
    for (int fx = threadIdx.x; fx < S2; fx += LB_THREADS) {
        float px[16];
#pragma unroll
        for (int k = 12; k < 16; ++k) px[k] = 0.0f;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int x = 2 * fx + dx;
            Tap tx;
            if (!area2x) tx = linear_tap(min(x, rw - 1), W, scale_x);
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const int y = 2 * fy + dy;
                const int patch = dx * 2 + dy;   // (tl, bl, tr, br) = (dx,dy) (0,0),(0,1),(1,0),(1,1)
                float v[3] = {padf, padf, padf};
                if (x < rw && y < rh) {
                    const unsigned char* r0 = smem + (2 * dy) * pitch + al[2 * dy];
                    const unsigned char* r1 = need[2 * dy + 1] ? smem + (2 * dy + 1) * pitch + al[2 * dy + 1] : r0;
                    if (area2x) {
                        const unsigned char* p0 = r0 + 6 * x;
                        const unsigned char* p1 = r1 + 6 * x;
#pragma unroll
                        for (int c = 0; c < 3; ++c) v[c] = (float)((p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2);
                    } else {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int h0 = r0[3 * tx.s0 + c] * tx.w0 + r0[3 * tx.s1 + c] * tx.w1;
                            const int h1 = r1[3 * tx.s0 + c] * tx.w0 + r1[3 * tx.s1 + c] * tx.w1;
                            const int o = (((ty[dy].w0 * (h0 >> 4)) >> 16) + ((ty[dy].w1 * (h1 >> 4)) >> 16) + 2) >> 2;
                            v[c] = (float)min(max(o, 0), 255);
                        }
                    }
                }
                px[patch * 3 + 0] = v[c0]; px[patch * 3 + 1] = v[1]; px[patch * 3 + 2] = v[c2];
            }
        }
        OutT o16[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) o16[k] = cvt_out<OutT>(px[k]);
        uint4* d4 = reinterpret_cast<uint4*>(orow + (size_t)fx * pix_pitch);
        const uint4* s4 = reinterpret_cast<const uint4*>(o16);
#pragma unroll
        for (int k = 0; k < (int)(16 * sizeof(OutT) / 16); ++k) d4[k] = s4[k];
    }
}

}  // namespace

extern "C" int tk_letterbox_u8(const unsigned char* src, int n_frames, int H, int W, long long frame_stride_bytes,
                               void* dst, int out_dtype, int out_layout, int S, int pad_value, int swap_rb,
                               double* ratio_out, void* stream) {
    if (!src || !dst || n_frames <= 0 || H <= 0 || W <= 0 || S <= 0) return TK_ERR_ARG;
    if (out_dtype != TK_DTYPE_F32 && out_dtype != TK_DTYPE_BF16) return TK_ERR_ARG;
    // rtmlib: ratio = min(S/h, S/w) in Python floats; resized size = int(dim * ratio)
    const double ratio = fmin((double)S / (double)H, (double)S / (double)W);
    const int rw = (int)((double)W * ratio), rh = (int)((double)H * ratio);
    if (rw <= 0 || rh <= 0 || rw > S || rh > S) return TK_ERR_ARG;
    if (out_layout < 0 || out_layout > 3 || (out_layout >= 2 && (S & 1))) return TK_ERR_ARG;
    if (ratio_out) *ratio_out = ratio;
    const double scale_x = 1.0 / ((double)rw / (double)W), scale_y = 1.0 / ((double)rh / (double)H);
    const int area2x = (W == 2 * rw && H == 2 * rh) ? 1 : 0;
    const size_t pitch = ((size_t)W * 3 + 31 + 15) & ~(size_t)15;
    const size_t smem = 2 * pitch + 16;
    if (smem > 200 * 1024) return TK_ERR_CAPACITY;
    if (((size_t)src & 15) != 0) return TK_ERR_ARG;  // bulk copies round spans down to 16 B
    const unsigned char* src_end = src + (size_t)(n_frames - 1) * (size_t)frame_stride_bytes + (size_t)H * W * 3;
    cudaStream_t st = (cudaStream_t)stream;
    if (out_layout >= 2) {
        const int pix_pitch = out_layout == 3 ? 32 : 16;
        const size_t smem4 = 4 * pitch + 16;
        if (smem4 > 200 * 1024) return TK_ERR_CAPACITY;
        dim3 grid2(S / 2, n_frames);
        if (out_dtype == TK_DTYPE_F32) {
            TK_CUDA_TRY(cudaFuncSetAttribute(letterbox_focus16_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
            letterbox_focus16_kernel<float><<<grid2, LB_THREADS, smem4, st>>>(src, (size_t)frame_stride_bytes, H, W, (float*)dst, S, rw, rh,
                                                                              scale_x, scale_y, area2x, pad_value, swap_rb, src_end, pix_pitch);
        } else {
            TK_CUDA_TRY(cudaFuncSetAttribute(letterbox_focus16_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
            letterbox_focus16_kernel<__nv_bfloat16><<<grid2, LB_THREADS, smem4, st>>>(src, (size_t)frame_stride_bytes, H, W,
                                                                                      (__nv_bfloat16*)dst, S, rw, rh, scale_x, scale_y,
                                                                                      area2x, pad_value, swap_rb, src_end, pix_pitch);
        }
        TK_CUDA_TRY(cudaGetLastError());
        return TK_OK;
    }
    dim3 grid(S, n_frames);
    if (out_dtype == TK_DTYPE_F32) {
        TK_CUDA_TRY(cudaFuncSetAttribute(letterbox_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        letterbox_kernel<float><<<grid, LB_THREADS, smem, st>>>(src, (size_t)frame_stride_bytes, H, W, (float*)dst, S, rw, rh,
                                                               scale_x, scale_y, area2x, pad_value, swap_rb, out_layout, src_end);
    } else {
        TK_CUDA_TRY(cudaFuncSetAttribute(letterbox_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        letterbox_kernel<__nv_bfloat16><<<grid, LB_THREADS, smem, st>>>(src, (size_t)frame_stride_bytes, H, W,
                                                                       (__nv_bfloat16*)dst, S, rw, rh, scale_x, scale_y,
                                                                       area2x, pad_value, swap_rb, out_layout, src_end);
    }
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// ReID crop gather: crop (StrongSORT rule) -> PIL-exact antialiased bilinear resize to out_h x out_w -> /255 -> mean/std
// Replaces the per-crop CPU path /root/reference/plugins/track/strong_sort/strong_sort.py:102-108,135-145 +
// /root/reference/plugins/track/strong_sort/reid_multibackend.py:45-52,184-195 (T.Resize on a PIL image, ToTensor,
// Normalize). Pillow's resampler is two-pass (horizontal then vertical) with 22-bit fixed-point coefficients and a
// uint8 intermediate image; the kernel evaluates out(y,x) = clip8(sum_ky kv[ky] * clip8(sum_kx kh[kx] * src)) directly,
// re-deriving the few horizontal taps per vertical tap instead of materialising the intermediate — integer-exact.
// One CTA per (output row, crop): the vertical taps are shared by the row, each thread owns one output column.
namespace {

constexpr int CR_ROWS = 16;   // output rows per CTA of the crop / frame resize kernels
constexpr int CR_HROWS = 48; // source rows resampled horizontally into shared memory per CTA (separable crop path)
constexpr int CR_HCOLS = 128; // output columns of the separable crop path
constexpr int CR_KMAX = 64;   // taps per axis: 2*ceil(scale)+1 -> crops up to ~31x the output size (a full 4K row into 128 px)

// Pillow precompute_coeffs + normalize_coeffs_8bpc for one output index (bilinear filter, support 1)
__device__ __forceinline__ int pil_taps(int in_size, int out_size, int xx, int* k_int, int& xmin_out) {
    const double scale = (double)in_size / (double)out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    const double center = __dmul_rn((double)xx + 0.5, scale);
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    if (xmax > CR_KMAX) xmax = CR_KMAX;
    double w[CR_KMAX];
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
        const double a = fabs(__dmul_rn(__dadd_rn((double)(x + xmin) - center, 0.5), ss));
        const double v = a < 1.0 ? 1.0 - a : 0.0;
        w[x] = v;
        ww += v;
    }
    for (int x = 0; x < xmax; ++x) {
        const double v = ww != 0.0 ? w[x] / ww : w[x];
        k_int[x] = v < 0 ? (int)(-0.5 + v * 4194304.0) : (int)(0.5 + v * 4194304.0);
    }
    xmin_out = xmin;
    return xmax;
}

__device__ __forceinline__ int clip8(int v) { v >>= 22; return v < 0 ? 0 : (v > 255 ? 255 : v); }

template <typename OutT>
__global__ void __launch_bounds__(128)
crop_resize_norm_kernel(const unsigned char* __restrict__ frames, size_t frame_stride, int H, int W,
                        const double* __restrict__ dets, const int* __restrict__ det_frame, OutT* __restrict__ out,
                        int out_h, int out_w, float m0, float m1, float m2, float s0, float s1, float s2, int nhwc,
                        int crop_rule) {
    __shared__ int kv[CR_ROWS][CR_KMAX];
    __shared__ int s_ymin[CR_ROWS], s_ny[CR_ROWS];
    const int n = blockIdx.y, y0 = blockIdx.x * CR_ROWS;
    const double* d = dets + (size_t)n * 7;
    int x1, x2, y1, y2;
    if (crop_rule == TK_CROP_RULE_LTWH_ROUNDED) {
        // ReID wrapper rule (kpreid_api.py:118-121 -> utils/__init__.py:47-48 -> coordinates.py:216-267): the detector's float32
        // bbox_ltwh is clipped in place (sanitize_bbox_ltwh), turned into ltrb in float32 and rounded half-to-even; crop = [t:b, l:r]
        const float l = fmaxf(0.0f, fminf((float)d[0], (float)(W - 2))), t = fmaxf(0.0f, fminf((float)d[1], (float)(H - 2)));
        const float w = fmaxf(1.0f, fminf((float)d[2], __fsub_rn((float)(W - 1), l)));
        const float h = fmaxf(1.0f, fminf((float)d[3], __fsub_rn((float)(H - 1), t)));
        x1 = __float2int_rn(l); y1 = __float2int_rn(t);
        x2 = __float2int_rn(__fadd_rn(l, w)); y2 = __float2int_rn(__fadd_rn(t, h));
    } else if (crop_rule == TK_CROP_RULE_XYXY_INT) {
        // Deep OC-SORT / BoT-SORT rule (deep_oc_sort/ocsort.py:560-565): box.astype(int), then the NumPy slice img[y1:y2, x1:x2]
        x1 = max((int)d[0], 0); y1 = max((int)d[1], 0);
        x2 = min(max((int)d[2], 0), W); y2 = min(max((int)d[3], 0), H);
    } else {
        // StrongSORT crop rule (strong_sort.py:102-108): centre box -> int() truncation -> clip
        const double cx = (d[0] + d[2]) / 2, cy = (d[1] + d[3]) / 2, bw = d[2] - d[0], bh = d[3] - d[1];
        x1 = max((int)(cx - bw / 2), 0); x2 = min((int)(cx + bw / 2), W - 1);
        y1 = max((int)(cy - bh / 2), 0); y2 = min((int)(cy + bh / 2), H - 1);
    }
    const int cw = x2 - x1, ch = y2 - y1;
    const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
    const unsigned char* img = frames + (size_t)det_frame[n] * frame_stride;
    // the vertical taps of the CTA's rows (one thread each); the horizontal taps of a column are derived once per thread and
    // reused for all CR_ROWS rows (they only depend on the crop width and the column: ncu showed the one-row version
    // issue-bound on re-deriving them in double precision for every output pixel)
    if (threadIdx.x < CR_ROWS && y0 + threadIdx.x < out_h && cw > 0 && ch > 0) {
        int ym;
        s_ny[threadIdx.x] = pil_taps(ch, out_h, y0 + threadIdx.x, kv[threadIdx.x], ym);
        s_ymin[threadIdx.x] = ym;
    }
    __syncthreads();
    // Separable form (Pillow itself runs the horizontal pass first and rounds it to 8 bits): the CTA's output rows need the source
    // rows [ylo, yhi); each of them is resampled horizontally ONCE into shared memory (uint8, like Pillow's intermediate image)
    // instead of once per output row whose vertical support contains it (x2-3 fewer source loads at the usual scales), then the
    // vertical pass runs out of shared memory. Falls back to the direct form when the rows do not fit (strong down-scaling).
    __shared__ unsigned char hbuf[CR_HROWS][CR_HCOLS][3];
    __shared__ int s_lo, s_hi;
    const int n_rows = min(CR_ROWS, out_h - y0);
    if (threadIdx.x == 0) {
        int lo = 0x7fffffff, hi = 0;
        if (cw > 0 && ch > 0)
            for (int r = 0; r < n_rows; ++r) { lo = min(lo, s_ymin[r]); hi = max(hi, s_ymin[r] + s_ny[r]); }
        s_lo = lo; s_hi = hi;
    }
    __syncthreads();
    const int ylo = s_lo, n_src = s_hi - s_lo;
    const bool separable = cw > 0 && ch > 0 && n_src <= CR_HROWS && out_w <= CR_HCOLS;
    for (int xx0 = 0; xx0 < out_w; xx0 += blockDim.x) {
        const int xx = xx0 + threadIdx.x;
        int kh[CR_KMAX], xmin = 0, nx = 0;
        if (xx < out_w && cw > 0 && ch > 0) nx = pil_taps(cw, out_w, xx, kh, xmin);
        if (separable && xx < out_w) {
            for (int rr = 0; rr < n_src; ++rr) {
                const unsigned char* row = img + ((size_t)(y1 + ylo + rr) * W + (x1 + xmin)) * 3;
                int h[3] = {1 << 21, 1 << 21, 1 << 21};
                for (int kx = 0; kx < nx; ++kx) {
                    h[0] += row[kx * 3 + 0] * kh[kx]; h[1] += row[kx * 3 + 1] * kh[kx]; h[2] += row[kx * 3 + 2] * kh[kx];
                }
                hbuf[rr][xx][0] = (unsigned char)clip8(h[0]); hbuf[rr][xx][1] = (unsigned char)clip8(h[1]); hbuf[rr][xx][2] = (unsigned char)clip8(h[2]);
            }
        }
        // (no barrier needed: a thread only reads the hbuf column it wrote)
        float o2[2][3];
        for (int r = 0; r < n_rows; ++r) {
            const int yy = y0 + r;
            int v[3] = {0, 0, 0};
            if (xx < out_w && cw > 0 && ch > 0) {
                int acc[3] = {1 << 21, 1 << 21, 1 << 21};
                if (separable) {
                    const int rb = s_ymin[r] - ylo;
                    for (int ky = 0; ky < s_ny[r]; ++ky) {
                        const int w = kv[r][ky];
                        acc[0] += hbuf[rb + ky][xx][0] * w; acc[1] += hbuf[rb + ky][xx][1] * w; acc[2] += hbuf[rb + ky][xx][2] * w;
                    }
                } else {
                    for (int ky = 0; ky < s_ny[r]; ++ky) {
                        const unsigned char* row = img + ((size_t)(y1 + s_ymin[r] + ky) * W + (x1 + xmin)) * 3;
                        int h[3] = {1 << 21, 1 << 21, 1 << 21};
                        // (when a size already matches, Pillow skips that pass; the taps then are {1<<22, 0}, i.e. the identity)
                        for (int kx = 0; kx < nx; ++kx) {
                            h[0] += row[kx * 3 + 0] * kh[kx]; h[1] += row[kx * 3 + 1] * kh[kx]; h[2] += row[kx * 3 + 2] * kh[kx];
                        }
#pragma unroll
                        for (int c = 0; c < 3; ++c) acc[c] += clip8(h[c]) * kv[r][ky];
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = clip8(acc[c]);
            }
            float o[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float t = __fdiv_rn((float)v[c], 255.0f);                 // ToTensor
                o[c] = __fdiv_rn(__fsub_rn(t, mean[c]), stdv[c]);               // Normalize
            }
            if (nhwc == TK_CROP_LAYOUT_S2D16 && sizeof(OutT) == 2 && (n_rows & 1) == 0) {
                // 2x2 space-to-depth, 16-channel pitch: the quad (rows yy, yy+1; columns xx, xx+1) is 12 contiguous bf16 of one
                // 32-byte group -> the even lane collects its neighbour's values and writes 16 + 8 bytes instead of 12 2-byte stores
#pragma unroll
                for (int c = 0; c < 3; ++c) o2[r & 1][c] = o[c];
                if (r & 1) {
                    float nb[2][3];
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int c = 0; c < 3; ++c) nb[q][c] = __shfl_down_sync(0xffffffffu, o2[q][c], 1);
                    if (!(xx & 1) && xx < out_w) {
                        const size_t g = (((size_t)n * (out_h / 2 + 3) + ((yy - 1) >> 1) + 2) * (out_w / 2 + 3) + (xx >> 1) + 2) << 4;
                        __nv_bfloat162 w01 = __floats2bfloat162_rn(o2[0][0], o2[0][1]), w23 = __floats2bfloat162_rn(o2[0][2], nb[0][0]);
                        __nv_bfloat162 w45 = __floats2bfloat162_rn(nb[0][1], nb[0][2]), w67 = __floats2bfloat162_rn(o2[1][0], o2[1][1]);
                        __nv_bfloat162 w89 = __floats2bfloat162_rn(o2[1][2], nb[1][0]), wab = __floats2bfloat162_rn(nb[1][1], nb[1][2]);
                        uint4 a4; uint2 b2;
                        a4.x = *reinterpret_cast<unsigned*>(&w01); a4.y = *reinterpret_cast<unsigned*>(&w23);
                        a4.z = *reinterpret_cast<unsigned*>(&w45); a4.w = *reinterpret_cast<unsigned*>(&w67);
                        b2.x = *reinterpret_cast<unsigned*>(&w89); b2.y = *reinterpret_cast<unsigned*>(&wab);
                        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + g) = a4;
                        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + g + 8) = b2;
                    }
                }
                continue;
            }
            if (xx >= out_w) continue;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                size_t idx;
                if (nhwc == TK_CROP_LAYOUT_S2D16)   // 2x2 space-to-depth, 16-channel pitch, zero border of 2 before / 1 after (see trackkern.h)
                    idx = ((((size_t)n * (out_h / 2 + 3) + (yy >> 1) + 2) * (out_w / 2 + 3) + (xx >> 1) + 2) << 4) + (((yy & 1) * 2 + (xx & 1)) * 3 + c);
                else
                    idx = nhwc ? (((size_t)n * out_h + yy) * out_w + xx) * nhwc + c   // nhwc = channel pitch (3, or 8 with zero padding)
                               : (((size_t)n * 3 + c) * out_h + yy) * out_w + xx;
                out[idx] = cvt_out<OutT>(o[c]);
            }
        }
    }
}

// Whole-frame Pillow-exact antialiased bilinear resize (no aspect preservation) + rescale: the RT-DETR image processor
// (transformers RTDetrImageProcessor: resize to 640x640, x 1/255, no normalisation) behind
// /root/reference/tracklab/wrappers/bbox_detector/transformers_api.py:32. One CTA per (output row, frame); planar output.
template <typename OutT>
__global__ void __launch_bounds__(128)
resize_frames_kernel(const unsigned char* __restrict__ frames, size_t frame_stride, int H, int W, OutT* __restrict__ out,
                     int out_h, int out_w, float scale) {
    __shared__ int kv[CR_ROWS][CR_KMAX];
    __shared__ int s_ymin[CR_ROWS], s_ny[CR_ROWS];
    const int n = blockIdx.y, y0 = blockIdx.x * CR_ROWS;
    const unsigned char* img = frames + (size_t)n * frame_stride;
    if (threadIdx.x < CR_ROWS && y0 + threadIdx.x < out_h) {
        int ym;
        s_ny[threadIdx.x] = pil_taps(H, out_h, y0 + threadIdx.x, kv[threadIdx.x], ym);
        s_ymin[threadIdx.x] = ym;
    }
    __syncthreads();
    for (int xx = threadIdx.x; xx < out_w; xx += blockDim.x) {
        int kh[CR_KMAX], xmin;
        const int nx = pil_taps(W, out_w, xx, kh, xmin);   // once per column, reused for the CTA's rows
        for (int r = 0; r < CR_ROWS && y0 + r < out_h; ++r) {
            int acc[3] = {1 << 21, 1 << 21, 1 << 21};
            for (int ky = 0; ky < s_ny[r]; ++ky) {
                const unsigned char* row = img + ((size_t)(s_ymin[r] + ky) * W + xmin) * 3;
                int h[3] = {1 << 21, 1 << 21, 1 << 21};
                for (int kx = 0; kx < nx; ++kx) {
                    h[0] += row[kx * 3 + 0] * kh[kx]; h[1] += row[kx * 3 + 1] * kh[kx]; h[2] += row[kx * 3 + 2] * kh[kx];
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[c] += clip8(h[c]) * kv[r][ky];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
                out[(((size_t)n * 3 + c) * out_h + y0 + r) * out_w + xx] = cvt_out<OutT>(__fmul_rn((float)clip8(acc[c]), scale));
        }
    }
}

}  // namespace

extern "C" int tk_resize_frames_u8(const unsigned char* frames, int n_frames, int H, int W, long long frame_stride_bytes, void* out,
                                   int out_dtype, int out_h, int out_w, float scale, void* stream) {
    if (!frames || !out || n_frames < 0 || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0) return TK_ERR_ARG;
    if (out_dtype != TK_DTYPE_F32 && out_dtype != TK_DTYPE_BF16) return TK_ERR_ARG;
    if (n_frames == 0) return TK_OK;
    if (2 * ((W + out_w - 1) / out_w) + 1 > CR_KMAX || 2 * ((H + out_h - 1) / out_h) + 1 > CR_KMAX) return TK_ERR_CAPACITY;
    dim3 grid((out_h + CR_ROWS - 1) / CR_ROWS, n_frames);
    if (out_dtype == TK_DTYPE_F32)
        resize_frames_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(frames, (size_t)frame_stride_bytes, H, W, (float*)out, out_h, out_w, scale);
    else
        resize_frames_kernel<__nv_bfloat16><<<grid, 128, 0, (cudaStream_t)stream>>>(frames, (size_t)frame_stride_bytes, H, W, (__nv_bfloat16*)out, out_h,
                                                                                  out_w, scale);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

extern "C" int tk_crop_resize_norm(const unsigned char* frames, int H, int W, long long frame_stride_bytes, const double* dets,
                                   const int* det_frame, int n_dets, void* out, int out_dtype, int out_nhwc, int out_h, int out_w,
                                   const float* mean3, const float* std3, void* stream) {
    return tk_crop_resize_norm_ex(frames, H, W, frame_stride_bytes, dets, det_frame, n_dets, out, out_dtype, out_nhwc, out_h, out_w,
                                  mean3, std3, TK_CROP_RULE_STRONGSORT, stream);
}

extern "C" int tk_crop_resize_norm_ex(const unsigned char* frames, int H, int W, long long frame_stride_bytes, const double* dets,
                                      const int* det_frame, int n_dets, void* out, int out_dtype, int out_nhwc, int out_h, int out_w,
                                      const float* mean3, const float* std3, int crop_rule, void* stream) {
    if (crop_rule != TK_CROP_RULE_STRONGSORT && crop_rule != TK_CROP_RULE_LTWH_ROUNDED && crop_rule != TK_CROP_RULE_XYXY_INT) return TK_ERR_ARG;
    if (!frames || !dets || !det_frame || !out || !mean3 || !std3 || n_dets < 0 || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0) return TK_ERR_ARG;
    if (out_dtype != TK_DTYPE_F32 && out_dtype != TK_DTYPE_BF16) return TK_ERR_ARG;
    if (out_nhwc == TK_CROP_LAYOUT_S2D16 && ((out_h | out_w) & 1)) return TK_ERR_ARG;
    if (n_dets == 0) return TK_OK;
    if (2 * ((W + out_w - 1) / out_w) + 1 > CR_KMAX || 2 * ((H + out_h - 1) / out_h) + 1 > CR_KMAX) return TK_ERR_CAPACITY;
    dim3 grid((out_h + CR_ROWS - 1) / CR_ROWS, n_dets);
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == TK_DTYPE_F32)
        crop_resize_norm_kernel<float><<<grid, 128, 0, st>>>(frames, (size_t)frame_stride_bytes, H, W, dets, det_frame, (float*)out, out_h,
                                                            out_w, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], out_nhwc, crop_rule);
    else
        crop_resize_norm_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(frames, (size_t)frame_stride_bytes, H, W, dets, det_frame,
                                                                    (__nv_bfloat16*)out, out_h, out_w, mean3[0], mean3[1], mean3[2],
                                                                    std3[0], std3[1], std3[2], out_nhwc, crop_rule);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}
