// RT-DETR post-processing on device: sigmoid focal scores, top-Q over Q x C, box decoding, threshold, class filter,
// sanitising and ltwh conversion — one CTA per image.
//
// Replaces, behind /root/reference/tracklab/wrappers/bbox_detector/transformers_api.py:33-54,
//   transformers RTDetrImageProcessor.post_process_object_detection (use_focal_loss=True):
//       boxes = center_to_corners(pred_boxes) * [W,H,W,H]            (float32)
//       scores, index = topk(sigmoid(logits).flatten(1), Q); labels = index % C; query = index // C
//       keep score > threshold
//   the wrapper's `label == 0` filter and ltrb_to_ltwh(box.numpy(), (W, H))
//       (/root/reference/tracklab/utils/coordinates.py:270-295,318-328: clamp l,t to [0, dim-2], r,b to [1, dim-1], float32).
// Selection: 4-pass radix select of the Q-th largest score key (positive floats order like their bit patterns), ties at
// the threshold key taken in ascending flat index, then a bitonic sort of the selected (score desc, flat index asc).
// torch.topk leaves the order of exactly equal scores unspecified; this kernel's order is deterministic.
#include "tk_common.cuh"
#include "trackkern.h"

namespace {

using namespace tk;

constexpr int RD_THREADS = 256;
constexpr int RD_QMAX = 1024;

__device__ __forceinline__ unsigned score_key(float logit) {
    const float s = 1.0f / (1.0f + expf(-logit));   // torch.sigmoid in float32
    return __float_as_uint(s);
}

__global__ void __launch_bounds__(RD_THREADS)
rtdetr_decode_kernel(const float* __restrict__ logits, const float* __restrict__ boxes, int Q, int C, float img_w, float img_h,
                     float threshold, int keep_label, double* __restrict__ rows_out, int* __restrict__ counts_out) {
    __shared__ unsigned hist[256];
    __shared__ unsigned long long sel[RD_QMAX];
    __shared__ unsigned s_prefix, s_mask, s_need, s_cnt;
    __shared__ unsigned char keep[RD_QMAX];
    const int img = blockIdx.x, tid = threadIdx.x, N = Q * C;
    const float* lg = logits + (size_t)img * N;
    if (tid == 0) { s_prefix = 0u; s_mask = 0u; s_need = (unsigned)min(Q, N); s_cnt = 0u; }
    __syncthreads();
    // ---- radix select: key of the Q-th largest score
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int b = tid; b < 256; b += RD_THREADS) hist[b] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix, mask = s_mask;
        for (int i = tid; i < N; i += RD_THREADS) {
            const unsigned k = score_key(lg[i]);
            if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned need = s_need, above = 0u;
            int b = 255;
            for (; b > 0; --b) { if (above + hist[b] >= need) break; above += hist[b]; }
            s_need = need - above;
            s_prefix = prefix | ((unsigned)b << shift);
            s_mask = mask | (255u << shift);
        }
        __syncthreads();
    }
    const unsigned T = s_prefix, need_eq = s_need;   // take every key > T and the first need_eq keys == T
    const int n_sel = min(Q, N);
    for (int i = tid; i < RD_QMAX; i += RD_THREADS) sel[i] = 0ull;
    __syncthreads();
    for (int i = tid; i < N; i += RD_THREADS) {
        const unsigned k = score_key(lg[i]);
        if (k > T) sel[atomicAdd(&s_cnt, 1u)] = ((unsigned long long)k << 32) | (unsigned)(0xffffffffu - (unsigned)i);
    }
    __syncthreads();
    if (tid < 32) {   // ties at the threshold key in ascending flat index
        unsigned got = 0u;
        const unsigned base = s_cnt;
        for (int i0 = 0; i0 < N && got < need_eq; i0 += 32) {
            const int i = i0 + tid;
            const bool eq = i < N && score_key(lg[i]) == T;
            const unsigned m = __ballot_sync(0xffffffffu, eq);
            const unsigned rank = got + __popc(m & ((1u << tid) - 1u));
            if (eq && rank < need_eq) sel[base + rank] = ((unsigned long long)T << 32) | (unsigned)(0xffffffffu - (unsigned)i);
            got += __popc(m);
        }
    }
    __syncthreads();
    // ---- bitonic sort, descending, of the padded selection
    int P = 1;
    while (P < n_sel) P <<= 1;
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += RD_THREADS) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = sel[i], b = sel[l];
                    const bool desc = (i & k) == 0;
                    if (desc ? a < b : a > b) { sel[i] = b; sel[l] = a; }
                }
            }
            __syncthreads();
        }
    // ---- threshold + class filter, ordered compaction, decode
    for (int p = tid; p < n_sel; p += RD_THREADS) {
        const unsigned k = (unsigned)(sel[p] >> 32);
        const int i = (int)(0xffffffffu - (unsigned)(sel[p] & 0xffffffffull));
        const float sc = __uint_as_float(k);
        keep[p] = (sc > threshold) && (keep_label < 0 || (i % C) == keep_label);
    }
    __syncthreads();
    if (tid < 32) {
        int base = 0;
        for (int p0 = 0; p0 < n_sel; p0 += 32) {
            const int p = p0 + tid;
            const bool f = p < n_sel && keep[p];
            const unsigned m = __ballot_sync(0xffffffffu, f);
            if (f) {
                const int o = base + __popc(m & ((1u << tid) - 1u));
                const unsigned k = (unsigned)(sel[p] >> 32);
                const int i = (int)(0xffffffffu - (unsigned)(sel[p] & 0xffffffffull));
                const int q = i / C;
                const float* b = boxes + ((size_t)img * Q + q) * 4;
                // center_to_corners_format, then * [W,H,W,H], all float32
                float l = __fmul_rn(__fsub_rn(b[0], __fmul_rn(0.5f, b[2])), img_w), t = __fmul_rn(__fsub_rn(b[1], __fmul_rn(0.5f, b[3])), img_h);
                float r = __fmul_rn(__fadd_rn(b[0], __fmul_rn(0.5f, b[2])), img_w), bt = __fmul_rn(__fadd_rn(b[1], __fmul_rn(0.5f, b[3])), img_h);
                // sanitize_bbox_ltrb on the float32 array (coordinates.py:288-292), then ltwh in float32
                l = fmaxf(0.0f, fminf(l, img_w - 2.0f)); t = fmaxf(0.0f, fminf(t, img_h - 2.0f));
                r = fmaxf(1.0f, fminf(r, img_w - 1.0f)); bt = fmaxf(1.0f, fminf(bt, img_h - 1.0f));
                double* row = rows_out + ((size_t)img * Q + o) * 6;
                row[0] = (double)l; row[1] = (double)t; row[2] = (double)__fsub_rn(r, l); row[3] = (double)__fsub_rn(bt, t);
                row[4] = (double)__uint_as_float(k);
                row[5] = (double)(keep_label < 0 ? i : q);
            }
            base += __popc(m);
        }
        if (tid == 0) counts_out[img] = base;
    }
}

}  // namespace

extern "C" int tk_rtdetr_decode(const float* logits, const float* boxes, int n_images, int n_queries, int n_classes, int img_w, int img_h,
                                float threshold, int keep_label, double* rows_out, int* counts_out, void* stream) {
    if (!logits || !boxes || !rows_out || !counts_out || n_images < 0 || n_queries <= 0 || n_classes <= 0 || img_w <= 2 || img_h <= 2)
        return TK_ERR_ARG;
    if (n_queries > RD_QMAX) return TK_ERR_CAPACITY;
    if (n_images == 0) return TK_OK;
    rtdetr_decode_kernel<<<n_images, RD_THREADS, 0, (cudaStream_t)stream>>>(logits, boxes, n_queries, n_classes, (float)img_w, (float)img_h,
                                                                           threshold, keep_label, rows_out, counts_out);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}
