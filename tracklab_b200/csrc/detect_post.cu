// Detector post-processing: YOLOX decode + score threshold + class-aware NMS, and the packing of the
// surviving boxes into the tracker-input rows the reference wrappers build.
//
// Replaces (rtmlib 0.0.13, un-vendored — SURVEY.md §3.2 [3P-memory], restated in oracle/yolox_post_np.py):
//   YOLOX.postprocess: (xy + grid) * stride, exp(wh) * stride, score = obj * cls, cxcywh -> xyxy, / ratio,
//   multiclass_nms(nms_thr=0.45, score_thr=0.7) with the "+1" float32 overlap and `ovr <= nms_thr` keep rule
// and the wrapper /root/reference/tracklab/wrappers/bbox_detector/rtmlib_api.py:27-46
//   (ltrb_to_ltwh with clipping, /root/reference/tracklab/utils/coordinates.py:270-295,318-328;
//    bbox_conf hard-coded to 1.0 at rtmlib_api.py:38; running detection id :42-45)
// followed by the tracker wrapper's row layout /root/reference/tracklab/wrappers/track/oc_sort_api.py:33-47.
//
// One CTA per image: the 8400x(5+nc) prediction block is streamed once with coalesced loads (HBM-bound,
// 8400*(5+nc)*4 B per image), candidates above the threshold are compacted into shared memory, sorted by
// score with a bitonic network, the suppression relation is built as a bit matrix in shared memory and a
// single warp walks it. Nothing goes back to the host.
#include <cuda_bf16.h>
#include "tk_common.cuh"
#include "trackkern.h"

namespace {

constexpr int NMS_THREADS = 256;
constexpr int NMS_CAP = 1024;  // candidates above the score threshold per image

struct Cand { float score; int key; };  // key = anchor * nc + class

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// anchor index -> (grid x, grid y, stride) for strides 8/16/32 on an SxS input (YOLOX head order)
__device__ __forceinline__ void anchor_geom(int a, int S, int& gx, int& gy, int& stride) {
    const int n8 = (S / 8) * (S / 8), n16 = (S / 16) * (S / 16);
    int w;
    if (a < n8) { stride = 8; w = S / 8; }
    else if (a < n8 + n16) { a -= n8; stride = 16; w = S / 16; }
    else { a -= n8 + n16; stride = 32; w = S / 32; }
    gy = a / w; gx = a % w;
}

template <typename T>
__global__ void __launch_bounds__(NMS_THREADS)
yolox_nms_kernel(const T* __restrict__ pred, int A, int nc, int S, int logits, float ratio, float score_thr, float nms_thr,
                 int max_out, float* __restrict__ out_boxes, float* __restrict__ out_scores, int* __restrict__ out_cls,
                 int* __restrict__ out_count, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem[];
    Cand* cand = (Cand*)smem;                                 // [NMS_CAP]
    float4* box = (float4*)(cand + NMS_CAP);                  // [NMS_CAP]
    unsigned* mask = (unsigned*)(box + NMS_CAP);              // [n][words]
    __shared__ int s_n;
    __shared__ int s_keep[NMS_CAP];
    __shared__ int s_nkeep;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int C = 5 + nc;
    const T* p = pred + (size_t)b * A * C;
    if (tid == 0) s_n = 0;
    __syncthreads();
    // ---- 1. threshold + compaction ---------------------------------------------------------------
    for (int a = tid; a < A; a += NMS_THREADS) {
        const T* row = p + (size_t)a * C;
        float obj = ldf(row + 4);
        if (logits) obj = sigmoidf_(obj);
        for (int c = 0; c < nc; ++c) {
            float cl = ldf(row + 5 + c);
            if (logits) cl = sigmoidf_(cl);
            const float sc = __fmul_rn(obj, cl);
            if (sc > score_thr) {
                const int k = atomicAdd(&s_n, 1);
                if (k < NMS_CAP) { cand[k].score = sc; cand[k].key = a * nc + c; }
            }
        }
    }
    __syncthreads();
    int n = s_n;
    if (n > NMS_CAP) { if (tid == 0) atomicOr(status, TK_DEV_OVERFLOW_DETS); n = NMS_CAP; }
    // ---- 2. bitonic sort: score descending, key ascending on ties -----------------------------------
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = n + tid; i < np2; i += NMS_THREADS) { cand[i].score = -1.0f; cand[i].key = 0x7fffffff; }
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < np2; i += NMS_THREADS) {
                const int l = i ^ j;
                if (l > i) {
                    const Cand x = cand[i], y = cand[l];
                    const bool x_first = (x.score > y.score) || (x.score == y.score && x.key < y.key);
                    const bool up = ((i & k) == 0);
                    if (up ? !x_first : x_first) { cand[i] = y; cand[l] = x; }
                }
            }
            __syncthreads();
        }
    }
    // ---- 3. decode the candidates' boxes (float32, original-image coordinates) -------------------------
    for (int i = tid; i < n; i += NMS_THREADS) {
        const int a = cand[i].key / nc;
        const T* row = p + (size_t)a * C;
        int gx, gy, st;
        anchor_geom(a, S, gx, gy, st);
        const float fs = (float)st;
        const float cx = __fmul_rn(__fadd_rn(ldf(row + 0), (float)gx), fs);
        const float cy = __fmul_rn(__fadd_rn(ldf(row + 1), (float)gy), fs);
        const float w = __fmul_rn(expf(ldf(row + 2)), fs);
        const float h = __fmul_rn(expf(ldf(row + 3)), fs);
        float4 bx;
        bx.x = __fdiv_rn(__fsub_rn(cx, __fdiv_rn(w, 2.0f)), ratio);
        bx.y = __fdiv_rn(__fsub_rn(cy, __fdiv_rn(h, 2.0f)), ratio);
        bx.z = __fdiv_rn(__fadd_rn(cx, __fdiv_rn(w, 2.0f)), ratio);
        bx.w = __fdiv_rn(__fadd_rn(cy, __fdiv_rn(h, 2.0f)), ratio);
        box[i] = bx;
    }
    __syncthreads();
    // ---- 4. suppression bit matrix: bit j of row i set when j > i, same class, overlap > nms_thr -----
    const int words = (n + 31) >> 5;
    for (int e = tid; e < n * words; e += NMS_THREADS) {
        const int i = e / words, wd = e % words;
        const float4 bi = box[i];
        const int ci = cand[i].key % nc;
        const float ai = __fmul_rn(__fadd_rn(__fsub_rn(bi.z, bi.x), 1.0f), __fadd_rn(__fsub_rn(bi.w, bi.y), 1.0f));
        unsigned bits = 0u;
        const int j0 = wd * 32;
        for (int t = 0; t < 32; ++t) {
            const int j = j0 + t;
            if (j <= i || j >= n) continue;
            if ((cand[j].key % nc) != ci) continue;
            const float4 bj = box[j];
            const float aj = __fmul_rn(__fadd_rn(__fsub_rn(bj.z, bj.x), 1.0f), __fadd_rn(__fsub_rn(bj.w, bj.y), 1.0f));
            const float iw = fmaxf(0.0f, __fadd_rn(__fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)), 1.0f));
            const float ih = fmaxf(0.0f, __fadd_rn(__fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)), 1.0f));
            const float inter = __fmul_rn(iw, ih);
            const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, aj), inter));
            if (!(ovr <= nms_thr)) bits |= 1u << t;
        }
        mask[(size_t)i * words + wd] = bits;
    }
    __syncthreads();
    // ---- 5. greedy walk by one warp ---------------------------------------------------------------------
    if (tid < 32) {
        unsigned removed = 0u;  // lane l owns words l, l+32 (n <= 1024 -> words <= 32)
        int nk = 0;
        for (int i = 0; i < n; ++i) {
            const unsigned r = __shfl_sync(0xffffffffu, removed, i >> 5);
            if ((r >> (i & 31)) & 1u) continue;
            if (tid == 0) s_keep[nk] = i;
            ++nk;
            if (tid < words) removed |= mask[(size_t)i * words + tid];
        }
        if (tid == 0) s_nkeep = nk;
    }
    __syncthreads();
    int nk = s_nkeep;
    if (nk > max_out) { if (tid == 0) atomicOr(status, TK_DEV_OVERFLOW_OUT); nk = max_out; }
    for (int k = tid; k < nk; k += NMS_THREADS) {
        const int i = s_keep[k];
        const float4 bx = box[i];
        float* ob = out_boxes + ((size_t)b * max_out + k) * 4;
        ob[0] = bx.x; ob[1] = bx.y; ob[2] = bx.z; ob[3] = bx.w;
        out_scores[(size_t)b * max_out + k] = cand[i].score;
        out_cls[(size_t)b * max_out + k] = cand[i].key % nc;
    }
    if (tid == 0) out_count[b] = nk;
}

// boxes float32 [B, K, 4] xyxy + counts -> tracker rows float64 [N,7] and frame offsets.
// ltrb_to_ltwh(bbox, (W,H)) then ltwh_to_ltrb, all in float32 exactly like the float32 ndarray the
// wrappers pass around; conf = fixed_conf when >= 0 (rtmlib_api.py:38) else the detector score.
__global__ void pack_detections_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                       const int* __restrict__ cls, const int* __restrict__ counts, int B, int K,
                                       int keep_class, float W, float H, double fixed_conf, double category_id,
                                       int* __restrict__ cursor, double* __restrict__ dets, int* __restrict__ offsets_all,
                                       int dets_cap, int frames_cap, int* __restrict__ status, int row_format,
                                       int* __restrict__ frame_of_row) {
    __shared__ int s_off[1025];
    __shared__ int s_cnt[1024];
    const int tid = threadIdx.x;
    // per-image number of rows of the kept class
    for (int b = tid; b < B; b += blockDim.x) {
        int n = 0;
        for (int k = 0; k < counts[b]; ++k) n += (keep_class < 0 || cls[(size_t)b * K + k] == keep_class) ? 1 : 0;
        s_cnt[b] = n;
    }
    __syncthreads();
    __shared__ int s_frame0;
    if (tid == 0) {
        int acc = cursor[0];
        s_frame0 = cursor[1];
        for (int b = 0; b < B; ++b) { s_off[b] = acc; acc += s_cnt[b]; }
        s_off[B] = acc;
        if (acc > dets_cap || s_frame0 + B > frames_cap) atomicOr(status, TK_DEV_OVERFLOW_OUT);
    }
    __syncthreads();
    if (s_off[B] > dets_cap || s_frame0 + B > frames_cap) return;
    int* offsets = offsets_all + s_frame0;
    for (int b = tid; b <= B; b += blockDim.x) offsets[b] = s_off[b];
    __syncthreads();
    if (tid == 0) { cursor[0] = s_off[B]; cursor[1] = s_frame0 + B; }   // next call appends after this batch
    for (int b = tid; b < B; b += blockDim.x) {
        int r = s_off[b];
        for (int k = 0; k < counts[b]; ++k) {
            if (!(keep_class < 0 || cls[(size_t)b * K + k] == keep_class)) continue;
            const float* bx = boxes + ((size_t)b * K + k) * 4;
            // sanitize_bbox_ltrb (coordinates.py:270-295)
            const float l = fmaxf(0.0f, fminf(bx[0], W - 2.0f));
            const float t = fmaxf(0.0f, fminf(bx[1], H - 2.0f));
            const float rr = fmaxf(1.0f, fminf(bx[2], W - 1.0f));
            const float bb = fmaxf(1.0f, fminf(bx[3], H - 1.0f));
            const float w = __fsub_rn(rr, l), h = __fsub_rn(bb, t);      // ltrb_to_ltwh (coordinates.py:318-328)
            double* d = dets + (size_t)r * 7;
            d[0] = (double)l; d[1] = (double)t;
            if (row_format == TK_ROWS_LTWH) {      // the detector's bbox_ltwh column as is (ReID / BPBReID wrappers read it)
                d[2] = (double)w; d[3] = (double)h;
            } else {
                d[2] = (double)__fadd_rn(l, w); d[3] = (double)__fadd_rn(t, h);  // ltwh_to_ltrb (coordinates.py:257-267)
            }
            if (frame_of_row) frame_of_row[r] = b;   // image index inside this batch (what tk_crop_resize_norm gathers from)
            d[4] = fixed_conf >= 0.0 ? fixed_conf : (double)scores[(size_t)b * K + k];
            d[5] = category_id;
            d[6] = (double)r;
            ++r;
        }
    }
}

}  // namespace

extern "C" {

int tk_yolox_nms(const void* pred, int pred_dtype, int n_images, int n_anchors, int n_classes, int input_size,
                 int logits, float ratio, float score_thr, float nms_thr, int max_out, float* out_boxes,
                 float* out_scores, int* out_cls, int* out_count, int* status_dev, void* stream) {
    if (!pred || !out_boxes || !out_scores || !out_cls || !out_count || !status_dev) return TK_ERR_ARG;
    if (n_images <= 0 || n_classes <= 0 || max_out <= 0 || max_out > NMS_CAP) return TK_ERR_ARG;
    const int S = input_size;
    if (n_anchors != (S / 8) * (S / 8) + (S / 16) * (S / 16) + (S / 32) * (S / 32)) return TK_ERR_ARG;
    const size_t smem = sizeof(Cand) * NMS_CAP + sizeof(float4) * NMS_CAP + sizeof(unsigned) * NMS_CAP * (NMS_CAP / 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (pred_dtype == TK_DTYPE_F32) {
        TK_CUDA_TRY(cudaFuncSetAttribute(yolox_nms_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        yolox_nms_kernel<float><<<n_images, NMS_THREADS, smem, st>>>((const float*)pred, n_anchors, n_classes, S, logits, ratio,
                                                                     score_thr, nms_thr, max_out, out_boxes, out_scores,
                                                                     out_cls, out_count, status_dev);
    } else if (pred_dtype == TK_DTYPE_BF16) {
        TK_CUDA_TRY(cudaFuncSetAttribute(yolox_nms_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        yolox_nms_kernel<__nv_bfloat16><<<n_images, NMS_THREADS, smem, st>>>((const __nv_bfloat16*)pred, n_anchors, n_classes, S,
                                                                             logits, ratio, score_thr, nms_thr, max_out,
                                                                             out_boxes, out_scores, out_cls, out_count, status_dev);
    } else {
        return TK_ERR_ARG;
    }
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_pack_detections_ex(const float* boxes, const float* scores, const int* cls, const int* counts, int n_images,
                          int max_per_image, int keep_class, int img_w, int img_h, double fixed_conf, double category_id,
                          int* cursor_dev, double* dets_out, int* offsets_out, int dets_cap, int frames_cap,
                          int* status_dev, int row_format, int* frame_of_row_out, void* stream) {
    if (!boxes || !scores || !cls || !counts || !dets_out || !offsets_out || !cursor_dev || !status_dev) return TK_ERR_ARG;
    if (n_images <= 0 || n_images > 1024) return TK_ERR_ARG;
    if (row_format != TK_ROWS_LTRB && row_format != TK_ROWS_LTWH) return TK_ERR_ARG;
    pack_detections_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(boxes, scores, cls, counts, n_images, max_per_image, keep_class,
                                                              (float)img_w, (float)img_h, fixed_conf, category_id,
                                                              cursor_dev, dets_out, offsets_out, dets_cap, frames_cap, status_dev,
                                                              row_format, frame_of_row_out);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_pack_detections(const float* boxes, const float* scores, const int* cls, const int* counts, int n_images,
                       int max_per_image, int keep_class, int img_w, int img_h, double fixed_conf, double category_id,
                       int* cursor_dev, double* dets_out, int* offsets_out, int dets_cap, int frames_cap,
                       int* status_dev, void* stream) {
    return tk_pack_detections_ex(boxes, scores, cls, counts, n_images, max_per_image, keep_class, img_w, img_h, fixed_conf,
                                 category_id, cursor_dev, dets_out, offsets_out, dets_cap, frames_cap, status_dev, TK_ROWS_LTRB,
                                 nullptr, stream);
}

}  // extern "C"
