/* libtrackkern — C ABI of the B200-native TrackLab tracking hot path.
 *
 * Every entry point replaces a piece of the reference's per-frame Python path (cited per function).
 * Conventions: plain pointers and sizes only; pointers named *_dev / documented "device" are CUDA device
 * pointers on the current device; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 * functions return 0 on success or a negative TK_ERR_* code, never throw, and never allocate after
 * `*_create`. Work is enqueued asynchronously on `stream` unless stated otherwise.
 */
#ifndef TRACKKERN_H
#define TRACKKERN_H

#ifdef __cplusplus
extern "C" {
#endif

#define TK_ABI_VERSION 1

int tk_abi_version(void);
/* last CUDA runtime error code seen by the library (cudaError_t as int), 0 if none */
int tk_last_cuda_error(void);


/* overlap variants (oc_sort/association.py) */
#define TK_ASSO_IOU 0
#define TK_ASSO_GIOU 1
#define TK_ASSO_DIOU 2
#define TK_ASSO_CIOU 3
#define TK_ASSO_CT_DIST 4   /* centre distance rescaled by the matrix maximum (association.py:150-171) */

/* element types for tensor arguments */
#define TK_DTYPE_F32 0
#define TK_DTYPE_BF16 1

/* device-side status bits (per video / per call status words) */
#define TK_STATUS_OVERFLOW_TRACKS 1
#define TK_STATUS_OVERFLOW_DETS 2
#define TK_STATUS_LAP_INFEASIBLE 4
#define TK_STATUS_OVERFLOW_OUT 8
#define TK_STATUS_BAD_CHOLESKY 16

/* ---- Detector pre-processing: letterbox ---------------------------------------------------------
 * Replaces rtmlib YOLOX.preprocess (ratio = min(S/h,S/w); cv2.resize INTER_LINEAR; 114-padded SxS canvas,
 * top-left paste; HWC uint8 -> CHW float) that runs behind
 *   /root/reference/tracklab/wrappers/bbox_detector/rtmlib_api.py:19-30
 * src: device uint8 [n_frames, H, W, 3] (16-byte aligned, frame pitch `frame_stride_bytes`);
 * dst: device tensor of out_dtype; out_layout 0 = planar [n,3,S,S], 1 = channels-last [n,S,S,3],
 * 2 = YOLOX Focus space-to-depth pre-applied, channels-last [n,S/2,S/2,16] (12 used, 4 zero pad channels the
 * caller must have zeroed once); 3 = the same with a 32-channel pixel pitch [n,S/2,S/2,32] (channels 16..31 caller-zeroed:
 * cuDNN's sm_100 kernels need 32 input channels, the 16-channel stem convolution falls back to a 2.6x slower sm_80 kernel);
 * swap_rb=1 turns the engine's RGB frames
 * (/root/reference/tracklab/utils/cv2.py:54-66) into the BGR order cv2.imread feeds the detector
 * (rtmlib_api.py:28). *ratio_out (host, optional) receives the letterbox ratio.
 */
int tk_letterbox_u8(const unsigned char* src, int n_frames, int H, int W, long long frame_stride_bytes,
                    void* dst, int out_dtype, int out_layout, int S, int pad_value, int swap_rb, double* ratio_out,
                    void* stream);

/* ---- ReID input: crop gather + resize + normalise -------------------------------------------------------
 * Replaces the per-crop CPU path of the in-tracker ReID:
 *   /root/reference/plugins/track/strong_sort/strong_sort.py:102-108,135-145       (int()-truncated, clipped crop)
 *   /root/reference/plugins/track/strong_sort/reid_multibackend.py:45-52,184-195   (PIL Resize((256,128)) bilinear with
 *                                                                                   antialias, ToTensor, Normalize)
 * frames: device uint8 [F,H,W,3] RGB; dets: device double [N,7] wrapper rows (only l,t,r,b are read); det_frame:
 * device int[N] frame index of each row; out: device [N,3,out_h,out_w] (out_nhwc=P>0: channels-last [N,out_h,out_w,P], P = channel pitch 3 or
 * e.g. 8 with caller-zeroed padding channels) of out_dtype.
 * mean3/std3: HOST float[3]. Pixel values are integer-exact vs Pillow's resampler.
 * out_nhwc = TK_CROP_LAYOUT_S2D16: the ResNet stem layout. out is [N, out_h/2 + 3, out_w/2 + 3, 16] (caller-zeroed once):
 * pixel (y, x, c) goes to row y/2 + 2, column x/2 + 2, channel ((y&1)*2 + (x&1))*3 + c, so that the 7x7 stride-2 pad-3
 * stem convolution (/root/reference/plugins/track/strong_sort/deep/models/resnet.py:342-348) becomes a 4x4 stride-1
 * convolution on 16 channels without padding (same sums, 5x faster in cuDNN than 7x7/2 on a 3->8 channel input).
 */
#define TK_CROP_LAYOUT_S2D16 (-16)
int tk_crop_resize_norm(const unsigned char* frames, int H, int W, long long frame_stride_bytes, const double* dets,
                        const int* det_frame, int n_dets, void* out, int out_dtype, int out_nhwc, int out_h, int out_w,
                        const float* mean3, const float* std3, void* stream);
/* crop_rule selects how a row becomes a pixel rectangle:
 *   TK_CROP_RULE_STRONGSORT    rows [l,t,r,b,..]: centre box, int() truncation, clip (strong_sort.py:102-108) — tk_crop_resize_norm
 *   TK_CROP_RULE_LTWH_ROUNDED  rows [l,t,w,h,..] (the detector's float32 bbox_ltwh): the ReID wrapper's rule
 *                              /root/reference/tracklab/wrappers/reid/kpreid_api.py:118-121 = sanitize_bbox_ltwh + ltwh_to_ltrb in
 *                              float32 + round-half-even (/root/reference/tracklab/utils/coordinates.py:216-267), crop = image[t:b, l:r]
 */
#define TK_CROP_RULE_STRONGSORT 0
#define TK_CROP_RULE_LTWH_ROUNDED 1
#define TK_CROP_RULE_XYXY_INT 2   /* rows [l,t,r,b,..]: box.astype(int) + NumPy slice (deep_oc_sort/ocsort.py:560-565, bot_sort `_get_features`) */
int tk_crop_resize_norm_ex(const unsigned char* frames, int H, int W, long long frame_stride_bytes, const double* dets,
                           const int* det_frame, int n_dets, void* out, int out_dtype, int out_nhwc, int out_h, int out_w,
                           const float* mean3, const float* std3, int crop_rule, void* stream);

/* ---- RT-DETR detector pre/post-processing ---------------------------------------------------------------------
 * Replace transformers' RTDetrImageProcessor around the model call of
 * /root/reference/tracklab/wrappers/bbox_detector/transformers_api.py:31-54.
 *   tk_resize_frames_u8  frames uint8 [n,H,W,3] -> out [n,3,out_h,out_w] (float32 / bf16) = resize(frame) * scale:
 *                        Pillow's antialiased bilinear resampler evaluated exactly (the processor of the pinned
 *                        transformers 4.52; the torchvision backend of 5.x differs from Pillow by at most one 8-bit step).
 *   tk_rtdetr_decode     post_process_object_detection (sigmoid focal scores, top-Q over Q*C, cxcywh -> absolute xyxy in
 *                        float32, score > threshold) + the wrapper's class filter and sanitize/ltwh conversion
 *                        (utils/coordinates.py:270-328): logits [n,Q,C] float32, boxes [n,Q,4] float32 ->
 *                        rows float64 [n, Q, 6] = [l, t, w, h, score, query index] in descending score order and
 *                        counts int32 [n]. keep_label < 0 keeps every class (the 6th column then holds label * Q + query).
 */
int tk_resize_frames_u8(const unsigned char* frames, int n_frames, int H, int W, long long frame_stride_bytes, void* out,
                        int out_dtype, int out_h, int out_w, float scale, void* stream);
int tk_rtdetr_decode(const float* logits, const float* boxes, int n_images, int n_queries, int n_classes, int img_w, int img_h,
                     float threshold, int keep_label, double* rows_out, int* counts_out, void* stream);

/* ---- Pooling passes of the ReID backbone (channels-last bf16) ----------------------------------------------
 * tk_maxpool3x3s2_nhwc: src [N,H,W,C] -> dst [N,(H+1)/2,(W+1)/2,C], 3x3 window, stride 2, padding 1 (-inf), i.e.
 *                       nn.MaxPool2d(3, 2, 1) after the stem (/root/reference/plugins/track/strong_sort/deep/models/resnet.py:349)
 * tk_avgpool_nhwc:      src [N,HW,C] bf16 -> dst [N,C] float32, mean over the HW positions accumulated in float32
 *                       (global average pool + flatten, resnet.py:355-356).  C must be a multiple of 8.
 */
int tk_maxpool3x3s2_nhwc(const void* src, int n, int H, int W, int C, void* dst, void* stream);
int tk_avgpool_nhwc(const void* src, int n, int HW, int C, float* dst, void* stream);

/* ---- Detector post-processing: YOLOX decode + threshold + class-aware NMS --------------------------
 * Replaces rtmlib YOLOX.postprocess / multiclass_nms behind rtmlib_api.py:30 (score_thr 0.7, nms_thr 0.45).
 * pred: device [n_images, n_anchors, 5+n_classes] (reg xywh raw, obj, cls; logits=1 -> sigmoid applied here),
 * outputs per image: up to max_out (<=1024) boxes float32 xyxy in original-image pixels, scores, classes,
 * count; sorted by score descending. status_dev: device int, TK_STATUS_* bits are OR-ed in.
 */
int tk_yolox_nms(const void* pred, int pred_dtype, int n_images, int n_anchors, int n_classes, int input_size,
                 int logits, float ratio, float score_thr, float nms_thr, int max_out, float* out_boxes,
                 float* out_scores, int* out_cls, int* out_count, int* status_dev, void* stream);

/* ---- Detector wrapper row packing --------------------------------------------------------------------
 * Replaces RTMLibDetector.process' per-box Series construction (rtmlib_api.py:31-46: clip with
 * ltrb_to_ltwh(bbox,(W,H)), bbox_conf = 1.0, running id) and the tracker wrappers' row builder
 * (/root/reference/tracklab/wrappers/track/oc_sort_api.py:33-47): writes double[.,7] rows
 * [l,t,r,b,conf,cls,det_id]. cursor_dev is a device int[2] = {next row, next frame}, read and advanced by the
 * call (so consecutive batches of one video chain without host involvement, CUDA-graph friendly): rows are
 * appended at dets_out[cursor[0]..], frame offsets written to offsets_out[cursor[1] .. cursor[1]+n_images].
 * fixed_conf < 0 keeps the detector score. keep_class < 0 keeps every class.
 */
int tk_pack_detections(const float* boxes, const float* scores, const int* cls, const int* counts, int n_images,
                       int max_per_image, int keep_class, int img_w, int img_h, double fixed_conf, double category_id,
                       int* cursor_dev, double* dets_out, int* offsets_out, int dets_cap, int frames_cap,
                       int* status_dev, void* stream);
/* Same, for the connected detect -> ReID -> associate pipeline: row_format TK_ROWS_LTRB writes [l,t,r,b,...] (the tracker
 * wrappers' rows), TK_ROWS_LTWH writes [l,t,w,h,...] = the detector's float32 bbox_ltwh column as is (what the ReID wrapper
 * kpreid_api.py:115-144 and bpbreid_strong_sort_api.py:73-90 read); frame_of_row_out (nullable, int[dets_cap]) receives for
 * every appended row the index of its image inside this batch (the det_frame input of tk_crop_resize_norm). */
#define TK_ROWS_LTRB 0
#define TK_ROWS_LTWH 1
int tk_pack_detections_ex(const float* boxes, const float* scores, const int* cls, const int* counts, int n_images,
                          int max_per_image, int keep_class, int img_w, int img_h, double fixed_conf, double category_id,
                          int* cursor_dev, double* dets_out, int* offsets_out, int dets_cap, int frames_cap,
                          int* status_dev, int row_format, int* frame_of_row_out, void* stream);

/* ---- Backbone epilogues (channels-last bf16; the convolutions themselves stay in cuDNN) ----------------
 * Replace the separate bias / activation / concat / max-pool / up-sampling passes the reference's runtimes
 * execute between convolutions (rtmlib_api.py:19-30 -> onnxruntime; strong_sort/deep/models/resnet.py:342-361).
 * act: 0 none, 1 SiLU, 2 ReLU, 3 ReLU applied after the residual add. Pitches/offsets in elements, multiples of 8.
 */
int tk_bias_act_nhwc(const void* src, const float* bias, void* dst, const void* residual, long long n_pixels, int channels,
                     int dst_pitch, int dst_offset, int res_pitch, int res_offset, int act, void* stream);
/* 1x1 convolution with the epilogue fused in (hand-written sm_100a GEMM: TMA tensor-map loads of the NHWC activation tile,
 * tcgen05.mma into a TMEM accumulator, bias + activation (+ residual) in the TMEM read-out, bf16 straight into the concat slice):
 *   dst[m, dst_off + n] = act(sum_k x[m, k] * w[n, k] + bias[n]) (+ residual[m, res_off + n]),  m < M = images * H * W
 * x: bf16 [M, x_pitch] (first K channels of every pixel are read), w: bf16 [N, K], bias: float [N] or NULL, dst / residual: bf16
 * rows of dst_pitch / res_pitch channels. K, pitches and offsets multiples of 8, N a multiple of 16; all pointers 16-byte aligned.
 * Replaces "cuDNN 1x1 convolution + tk_bias_act_nhwc" for the 1x1 layers of the YOLOX / ResNet-50 executors. */
#define TK_ACT_NONE 0
#define TK_ACT_SILU 1
#define TK_ACT_RELU 2
#define TK_ACT_RELU_AFTER_RESIDUAL 3
int tk_conv1x1_bias_act_bf16(const void* x, long long M, int K, int x_pitch, const void* w, int N, const float* bias, void* dst,
                             int dst_pitch, int dst_off, const void* residual, int res_pitch, int res_off, int act, void* stream);
/* 3x3 convolution, stride 1, padding 1, same fused epilogue (hand-written sm_100a implicit GEMM: the im2col gather is done by the
 * TMA engine — one 4-D tensor-map box per tap, out-of-image rows arrive as zeros — tcgen05.mma into TMEM, TMA store of the tile):
 *   dst[b, y, x, dst_off + n] = act(sum_{ky,kx,c} x[b, y+ky-1, x+kx-1, c] * w[n, ky, kx, c] + bias[n]) (+ residual[b, y, x, res_off + n])
 * x: bf16 NHWC [n_images, H, W, Cin] contiguous; w: bf16 [N][3][3][Cin] (a channels-last PyTorch weight); Cin and N multiples of 16.
 * Replaces "cuDNN 3x3 convolution + tk_bias_act_nhwc" in the YOLOX executor. */
int tk_conv3x3_bias_act_bf16(const void* x, int n_images, int H, int W, int Cin, const void* w, int N, const float* bias, void* dst,
                             int dst_pitch, int dst_off, const void* residual, int res_pitch, int res_off, int act, void* stream);
int tk_spp_nhwc(const void* x, void* dst, int n_images, int H, int W, int channels, int dst_pitch, int dst_offset, void* stream);
int tk_upsample2x_nhwc(const void* src, int src_pitch, int src_offset, void* dst, int n_images, int h, int w, int channels,
                       int dst_pitch, int dst_offset, void* stream);

/* ---- ByteTrack: whole-video association -------------------------------------------------------
 * Replaces BYTETracker.update called once per frame by the wrapper:
 *   /root/reference/plugins/track/byte_track/byte_tracker.py:167-320  (update)
 *   /root/reference/tracklab/wrappers/track/byte_track_api.py:50-76   (per-frame filter + row layout)
 * Hyper-parameters: /root/reference/tracklab/configs/modules/track/byte_track.yaml:4-10.
 */
typedef struct {
    double track_thresh;   /* byte_track.yaml: track_thresh (0.6) */
    double match_thresh;   /* match_thresh (0.8) */
    double min_confidence; /* wrapper filter `conf > min_confidence` (0.4), byte_track_api.py:54 */
    int track_buffer;      /* track_buffer (30) */
    int frame_rate;        /* frame_rate (30) */
    int first_id;          /* first track id handed out (reference: 1 + process-global BaseTrack._count) */
} tk_bytetrack_params;

/* n_seq independent videos tracked side by side (one CTA each); cap_tracks = max live tracks per video
 * (tracked + lost), cap_dets = max detections per frame; both <= 256. */
int tk_bytetrack_create(const tk_bytetrack_params* p, int n_seq, int cap_tracks, int cap_dets, void** handle);
/* keep_id_counter != 0 continues the id numbering across videos like the reference's process-global
 * BaseTrack._count (basetrack.py:13,35-37); 0 restarts at first_id. */
int tk_bytetrack_reset(void* handle, int keep_id_counter, void* stream);
/* Run `n_frames` consecutive frames of every video, continuing from the current tracker state.
 *   dets      device double[N,7] = [l,t,r,b,conf,cls,det_id] rows (oc_sort_api.py:33-47 layout)
 *   offsets   device int[n_seq*(n_frames+1)]: video s, frame f owns rows offsets[s*(n_frames+1)+f .. +f+1)
 *   out_rows  device double[.,8] = [x1,y1,x2,y2,track_id,cls,score,det_id] (byte_tracker.py:301-318)
 *   out_start device int[n_seq]: first output row of video s; rows are appended after out_count[s]
 *   out_frame_count device int[n_seq*n_frames]: rows emitted per frame
 *   out_count device int[n_seq]: in/out running number of rows of video s (zero it before the first chunk)
 */
int tk_bytetrack_run(void* handle, const double* dets, const int* offsets, int n_frames, double* out_rows,
                     const int* out_start, int* out_frame_count, int* out_count, void* stream);
/* Copies the per-video device status words (0 = ok, TK_DEV_* bits otherwise) to host; synchronises `stream`. */
int tk_bytetrack_status(void* handle, int* status_host, void* stream);
int tk_bytetrack_destroy(void* handle);

/* ---- OC-SORT: whole-video association ------------------------------------------------------------
 * Replaces OCSort.update called once per frame by the wrapper:
 *   /root/reference/plugins/track/oc_sort/ocsort.py:203-334            (update)
 *   /root/reference/tracklab/wrappers/track/oc_sort_api.py:50-76       (per-frame filter + row layout)
 * Hyper-parameters: /root/reference/tracklab/configs/modules/track/oc_sort.yaml:4-14.
 * Same calling convention as tk_bytetrack_*; output rows are [x1,y1,x2,y2,id+1,cls,conf,det_id].
 */
typedef struct {
    double det_thresh;     /* oc_sort.yaml: det_thresh (0) */
    double iou_threshold;  /* iou_threshold (0.2213...) */
    double inertia;        /* inertia (0.3941...) — weight of the velocity-direction term */
    double min_confidence; /* wrapper filter (0.4), oc_sort_api.py:54 */
    int max_age;           /* 50 */
    int min_hits;          /* 1 */
    int delta_t;           /* 1 (<= 8) */
    int asso_func;         /* TK_ASSO_* used by the BYTE / OCR rounds (ocsort.py:266,287); round 1 is plain IoU */
    int use_byte;          /* false */
} tk_ocsort_params;

int tk_ocsort_create(const tk_ocsort_params* p, int n_seq, int cap_tracks, int cap_dets, void** handle);
int tk_ocsort_reset(void* handle, int keep_id_counter, void* stream);
int tk_ocsort_run(void* handle, const double* dets, const int* offsets, int n_frames, double* out_rows,
                  const int* out_start, int* out_frame_count, int* out_count, void* stream);
int tk_ocsort_status(void* handle, int* status_host, void* stream);
int tk_ocsort_destroy(void* handle);

/* ---- StrongSORT (DeepSORT lineage): whole-video association with externally supplied appearance features ------
 * Replaces StrongSORT.update called once per frame by the wrapper, minus the in-tracker ReID forward (a separate
 * stage here) and ECC (out of scope):
 *   /root/reference/plugins/track/strong_sort/strong_sort.py:41-85                 (update)
 *   /root/reference/plugins/track/strong_sort/sort/tracker.py:53-59,80-115,151-193 (predict / update / _match)
 *   /root/reference/tracklab/wrappers/track/strong_sort_api.py:66-93              (per-frame filter + row layout)
 * Hyper-parameters: /root/reference/tracklab/configs/modules/track/strong_sort.yaml:8-23.
 * features: device float32 [N, feature_dim], row i belongs to dets row i (raw, un-normalised ReID outputs).
 * Output rows [x1,y1,x2,y2,track_id,cls,conf,det_id] with int()-truncated, clipped boxes (strong_sort.py:110-121);
 * a track is reported while time_since_update <= 1, so up to 2x the detections of a frame may come out:
 * out_capacity_rows bounds the rows of one video (TK_STATUS_OVERFLOW_OUT otherwise).
 * Each video is served by a cooperative group of `ctas_per_video` CTAs (default 8); n_seq*ctas_per_video <= 148.
 */
typedef struct {
    double max_dist;        /* strong_sort.yaml: max_dist (0.1594...) — appearance matching threshold */
    double max_iou_dist;    /* max_iou_dist (0.5431...) */
    double mc_lambda;       /* 0.995: weight of appearance vs Mahalanobis distance */
    double ema_alpha;       /* 0.8962...: EMA of the appearance feature */
    double min_confidence;  /* wrapper filter (0.4) */
    int max_age;            /* 40 */
    int n_init;             /* 3 */
    int nn_budget;          /* 100: gallery length per track */
    int max_unmatched_preds;/* must be 0 (reference YAML) */
    int feature_dim;        /* E */
    int image_width;        /* for the output clipping (strong_sort.py:112-118) */
    int image_height;
    int ctas_per_video;     /* 0 = default */
} tk_strongsort_params;

int tk_strongsort_create(const tk_strongsort_params* p, int n_seq, int cap_tracks, int cap_dets, void** handle);
int tk_strongsort_reset(void* handle, int keep_id_counter, void* stream);
int tk_strongsort_run(void* handle, const double* dets, const float* features, const int* offsets, int n_frames, double* out_rows,
                      const int* out_start, int* out_frame_count, int* out_count, int out_capacity_rows, void* stream);
/* Same with camera-motion compensation (cfg.ecc of /root/reference/tracklab/wrappers/track/strong_sort_api.py:62-65 ->
 * Tracker.camera_update -> Track.camera_update, sort/track.py:224-239): warps = float [n_seq][n_frames][6], the 2x3 matrix of every
 * frame of this call (tk_ecc_euclidean); it is applied to the box of EVERY track before the frame is processed (also on frames
 * without detections). A NaN first entry means "no previous frame / ECC failed": tracks are left alone. NULL = tk_strongsort_run. */
int tk_strongsort_run_cmc(void* handle, const double* dets, const float* features, const int* offsets, int n_frames, const float* warps,
                          double* out_rows, const int* out_start, int* out_frame_count, int* out_count, int out_capacity_rows, void* stream);

/* ---- ECC camera-motion estimation (sort/track.py:129-214: gray, x0.1 bilinear, cv2.findTransformECC MOTION_EUCLIDEAN) -------------
 *   tk_ecc_small_size  (H, W, scale) -> size of the down-scaled gray image (cv2.resize with fx = fy = scale: cvRound)
 *   tk_ecc_gray_small  frames uint8 [n,H,W,3] (the RGB frames the wrapper loads; the plugin runs COLOR_BGR2GRAY on them as they
 *                      are) -> uint8 [n,h,w]: OpenCV's fixed-point gray conversion + INTER_LINEAR resize, bit-equal to cv2
 *   tk_ecc_euclidean   small uint8 [n,h,w]: for every consecutive pair (i-1, i) the forward-additive ECC iteration of OpenCV
 *                      (max_iter, eps on the change of the correlation coefficient) -> warps_out float [n][6] (row 0 = NaN: the first
 *                      image has no predecessor; translation already divided by `scale` like track.py:196-198), rho_out double [n],
 *                      ok_out int [n] (0: did not converge / singular -> the caller stores NaN in the warp, "ecc transform failed") */
int tk_ecc_small_size(int H, int W, double scale, int* h_out, int* w_out);
int tk_ecc_gray_small(const unsigned char* frames, int n_frames, int H, int W, long long frame_stride_bytes, double scale,
                      unsigned char* out, void* stream);
int tk_ecc_euclidean(const unsigned char* small_images, int n_images, int h, int w, int max_iter, double eps, double scale,
                     float* warps_out, double* rho_out, int* ok_out, void* stream);
int tk_strongsort_status(void* handle, int* status_host, void* stream);
int tk_strongsort_destroy(void* handle);

/* ---- BPBReID-StrongSORT: part-based appearance, visibility-aware EMA ------------------------------------------
 * Replaces bpbreid_strong_sort.StrongSORT.update driven once per frame by
 * /root/reference/tracklab/wrappers/track/bpbreid_strong_sort_api.py:73-118
 * (/root/reference/plugins/track/bpbreid_strong_sort/strong_sort.py:53-141, sort/tracker.py:123-167,242-333) for
 * matching_strategy "strong_sort_matching" + motion_criterium "iou".
 *   dets        float64 [N,7] rows [l, t, w, h, bbox_conf, class, det id]   (ltwh, as the wrapper stacks bbox_ltwh)
 *   features    float32 [N, n_parts, feature_dim] (feature_dim a multiple of 4, base 16-byte aligned);
 *   visibility  float32 [N, n_parts] (booleans as 0/1)
 *   cap_tracks  may be far larger than cap_dets (<= 256): with n_init 0 / max_age 300 every false positive is a
 *               confirmed track for 300 frames; only tracks that pass the Mahalanobis gate enter the assignment.
 *   out_rows    float64 [.,14] = [track_id, track_bbox_kf_ltwh(4), track_bbox_pred_kf_ltwh(4) (NaN at birth),
 *               matched_with stage (0 none, 1 'R', 2 'S'), matched distance (NaN when none), hits, age, det id];
 *               only confirmed tracks updated in the frame are emitted (time_since_update 0, state 'c').
 * The per-detection `costs` visualisation dictionaries (sort/tracker.py:365-407) are not produced.
 */
typedef struct tk_bpbreid_params {
    double max_dist;            /* 0.5   */
    double max_iou_distance;    /* 0.8   */
    double mc_lambda;           /* 0.995 */
    double ema_alpha;           /* 0.9   */
    double min_bbox_confidence; /* 0.0   */
    int max_age;                /* 300   */
    int n_init;                 /* 0     */
    int max_kalman_prediction_without_update; /* 7 */
    int n_parts;
    int feature_dim;
    int ctas_per_video;         /* 0 = default */
    int matching_strategy;      /* TK_BPBREID_STRONG_SORT_MATCHING (two stages, the YAML) or TK_BPBREID_BOT_SORT_MATCHING: one stage over all
                                   tracks on (w_kfgd * pos + w_reid * app + w_st * st) / sum(w), sort/tracker.py:335-363,169-240 */
    double gating_thres_factor; /* 1.0 */
    double w_kfgd;              /* 1.0 (must be > 0 on device: the position gate prunes the pairs) */
    double w_reid;              /* 1.0 */
    double w_st;                /* 1.0 */
} tk_bpbreid_params;
#define TK_BPBREID_STRONG_SORT_MATCHING 0
#define TK_BPBREID_BOT_SORT_MATCHING 1

int tk_bpbreid_create(const tk_bpbreid_params* p, int n_seq, int cap_tracks, int cap_dets, void** handle);
int tk_bpbreid_reset(void* handle, int keep_id_counter, void* stream);
int tk_bpbreid_run(void* handle, const double* dets, const float* features, const float* visibility, const int* offsets, int n_frames,
                   double* out_rows, const int* out_start, int* out_frame_count, int* out_count, int out_capacity_rows, void* stream);
int tk_bpbreid_status(void* handle, int* status_host, void* stream);
int tk_bpbreid_destroy(void* handle);

/* ---- Stateless batched cost matrices + assignment (building blocks; stress sweep of BASELINE configs[4]) ------
 * All tensors device, row-major, `n_problems` independent problems stacked on the leading axis.
 *   tk_iou_matrix   a [B,N,4], b [B,M,4] float64 x1y1x2y2 -> out [B,N,M]; variant TK_ASSO_*
 *                   (/root/reference/plugins/track/oc_sort/association.py:5-147)
 *   tk_iou_p1_f32   float32 tlbr boxes -> 1 - IoU with +1-pixel extents, float32
 *                   (/root/reference/plugins/track/byte_track/matching.py:51-89,182-218)
 *   tk_cosine_dist  a [B,N,E], b [B,M,E] float32 -> out [B,N,M] float64 = 1 - normalised dot product computed in
 *                   float32 (/root/reference/plugins/track/strong_sort/sort/nn_matching.py:30-49,144-161);
 *                   norm_scratch: float32 [B*(N+M)]
 *   tk_lap_batched  cost [B,N,M] float64 -> x [B,N] (column of each row or -1), y [B,M]; has_limit=1 gives
 *                   lap.lapjv(cost, extend_cost=True, cost_limit=L) semantics (byte_track/matching.py:37-48),
 *                   has_limit=0 assigns all min(N,M) pairs (oc_sort/association.py:187-191, scipy LSA)
 *   tk_part_dist    a [B,N,K,E], va [B,N,K], b [B,M,K,E], vb [B,M,K] float32 -> out [B,N,M] float32: visibility-weighted mean
 *                   over the K parts of the Euclidean distance between L2-normalised part embeddings, halved
 *                   (/root/reference/plugins/track/bpbreid_strong_sort/sort/nn_matching.py:99-135; the torchreid function it
 *                   calls is restated, see oracle/bpbreid_np.py); norm_scratch: float32 [B*(N+M)*K]
 *   tk_kf_gate      squared Mahalanobis gating distances of xyah Kalman filters: mean [T,8], cov [T,8,8], z [D,4] float64
 *                   -> out [T,D]; aspect_const 1 = ByteTrack / StrongSORT measurement noise (1e-1 on the aspect ratio,
 *                   byte_track/kalman_filter.py:228-269, strong_sort/sort/kalman_filter.py:176-214), 0 = BPBReID (all
 *                   terms scale with the height, bpbreid_strong_sort/sort/kalman_filter.py:168-227)
 */
int tk_part_dist(const float* a, const float* va, const float* b, const float* vb, float* out, float* norm_scratch, int n_problems, int N,
                 int M, int K, int E, void* stream);
int tk_kf_gate(const double* mean, const double* cov, const double* z, double* out, int n_tracks, int n_dets, int aspect_const,
               int* status_dev, void* stream);
/* Stateless Kalman steps of SURVEY.md 8b's proposed ABI, one thread per track, in place on mean [n,8] / cov [n,8,8] (float64):
 * model 0 = ByteTrack xyah (/root/reference/plugins/track/byte_track/kalman_filter.py:88-124 predict, :194-226 update),
 * model 1 = BoT-SORT xywh (/root/reference/plugins/track/bot_sort/kalman_filter.py:88-124, :194-226); z [n,4] = one measurement per track. */
int tk_kf_predict(double* mean, double* cov, int n_tracks, int model, void* stream);
int tk_kf_update(double* mean, double* cov, const double* z, int n_tracks, int model, int* status_dev, void* stream);
/* OC-SORT velocity-direction-consistency cost (/root/reference/plugins/track/oc_sort/association.py:175-184,246-266): dets [D,6]
 * (x1,y1,x2,y2,score,cls), prev_obs [T,5] (k_previous_obs, score < 0 = placeholder), velocities [T,2] (dy, dx) -> out [D,T] =
 * valid * (pi/2 - |acos(clip(v . dir))|) / pi * inertia * dets[:, weight_col] (the reference multiplies with column 5, the class). */
int tk_vdc_cost(const double* dets6, const double* prev_obs5, const double* velocities2, double* out, int n_dets, int n_tracks, double inertia,
                int weight_col, void* stream);
int tk_iou_matrix(const double* a, const double* b, double* out, int n_problems, int N, int M, int variant, void* stream);
int tk_iou_p1_f32(const float* a_tlbr, const float* b_tlbr, float* dist_out, int n_problems, int N, int M, void* stream);
int tk_cosine_dist(const float* a, const float* b, double* out, float* norm_scratch, int n_problems, int N, int M, int E,
                   void* stream);
int tk_lap_batched(const double* cost, int n_problems, int N, int M, double cost_limit, int has_limit, int* x_out, int* y_out,
                   int* status_dev, void* stream);
/* scipy.optimize.linear_sum_assignment(cost) restated operation by operation INCLUDING its tie-breaking (the solver of
 * /root/reference/plugins/track/strong_sort/sort/linear_assignment.py:55 and bpbreid_strong_sort/sort/linear_assignment.py:56,
 * whose equal `max_distance + 1e-5` entries make the tie order decide the birth order of tracks): cost [B,N,M] float64 ->
 * x [B,N] = column of each row (or -1 when N > M leaves it unassigned), y [B,M] = row of each column (or -1). The whole-video
 * StrongSORT / BPBReID kernels use the same routine in shared memory (csrc/lsap_scipy.cuh). */
int tk_lsap_scipy_batched(const double* cost, int n_problems, int N, int M, int* x_out, int* y_out, int* status_dev, void* stream);

/* ---- Deep OC-SORT: whole-video association with externally supplied embeddings and camera-motion affines ----
 * Replaces OCSort.update of the deep_oc_sort plugin called once per frame by the wrapper, minus the in-tracker ReID forward
 * (`_get_features`, a separate stage here) and the camera-motion estimator (`CMCComputer.compute_affine`: its 2x3 result is an input):
 *   /root/reference/plugins/track/deep_oc_sort/ocsort.py:374-542        (update), :96-304 (KalmanBoxTracker, 8-d filter)
 *   /root/reference/plugins/track/deep_oc_sort/association.py:202-212,263-360
 *   /root/reference/plugins/track/deep_oc_sort/kalmanfilter.py:340-379,383-481,483-569
 *   /root/reference/tracklab/wrappers/track/deep_oc_sort_api.py:57-91   (per-frame filter + row layout)
 * Hyper-parameters: /root/reference/tracklab/configs/modules/track/deep_oc_sort.yaml (new_kf_off must be false).
 * dets [N,7] float64 (x1,y1,x2,y2,score,cls,det_id), embeddings float32 [N, feature_dim] (row i belongs to dets row i, as the
 * plugin's ReID returns them: not normalised by the tracker), affines float64 [n_seq, n_frames, 2, 3] (NULL when cmc_off),
 * offsets int32 [n_seq, n_frames+1]. Output rows [x1,y1,x2,y2,track_id,cls,conf,det_id], at most one per detection of a frame.
 * cap_tracks + cap_dets <= 512. lap 0.5.12 (the solver the plugin imports) is not vendored: its published (n+m)^2 extension is
 * solved with scipy's algorithm incl. its tie-breaking, like the stand-in the goldens were generated with. */
typedef struct {
    double det_thresh;        /* 0 */
    double iou_threshold;     /* 0.2213... */
    double inertia;           /* 0.3941... */
    double min_confidence;    /* wrapper filter (0.4), deep_oc_sort_api.py:65 */
    double w_association_emb; /* 0.75 */
    double alpha_fixed_emb;   /* 0.95 */
    double aw_param;          /* 0.5 */
    int max_age;              /* 50 */
    int min_hits;             /* 1 */
    int delta_t;              /* 1 (<= 7) */
    int asso_func;            /* TK_ASSO_* of the second (OCR) round; round 1 is plain IoU */
    int embedding_off;        /* false */
    int cmc_off;              /* false */
    int aw_off;               /* false */
    int feature_dim;          /* E */
} tk_deepocsort_params;

int tk_deepocsort_create(const tk_deepocsort_params* p, int n_seq, int cap_tracks, int cap_dets, void** handle);
int tk_deepocsort_reset(void* handle, void* stream);
int tk_deepocsort_run(void* handle, const double* dets, const float* embeddings, const double* affines, const int* offsets, int n_frames,
                      double* out_rows, const int* out_start, int* out_frame_count, int* out_count, int out_capacity_rows, void* stream);
int tk_deepocsort_status(void* handle, int* status_host, void* stream);
int tk_deepocsort_destroy(void* handle);

/* ---- BoT-SORT: whole-video association with externally supplied embeddings and camera-motion warps ---------
 * Replaces BoTSORT.update of the bot_sort plugin called once per frame by the wrapper, minus the in-tracker ReID forward
 * (`_get_features`) and the camera-motion estimator (`GMC.apply`: its 2x3 result is an input):
 *   /root/reference/plugins/track/bot_sort/bot_sort.py:275-485 (update), :15-240 (STrack), :507-545 (list helpers)
 *   /root/reference/plugins/track/bot_sort/matching.py:37-48,72-89,127-195,198-233
 *   /root/reference/plugins/track/bot_sort/kalman_filter.py:55-268
 *   /root/reference/tracklab/wrappers/track/bot_sort_api.py:57-87 (per-frame filter + row layout)
 * Hyper-parameters: /root/reference/tracklab/configs/modules/track/bot_sort.yaml. dets [N,7] float64, embeddings float32
 * [N, feature_dim] (raw backbone outputs; the tracker normalises them), warps float64 [n_seq, n_frames, 2, 3], offsets int32
 * [n_seq, n_frames+1]. Output rows [x1,y1,x2,y2,track_id,cls,score,det_id]. The class histogram of the plugin (update_cls) is
 * reduced to the class of the last matched detection. */
typedef struct {
    double track_high_thresh; /* 0.45 */
    double new_track_thresh;  /* 0.6 */
    double match_thresh;      /* 0.8 */
    double proximity_thresh;  /* 0.5 */
    double appearance_thresh; /* 0.25 */
    double lambda_;           /* 0.985 */
    double min_confidence;    /* wrapper filter (0.4), bot_sort_api.py:65 */
    int track_buffer;         /* 30 */
    int frame_rate;           /* 30 */
    int feature_dim;          /* E */
} tk_botsort_params;

int tk_botsort_create(const tk_botsort_params* p, int n_seq, int cap_tracks, int cap_dets, void** handle);
int tk_botsort_reset(void* handle, int keep_id_counter, void* stream);
int tk_botsort_run(void* handle, const double* dets, const float* embeddings, const double* warps, const int* offsets, int n_frames,
                   double* out_rows, const int* out_start, int* out_frame_count, int* out_count, void* stream);
int tk_botsort_status(void* handle, int* status_host, void* stream);
int tk_botsort_destroy(void* handle);

/* ---- HOTA of one sequence on the device -----------------------------------------------------------------------
 * Replaces HOTA.eval_sequence of the TrackEval fork vendored in the reference
 * (/root/reference/plugins/eval/PoseTrack21/posetrack21/posetrack21/trackeval/metrics/hota.py:28-154, final fields :205-221) with the
 * box similarity of its MOT dataset (trackeval/datasets/_base_dataset.py:244-282, box_format 'xywh'; posetrack_mot.py:479).
 * gt / tracker rows are frame-major: boxes [rows,4] float64 xywh, ids [rows] int32 already mapped to 0..n_ids-1 (unique within a
 * frame, as the reference's _check_unique_ids demands), offsets [n_frames+1] int32 — all device pointers. `alphas_host` (host
 * pointer, <= 32 values; the reference uses np.arange(0.05, 0.99, 0.05)). `pairs_cap` >= sum over frames of gt rows x tracker rows.
 * out [TK_HOTA_FIELDS, n_alphas] float64 (device), rows in the order below. The FragA field of the fork is not computed.
 * Integer fields (TP/FN/FP, the assignments) are exact; the float fields follow the reference's own summation order. */
#define TK_HOTA_HOTA 0
#define TK_HOTA_DETA 1
#define TK_HOTA_ASSA 2
#define TK_HOTA_DETRE 3
#define TK_HOTA_DETPR 4
#define TK_HOTA_ASSRE 5
#define TK_HOTA_ASSPR 6
#define TK_HOTA_LOCA 7
#define TK_HOTA_TP 8
#define TK_HOTA_FN 9
#define TK_HOTA_FP 10
#define TK_HOTA_FIELDS 11
int tk_hota_workspace_bytes(int n_frames, int n_gt_ids, int n_tr_ids, int n_alphas, long long pairs_cap, long long* bytes_out);
int tk_hota_sequence(const double* gt_boxes_xywh, const int* gt_ids, const int* gt_offsets, long long n_gt_rows,
                     const double* tr_boxes_xywh, const int* tr_ids, const int* tr_offsets, long long n_tr_rows, int n_frames,
                     int n_gt_ids, int n_tr_ids, int max_gt_per_frame, int max_tr_per_frame, const double* alphas_host,
                     int n_alphas, long long pairs_cap, void* workspace, long long workspace_bytes, double* out, int* status_dev,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRACKKERN_H */
