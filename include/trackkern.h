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
int tk_bytetrack_reset(void* handle, void* stream);
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

#ifdef __cplusplus
}
#endif
#endif /* TRACKKERN_H */
