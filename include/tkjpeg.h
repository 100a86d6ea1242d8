/* libtkjpeg.so - frame ingest on the device: batched JPEG decode with nvJPEG into the uint8 [n, H, W, 3] RGB frame
 * tensor of the pipeline. Replaces cv2_load_image (/root/reference/tracklab/utils/cv2.py:34-66) for JPEG files of image-folder
 * datasets. extern "C", plain pointers, int error codes; a separate shared object next to libtrackkern.so (depends on libnvjpeg). */
#ifndef TKJPEG_H
#define TKJPEG_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TKJ_OK 0
#define TKJ_ERR_ARG -1
#define TKJ_ERR_NVJPEG -2   /* tk_jpeg_last_status() holds the nvjpegStatus_t */
#define TKJ_ERR_SIZE -3     /* an image of the batch does not have the frame size */

/* prefer_hardware: try the hardware JPEG engine first (falls back to the default GPU/hybrid backend). */
int tk_jpeg_create(int prefer_hardware, void** handle);
int tk_jpeg_backend(void* handle);       /* 1 = hardware engine, 0 = default backend */
int tk_jpeg_last_status(void* handle);
int tk_jpeg_info(void* handle, const unsigned char* data, size_t length, int* width, int* height, int* components);
/* data / lengths: host arrays of n compressed images (host memory); out_dev: device uint8, image i at out_dev + i * frame_stride_bytes,
 * rows of W * 3 bytes, RGB interleaved. Asynchronous on `stream`. */
int tk_jpeg_decode_batch(void* handle, const unsigned char* const* data, const size_t* lengths, int n, unsigned char* out_dev, int H,
                         int W, long long frame_stride_bytes, void* stream);
int tk_jpeg_destroy(void* handle);

#ifdef __cplusplus
}
#endif
#endif
