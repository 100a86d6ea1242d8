// Frame ingest on the device (image-folder datasets): batched JPEG decode with nvJPEG straight into the uint8
// [n, H, W, 3] RGB frame tensor the letterbox / crop kernels read.
//
// Replaces cv2_load_image (/root/reference/tracklab/utils/cv2.py:34-66: cv2.imread + BGR->RGB on one host core per image, then the
// collate / H2D copy of the decoded frame) for JPEG files: the host only reads the compressed bytes (~10x smaller than the frame), the
// Huffman / IDCT / colour conversion run on the GPU (nvjpegDecodeBatched, hardware engine when the library offers it). JPEG decoders
// are not bit-identical to each other (IDCT and chroma up-sampling differ by design): parity with libjpeg-turbo is +-a few
// intensity levels, measured in tests/test_jpeg_gpu.py.
// Separate shared object (libtkjpeg.so) so that libtrackkern.so keeps libcudart as its only dependency.
#include <cuda_runtime.h>
#include <nvjpeg.h>

#include <cstdio>
#include <vector>

#include "tkjpeg.h"

namespace {
struct JpegHandle {
    nvjpegHandle_t lib = nullptr;
    nvjpegJpegState_t state = nullptr;
    int batch = 0;        // batch size nvjpegDecodeBatchedInitialize was last called with
    int backend = 0;
    int last_status = 0;
};
}  // namespace

extern "C" {

int tk_jpeg_create(int prefer_hardware, void** handle) {
    if (!handle) return TKJ_ERR_ARG;
    JpegHandle* h = new JpegHandle();
    nvjpegStatus_t st = NVJPEG_STATUS_NOT_INITIALIZED;
    if (prefer_hardware) {
        st = nvjpegCreateEx(NVJPEG_BACKEND_HARDWARE, nullptr, nullptr, 0, &h->lib);
        if (st == NVJPEG_STATUS_SUCCESS) h->backend = 1;
    }
    if (st != NVJPEG_STATUS_SUCCESS) {
        st = nvjpegCreateEx(NVJPEG_BACKEND_DEFAULT, nullptr, nullptr, 0, &h->lib);
        h->backend = 0;
    }
    if (st != NVJPEG_STATUS_SUCCESS) { h->last_status = (int)st; delete h; return TKJ_ERR_NVJPEG; }
    st = nvjpegJpegStateCreate(h->lib, &h->state);
    if (st != NVJPEG_STATUS_SUCCESS) { nvjpegDestroy(h->lib); delete h; return TKJ_ERR_NVJPEG; }
    *handle = h;
    return TKJ_OK;
}

int tk_jpeg_backend(void* handle) { return handle ? ((JpegHandle*)handle)->backend : TKJ_ERR_ARG; }
int tk_jpeg_last_status(void* handle) { return handle ? ((JpegHandle*)handle)->last_status : TKJ_ERR_ARG; }

int tk_jpeg_info(void* handle, const unsigned char* data, size_t length, int* width, int* height, int* components) {
    if (!handle || !data || !width || !height) return TKJ_ERR_ARG;
    JpegHandle* h = (JpegHandle*)handle;
    int nc = 0, w[NVJPEG_MAX_COMPONENT], hh[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t ss;
    const nvjpegStatus_t st = nvjpegGetImageInfo(h->lib, data, length, &nc, &ss, w, hh);
    if (st != NVJPEG_STATUS_SUCCESS) { h->last_status = (int)st; return TKJ_ERR_NVJPEG; }
    *width = w[0]; *height = hh[0];
    if (components) *components = nc;
    return TKJ_OK;
}

int tk_jpeg_decode_batch(void* handle, const unsigned char* const* data, const size_t* lengths, int n, unsigned char* out_dev, int H,
                         int W, long long frame_stride_bytes, void* stream) {
    if (!handle || !data || !lengths || !out_dev || n <= 0 || H <= 0 || W <= 0 || frame_stride_bytes < (long long)H * W * 3) return TKJ_ERR_ARG;
    JpegHandle* h = (JpegHandle*)handle;
    for (int i = 0; i < n; ++i) {           // every image must have the frame size (one video)
        int w = 0, hh = 0;
        const int e = tk_jpeg_info(handle, data[i], lengths[i], &w, &hh, nullptr);
        if (e != TKJ_OK) return e;
        if (w != W || hh != H) return TKJ_ERR_SIZE;
    }
    nvjpegStatus_t st;
    if (h->batch != n) {
        st = nvjpegDecodeBatchedInitialize(h->lib, h->state, n, 1, NVJPEG_OUTPUT_RGBI);
        if (st != NVJPEG_STATUS_SUCCESS) { h->last_status = (int)st; return TKJ_ERR_NVJPEG; }
        h->batch = n;
    }
    std::vector<nvjpegImage_t> dst(n);
    for (int i = 0; i < n; ++i) {
        for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) { dst[i].channel[c] = nullptr; dst[i].pitch[c] = 0; }
        dst[i].channel[0] = out_dev + (size_t)i * frame_stride_bytes;
        dst[i].pitch[0] = (size_t)W * 3;
    }
    st = nvjpegDecodeBatched(h->lib, h->state, data, lengths, dst.data(), (cudaStream_t)stream);
    if (st != NVJPEG_STATUS_SUCCESS) { h->last_status = (int)st; return TKJ_ERR_NVJPEG; }
    return TKJ_OK;
}

int tk_jpeg_destroy(void* handle) {
    if (!handle) return TKJ_ERR_ARG;
    JpegHandle* h = (JpegHandle*)handle;
    if (h->state) nvjpegJpegStateDestroy(h->state);
    if (h->lib) nvjpegDestroy(h->lib);
    delete h;
    return TKJ_OK;
}

}  // extern "C"
