"""Frame ingest on the device (image-folder datasets): JPEG files -> uint8 [n,H,W,3] RGB frames in HBM.

Replaces ``cv2_load_image`` (/root/reference/tracklab/utils/cv2.py:34-66: cv2.imread + BGR->RGB on the host, one image at a time, then
the collate copy and the H2D transfer of the raw frame) for JPEG files: the host reads the compressed bytes only, nvJPEG decodes the
batch on the GPU (C ABI: libtkjpeg.so, include/tkjpeg.h). Other formats (PNG, ...) keep the host decoder: ``load_frames`` picks per batch.
JPEG decoders are not pixel-identical: on 4:4:4 files nvJPEG and libjpeg-turbo differ by IDCT rounding only (mean |diff| < 0.5 level),
on 4:2:0 files by their chroma up-sampling filters as well (mean 1.5 levels, tens of levels on sharp colour edges of the synthetic
frames; tests/test_jpeg_gpu.py). The modules therefore default to ``decode="cv2"`` (the reference's pixels) and take
``decode="nvjpeg"`` / ``"auto"`` as an opt-in for throughput.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtkjpeg.so")
_jlib = None


def _load():
    global _jlib
    if _jlib is None:
        if not os.path.exists(LIB_PATH):
            raise _lib.TrackKernError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        for cand in ("/usr/local/cuda/lib64/libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so"):    # the rpath covers it; be explicit anyway
            if os.path.exists(cand):
                try:
                    ctypes.CDLL(cand, mode=ctypes.RTLD_GLOBAL)
                    break
                except OSError:
                    pass
        lib = ctypes.CDLL(LIB_PATH)
        vp, ci, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        lib.tk_jpeg_create.argtypes = [ci, ctypes.POINTER(vp)]; lib.tk_jpeg_create.restype = ci
        lib.tk_jpeg_backend.argtypes = [vp]; lib.tk_jpeg_backend.restype = ci
        lib.tk_jpeg_last_status.argtypes = [vp]; lib.tk_jpeg_last_status.restype = ci
        lib.tk_jpeg_info.argtypes = [vp, vp, sz, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci)]; lib.tk_jpeg_info.restype = ci
        lib.tk_jpeg_decode_batch.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(sz), ci, vp, ci, ci, ctypes.c_longlong, vp]
        lib.tk_jpeg_decode_batch.restype = ci
        lib.tk_jpeg_destroy.argtypes = [vp]; lib.tk_jpeg_destroy.restype = ci
        _jlib = lib
    return _jlib


class JpegDecoderDevice:
    """nvJPEG batched decoder bound to one device (C ABI: tk_jpeg_*)."""

    def __init__(self, device="cuda:0", prefer_hardware: bool = True):
        if not torch.cuda.is_available():
            raise _lib.TrackKernError("JpegDecoderDevice needs a CUDA device (no CPU path)")
        self.lib = _load()
        self.device = torch.device(device)
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.tk_jpeg_create(int(prefer_hardware), ctypes.byref(self.handle))
        if rc != 0:
            raise _lib.TrackKernError(f"tk_jpeg_create failed ({rc})")
        self.backend = "hardware" if self.lib.tk_jpeg_backend(self.handle) == 1 else "default"

    def size(self, data: bytes):
        w, h, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        buf = (ctypes.c_ubyte * len(data)).from_buffer_copy(data)
        rc = self.lib.tk_jpeg_info(self.handle, ctypes.cast(buf, ctypes.c_void_p), len(data), ctypes.byref(w), ctypes.byref(h), ctypes.byref(c))
        if rc != 0:
            raise _lib.TrackKernError(f"tk_jpeg_info failed ({rc}, nvjpeg status {self.lib.tk_jpeg_last_status(self.handle)})")
        return h.value, w.value

    def decode(self, blobs, out: torch.Tensor | None = None) -> torch.Tensor:
        """blobs: list of ``bytes`` (compressed JPEG files of equal frame size) -> uint8 [n,H,W,3] RGB on the device (asynchronous on the
        current stream; the host buffers are kept alive until the stream is synchronised here)."""
        n = len(blobs)
        H, W = self.size(blobs[0])
        if out is None:
            out = torch.empty((n, H, W, 3), dtype=torch.uint8, device=self.device)
        assert out.is_cuda and out.dtype == torch.uint8 and tuple(out.shape) == (n, H, W, 3) and out.is_contiguous()
        bufs = [(ctypes.c_ubyte * len(b)).from_buffer_copy(b) for b in blobs]
        ptrs = (ctypes.c_void_p * n)(*[ctypes.cast(b, ctypes.c_void_p) for b in bufs])
        lens = (ctypes.c_size_t * n)(*[len(b) for b in blobs])
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream()
            rc = self.lib.tk_jpeg_decode_batch(self.handle, ptrs, lens, n, out.data_ptr(), H, W, out.stride(0), ctypes.c_void_p(st.cuda_stream))
            if rc != 0:
                raise _lib.TrackKernError(f"tk_jpeg_decode_batch failed ({rc}, nvjpeg status {self.lib.tk_jpeg_last_status(self.handle)})")
            st.synchronize()          # nvJPEG reads the host bit-streams asynchronously
        return out

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.tk_jpeg_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_decoders = {}


def load_frames(paths, device, decode: str = "cv2") -> torch.Tensor:
    """Image files of one video -> uint8 [n,H,W,3] RGB frames on ``device``. JPEG files go through nvJPEG (``decode`` "auto" /
    "nvjpeg"), everything else (and ``decode="cv2"``) through cv2.imread + BGR->RGB + one H2D copy, like cv2_load_image."""
    paths = [str(p) for p in paths]
    dev = torch.device(device)
    is_jpeg = all(p.lower().endswith((".jpg", ".jpeg")) for p in paths)
    if decode == "nvjpeg" and not is_jpeg:
        raise _lib.TrackKernError("decode='nvjpeg' needs .jpg / .jpeg files")
    if is_jpeg and decode in ("auto", "nvjpeg"):
        try:
            dec = _decoders.get(str(dev))
            if dec is None:
                dec = _decoders[str(dev)] = JpegDecoderDevice(dev)
            blobs = [open(p, "rb").read() for p in paths]
            return dec.decode(blobs)
        except (_lib.TrackKernError, OSError):
            if decode == "nvjpeg":
                raise
    import cv2
    batch = np.stack([cv2.cvtColor(cv2.imread(p), cv2.COLOR_BGR2RGB) for p in paths])     # cv2_load_image (utils/cv2.py:54-66)
    return torch.from_numpy(batch).to(dev)
