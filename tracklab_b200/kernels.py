"""Python entry points of the stateless libtrackkern kernels (device tensors in, device tensors out)."""
from __future__ import annotations

import ctypes

import torch

from . import _lib

F32, BF16 = 0, 1


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dtype_code(t: torch.dtype) -> int:
    if t == torch.float32:
        return F32
    if t == torch.bfloat16:
        return BF16
    raise _lib.TrackKernError(f"unsupported dtype {t}")


def _cuda(t: torch.Tensor, name: str):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
        raise _lib.TrackKernError(f"{name} must be a contiguous CUDA tensor (no CPU path)")


def letterbox(frames: torch.Tensor, size: int = 640, out_dtype=torch.bfloat16, pad: int = 114, swap_rb: bool = True,
              out: torch.Tensor | None = None, channels_last: bool = False):
    """uint8 [B,H,W,3] -> [B,3,size,size]; returns (tensor, ratio). C ABI: tk_letterbox_u8."""
    lib = _lib.load()
    _cuda(frames, "frames")
    assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3
    B, H, W, _ = frames.shape
    if out is None:
        out = torch.empty((B, 3, size, size), dtype=out_dtype, device=frames.device,
                          memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    nhwc = out.is_contiguous(memory_format=torch.channels_last) and not out.is_contiguous()
    assert nhwc or out.is_contiguous()
    ratio = ctypes.c_double()
    with torch.cuda.device(frames.device):
        _lib.check(lib.tk_letterbox_u8(frames.data_ptr(), B, H, W, frames.stride(0), out.data_ptr(), _dtype_code(out.dtype),
                                       int(nhwc), size, pad, int(swap_rb), ctypes.byref(ratio), _stream()), "tk_letterbox_u8")
    return out, ratio.value


def yolox_nms(pred: torch.Tensor, ratio: float, input_size: int = 640, logits: bool = True, score_thr: float = 0.7,
              nms_thr: float = 0.45, max_out: int = 256, status: torch.Tensor | None = None):
    """pred [B, A, 5+nc] (f32/bf16) -> (boxes f32[B,max_out,4], scores f32[B,max_out], cls i32[B,max_out], count i32[B], status)."""
    lib = _lib.load()
    _cuda(pred, "pred")
    B, A, C = pred.shape
    dev = pred.device
    boxes = torch.empty((B, max_out, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((B, max_out), dtype=torch.float32, device=dev)
    cls = torch.empty((B, max_out), dtype=torch.int32, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    if status is None:
        status = torch.zeros((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.tk_yolox_nms(pred.data_ptr(), _dtype_code(pred.dtype), B, A, C - 5, input_size, int(logits),
                                    float(ratio), float(score_thr), float(nms_thr), max_out, boxes.data_ptr(),
                                    scores.data_ptr(), cls.data_ptr(), count.data_ptr(), status.data_ptr(), _stream()),
                   "tk_yolox_nms")
    return boxes, scores, cls, count, status


def pack_detections(boxes, scores, cls, count, width: int, height: int, cursor: torch.Tensor, dets_out: torch.Tensor,
                    offsets_out: torch.Tensor, status: torch.Tensor, keep_class: int = 0, fixed_conf: float = 1.0,
                    category_id: float = 1.0):
    """NMS output -> tracker rows float64[.,7] appended at row cursor[0]; frame offsets written at
    offsets_out[cursor[1]:cursor[1]+B+1]; cursor (device int32[2]) is advanced by the kernel."""
    lib = _lib.load()
    B, K = scores.shape
    with torch.cuda.device(boxes.device):
        _lib.check(lib.tk_pack_detections(boxes.data_ptr(), scores.data_ptr(), cls.data_ptr(), count.data_ptr(), B, K,
                                          keep_class, width, height, float(fixed_conf), float(category_id),
                                          cursor.data_ptr(), dets_out.data_ptr(), offsets_out.data_ptr(),
                                          dets_out.shape[0], offsets_out.shape[0] - 1, status.data_ptr(), _stream()),
                   "tk_pack_detections")
    return dets_out, offsets_out
