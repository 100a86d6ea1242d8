"""Python entry points of the stateless libtrackkern kernels (device tensors in, device tensors out)."""
from __future__ import annotations

import ctypes

import torch

from . import _lib

F32, BF16 = 0, 1
CROP_RULE_STRONGSORT, CROP_RULE_LTWH_ROUNDED, CROP_RULE_XYXY_INT = 0, 1, 2   # TK_CROP_RULE_* (``ltwh_rows`` of crop_resize_norm accepts them)

LAUNCHES = 0   # number of libtrackkern kernel launches issued through this module (bench.py reports it)


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


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
              out: torch.Tensor | None = None, channels_last: bool = False, focus16: bool = False):
    """uint8 [B,H,W,3] -> [B,3,size,size] (or, focus16, the Focus-unfolded [B,16,size/2,size/2] channels-last tensor);
    returns (tensor, ratio). C ABI: tk_letterbox_u8."""
    lib = _lib.load()
    _cuda(frames, "frames")
    assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3
    B, H, W, _ = frames.shape
    if out is None:
        out = torch.empty((B, 3, size, size), dtype=out_dtype, device=frames.device,
                          memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    if focus16:
        assert out.shape[1] in (16, 32) and out.shape == (B, out.shape[1], size // 2, size // 2)
        assert out.is_contiguous(memory_format=torch.channels_last)
        nhwc = 2 if out.shape[1] == 16 else 3   # pitch 32: channels 16..31 stay as the caller zeroed them
    else:
        nhwc = int(out.is_contiguous(memory_format=torch.channels_last) and not out.is_contiguous())
        assert nhwc or out.is_contiguous()
    ratio = ctypes.c_double()
    with torch.cuda.device(frames.device):
        _lib.check(lib.tk_letterbox_u8(frames.data_ptr(), B, H, W, frames.stride(0), out.data_ptr(), _dtype_code(out.dtype),
                                       int(nhwc), size, pad, int(swap_rb), ctypes.byref(ratio), _stream()), "tk_letterbox_u8"); _count()
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
                   "tk_yolox_nms"); _count()
    return boxes, scores, cls, count, status


def pack_detections(boxes, scores, cls, count, width: int, height: int, cursor: torch.Tensor, dets_out: torch.Tensor,
                    offsets_out: torch.Tensor, status: torch.Tensor, keep_class: int = 0, fixed_conf: float = 1.0,
                    category_id: float = 1.0, ltwh: bool = False, frame_of_row: torch.Tensor | None = None):
    """NMS output -> tracker rows float64[.,7] appended at row cursor[0]; frame offsets written at
    offsets_out[cursor[1]:cursor[1]+B+1]; cursor (device int32[2]) is advanced by the kernel. ``ltwh`` writes [l,t,w,h,..]
    rows (the detector's bbox_ltwh column) instead of [l,t,r,b,..]; ``frame_of_row`` (int32[dets_cap]) receives the batch-local
    image index of every appended row (C ABI: tk_pack_detections_ex)."""
    lib = _lib.load()
    B, K = scores.shape
    if frame_of_row is not None:
        assert frame_of_row.dtype == torch.int32 and frame_of_row.numel() >= dets_out.shape[0]
    with torch.cuda.device(boxes.device):
        _lib.check(lib.tk_pack_detections_ex(boxes.data_ptr(), scores.data_ptr(), cls.data_ptr(), count.data_ptr(), B, K,
                                             keep_class, width, height, float(fixed_conf), float(category_id),
                                             cursor.data_ptr(), dets_out.data_ptr(), offsets_out.data_ptr(),
                                             dets_out.shape[0], offsets_out.shape[0] - 1, status.data_ptr(), int(bool(ltwh)),
                                             frame_of_row.data_ptr() if frame_of_row is not None else None, _stream()),
                   "tk_pack_detections_ex"); _count()
    return dets_out, offsets_out


def bias_act(src: torch.Tensor, bias: torch.Tensor, dst: torch.Tensor, dst_offset: int = 0, act: int = 1,
             residual: torch.Tensor | None = None, res_offset: int = 0):
    """dst[:, off:off+C] = act(src + bias) (+ residual). All tensors bf16 channels-last [B,C,H,W]; dst/residual may be wider
    (concat buffers). C ABI: tk_bias_act_nhwc."""
    lib = _lib.load()
    B, C, H, W = src.shape
    assert src.dtype == torch.bfloat16 and src.is_contiguous(memory_format=torch.channels_last) and bias.dtype == torch.float32
    assert dst.is_contiguous(memory_format=torch.channels_last) and dst.shape[0] == B and dst.shape[2:] == src.shape[2:]
    rp, r = (residual.shape[1], residual.data_ptr()) if residual is not None else (0, None)
    with torch.cuda.device(src.device):
        _lib.check(lib.tk_bias_act_nhwc(src.data_ptr(), bias.data_ptr(), dst.data_ptr(), r, B * H * W, C, dst.shape[1], dst_offset,
                                        rp, res_offset, act, _stream()), "tk_bias_act_nhwc"); _count()
    return dst


def conv1x1_bias_act(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None, dst: torch.Tensor | None = None, dst_offset: int = 0,
                     act: int = 1, residual: torch.Tensor | None = None, res_offset: int = 0):
    """1x1 convolution with fused epilogue on the tcgen05 path (C ABI: tk_conv1x1_bias_act_bf16).
    x bf16 channels-last [B,K,H,W] (or [M,K]); w bf16 [N,K]; bias float32 [N]; dst / residual bf16 channels-last, possibly wider
    (concat buffers). act: 0 none, 1 SiLU, 2 ReLU, 3 ReLU after the residual add. Returns dst."""
    lib = _lib.load()
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and w.is_contiguous() and w.dim() == 2
    if x.dim() == 4:
        B, K, H, W = x.shape
        assert x.is_contiguous(memory_format=torch.channels_last)
        M = B * H * W
    else:
        M, K = x.shape
        assert x.is_contiguous()
        B = H = W = None
    N = w.shape[0]
    assert w.shape[1] == K
    if dst is None:
        dst = (torch.empty((B, N, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last) if B is not None
               else torch.empty((M, N), dtype=torch.bfloat16, device=x.device))
    if dst.dim() == 4:
        assert dst.is_contiguous(memory_format=torch.channels_last) and dst.shape[0] * dst.shape[2] * dst.shape[3] == M
        dp = dst.shape[1]
    else:
        assert dst.is_contiguous() and dst.shape[0] == M
        dp = dst.shape[1]
    rp, r = 0, None
    if residual is not None:
        assert residual.dtype == torch.bfloat16
        rp = residual.shape[1]
        r = residual.data_ptr()
    with torch.cuda.device(x.device):
        _lib.check(lib.tk_conv1x1_bias_act_bf16(x.data_ptr(), M, K, K, w.data_ptr(), N, bias.data_ptr() if bias is not None else None,
                                                dst.data_ptr(), dp, dst_offset, r, rp, res_offset, act, _stream()),
                   "tk_conv1x1_bias_act_bf16"); _count()
    return dst


def conv3x3_bias_act(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None, dst: torch.Tensor | None = None, dst_offset: int = 0,
                     act: int = 1, residual: torch.Tensor | None = None, res_offset: int = 0):
    """3x3 / stride 1 / padding 1 convolution with fused epilogue on the tcgen05 path (C ABI: tk_conv3x3_bias_act_bf16).
    x bf16 channels-last [B,Cin,H,W]; w bf16 channels-last [N,Cin,3,3]; bias float32 [N]; dst / residual bf16 channels-last [B,*,H,W]."""
    lib = _lib.load()
    B, Cin, H, W = x.shape
    N = w.shape[0]
    assert x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last)
    assert w.dtype == torch.bfloat16 and tuple(w.shape[1:]) == (Cin, 3, 3) and w.is_contiguous(memory_format=torch.channels_last)
    if dst is None:
        dst = torch.empty((B, N, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    assert dst.is_contiguous(memory_format=torch.channels_last) and dst.shape[0] == B and tuple(dst.shape[2:]) == (H, W)
    rp, r = 0, None
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.is_contiguous(memory_format=torch.channels_last) and tuple(residual.shape[2:]) == (H, W)
        rp, r = residual.shape[1], residual.data_ptr()
    with torch.cuda.device(x.device):
        _lib.check(lib.tk_conv3x3_bias_act_bf16(x.data_ptr(), B, H, W, Cin, w.data_ptr(), N, bias.data_ptr() if bias is not None else None,
                                                dst.data_ptr(), dst.shape[1], dst_offset, r, rp, res_offset, act, _stream()),
                   "tk_conv3x3_bias_act_bf16"); _count()
    return dst


def ecc_gray_small(frames: torch.Tensor, scale: float = 0.1) -> torch.Tensor:
    """frames uint8 [n,H,W,3] (RGB, as the StrongSORT wrapper loads them) -> uint8 [n,h,w]: cv2.cvtColor(COLOR_BGR2GRAY) on the
    array as it is + cv2.resize(fx=fy=scale, INTER_LINEAR), bit-equal to OpenCV (C ABI: tk_ecc_gray_small)."""
    lib = _lib.load()
    _cuda(frames, "frames")
    n, H, W, _ = frames.shape
    h, w = ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.tk_ecc_small_size(H, W, float(scale), ctypes.byref(h), ctypes.byref(w)), "tk_ecc_small_size")
    out = torch.empty((n, h.value, w.value), dtype=torch.uint8, device=frames.device)
    with torch.cuda.device(frames.device):
        _lib.check(lib.tk_ecc_gray_small(frames.data_ptr(), n, H, W, frames.stride(0), float(scale), out.data_ptr(), _stream()),
                   "tk_ecc_gray_small"); _count()
    return out


def ecc_euclidean(small: torch.Tensor, max_iter: int = 100, eps: float = 1e-5, scale: float = 0.1):
    """small uint8 [n,h,w] -> (warps float32 [n,6] (row i: frame i-1 -> i; row 0 and failed pairs NaN), rho float64 [n], ok int32 [n]):
    cv2.findTransformECC(MOTION_EUCLIDEAN) of every consecutive pair, translation divided by ``scale`` (C ABI: tk_ecc_euclidean)."""
    lib = _lib.load()
    _cuda(small, "small")
    n, h, w = small.shape
    warps = torch.full((n, 6), float("nan"), dtype=torch.float32, device=small.device)
    rho = torch.zeros((n,), dtype=torch.float64, device=small.device)
    ok = torch.zeros((n,), dtype=torch.int32, device=small.device)
    with torch.cuda.device(small.device):
        _lib.check(lib.tk_ecc_euclidean(small.data_ptr(), n, h, w, max_iter, float(eps), float(scale), warps.data_ptr(), rho.data_ptr(),
                                        ok.data_ptr(), _stream()), "tk_ecc_euclidean"); _count()
    warps[ok == 0] = float("nan")      # "ecc transform failed": camera_update leaves the tracks alone (track.py:190-193,226-227)
    return warps, rho, ok


def resize_frames(frames: torch.Tensor, out_hw=(640, 640), out_dtype=torch.float32, scale: float = 1.0 / 255.0) -> torch.Tensor:
    """frames uint8 [n,H,W,3] -> [n,3,h,w] = PIL-bilinear(frame) * scale (RTDetrImageProcessor; C ABI: tk_resize_frames_u8)."""
    lib = _lib.load()
    _cuda(frames, "frames")
    n, H, W, _ = frames.shape
    out = torch.empty((n, 3, out_hw[0], out_hw[1]), dtype=out_dtype, device=frames.device)
    with torch.cuda.device(frames.device):
        _lib.check(lib.tk_resize_frames_u8(frames.data_ptr(), n, H, W, frames.stride(0), out.data_ptr(), _dtype_code(out_dtype), out_hw[0],
                                           out_hw[1], scale, _stream()), "tk_resize_frames_u8"); _count()
    return out


def rtdetr_decode(logits: torch.Tensor, boxes: torch.Tensor, image_wh, threshold: float, keep_label: int = 0):
    """logits float32 [n,Q,C], boxes float32 [n,Q,4] (cxcywh, relative) -> (rows float64 [n,Q,6] = [l,t,w,h,score,query],
    counts int32 [n]) in descending score order (C ABI: tk_rtdetr_decode)."""
    lib = _lib.load()
    _cuda(logits, "logits"); _cuda(boxes, "boxes")
    assert logits.dtype == torch.float32 and boxes.dtype == torch.float32 and logits.is_contiguous() and boxes.is_contiguous()
    n, Q, C = logits.shape
    rows = torch.empty((n, Q, 6), dtype=torch.float64, device=logits.device)
    counts = torch.empty((n,), dtype=torch.int32, device=logits.device)
    with torch.cuda.device(logits.device):
        _lib.check(lib.tk_rtdetr_decode(logits.data_ptr(), boxes.data_ptr(), n, Q, C, int(image_wh[0]), int(image_wh[1]), float(threshold),
                                        int(keep_label), rows.data_ptr(), counts.data_ptr(), _stream()), "tk_rtdetr_decode"); _count()
    return rows, counts


def maxpool3x3s2(x: torch.Tensor) -> torch.Tensor:
    """nn.MaxPool2d(3, 2, 1) on a bf16 channels-last [N,C,H,W] tensor (C ABI: tk_maxpool3x3s2_nhwc)."""
    lib = _lib.load()
    N, C, H, W = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last)
    out = torch.empty((N, C, (H + 1) // 2, (W + 1) // 2), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        _lib.check(lib.tk_maxpool3x3s2_nhwc(x.data_ptr(), N, H, W, C, out.data_ptr(), _stream()), "tk_maxpool3x3s2_nhwc"); _count()
    return out


def avgpool(x: torch.Tensor) -> torch.Tensor:
    """Global average pool of a bf16 channels-last [N,C,H,W] tensor -> float32 [N,C] (C ABI: tk_avgpool_nhwc)."""
    lib = _lib.load()
    N, C, H, W = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last)
    out = torch.empty((N, C), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.tk_avgpool_nhwc(x.data_ptr(), N, H * W, C, out.data_ptr(), _stream()), "tk_avgpool_nhwc"); _count()
    return out


def spp_pool(x: torch.Tensor, dst: torch.Tensor, dst_offset: int = 0):
    """dst[:, off:off+4C] = [x, maxpool5(x), maxpool9(x), maxpool13(x)] (C ABI: tk_spp_nhwc)."""
    lib = _lib.load()
    B, C, H, W = x.shape
    with torch.cuda.device(x.device):
        _lib.check(lib.tk_spp_nhwc(x.data_ptr(), dst.data_ptr(), B, H, W, C, dst.shape[1], dst_offset, _stream()), "tk_spp_nhwc"); _count()
    return dst


def upsample2x(src: torch.Tensor, dst: torch.Tensor, dst_offset: int = 0, src_offset: int = 0, channels: int | None = None):
    """dst[:, off:off+C, 2h, 2w] = nearest-2x(src[:, soff:soff+C]) (C ABI: tk_upsample2x_nhwc)."""
    lib = _lib.load()
    B, Cs, h, w = src.shape
    C = channels or Cs
    with torch.cuda.device(src.device):
        _lib.check(lib.tk_upsample2x_nhwc(src.data_ptr(), Cs, src_offset, dst.data_ptr(), B, h, w, C, dst.shape[1], dst_offset,
                                          _stream()), "tk_upsample2x_nhwc"); _count()
    return dst


def iou_matrix(a: torch.Tensor, b: torch.Tensor, variant: str = "iou"):
    """a [B,N,4], b [B,M,4] float64 -> [B,N,M] (C ABI: tk_iou_matrix)."""
    lib = _lib.load()
    _cuda(a, "a"); _cuda(b, "b")
    B, N, _ = a.shape
    M = b.shape[1]
    out = torch.empty((B, N, M), dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(lib.tk_iou_matrix(a.data_ptr(), b.data_ptr(), out.data_ptr(), B, N, M, _lib.ASSO_CODES[variant], _stream()),
                   "tk_iou_matrix"); _count()
    return out


def iou_p1_dist(a: torch.Tensor, b: torch.Tensor):
    """float32 tlbr boxes a [B,N,4], b [B,M,4] -> 1 - IoU(+1 px) float32 [B,N,M] (C ABI: tk_iou_p1_f32)."""
    lib = _lib.load()
    _cuda(a, "a"); _cuda(b, "b")
    B, N, _ = a.shape
    M = b.shape[1]
    out = torch.empty((B, N, M), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(lib.tk_iou_p1_f32(a.data_ptr(), b.data_ptr(), out.data_ptr(), B, N, M, _stream()), "tk_iou_p1_f32"); _count()
    return out


def cosine_dist(a: torch.Tensor, b: torch.Tensor):
    """float32 features a [B,N,E], b [B,M,E] -> float64 [B,N,M] cosine distance (C ABI: tk_cosine_dist)."""
    lib = _lib.load()
    _cuda(a, "a"); _cuda(b, "b")
    B, N, E = a.shape
    M = b.shape[1]
    out = torch.empty((B, N, M), dtype=torch.float64, device=a.device)
    scratch = torch.empty((B * (N + M),), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(lib.tk_cosine_dist(a.data_ptr(), b.data_ptr(), out.data_ptr(), scratch.data_ptr(), B, N, M, E, _stream()),
                   "tk_cosine_dist"); _count(3)
    return out


def lsap_scipy_batched(cost: torch.Tensor, status: torch.Tensor | None = None):
    """cost float64 [B,N,M] -> (x int32 [B,N], y int32 [B,M], status): scipy.optimize.linear_sum_assignment including its
    tie-breaking (C ABI: tk_lsap_scipy_batched)."""
    lib = _lib.load()
    _cuda(cost, "cost")
    B, N, M = cost.shape
    x = torch.empty((B, N), dtype=torch.int32, device=cost.device)
    y = torch.empty((B, M), dtype=torch.int32, device=cost.device)
    if status is None:
        status = torch.zeros((1,), dtype=torch.int32, device=cost.device)
    with torch.cuda.device(cost.device):
        _lib.check(lib.tk_lsap_scipy_batched(cost.data_ptr(), B, N, M, x.data_ptr(), y.data_ptr(), status.data_ptr(), _stream()),
                   "tk_lsap_scipy_batched"); _count()
    return x, y, status


def lap_batched(cost: torch.Tensor, cost_limit: float | None = None, status: torch.Tensor | None = None):
    """cost float64 [B,N,M] -> (x int32 [B,N], y int32 [B,M], status) (C ABI: tk_lap_batched)."""
    lib = _lib.load()
    _cuda(cost, "cost")
    B, N, M = cost.shape
    x = torch.empty((B, N), dtype=torch.int32, device=cost.device)
    y = torch.empty((B, M), dtype=torch.int32, device=cost.device)
    if status is None:
        status = torch.zeros((1,), dtype=torch.int32, device=cost.device)
    with torch.cuda.device(cost.device):
        _lib.check(lib.tk_lap_batched(cost.data_ptr(), B, N, M, float(cost_limit or 0.0), int(cost_limit is not None),
                                      x.data_ptr(), y.data_ptr(), status.data_ptr(), _stream()), "tk_lap_batched"); _count()
    return x, y, status


def part_dist(a: torch.Tensor, va: torch.Tensor, b: torch.Tensor, vb: torch.Tensor) -> torch.Tensor:
    """Part-based appearance distance: a [B,N,K,E], va [B,N,K], b [B,M,K,E], vb [B,M,K] float32 -> [B,N,M] float32
    (C ABI: tk_part_dist)."""
    lib = _lib.load()
    for t in (a, va, b, vb):
        _cuda(t, "part_dist input")
        assert t.dtype == torch.float32 and t.is_contiguous()
    B, N, K, E = a.shape
    M = b.shape[1]
    out = torch.empty((B, N, M), dtype=torch.float32, device=a.device)
    scratch = torch.empty((B * (N + M) * K,), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(lib.tk_part_dist(a.data_ptr(), va.data_ptr(), b.data_ptr(), vb.data_ptr(), out.data_ptr(), scratch.data_ptr(), B, N, M, K, E,
                                    _stream()), "tk_part_dist"); _count()
    return out


def kf_predict(mean: torch.Tensor, cov: torch.Tensor, model: str = "bytetrack"):
    """In-place Kalman predict of n tracks (mean float64 [n,8], cov [n,8,8]); model "bytetrack" (xyah) or "botsort" (xywh). C ABI: tk_kf_predict."""
    lib = _lib.load()
    _cuda(mean, "mean"); _cuda(cov, "cov")
    with torch.cuda.device(mean.device):
        _lib.check(lib.tk_kf_predict(mean.data_ptr(), cov.data_ptr(), mean.shape[0], {"bytetrack": 0, "botsort": 1}[model], _stream()), "tk_kf_predict"); _count()
    return mean, cov


def kf_update(mean: torch.Tensor, cov: torch.Tensor, z: torch.Tensor, model: str = "bytetrack", status: torch.Tensor | None = None):
    """In-place Kalman measurement update with one measurement z [n,4] per track. C ABI: tk_kf_update."""
    lib = _lib.load()
    _cuda(mean, "mean"); _cuda(cov, "cov"); _cuda(z, "z")
    if status is None:
        status = torch.zeros((1,), dtype=torch.int32, device=mean.device)
    with torch.cuda.device(mean.device):
        _lib.check(lib.tk_kf_update(mean.data_ptr(), cov.data_ptr(), z.data_ptr(), mean.shape[0], {"bytetrack": 0, "botsort": 1}[model],
                                    status.data_ptr(), _stream()), "tk_kf_update"); _count()
    return mean, cov, status


def vdc_cost(dets6: torch.Tensor, prev_obs5: torch.Tensor, velocities2: torch.Tensor, inertia: float, weight_col: int = 5) -> torch.Tensor:
    """OC-SORT velocity-direction-consistency cost [D,T] (C ABI: tk_vdc_cost)."""
    lib = _lib.load()
    _cuda(dets6, "dets"); _cuda(prev_obs5, "prev_obs"); _cuda(velocities2, "velocities")
    D, T = dets6.shape[0], prev_obs5.shape[0]
    out = torch.empty((D, T), dtype=torch.float64, device=dets6.device)
    with torch.cuda.device(dets6.device):
        _lib.check(lib.tk_vdc_cost(dets6.data_ptr(), prev_obs5.data_ptr(), velocities2.data_ptr(), out.data_ptr(), D, T, float(inertia),
                                   int(weight_col), _stream()), "tk_vdc_cost"); _count()
    return out


def kf_gate(mean: torch.Tensor, cov: torch.Tensor, z: torch.Tensor, aspect_const: bool = True, status: torch.Tensor | None = None):
    """Squared Mahalanobis gating distances: mean [T,8], cov [T,8,8], z [D,4] float64 -> [T,D] (C ABI: tk_kf_gate)."""
    lib = _lib.load()
    for t in (mean, cov, z):
        _cuda(t, "kf_gate input")
        assert t.dtype == torch.float64 and t.is_contiguous()
    T, D = mean.shape[0], z.shape[0]
    out = torch.empty((T, D), dtype=torch.float64, device=mean.device)
    if status is None:
        status = torch.zeros((1,), dtype=torch.int32, device=mean.device)
    with torch.cuda.device(mean.device):
        _lib.check(lib.tk_kf_gate(mean.data_ptr(), cov.data_ptr(), z.data_ptr(), out.data_ptr(), T, D, int(aspect_const), status.data_ptr(),
                                  _stream()), "tk_kf_gate"); _count()
    return out, status


REID_MEAN = (0.485, 0.456, 0.406)
REID_STD = (0.229, 0.224, 0.225)


def crop_resize_norm(frames: torch.Tensor, dets: torch.Tensor, det_frame: torch.Tensor, out_hw=(256, 128),
                     out_dtype=torch.float32, channels_last: bool = False, mean=REID_MEAN, std=REID_STD, pad_channels_to: int = 3,
                     s2d16_out: torch.Tensor | None = None, out: torch.Tensor | None = None, ltwh_rows: bool = False):
    """frames uint8 [F,H,W,3], dets float64 [N,7], det_frame int32 [N] -> ReID input [N,3,h,w] (C ABI: tk_crop_resize_norm_ex).
    ltwh_rows: rows are [l,t,w,h,..] and the crop follows the ReID wrapper's rounded/clipped rule (kpreid_api.py:118-121)
    instead of StrongSORT's centre/int() rule on [l,t,r,b,..] rows.
    pad_channels_to=8 (channels-last only) returns [N,8,h,w] with zero channels 3..7 for the fused backbone."""
    lib = _lib.load()
    _cuda(frames, "frames"); _cuda(dets, "dets"); _cuda(det_frame, "det_frame")
    F, H, W, _ = frames.shape
    N = dets.shape[0]
    if s2d16_out is not None:
        # stem layout (TK_CROP_LAYOUT_S2D16): s2d16_out is a zero-initialised channels-last [>=N, 16, h/2+3, w/2+3] buffer
        assert s2d16_out.shape[0] >= N and tuple(s2d16_out.shape[1:]) == (16, out_hw[0] // 2 + 3, out_hw[1] // 2 + 3)
        assert s2d16_out.is_contiguous(memory_format=torch.channels_last)
        m = (ctypes.c_float * 3)(*mean)
        sd = (ctypes.c_float * 3)(*std)
        with torch.cuda.device(frames.device):
            _lib.check(lib.tk_crop_resize_norm_ex(frames.data_ptr(), H, W, frames.stride(0), dets.data_ptr(), det_frame.data_ptr(), N,
                                                  s2d16_out.data_ptr(), _dtype_code(s2d16_out.dtype), -16, out_hw[0], out_hw[1], m, sd,
                                                  int(ltwh_rows), _stream()), "tk_crop_resize_norm_ex"); _count()
        return s2d16_out
    if out is not None:   # caller-owned (bucket-sized) buffer: the first N crops are written
        assert out.shape[0] >= N and out.dtype == out_dtype and pad_channels_to == 3
        assert out.is_contiguous(memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    elif pad_channels_to != 3:
        assert channels_last
        out = torch.zeros((N, pad_channels_to, out_hw[0], out_hw[1]), dtype=out_dtype,
                          device=frames.device).contiguous(memory_format=torch.channels_last)
    else:
        out = torch.empty((N, 3, out_hw[0], out_hw[1]), dtype=out_dtype, device=frames.device,
                          memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    m = (ctypes.c_float * 3)(*mean)
    sd = (ctypes.c_float * 3)(*std)
    with torch.cuda.device(frames.device):
        _lib.check(lib.tk_crop_resize_norm_ex(frames.data_ptr(), H, W, frames.stride(0), dets.data_ptr(), det_frame.data_ptr(), N,
                                              out.data_ptr(), _dtype_code(out_dtype), (pad_channels_to if channels_last else 0), out_hw[0], out_hw[1], m, sd,
                                              int(ltwh_rows), _stream()), "tk_crop_resize_norm_ex"); _count()
    return out
