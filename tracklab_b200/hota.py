"""HOTA of tracker rows on the device (C ABI: tk_hota_sequence).

Mirrors HOTA.eval_sequence of the TrackEval fork vendored in the reference
(/root/reference/plugins/eval/PoseTrack21/posetrack21/posetrack21/trackeval/metrics/hota.py:28-154) on the inputs its MOT dataset
class prepares (posetrack_mot.py:479: per-frame gt / tracker boxes in xywh, ids mapped to 0..n-1): frame-major rows that are
already on the device (the trackers' output rows), no MOT text files, no per-frame Python.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .kernels import _count, _cuda, _stream

ALPHAS = np.arange(0.05, 0.99, 0.05)          # HOTA.array_labels (hota.py:18)
FIELDS = ("HOTA", "DetA", "AssA", "DetRe", "DetPr", "AssRe", "AssPr", "LocA", "HOTA_TP", "HOTA_FN", "HOTA_FP")


class HotaDevice:
    """Workspace + launch of the device HOTA for sequences up to the capacities given (re-usable across videos)."""

    def __init__(self, n_frames: int, n_gt_ids: int, n_tr_ids: int, pairs_cap: int, max_gt_per_frame: int, max_tr_per_frame: int,
                 device="cuda:0", alphas=ALPHAS):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.TrackKernError("HotaDevice needs a CUDA device (no CPU path)")
        self.device = torch.device(device)
        self.cap = (int(n_frames), int(n_gt_ids), int(n_tr_ids), int(pairs_cap), int(max_gt_per_frame), int(max_tr_per_frame))
        self.alphas = np.ascontiguousarray(alphas, dtype=np.float64)
        nbytes = ctypes.c_longlong(0)
        _lib.check(self.lib.tk_hota_workspace_bytes(self.cap[0], self.cap[1], self.cap[2], len(self.alphas), self.cap[3],
                                                    ctypes.byref(nbytes)), "tk_hota_workspace_bytes")
        self.workspace = torch.empty((max(int(nbytes.value), 16),), dtype=torch.uint8, device=self.device)
        self.out = torch.zeros((len(FIELDS), len(self.alphas)), dtype=torch.float64, device=self.device)
        self.status = torch.zeros((1,), dtype=torch.int32, device=self.device)

    def run(self, gt_boxes_xywh, gt_ids, gt_offsets, tr_boxes_xywh, tr_ids, tr_offsets, n_gt_ids: int, n_tr_ids: int):
        """boxes float64 [rows,4] xywh, ids int32 [rows] in 0..n_ids-1, offsets int32 [F+1] (device tensors). Returns the device
        tensor [11, n_alphas] (rows = FIELDS); nothing is synchronised."""
        F = gt_offsets.numel() - 1
        if tr_offsets.numel() - 1 != F or F > self.cap[0] or n_gt_ids > self.cap[1] or n_tr_ids > self.cap[2]:
            raise _lib.TrackKernError("HotaDevice.run: sequence exceeds the capacities of this workspace")
        for t, nm, dt in ((gt_boxes_xywh, "gt_boxes", torch.float64), (gt_ids, "gt_ids", torch.int32), (gt_offsets, "gt_offsets", torch.int32),
                          (tr_boxes_xywh, "tr_boxes", torch.float64), (tr_ids, "tr_ids", torch.int32), (tr_offsets, "tr_offsets", torch.int32)):
            _cuda(t, nm)
            if t.dtype != dt:
                raise _lib.TrackKernError(f"{nm} must be {dt}")
        self.status.zero_()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tk_hota_sequence(
                gt_boxes_xywh.data_ptr(), gt_ids.data_ptr(), gt_offsets.data_ptr(), gt_ids.numel(),
                tr_boxes_xywh.data_ptr(), tr_ids.data_ptr(), tr_offsets.data_ptr(), tr_ids.numel(), F, int(n_gt_ids), int(n_tr_ids),
                self.cap[4], self.cap[5], self.alphas.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(self.alphas), self.cap[3],
                self.workspace.data_ptr(), self.workspace.numel(), self.out.data_ptr(), self.status.data_ptr(), _stream()),
                "tk_hota_sequence")
            _count(6)
        return self.out

    def check_status(self):
        st = int(self.status.item())
        if st:
            raise _lib.TrackKernError("tk_hota_sequence: " + _lib.status_text(st))

    def result(self) -> dict:
        """Host dictionary like the reference's `res` (arrays over the alphas) + the alpha-averaged summary values."""
        self.check_status()
        o = self.out.cpu().numpy()
        res = {k: o[i].copy() for i, k in enumerate(FIELDS)}
        return res


def frame_major(frame_idx: torch.Tensor, n_frames: int):
    """Stable order of rows by frame + int32 offsets [n_frames+1] (device)."""
    order = torch.argsort(frame_idx, stable=True)
    counts = torch.bincount(frame_idx.to(torch.int64), minlength=n_frames)
    off = torch.zeros((n_frames + 1,), dtype=torch.int32, device=frame_idx.device)
    off[1:] = torch.cumsum(counts, 0).to(torch.int32)
    return order, off


def hota_of_rows(gt_boxes_ltwh: torch.Tensor, gt_ids: torch.Tensor, gt_frame: torch.Tensor, tr_boxes_ltwh: torch.Tensor,
                 tr_ids: torch.Tensor, tr_frame: torch.Tensor, n_frames: int) -> dict:
    """Convenience path: arbitrary integer ids and row -> frame indices (device tensors) -> contiguous ids (the order of
    np.unique, like the dataset class of the reference), frame-major rows, one tk_hota_sequence call."""
    dev = gt_boxes_ltwh.device
    go, goff = frame_major(gt_frame, n_frames)
    to, toff = frame_major(tr_frame, n_frames)
    gu, gi = torch.unique(gt_ids[go], return_inverse=True)
    tu, ti = torch.unique(tr_ids[to], return_inverse=True)
    ng = (goff[1:] - goff[:-1]); nt = (toff[1:] - toff[:-1])
    pairs = int((ng.to(torch.int64) * nt.to(torch.int64)).sum().item())
    h = HotaDevice(n_frames, max(int(gu.numel()), 1), max(int(tu.numel()), 1), max(pairs, 1), int(ng.max().item()) if n_frames else 0,
                   int(nt.max().item()) if n_frames else 0, device=dev)
    h.run(gt_boxes_ltwh[go].to(torch.float64).contiguous(), gi.to(torch.int32).contiguous(), goff,
          tr_boxes_ltwh[to].to(torch.float64).contiguous(), ti.to(torch.int32).contiguous(), toff, int(gu.numel()), int(tu.numel()))
    return h.result()
