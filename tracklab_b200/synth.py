"""Synthetic input generators shared by the oracle, the parity tests and bench.py.

Specification: SURVEY.md Appendix C (identities with reflection at the borders, 10 % misses,
2 % false positives, occlusion bursts, running detection-id counter as the detector wrappers
do, /root/reference/tracklab/wrappers/bbox_detector/rtmlib_api.py:42-45). Everything is drawn
from ``np.random.default_rng(seed)`` so the CPU oracle and the GPU path see identical inputs.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class SyntheticVideo:
    """Detections of one synthetic video in the tracker-wrapper input layout.

    ``dets`` is ``float64[N, 7] = [l, t, r, b, conf, cls, det_id]`` (the rows the reference
    wrappers build in ``preprocess``, /root/reference/tracklab/wrappers/track/oc_sort_api.py:33-47),
    ``offsets`` is ``int32[F+1]`` with frame ``f`` owning rows ``offsets[f]:offsets[f+1]``.
    """

    width: int
    height: int
    n_frames: int
    dets: np.ndarray
    offsets: np.ndarray
    gt_identity: np.ndarray  # int32[N], -1 for false positives
    embeddings: np.ndarray | None = None  # float32[N, E] or [N, K, E]
    visibility: np.ndarray | None = None  # float32[N, K]
    seed: int = 0
    meta: dict = field(default_factory=dict)

    def frame(self, f: int) -> np.ndarray:
        return self.dets[self.offsets[f]:self.offsets[f + 1]]

    @property
    def n_dets(self) -> int:
        return int(self.dets.shape[0])


def make_video(seed: int, n_frames: int = 64, n_ids: int = 20, width: int = 1920, height: int = 1080,
               p_detect: float = 0.9, fp_rate: float = 0.02, occlusion: bool = True,
               conf_range=(0.45, 1.0), emb_dim: int | None = None, n_parts: int | None = None,
               first_det_id: int = 0, p_visible: float = 0.85, camera_drift: bool = False) -> SyntheticVideo:
    """camera_drift: the whole scene (boxes here, background in make_frames) is displaced by an integer offset per frame,
    (round(12 sin(t / 8)), round(8 cos(t / 11))) px — exercises the ECC camera compensation of StrongSORT (cfg.ecc)."""
    rng = np.random.default_rng(seed)
    W, H = float(width), float(height)
    c = rng.uniform([0.05 * W, 0.1 * H], [0.95 * W, 0.9 * H], size=(n_ids, 2))
    v = rng.normal(0.0, 0.002 * W, size=(n_ids, 2))
    w = rng.uniform(0.02 * W, 0.06 * W, size=n_ids)
    h = rng.uniform(2.0, 3.0, size=n_ids) * w
    # occlusion bursts: once per 150 frames every identity vanishes for U{3..10} frames
    hidden = np.zeros((n_frames, n_ids), dtype=bool)
    if occlusion:
        for p in range(n_ids):
            for start in range(0, n_frames, 150):
                span = min(150, n_frames - start)
                if span < 16:
                    continue
                length = int(rng.integers(3, 11))
                t0 = start + int(rng.integers(2, max(3, span - length - 1)))
                hidden[t0:t0 + length, p] = True
    proto = None
    if emb_dim is not None:
        shape = (n_ids, emb_dim) if n_parts is None else (n_ids, n_parts, emb_dim)
        proto = rng.normal(0.0, 1.0, size=shape)

    rows, offsets, gt, embs, viss = [], [0], [], [], []
    det_id = first_det_id
    for f in range(n_frames):
        c = c + v
        for ax, lim in ((0, W), (1, H)):
            lo = c[:, ax] < 0.02 * lim
            hi = c[:, ax] > 0.98 * lim
            v[lo | hi, ax] *= -1.0
            c[:, ax] = np.clip(c[:, ax], 0.02 * lim, 0.98 * lim)
        emit = (rng.uniform(size=n_ids) < p_detect) & ~hidden[f]
        jit = rng.normal(0.0, 1.0, size=(n_ids, 4))
        conf = rng.uniform(conf_range[0], conf_range[1], size=n_ids)
        n_fp = int(rng.binomial(n_ids, fp_rate))
        frame_rows = []
        for p in np.nonzero(emit)[0]:
            l = c[p, 0] - w[p] / 2 + jit[p, 0]
            t = c[p, 1] - h[p] / 2 + jit[p, 1]
            r = c[p, 0] + w[p] / 2 + jit[p, 2]
            b = c[p, 1] + h[p] / 2 + jit[p, 3]
            frame_rows.append((l, t, r, b, conf[p], 1.0, int(p)))
        for _ in range(n_fp):
            fw = rng.uniform(0.02 * W, 0.06 * W)
            fh = rng.uniform(2.0, 3.0) * fw
            fx = rng.uniform(0.0, W - fw)
            fy = rng.uniform(0.0, max(1.0, H - fh))
            frame_rows.append((fx, fy, fx + fw, fy + fh, rng.uniform(0.1, 0.6), 1.0, -1))
        if camera_drift:
            ddx, ddy = camera_offset(f)
            frame_rows = [(l + ddx, t + ddy, r + ddx, b + ddy, s, k, p) for (l, t, r, b, s, k, p) in frame_rows]
        for (l, t, r, b, s, k, p) in frame_rows:
            # clip like the detector wrappers do (coordinates.py:270-295) so boxes stay in the image
            l = min(max(l, 0.0), W - 2.0)
            t = min(max(t, 0.0), H - 2.0)
            r = min(max(r, l + 1.0), W - 1.0)
            b = min(max(b, t + 1.0), H - 1.0)
            rows.append((l, t, r, b, s, k, float(det_id)))
            gt.append(p)
            if proto is not None:
                if p >= 0:
                    e = proto[p] + 0.15 * rng.normal(0.0, 1.0, size=proto[p].shape)
                else:
                    e = rng.normal(0.0, 1.0, size=proto[0].shape)
                embs.append(e.astype(np.float32))
                if n_parts is not None:
                    vis = (rng.uniform(size=n_parts) < p_visible).astype(np.float32)
                    vis[0] = 1.0
                    viss.append(vis)
            det_id += 1
        offsets.append(len(rows))
    dets = np.asarray(rows, dtype=np.float64).reshape(-1, 7)
    return SyntheticVideo(
        width=width, height=height, n_frames=n_frames, dets=dets,
        offsets=np.asarray(offsets, dtype=np.int32), gt_identity=np.asarray(gt, dtype=np.int32),
        embeddings=(np.stack(embs) if embs else None),
        visibility=(np.stack(viss) if viss else None), seed=seed,
        meta=dict(n_ids=n_ids, p_detect=p_detect, fp_rate=fp_rate, occlusion=occlusion, camera_drift=camera_drift),
    )


def camera_offset(f: int):
    """Integer scene displacement of frame ``f`` of a camera_drift video."""
    return int(np.rint(12.0 * np.sin(f / 8.0))), int(np.rint(8.0 * np.cos(f / 11.0)))


def make_frames(video: SyntheticVideo, f0: int, f1: int, device="cpu"):
    """RGB uint8 frames ``[f1-f0, H, W, 3]`` for frames ``f0:f1`` of ``video`` (torch tensor).

    Integer arithmetic only (Appendix C): a per-video low-frequency background plus one textured
    rectangle per detection, so the same bytes come out on CPU and on the GPU.
    """
    import torch

    H, W = video.height, video.width
    g = np.random.default_rng(video.seed + 7919)
    coarse = g.integers(40, 200, size=(H // 120 + 2, W // 120 + 2, 3), dtype=np.int64)
    tex = g.integers(0, 256, size=(64, 8, 8, 3), dtype=np.int64)
    dev = torch.device(device)
    coarse_t = torch.from_numpy(coarse).to(dev)
    ys = torch.arange(H, device=dev)
    xs = torch.arange(W, device=dev)
    # bilinear-free integer blend: nearest coarse cell + position hash
    drift = bool(video.meta.get("camera_drift", False))
    if drift:   # background on a canvas with a 32-px margin; every frame is a window displaced by the camera offset
        ye = torch.arange(H + 64, device=dev)
        xe = torch.arange(W + 64, device=dev)
        big = coarse_t[(ye // 120).clamp(max=coarse.shape[0] - 1)[:, None], (xe // 120).clamp(max=coarse.shape[1] - 1)[None, :]]
        big = (big + ((ye[:, None] * 7 + xe[None, :] * 13) % 32)[..., None]).clamp(0, 255).to(torch.uint8)
        out = torch.stack([big[32 - camera_offset(f)[1]:32 - camera_offset(f)[1] + H, 32 - camera_offset(f)[0]:32 - camera_offset(f)[0] + W]
                           for f in range(f0, f1)]).contiguous()
        tex_t = torch.from_numpy(tex).to(dev).to(torch.uint8)
    else:
        bg = coarse_t[(ys // 120)[:, None], (xs // 120)[None, :]]
        bg = (bg + ((ys[:, None] * 7 + xs[None, :] * 13) % 32)[..., None]).clamp(0, 255).to(torch.uint8)
        tex_t = torch.from_numpy(tex).to(dev).to(torch.uint8)
        out = bg.unsqueeze(0).repeat(f1 - f0, 1, 1, 1)
    for i, f in enumerate(range(f0, f1)):
        rows = video.frame(f)
        ids = video.gt_identity[video.offsets[f]:video.offsets[f + 1]]
        for (l, t, r, b, *_), p in zip(rows, ids):
            x0, y0, x1, y1 = int(l), int(t), int(r), int(b)
            if x1 <= x0 or y1 <= y0:
                continue
            ty = ((torch.arange(y0, y1, device=dev) - y0) * 8 // max(1, y1 - y0)).clamp(0, 7)
            tx = ((torch.arange(x0, x1, device=dev) - x0) * 8 // max(1, x1 - x0)).clamp(0, 7)
            out[i, y0:y1, x0:x1] = tex_t[int(p) % 64][ty[:, None], tx[None, :]]
    return out
