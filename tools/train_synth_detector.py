"""Train the YOLOX-s/m restatement on the synthetic generator so that the detector LOCALISES the synthetic targets.

Why: there is no network for the COCO weights the reference downloads (rtmlib_api.py:19-25), and a seeded random-weight
detector cannot localise anything — with it the detect -> ReID -> associate chain would track noise. A few hundred steps on
frames drawn by tracklab_b200.synth (every generator row is drawn as a textured rectangle) give a detector whose rows are
the generator's boxes to ~1 px, so the CONNECTED chain (tracker consumes the detector's own rows) is a realistic
~38 det/frame stream on both arms. Provenance of weights/yolox_<variant>_synth.pt = this script + its seed.

    python tools/train_synth_detector.py --variant s --steps 600 --out weights/yolox_s_synth.pt

Recipe: BatchNorm after every convolution during training (folded into the bias afterwards -> the plain inference module
tracklab_b200.nets.yolox.YOLOX); loss = YOLOX's (IoU loss on positives, objectness BCE on all anchors, IoU-aware class BCE
on positives) with a fixed centre-sampling assignment instead of SimOTA; Adam, cosine schedule. Frames go through the same
letterbox the detector stage uses. Prints one JSON line with recall / precision / mean IoU of the folded model on a held-out video.
"""
import argparse
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from tracklab_b200.nets import yolox as ymod
from tracklab_b200.synth import make_frames, make_video


class ConvBNAct(ymod.ConvAct):
    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


def build_trainable(variant, seed):
    m = ymod.build_yolox(variant, 1, seed, prior_prob=0.01)
    for mod in m.modules():
        if type(mod) is ymod.ConvAct:
            mod.bn = nn.BatchNorm2d(mod.conv.out_channels, eps=1e-3, momentum=0.03)
            mod.act = nn.SiLU()
            mod.__class__ = ConvBNAct
    return m.train()


def fold(m_bn, variant):
    """ConvBNAct -> ConvAct with w' = w * g / sqrt(var + eps), b' = beta + (b - mean) * g / sqrt(var + eps)."""
    out = ymod.build_yolox(variant, 1, 0)
    src = dict(m_bn.named_modules())
    with torch.no_grad():
        for name, mod in out.named_modules():
            s = src.get(name)
            if isinstance(mod, ymod.ConvAct):
                k = s.bn.weight / torch.sqrt(s.bn.running_var + s.bn.eps)
                mod.conv.weight.copy_(s.conv.weight * k[:, None, None, None])
                mod.conv.bias.copy_(s.bn.bias + (s.conv.bias - s.bn.running_mean) * k)
            elif isinstance(mod, nn.Conv2d) and name.split(".")[0] in ("cls_preds", "reg_preds", "obj_preds"):
                mod.weight.copy_(s.weight); mod.bias.copy_(s.bias)
    return out.eval()


def grids(h, w, device):
    gs, ss = [], []
    for s in (8, 16, 32):
        ny, nx = h // s, w // s
        yv, xv = torch.meshgrid(torch.arange(ny, device=device), torch.arange(nx, device=device), indexing="ij")
        gs.append(torch.stack((xv, yv), 2).reshape(-1, 2).float())
        ss.append(torch.full((ny * nx,), float(s), device=device))
    return torch.cat(gs), torch.cat(ss)


def box_iou_pair(a, b):
    lt = torch.maximum(a[:, :2], b[:, :2]); rb = torch.minimum(a[:, 2:], b[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    ua = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]) + (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) - inter
    return inter / ua.clamp(min=1e-9)


def yolox_loss(raw, gts, h, w):
    """raw [B, A, 6] for an h x w input; gts: list of [G,4] xyxy in input pixels."""
    B, A, _ = raw.shape
    g, s = grids(h, w, raw.device)
    ctr = (g + 0.5) * s[:, None]
    xy = (raw[..., :2].float() + g) * s[:, None]
    wh = torch.exp(raw[..., 2:4].float().clamp(max=8.0)) * s[:, None]
    pred = torch.cat((xy - wh / 2, xy + wh / 2), -1)
    l_iou = raw.new_zeros((), dtype=torch.float32); l_obj = l_iou.clone(); l_cls = l_iou.clone(); n_fg = 0
    for b in range(B):
        gt = gts[b]
        obj_t = torch.zeros(A, device=raw.device)
        if len(gt):
            gc = (gt[:, :2] + gt[:, 2:]) / 2
            gsz = (gt[:, 2:] - gt[:, :2]).max(1).values
            in_box = ((ctr[None, :, 0] > gt[:, None, 0]) & (ctr[None, :, 0] < gt[:, None, 2])
                      & (ctr[None, :, 1] > gt[:, None, 1]) & (ctr[None, :, 1] < gt[:, None, 3]))
            near = ((ctr[None, :, 0] - gc[:, None, 0]).abs() <= 1.5 * s[None]) & ((ctr[None, :, 1] - gc[:, None, 1]).abs() <= 1.5 * s[None])
            lo = torch.where(s == 8, 0.0, torch.where(s == 16, 32.0, 96.0))
            hi = torch.where(s == 8, 64.0, torch.where(s == 16, 160.0, 1e9))
            lvl = (gsz[:, None] > lo[None]) & (gsz[:, None] <= hi[None])
            m = in_box & near & lvl                                         # [G, A]
            area = ((gt[:, 2] - gt[:, 0]) * (gt[:, 3] - gt[:, 1]))[:, None].expand_as(m)
            cost = torch.where(m, area, torch.full_like(area, float("inf")))
            best = cost.argmin(0)
            fg = m.any(0)
            if fg.any():
                tb = gt[best[fg]]
                iou = box_iou_pair(pred[b, fg], tb)
                l_iou = l_iou + (1.0 - iou ** 2).sum()
                l_cls = l_cls + F.binary_cross_entropy_with_logits(raw[b, fg, 5].float(), iou.detach(), reduction="sum")
                obj_t[fg] = 1.0
                n_fg += int(fg.sum())
        l_obj = l_obj + F.binary_cross_entropy_with_logits(raw[b, :, 4].float(), obj_t, reduction="sum")
    n = max(1, n_fg)
    return (5.0 * l_iou + l_obj + l_cls) / n, (l_iou.item() / n, l_obj.item() / n, l_cls.item() / n, n_fg)


def letterbox_batch(frames, dev):
    """uint8 [n,H,W,3] RGB -> float32 [n,3,640,640] BGR 0..255 (same arithmetic as the detector stage), ratio."""
    if dev.type == "cuda":
        from tracklab_b200 import kernels
        x, ratio = kernels.letterbox(frames, 640, torch.float32, swap_rb=True)
        return x, ratio
    from oracle.preprocess_np import letterbox_yolox
    xs = []
    for f in frames.numpy():
        x, ratio = letterbox_yolox(np.ascontiguousarray(f[..., ::-1]), 640)
        xs.append(torch.from_numpy(x))
    return torch.stack(xs), ratio


def sample_batch(rng, dev, n_videos, frames_per_video, n_ids):
    xs, gts = [], []
    for _ in range(n_videos):
        seed = int(rng.integers(1_000_000, 2_000_000))
        nf = int(rng.integers(frames_per_video, 40))
        v = make_video(seed=seed, n_frames=nf, n_ids=int(rng.integers(max(4, n_ids - 24), n_ids + 8)))
        pick = rng.choice(nf, size=frames_per_video, replace=False)
        for f in pick:
            fr = make_frames(v, int(f), int(f) + 1, device=dev)
            xs.append(fr)
            gts.append(torch.from_numpy(v.frame(int(f))[:, :4].astype(np.float32)).to(dev))
    frames = torch.cat(xs)
    x, ratio = letterbox_batch(frames if dev.type == "cuda" else frames.cpu(), dev)
    return x.to(dev), [g * float(ratio) for g in gts], ratio


@torch.no_grad()
def evaluate(model, dev, seed=2000, n_frames=24, score_thr=0.7, nms_thr=0.45):
    """Folded model (fp32) on a held-out video: NumPy decode/NMS oracle, recall / precision at IoU >= 0.5 and mean matched IoU."""
    from oracle.yolox_post_np import yolox_postprocess
    v = make_video(seed=seed, n_frames=n_frames, n_ids=44)
    tp = fp = fn = 0
    ious = []
    model = model.to(dev).float().eval()
    for f in range(0, n_frames, 2):
        fr = make_frames(v, f, f + 1, device=dev)
        x, ratio = letterbox_batch(fr if dev.type == "cuda" else fr.cpu(), dev)
        raw = model(x.to(dev))[0].float().cpu().numpy()
        raw[:, 4:] = 1.0 / (1.0 + np.exp(-raw[:, 4:]))
        boxes, scores, cls = yolox_postprocess(raw, np.float32(ratio), 640, score_thr, nms_thr)
        gt = v.frame(f)[:, :4]
        if len(boxes) == 0:
            fn += len(gt); continue
        b = torch.from_numpy(boxes.astype(np.float64)); g = torch.from_numpy(gt)
        lt = torch.maximum(b[:, None, :2], g[None, :, :2]); rb = torch.minimum(b[:, None, 2:], g[None, :, 2:])
        wh = (rb - lt).clamp(min=0); inter = wh[..., 0] * wh[..., 1]
        iou = inter / ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))[:, None].add(((g[:, 2] - g[:, 0]) * (g[:, 3] - g[:, 1]))[None]).sub(inter)
        best_g = iou.max(0).values
        tp_f = int((best_g >= 0.5).sum()); tp += tp_f; fn += len(gt) - tp_f
        fp += int((iou.max(1).values < 0.5).sum())
        ious.extend(best_g[best_g >= 0.5].tolist())
    return {"recall": tp / max(1, tp + fn), "precision": tp / max(1, tp + fp), "mean_iou": float(np.mean(ious)) if ious else 0.0,
            "gt": tp + fn}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="s"); ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--videos", type=int, default=4); ap.add_argument("--frames-per-video", type=int, default=4)
    ap.add_argument("--lr", type=float, default=2e-3); ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--rows", type=int, default=384, help="letterbox rows kept for training (content = 360 of 640)")
    ap.add_argument("--out", default=None); ap.add_argument("--log-every", type=int, default=25)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    a = ap.parse_args()
    dev = torch.device(a.device)
    torch.manual_seed(a.seed)
    rng = np.random.default_rng(a.seed)
    model = build_trainable(a.variant, a.seed).to(dev)
    if dev.type == "cuda":
        model = model.to(memory_format=torch.channels_last)
        torch.backends.cudnn.benchmark = True
    opt = torch.optim.AdamW(model.parameters(), lr=a.lr, weight_decay=1e-4)
    warm = max(1, a.steps // 20)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda i: min(1.0, (i + 1) / warm) * 0.5 * (1 + math.cos(math.pi * min(1.0, i / a.steps))))
    t0 = time.time()
    for step in range(a.steps):
        x, gts, _ = sample_batch(rng, dev, a.videos, a.frames_per_video, 44)
        x = x[:, :, :a.rows]
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dev.type == "cuda"):
            raw = model(x.contiguous(memory_format=torch.channels_last) if dev.type == "cuda" else x)
        loss, parts = yolox_loss(raw, gts, a.rows, 640)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        opt.step(); sched.step()
        if step % a.log_every == 0 or step == a.steps - 1:
            print(f"step {step} loss {loss.item():.4f} iou {parts[0]:.3f} obj {parts[1]:.3f} cls {parts[2]:.3f} fg {parts[3]} "
                  f"t {time.time() - t0:.0f}s", flush=True)
    folded = fold(model.float().cpu(), a.variant)
    ev = evaluate(folded, dev)
    ev.update(variant=a.variant, steps=a.steps, seed=a.seed, train_seconds=time.time() - t0)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        sd = {k: v.detach().cpu().half() for k, v in folded.state_dict().items()}
        torch.save({"state_dict": sd, "variant": a.variant, "meta": ev}, a.out)
        ev["out"] = a.out; ev["bytes"] = os.path.getsize(a.out)
    print(json.dumps(ev), flush=True)


if __name__ == "__main__":
    main()
