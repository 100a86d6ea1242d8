"""Strict-parity probe (GPU): for every StrongSORT / BPBReID golden and oracle comparison, report whether the device rows
equal the reference rows WITHOUT the relabelling escape and WITHOUT the 1-pixel box tolerance.
Prints one JSON line per case: ids_exact, n_box_rows_differ, max_box_err. Usage: python tools/strict_parity_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tests.util import BPB_KEYS, load_bpbreid_golden, load_golden
from tracklab_b200.device_trackers import BpbreidStrongSortDevice, StrongSortDevice, rows_to_frames
from tracklab_b200.synth import make_video


def cmp8(name, rows, fr, ref, rf):
    out = {"case": name, "shape_equal": rows.shape == ref.shape}
    if rows.shape == ref.shape:
        ka, kb = np.lexsort((rows[:, 7], fr)), np.lexsort((ref[:, 7], rf))
        a, b = rows[ka], ref[kb]
        out["frames_equal"] = bool(np.array_equal(fr[ka], rf[kb]))
        out["det_ids_equal"] = bool(np.array_equal(a[:, 7], b[:, 7]))
        out["ids_exact"] = bool(np.array_equal(a[:, 4], b[:, 4]))
        d = np.abs(a[:, :4] - b[:, :4])
        out["n_box_rows_differ"] = int((d.max(1) > 0).sum()) if len(a) else 0
        out["max_box_err"] = float(d.max()) if len(a) else 0.0
        out["rows"] = int(len(a))
    print(json.dumps(out), flush=True)


def ss_device(video, hyper, min_conf, ncta=8, feats=None):
    trk = StrongSortDevice(video.embeddings.shape[1] if feats is None else feats.shape[1], **hyper, min_confidence=min_conf,
                           image_size=(video.width, video.height), ctas_per_video=ncta)
    dets = torch.from_numpy(video.dets).cuda()
    offs = torch.from_numpy(video.offsets.astype(np.int32))[None].cuda()
    rows, fc, cnt = trk.run(dets, offs, torch.from_numpy(video.embeddings).cuda() if feats is None else feats)
    trk.check_status()
    return rows_to_frames(rows, fc, torch.zeros(1, dtype=torch.int32))


for name in ("strongsort_s4000", "strongsort_budget8_s4001"):
    g = load_golden(name)
    v = make_video(**g["gen"])
    for ncta in (1, 8):
        r, f = ss_device(v, g["hyper"], g["min_conf"], ncta)
        cmp8(f"{name}/ncta{ncta}", r, f, g["rows"], g["frames"])

from oracle.strongsort_np import StrongSortOracle
v = make_video(seed=41, n_frames=120, n_ids=40, emb_dim=256)
hyper = dict(max_dist=0.16, max_iou_dist=0.55, max_age=30, max_unmatched_preds=0, n_init=3, nn_budget=50, mc_lambda=0.995, ema_alpha=0.9)
rr, rf = StrongSortOracle(**hyper, min_confidence=0.4, image_size=(v.width, v.height)).run_video(v.dets, v.offsets, v.embeddings)
r, f = ss_device(v, hyper, 0.4)
cmp8("strongsort_oracle_seed41", r, f, rr, rf)

# e2e with ReID
from tracklab_b200.reid import ReidStageDevice
from tracklab_b200.synth import make_frames
for gname in ("strongsort_e2e_s5000", "strongsort_e2e_s5001"):
    if not os.path.exists(os.path.join("tests", "golden", gname + ".npz")):
        continue
    g = load_golden(gname)
    v = make_video(**g["gen"])
    frames = make_frames(v, 0, v.n_frames, device="cuda")
    dets = torch.from_numpy(v.dets).cuda()
    det_frame = torch.from_numpy(np.repeat(np.arange(v.n_frames), np.diff(v.offsets)).astype(np.int32)).cuda()
    feats = {}
    for prec in ("fp32", "bf16"):
        feats[prec] = ReidStageDevice(precision=prec).features(frames, dets, det_frame)
        r, f = ss_device(v, g["hyper"], g["min_conf"], 8, feats[prec])
        cmp8(f"{gname}/{prec}", r, f, g["rows"], g["frames"])
    a = torch.nn.functional.normalize(feats["fp32"], dim=1)
    b = torch.nn.functional.normalize(feats["bf16"], dim=1)
    dd = (1 - a @ a.T) - (1 - b @ b.T)
    dc = (1 - a @ a.T) - (1 - a @ b.T)
    print(json.dumps({"case": gname + "/bf16_vs_fp32_cosine_distance", "max_abs_delta_both_bf16": float(dd.abs().max()),
                      "max_abs_delta_one_side": float(dc.abs().max()),
                      "cos_min": float((a * b).sum(1).min())}), flush=True)


def bp_device(v, hyper, ncta=8, cap=128):
    K, E = v.embeddings.shape[1:]
    trk = BpbreidStrongSortDevice(K, E, **{k: hyper[k] for k in BPB_KEYS}, ctas_per_video=ncta, cap_tracks=8 * cap, cap_dets=cap)
    d = v.dets.copy(); d[:, 2] -= d[:, 0]; d[:, 3] -= d[:, 1]
    rows, fc, cnt = trk.run(torch.from_numpy(d).cuda(), torch.from_numpy(v.offsets.astype(np.int32))[None].cuda(),
                            torch.from_numpy(v.embeddings).cuda(), torch.from_numpy(v.visibility.astype(np.float32)).cuda())
    trk.check_status()
    return rows_to_frames(rows, fc, torch.zeros(1, dtype=torch.int32))


for name, ncta in (("bpbreid_yaml_s6000", 8), ("bpbreid_tight_s6001", 1), ("bpbreid_tight_s6001", 24)):
    g = load_bpbreid_golden(name)
    v = make_video(**g["gen"])
    rows, fr = bp_device(v, g["hyper"], ncta)
    ref, rf = g["rows"], g["frames"]
    out = {"case": f"{name}/ncta{ncta}", "shape_equal": rows.shape == ref.shape}
    if rows.shape == ref.shape:
        ka, kb = np.lexsort((rows[:, 13], fr)), np.lexsort((ref[:, 13], rf))
        a, b = rows[ka], ref[kb]
        out["ids_exact"] = bool(np.array_equal(a[:, 0], b[:, 0]))
        out["int_cols_exact"] = bool(all(np.array_equal(a[:, c], b[:, c]) for c in (9, 11, 12, 13)))
        out["max_box_err"] = float(np.nanmax(np.abs(a[:, 1:9] - b[:, 1:9])))
        out["max_dist_err"] = float(np.nanmax(np.abs(a[:, 10] - b[:, 10])))
    print(json.dumps(out), flush=True)
