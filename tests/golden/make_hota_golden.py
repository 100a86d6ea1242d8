"""Golden for the HOTA path: the UNMODIFIED HOTA.eval_sequence + _calculate_box_ious of the TrackEval fork vendored in the reference,
run in the build container on boxes (tests/golden/hota_boxes_*.npz). Only hota.py / _base_metric.py / _base_dataset.py / _timing.py /
utils.py of the fork are executed; its package __init__ files import un-installable packages (shapely, ...) and are bypassed by
registering bare package objects.

    python tests/golden/make_hota_golden.py
"""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
TE = "/root/reference/plugins/eval/PoseTrack21/posetrack21/posetrack21/trackeval"


def load_trackeval():
    for name, path in (("trackeval", TE), ("trackeval.metrics", TE + "/metrics"), ("trackeval.datasets", TE + "/datasets")):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    hota = importlib.import_module("trackeval.metrics.hota")
    base = importlib.import_module("trackeval.datasets._base_dataset")
    return hota.HOTA, base._BaseDataset


def reference_hota(gt_boxes, gt_ids, gt_off, tr_boxes, tr_ids, tr_off):
    """boxes xywh; ids contiguous ints. Returns (res dict of the reference, list of similarity matrices)."""
    HOTA, Base = load_trackeval()
    F = len(gt_off) - 1
    data = {"gt_ids": [], "tracker_ids": [], "similarity_scores": []}
    for f in range(F):
        g = slice(gt_off[f], gt_off[f + 1]); t = slice(tr_off[f], tr_off[f + 1])
        data["gt_ids"].append(gt_ids[g].astype(int)); data["tracker_ids"].append(tr_ids[t].astype(int))
        data["similarity_scores"].append(Base._calculate_box_ious(gt_boxes[g], tr_boxes[t], box_format="xywh"))
    data["num_gt_dets"], data["num_tracker_dets"] = int(len(gt_ids)), int(len(tr_ids))
    data["num_gt_ids"] = int(gt_ids.max()) + 1 if len(gt_ids) else 0
    data["num_tracker_ids"] = int(tr_ids.max()) + 1 if len(tr_ids) else 0
    data["num_timesteps"] = F
    return HOTA().eval_sequence(data), data["similarity_scores"]


def case_from_tracker_golden(name, drop_frames=()):
    """gt = the generator's identities on the detections, tracker = the rows the reference ByteTrack plugin produced (golden)."""
    from tests.util import load_golden
    from tracklab_b200.synth import make_video
    g = load_golden(name)
    v = make_video(**g["gen"])
    F = v.n_frames
    gb, gi, go, tb, ti, to = [], [], [0], [], [], [0]
    uniq = {u: k for k, u in enumerate(np.unique(g["rows"][:, 4]))}
    gmap = {u: k for k, u in enumerate(np.unique(v.gt_identity[v.gt_identity >= 0]))}
    for f in range(F):
        sl = slice(v.offsets[f], v.offsets[f + 1])
        keep = v.gt_identity[sl] >= 0
        d = v.dets[sl][keep]
        if f in drop_frames:
            d = d[:0]; keep[:] = False
        gb.append(np.column_stack([d[:, 0], d[:, 1], d[:, 2] - d[:, 0], d[:, 3] - d[:, 1]]).reshape(-1, 4))
        gi.append(np.array([gmap[x] for x in v.gt_identity[sl][keep]], dtype=np.int32))
        r = g["rows"][g["frames"] == f]
        if (f + 7) in drop_frames:
            r = r[:0]
        tb.append(np.column_stack([r[:, 0], r[:, 1], r[:, 2] - r[:, 0], r[:, 3] - r[:, 1]]).reshape(-1, 4))
        ti.append(np.array([uniq[x] for x in r[:, 4]], dtype=np.int32))
        go.append(go[-1] + len(gi[-1])); to.append(to[-1] + len(ti[-1]))
    return (np.concatenate(gb), np.concatenate(gi), np.array(go, dtype=np.int32), np.concatenate(tb), np.concatenate(ti),
            np.array(to, dtype=np.int32))


def case_random(seed, F, n_ids, n_false, jitter):
    """Tie-free random case with id switches, false positives, misses, empty frames on both sides and a degenerate box."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(100, 1700, (n_ids, 2)); vel = rng.normal(0, 6, (n_ids, 2)); wh = rng.uniform(40, 160, (n_ids, 2))
    tid = np.arange(n_ids)
    gb, gi, go, tb, ti, to = [], [], [0], [], [], [0]
    nxt = n_ids
    for f in range(F):
        pos += vel
        present = rng.random(n_ids) < 0.9
        if f in (5, 17):
            present[:] = False
        b = np.column_stack([pos - wh / 2, wh])[present]
        gb.append(b); gi.append(np.arange(n_ids)[present].astype(np.int32))
        det = rng.random(n_ids) < 0.85
        if f in (9,):
            det[:] = False
        sw = rng.random(n_ids) < 0.02
        for k in np.nonzero(sw)[0]:
            tid[k] = nxt; nxt += 1
        m = present & det
        tbx = np.column_stack([pos - wh / 2 + rng.normal(0, jitter, (n_ids, 2)), wh * rng.uniform(0.9, 1.1, (n_ids, 2))])[m]
        fp = np.column_stack([rng.uniform(0, 1800, (n_false, 2)), rng.uniform(30, 120, (n_false, 2))])
        if f == 3:
            fp[0, 2:] = 0.0          # zero-area tracker box: the reference zeroes its intersections
        ids_fp = np.arange(nxt, nxt + n_false); nxt += n_false
        tb.append(np.vstack([tbx, fp])); ti.append(np.concatenate([tid[m], ids_fp]).astype(np.int32))
        go.append(go[-1] + len(gi[-1])); to.append(to[-1] + len(ti[-1]))
    ti = np.concatenate(ti)
    _, ti = np.unique(ti, return_inverse=True)
    gi = np.concatenate(gi)
    _, gi = np.unique(gi, return_inverse=True)
    return (np.concatenate(gb), gi.astype(np.int32), np.array(go, dtype=np.int32), np.concatenate(tb), ti.astype(np.int32),
            np.array(to, dtype=np.int32))


def main():
    cases = {
        "hota_boxes_bytetrack_c2": case_from_tracker_golden("bytetrack_c2_s2000", drop_frames=(20, 21)),
        "hota_boxes_random_s7": case_random(7, 60, 24, 3, 4.0),
        "hota_boxes_random_wide_s8": case_random(8, 25, 150, 10, 8.0),      # 150 ids: rows > 128 exercise the pairwise split
    }
    for name, (gb, gi, go, tb, ti, to) in cases.items():
        res, sims = reference_hota(gb, gi, go, tb, ti, to)
        out = {"gt_boxes": gb, "gt_ids": gi, "gt_off": go, "tr_boxes": tb, "tr_ids": ti, "tr_off": to,
               "sims": np.concatenate([s.ravel() for s in sims]) if name.endswith("random_s7") else np.zeros(0)}
        for k in ("HOTA", "DetA", "AssA", "DetRe", "DetPr", "AssRe", "AssPr", "LocA", "HOTA_TP", "HOTA_FN", "HOTA_FP"):
            out["ref_" + k] = np.asarray(res[k], dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "frames", len(go) - 1, "gt rows", len(gi), "tracker rows", len(ti), "HOTA", float(np.mean(res["HOTA"])),
              "DetA", float(np.mean(res["DetA"])), "AssA", float(np.mean(res["AssA"])))


if __name__ == "__main__":
    main()
