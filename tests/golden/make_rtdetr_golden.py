"""Golden for the RT-DETR post-processing path (build container only): random logits/boxes -> transformers'
post_process_object_detection -> the reference wrapper's loop (transformers_api.py:37-53 with coordinates.ltrb_to_ltwh)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = ["/root/reference"]
from transformers import RTDetrImageProcessor  # noqa: E402
from transformers.models.rt_detr.modeling_rt_detr import RTDetrObjectDetectionOutput  # noqa: E402
import importlib.util  # noqa: E402
spec = importlib.util.spec_from_file_location("coordinates", "/root/reference/tracklab/utils/coordinates.py")
coords = importlib.util.module_from_spec(spec); spec.loader.exec_module(coords)

rng = np.random.default_rng(7)
n, Q, C, W, H, thr = 3, 300, 80, 1920, 1080, 0.4
logits = rng.normal(-2.0, 1.5, size=(n, Q, C)).astype(np.float32)
logits[:, :, 0] += 1.5
cxcy = rng.uniform(-0.05, 1.05, size=(n, Q, 2)); wh = rng.uniform(0.01, 0.3, size=(n, Q, 2))
boxes = np.concatenate([cxcy, wh], axis=2).astype(np.float32)
out = RTDetrObjectDetectionOutput(logits=torch.from_numpy(logits), pred_boxes=torch.from_numpy(boxes))
res = RTDetrImageProcessor().post_process_object_detection(out, target_sizes=[(H, W)] * n, threshold=thr)
save = dict(logits=logits, boxes=boxes, W=W, H=H, threshold=thr)
for i, r in enumerate(res):
    rows = []
    for score, label, box in zip(r["scores"], r["labels"], r["boxes"]):
        if label == 0:
            ltwh = coords.ltrb_to_ltwh(box.numpy(), (W, H))
            rows.append([*ltwh, score.item()])
    save[f"rows_{i}"] = np.asarray(rows, dtype=np.float64).reshape(-1, 5)
    print(i, len(rows), "label-0 rows")
np.savez_compressed(os.path.join(HERE, "rtdetr_post.npz"), **save)
