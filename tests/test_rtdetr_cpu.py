"""CPU: the RT-DETR post-processing oracle against transformers' own post_process_object_detection (pip package in the image)
followed by the reference wrapper's conversion (golden from /root/reference's coordinates.py, tests/golden/rtdetr_post.npz)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_rtdetr_post_oracle_matches_transformers_and_wrapper_golden():
    torch = pytest.importorskip("torch")
    from transformers import RTDetrImageProcessor
    from transformers.models.rt_detr.modeling_rt_detr import RTDetrObjectDetectionOutput
    from oracle.rtdetr_post_np import post_process
    g = np.load(os.path.join(HERE, "golden", "rtdetr_post.npz"))
    logits, boxes = g["logits"], g["boxes"]
    W, H, thr = int(g["W"]), int(g["H"]), float(g["threshold"])
    out = RTDetrObjectDetectionOutput(logits=torch.from_numpy(logits), pred_boxes=torch.from_numpy(boxes))
    res = RTDetrImageProcessor().post_process_object_detection(out, target_sizes=[(H, W)] * len(logits), threshold=thr)
    for i in range(len(logits)):
        rows = post_process(logits[i], boxes[i], (W, H), thr, keep_label=0)
        keep = res[i]["labels"].numpy() == 0
        assert np.allclose(rows[:, 4], res[i]["scores"].numpy()[keep], atol=1e-7)           # transformers' scores, label 0, same order
        ref = g[f"rows_{i}"]                                                                   # wrapper golden: ltwh float32 + score
        assert rows.shape[0] == ref.shape[0] and np.array_equal(rows[:, :4], ref[:, :4])
        assert np.allclose(rows[:, 4], ref[:, 4], atol=1e-7)
