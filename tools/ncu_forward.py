"""One eager YOLOX-m detector forward (batch 20) + one eager ResNet-50 ReID forward (768 crops) inside a cudaProfilerStart/Stop range:
    ncu --profile-from-start off --metrics <dram bytes, duration, tensor pipe> ... python tools/ncu_forward.py
Gives per-launch DRAM traffic + tensor-pipe activity of every kernel of the two networks (repo kernels and cuDNN's)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tracklab_b200.detector import YoloxDetectorDevice, synth_weights_path
from tracklab_b200.reid import ReidStageDevice

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="m"); ap.add_argument("--batch", type=int, default=20); ap.add_argument("--crops", type=int, default=768)
a = ap.parse_args()
dev = torch.device("cuda:0")
det = YoloxDetectorDevice(a.variant, device=dev, batch=a.batch, frames_cap=64, dets_cap=1 << 14, use_graph=False, weights=synth_weights_path(a.variant))
reid = ReidStageDevice(device=dev, use_graphs=False)
buf = reid.fused.input_buffer(a.crops)
with torch.no_grad():
    for _ in range(3):
        det.fused(det.x); reid.fused(buf, n_valid=a.crops)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    det.fused(det.x); reid.fused(buf, n_valid=a.crops)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done", det.fused.tc_layers)
