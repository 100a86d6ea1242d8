"""Top CUDA kernels of one detector batch (torch.profiler table): python tools/profile_detector.py [batch]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from tracklab_b200.detector import YoloxDetectorDevice

B = int(sys.argv[1]) if len(sys.argv) > 1 else 50
det = YoloxDetectorDevice("s", batch=B, use_graph=False)
frames = torch.randint(0, 255, (B, 1080, 1920, 3), dtype=torch.uint8, device="cuda")
det.calibrate(frames)
for _ in range(5):
    det.reset(); det.detect_batch(frames)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        det.reset(); det.detect_batch(frames)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
