import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
from tracklab_b200.ingest import JpegDecoderDevice
from tracklab_b200.synth import make_frames, make_video
video = make_video(seed=3300, n_frames=3, n_ids=20)
frames = make_frames(video, 0, 3, device="cpu").numpy()
tmp = tempfile.mkdtemp(); paths = []
for f in range(3):
    p = os.path.join(tmp, f"{f}.jpg"); cv2.imwrite(p, frames[f][..., ::-1], [cv2.IMWRITE_JPEG_QUALITY, 92]); paths.append(p)
ref = np.stack([cv2.cvtColor(cv2.imread(p), cv2.COLOR_BGR2RGB) for p in paths]).astype(np.int16)
for hw in (True, False):
    dec = JpegDecoderDevice("cuda:0", prefer_hardware=hw)
    out = dec.decode([open(p, "rb").read() for p in paths]).cpu().numpy().astype(np.int16)
    d = np.abs(out - ref); d2 = np.abs(out[..., ::-1] - ref)
    print("prefer_hw", hw, "backend", dec.backend, "mean", d.mean(), "p99.9", np.percentile(d, 99.9), "max", d.max(), "| if BGR:", d2.mean(),
          "| per-channel mean", d.mean(axis=(0, 1, 2)), "frame means", out.mean(axis=(1, 2, 3)), ref.mean(axis=(1, 2, 3)))
try:
    import torchvision
    tv = torch.stack([torchvision.io.decode_jpeg(torchvision.io.read_file(p), device="cuda") for p in paths]).permute(0, 2, 3, 1).cpu().numpy().astype(np.int16)
    print("torchvision nvjpeg vs cv2: mean", np.abs(tv - ref).mean(), "max", np.abs(tv - ref).max())
except Exception as e:
    print("torchvision:", type(e).__name__, e)
