"""First frame where the StrongSORT device kernel and the oracle disagree (debug aid, GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.util import load_golden
from tracklab_b200.synth import make_video
from tracklab_b200.device_trackers import StrongSortDevice, rows_to_frames

np.set_printoptions(linewidth=220, precision=3, suppress=True)
name = sys.argv[1] if len(sys.argv) > 1 else "strongsort_s4000"
ncta = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = load_golden(name)
video = make_video(**g["gen"])
trk = StrongSortDevice(video.embeddings.shape[1], **g["hyper"], min_confidence=g["min_conf"], image_size=(video.width, video.height), ctas_per_video=ncta)
rows, fc, cnt = trk.run(torch.from_numpy(video.dets).cuda(), torch.from_numpy(video.offsets.astype(np.int32))[None].cuda(), torch.from_numpy(video.embeddings).cuda())
print("status", trk.status())
got, gf = rows_to_frames(rows, fc, torch.zeros(1, dtype=torch.int32))
ref, rf = g["rows"], g["frames"]
fwd = {}
for f in range(video.n_frames):
    a = got[gf == f]; b = ref[rf == f]
    a = a[np.argsort(a[:, 7])]; b = b[np.argsort(b[:, 7])]
    bad = a.shape != b.shape or not np.array_equal(a[:, 7], b[:, 7])
    if not bad:
        for x, y in zip(a[:, 4], b[:, 4]):
            if fwd.setdefault(x, y) != y: bad = True
    if bad:
        print("first differing frame", f, a.shape, b.shape)
        sa, sb = set(a[:, 7]), set(b[:, 7])
        print(" det ids only on device:", sorted(sa - sb), " only in reference:", sorted(sb - sa))
        ia = {d: t for d, t in zip(a[:, 7], a[:, 4])}; ib = {d: t for d, t in zip(b[:, 7], b[:, 4])}
        print(" id mismatches:", [(d, ia[d], ib[d], fwd.get(ia[d])) for d in sorted(sa & sb) if fwd.get(ia[d]) != ib[d]][:10])
        print(" device rows:\n", a[:, [4, 7, 6]].T[:, :40], "\n ref rows:\n", b[:, [4, 7, 6]].T[:, :40])
        break
else:
    print("equal up to relabelling; max box diff", np.abs(got[np.lexsort((got[:,7], gf))][:, :4] - ref[np.lexsort((ref[:,7], rf))][:, :4]).max())
