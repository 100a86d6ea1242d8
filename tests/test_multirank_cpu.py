"""CPU, world_size 2 over gloo: video sharding + the single all_gather of per-video metrics."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.bytetrack_np import ByteTrackOracle   # the CPU test drives the oracle tracker; the sharding logic is the product's
    from tracklab_b200 import dist as tdist
    from tracklab_b200.synth import make_video
    mine = tdist.shard_videos(4, rank, world)
    rows = []
    for v in mine:
        video = make_video(seed=2000 + v, n_frames=24, n_ids=10)
        out, _ = ByteTrackOracle().run_video(video.dets, video.offsets)
        rows.append([video.n_frames, video.n_dets, len(out), len(np.unique(out[:, 4])), 1.0 + rank])
    allm = tdist.gather_video_metrics(torch.tensor(rows, dtype=torch.float64))
    slow = tdist.max_over_ranks(10.0 * (rank + 1), "cpu")
    q.put((rank, mine, allm.numpy(), slow))
    dist.barrier()
    dist.destroy_process_group()


def test_video_sharding_and_metric_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2] and res[1][1] == [1, 3]                 # round-robin shards, disjoint and complete
    assert np.array_equal(res[0][2], res[1][2]) and res[0][2].shape == (2, 2, 5)   # every rank sees every video's metrics
    assert res[0][2][:, :, 0].sum() == 4 * 24 and (res[0][2][1, :, 4] == 2.0).all()
    assert res[0][3] == res[1][3] == 20.0                               # job time = slowest rank


def test_shard_videos_covers_everything():
    from tracklab_b200.dist import shard_videos
    for n, w in [(8, 8), (8, 4), (5, 2), (1, 4)]:
        got = sorted(i for r in range(w) for i in shard_videos(n, r, w))
        assert got == list(range(n))
