"""N > 1 path on CPU: two gloo ranks partition the streams, each reconstructs its
own frames with the sequential oracle (stand-in for its GPU), no data-path
collective; the job throughput uses the max over ranks of the elapsed time."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import _d1pkg
    _d1pkg.load_pkg()
    from dav1d_mirror_b200 import dist as D
    from dav1d_mirror_b200 import frame as F
    import refdsp
    import test_frame as T
    ref = refdsp.RefDSP()
    mine = D.streams_of_rank(5, world, rank)
    sums = []
    for s in mine:
        hf = F.HostFrame(128, 96, 0x3ff, 500 + s)
        _, _, planes = T.oracle_planes(ref, hf, 500 + s)
        sums.append(int(sum(int(p.astype(np.uint64).sum()) for p in planes)))
    dist.barrier()
    ms = 10.0 * (rank + 1)                      # pretend rank 1 is slower
    worst = D.max_over_ranks(dist, ms)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, sums))
    if rank == 0:
        out.put((worst, gathered))
    dist.destroy_process_group()


def test_two_rank_stream_partition_gloo():
    from importlib import import_module
    sys.path.insert(0, ROOT)
    import _d1pkg
    _d1pkg.load_pkg()
    D = import_module("dav1d_mirror_b200.dist")
    assert D.streams_of_rank(8, 2, 0) == [0, 2, 4, 6] and D.streams_of_rank(8, 2, 1) == [1, 3, 5, 7]
    assert sorted(D.streams_of_rank(5, 2, 0) + D.streams_of_rank(5, 2, 1)) == list(range(5))
    assert abs(D.job_throughput([4, 4], [10.0, 20.0]) - 400.0) < 1e-9
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    worst, gathered = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert worst == 20.0
    all_streams = sorted(s for mine, _ in gathered for s in mine)
    assert all_streams == [0, 1, 2, 3, 4]
    # every stream was reconstructed exactly once and deterministically
    sums = {s: v for mine, vals in gathered for s, v in zip(mine, vals)}
    assert len(sums) == 5 and all(v > 0 for v in sums.values())
