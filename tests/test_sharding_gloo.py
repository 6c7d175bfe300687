"""Multi-rank host logic on CPU: frame partition + the single gather of result rows, world_size 2, gloo."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, nframes, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from slv_b200.shard import frame_range, gather_rows
    lo, hi = frame_range(nframes, rank, world)
    rows = torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.ones(1, 16, dtype=torch.float64)
    out = gather_rows(rows, nframes, dst=0)
    if rank == 0:
        q.put(out.numpy())
    dist.destroy_process_group()


def test_frame_range_partitions():
    from slv_b200.shard import frame_range
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            blocks = [frame_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_gather_rows_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    nframes = 11        # ragged: 6 + 5
    procs = [ctx.Process(target=_worker, args=(r, 2, 29517, nframes, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out.shape == (11, 16) and np.array_equal(out[:, 0], np.arange(11.0))
