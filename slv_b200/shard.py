"""Frame sharding across ranks (one process per GPU).

Trajectory frames are independent -- the reference farms one frame per Parallel-Python job
(util/slv_md-run:155-169) -- so rank r evaluates the contiguous block [lo, hi) of frames with no
data-path communication, and ONE gather of the per-frame result rows ends the run.
"""


def frame_range(nframes, rank, world):
    """Contiguous, balanced block of frames of this rank: sizes differ by at most one."""
    base, rem = divmod(int(nframes), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(rows, nframes, dst=0, group=None):
    """Gather per-rank row blocks [n_r, ncol] (torch tensors on this rank's device) to `dst`.
    Returns the [nframes, ncol] tensor on dst, None elsewhere.  Blocks may be ragged by one row."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group); rank = dist.get_rank(group)
    sizes = [frame_range(nframes, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((nmax, rows.shape[1]), dtype=rows.dtype, device=rows.device)
    pad[:rows.shape[0]] = rows
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)], 0)
