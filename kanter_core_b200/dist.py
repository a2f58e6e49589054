"""Multi-GPU plumbing: one process per GPU over torch.distributed.

The per-pixel path shards two ways, neither needs a collective on
the data path:
  * independent graphs: round-robin whole graphs over ranks (shard_units);
  * one big image: horizontal strips (strip_rows); elementwise nodes need nothing,
    HeightToNormal needs ONE halo row from the strip above (ring_halo_rows: a
    point-to-point send/recv of w*4 bytes per strip boundary over NVLink).
torch.distributed is used for the rendezvous, the barrier, the max-over-ranks of
the timing and that one send/recv.  Works with the gloo backend on CPU tensors
(tests) and nccl on CUDA tensors (the GPU box).
"""
import os


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_units(n_units, rank, world):
    """Indices of the independent units (graphs) this rank evaluates: round-robin."""
    return list(range(rank, n_units, world))


def strip_rows(height, rank, world):
    """Rows [y0, y1) of rank's horizontal strip; strips differ by at most one row."""
    base, extra = divmod(height, world)
    y0 = rank * base + min(rank, extra)
    return y0, y0 + base + (1 if rank < extra else 0)


def max_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def ring_halo_rows(last_row):
    """Every rank passes the last row of its strip to the next rank (the last rank's
    row wraps around to rank 0: the reference samples row -1 as row H-1) and returns
    the row it received, i.e. the row directly above its own strip."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return last_row.clone()
    rank, world = dist.get_rank(), dist.get_world_size()
    recv = torch.empty_like(last_row)
    ops = [dist.P2POp(dist.isend, last_row, (rank + 1) % world), dist.P2POp(dist.irecv, recv, (rank - 1) % world)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return recv
