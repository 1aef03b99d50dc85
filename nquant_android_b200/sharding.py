"""Image-level sharding of a batch over ranks (one process per GPU). The path has no data-path
collective: each rank converts its own images; ranks only meet at a barrier and a MAX-reduce of the
elapsed time."""


def shard_range(n_images, rank, world):
    """Contiguous, balanced [begin, end) of the images rank `rank` owns."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_images, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def max_over_ranks(value, dist=None, device=None):
    """MAX all-reduce of a python float across the process group (identity without one)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
