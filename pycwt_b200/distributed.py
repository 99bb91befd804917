"""Multi-GPU plumbing for the batched transform: one process per GPU, channels
block-partitioned across ranks, no collective on the data path.  torch.distributed (NCCL on
GPUs, gloo in CPU tests) is used only to gather the *reduced* per-channel products (global
wavelet spectra [channels, scales]) and to agree on timings; coefficients stay sharded in
each GPU's HBM."""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of `n_items` owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local, n_total, dist=None, device=None):
    """All ranks contribute their block of rows (as produced by shard_range); every rank
    gets the full [n_total, ...] array back.  `dist` is torch.distributed (initialised)."""
    local = np.ascontiguousarray(local)
    if dist is None or dist.get_world_size() == 1:
        return local
    import torch
    world = dist.get_world_size()
    trail = local.shape[1:]
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    maxrows = max(hi - lo for lo, hi in sizes)
    buf = np.zeros((maxrows,) + trail, dtype=local.dtype)
    buf[:local.shape[0]] = local
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    parts = [o.cpu().numpy()[:hi - lo] for o, (lo, hi) in zip(outs, sizes)]
    return np.concatenate(parts, axis=0)


def max_over_ranks(value, dist=None, device=None):
    """Maximum of a python float over all ranks (timings are max-over-ranks)."""
    if dist is None or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def cwt_batch_sharded(X, dt, scales, family, param, precision, engine, dist=None, device=None):
    """Global wavelet spectra of every channel of X[channels, n0], computed by the rank that
    owns the channel and gathered on all ranks.  X may be the full array (each rank slices
    its block) -- only the [channels, scales] result crosses the interconnect."""
    rank = 0 if dist is None else dist.get_rank()
    world = 1 if dist is None else dist.get_world_size()
    lo, hi = shard_range(X.shape[0], rank, world)
    power, _ = engine.cwt_batch(X[lo:hi], dt, scales, family, param, precision, want_power=True)
    return gather_rows(power, X.shape[0], dist, device)
