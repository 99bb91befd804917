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


def sum_over_ranks(array, dist=None, device=None):
    """Element-wise sum of an integer/float array over all ranks (every rank gets the total)."""
    array = np.ascontiguousarray(array)
    if dist is None or dist.get_world_size() == 1:
        return array
    import torch
    t = torch.from_numpy(array.copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def surrogate_pair(seed, index, N, al1, al2):
    """Surrogate pair number `index` of a sharded Monte-Carlo run: white noise like the
    reference's `rednoise` output (helpers.py:146-173 filters a length-1 axis, SURVEY 8a row
    10), from an RNG stream keyed by (seed, index) so that the pair does not depend on which
    rank draws it."""
    rs = np.random.RandomState([int(seed) & 0x7fffffff, int(index)])

    def white(al):
        tau = 0 if al == 0 else int(np.ceil(-2 / np.log(np.abs(al))))
        return rs.randn(N + tau)[tau:]
    return white(al1), white(al2)


def wct_significance_sharded(al1, al2, dt, dj, s0, J, significance_level=0.95, wavelet='morlet',
                             mc_count=300, seed=0, engine=None, dist=None, device=None):
    """Monte-Carlo coherence significance (reference wavelet.py:531-647) with the surrogate
    pairs block-partitioned over the ranks: every rank accumulates the [S, 1000]
    int64 histograms of its pairs on its GPU, ONE all-reduce (sum, ~1 MB) combines them and
    every rank evaluates the percentiles.  The result is independent of the world size."""
    from . import wavelet as wv
    mother = wv._check_parameter_wavelet(wavelet)
    rank = 0 if dist is None else dist.get_rank()
    world = 1 if dist is None else dist.get_world_size()
    prob = wv._mc_problem(dt, dj, s0, J, mother)
    lo, hi = shard_range(mc_count, rank, world)
    hist = wv._mc_histogram(prob, dt, dj, mother,
                            lambda i: surrogate_pair(seed, i, prob['N'], al1, al2),
                            range(lo, hi), progress=False, engine=engine)
    hist = sum_over_ranks(hist, dist, device)
    return wv._mc_levels(prob, hist, significance_level)


def cwt_scale_sharded(signal, dt, scales, family, param, precision, engine, dist=None, device=None,
                      fetch=False):
    """One long signal, scales block-partitioned over the ranks.  Every rank
    holds the signal and runs its own forward transform (0.04 ms at N = 2^20 -- cheaper than
    broadcasting the 16 MiB spectrum), then transforms only its slab of scales, which stays
    resident in that GPU's HBM.  No collective on the data path; the per-scale global power
    (mean_n |W|^2, [S]) is all-gathered so that every rank sees the whole spectrum.

    Returns (lo, hi, global_power[S], W_slab or None): rows [lo, hi) are this rank's scales."""
    rank = 0 if dist is None else dist.get_rank()
    world = 1 if dist is None else dist.get_world_size()
    scales = np.ascontiguousarray(scales, dtype=np.float64)
    lo, hi = shard_range(scales.size, rank, world)
    W = None
    if hi > lo:
        W = engine.cwt(signal, dt, scales[lo:hi], family, param, precision, fetch=fetch)
        local = engine.global_power(hi - lo)
    else:
        local = np.zeros(0)
    power = gather_rows(local.reshape(-1, 1), scales.size, dist, device).ravel()
    return lo, hi, power, W
