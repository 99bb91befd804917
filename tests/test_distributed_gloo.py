"""World-size-2 gloo test (CPU) of the multi-GPU plumbing: channel sharding, gather of the
reduced per-channel spectra, max-over-ranks timing.  The per-rank GPU transform is replaced
by a stand-in so the test needs no device; the GPU path of the same code is covered by
bench.py --gpus N and tests/test_gpu_xwt_wct.py::test_cwt_batch_channels."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT
from pycwt_b200.distributed import shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 8192, 8193):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


class _FakeEngine(object):
    """Stands in for pycwt_b200.Engine on a box without a GPU: deterministic 'spectra'."""

    def cwt_batch(self, X, dt, scales, family, param, precision, want_power=True, want_w=False):
        return (X[:, :1] ** 2) * np.asarray(scales)[None, :] + X.sum(axis=1, keepdims=True), None


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from pycwt_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(0)
        X = rs.randn(7, 32)           # 7 channels over 2 ranks: 4 + 3
        sj = np.arange(1.0, 6.0)
        full = D.cwt_batch_sharded(X, 1.0, sj, 0, 6.0, 0, _FakeEngine(), dist)
        ref, _ = _FakeEngine().cwt_batch(X, 1.0, sj, 0, 6.0, 0)
        ok = full.shape == (7, 5) and np.array_equal(full, ref)
        t = D.max_over_ranks(1.0 + rank, dist)
        q.put((rank, bool(ok), t))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_gloo():
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, 2.0), (1, True, 2.0)]


def _mc_worker(rank, world, port, q):
    """Each rank accumulates coherence histograms of its surrogate pairs with the host-emulation
    build of the kernels (test infrastructure, no GPU here); one gloo all-reduce sums them."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from pycwt_b200 import distributed as D, _engine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = _engine.Engine(0, lib_path=os.path.join(ROOT, "tests", "_emu", "libcwtb200_emu.so"))
        sig = D.wct_significance_sharded(0.2, 0.1, 1.0, 0.5, 2.0, 8, 0.95, 'morlet', mc_count=5,
                                         seed=42, engine=eng, comm=D.TorchComm(dist))
        # one signal, scales sharded over the ranks; the global spectrum is gathered
        x = np.random.RandomState(5).randn(3000)
        sj = 2.0 * 2 ** (np.arange(13) / 2.0)
        rows, power, W = D.cwt_scale_sharded(x, 1.0, sj, 0, 6.0, 0, eng, D.TorchComm(dist), fetch=True)
        rb, pb, _ = D.cwt_scale_sharded(x, 1.0, sj, 0, 6.0, 0, eng, D.TorchComm(dist), layout='block')
        q.put((rank, sig.tolist(), rows.tolist(), power.tolist(), np.abs(W).sum(), rb.tolist(), pb.tolist()))
        eng.close()
    finally:
        dist.destroy_process_group()


def test_sharded_wct_significance_gloo():
    """World size 2 gives bit-identical significance levels to one process running every pair."""
    pytest.importorskip("torch")
    import torch.multiprocessing as mp
    from pycwt_b200 import build as _build, _engine, distributed as D
    lib = _build.build_emulation(os.path.join(ROOT, "tests", "_emu"))
    eng = _engine.Engine(0, lib_path=lib)
    single = D.wct_significance_sharded(0.2, 0.1, 1.0, 0.5, 2.0, 8, 0.95, 'morlet', mc_count=5,
                                        seed=42, engine=eng)
    eng.close()
    assert np.isnan(single).any() and np.isfinite(single).any()
    fin = single[np.isfinite(single)]
    assert ((fin > 0) & (fin < 1)).all()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_mc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = {g[0]: g for g in got}
    for r in (0, 1):
        assert np.array_equal(np.array(res[r][1]), single, equal_nan=True)
    # scale-sharded single signal: cyclic rows 0,2,4.. / 1,3,5.. (block: [0,7) / [7,13)), identical
    # gathered spectrum that matches the oracle
    from oracle import cwt_oracle as orc
    x = np.random.RandomState(5).randn(3000)
    sj = 2.0 * 2 ** (np.arange(13) / 2.0)
    Wr = orc.cwt(x, 1.0, wavelet=orc.Morlet(6), freqs=1 / (orc.Morlet(6).flambda() * sj))[0]
    ref = (np.abs(Wr) ** 2).mean(axis=1)
    assert res[0][2] == list(range(0, 13, 2)) and res[1][2] == list(range(1, 13, 2))
    assert res[0][5] == list(range(0, 7)) and res[1][5] == list(range(7, 13))
    for r in (0, 1):
        assert np.allclose(res[r][3], ref, rtol=1e-12) and np.allclose(res[r][6], ref, rtol=1e-12)
    assert abs(res[0][4] + res[1][4] - np.abs(Wr).sum()) < 1e-9 * np.abs(Wr).sum()


class _FakeCommEngine(object):
    """Engine stand-in recording the communicator calls of NcclComm (no NCCL on this box)."""

    def __init__(self):
        self.inited = None

    def comm_unique_id(self):
        return bytes(range(128))

    def comm_init(self, world, rank, uid):
        self.inited = (world, rank, bytes(uid))

    def comm_destroy(self):
        self.inited = None


def test_nccl_comm_id_exchange_through_rendezvous_file(tmp_path, monkeypatch):
    """NcclComm hands rank 0's 128-byte id to the other ranks through a rendezvous file
    (host-side logic only; the collectives themselves run on the GPU box)."""
    from pycwt_b200 import distributed as D
    monkeypatch.setenv("CWTB_COMM_DIR", str(tmp_path))
    monkeypatch.setenv("MASTER_PORT", "29517")
    e0, e1 = _FakeCommEngine(), _FakeCommEngine()
    c0 = D.NcclComm(e0, rank=0, world=2)
    c1 = D.NcclComm(e1, rank=1, world=2)
    assert e0.inited == (2, 0, bytes(range(128))) and e1.inited == (2, 1, bytes(range(128)))
    assert (c0.rank, c0.world, c1.rank) == (0, 2, 1)
    c0.close(), c1.close()
    assert e0.inited is None
    # a single process needs neither NCCL nor a file
    solo = D.NcclComm(_FakeCommEngine(), rank=0, world=1)
    assert solo.max(3.5) == 3.5
    assert np.array_equal(solo.allgather_rows(np.arange(6.0).reshape(3, 2), 3), np.arange(6.0).reshape(3, 2))
    assert np.array_equal(solo.allreduce_sum(np.arange(4)), np.arange(4))


def test_wct_halo_rows():
    """Rows of the scale boxcar's footprint: K = 14 (dj = 1/12) reaches 7 rows down and 6 up, like
    scipy's convolve2d(..., 'same') of the reference (mothers.py:96-102)."""
    from scipy.signal import convolve2d
    from pycwt_b200 import distributed as D
    for K in (1, 2, 5, 14, 15):
        T = np.zeros((40, 1))
        T[20] = 1.0
        out = convolve2d(T, np.ones((K, 1)), 'same')[:, 0]
        hit = np.flatnonzero(out)
        below, above = D.wct_halo(K)
        # an impulse at row 20 reaches outputs 20 - above .. 20 + below
        assert (hit.min(), hit.max()) == (20 - above, 20 + below), K


def _wct_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from pycwt_b200 import distributed as D, _engine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = _engine.Engine(0, lib_path=os.path.join(ROOT, "tests", "_emu", "libcwtb200_emu.so"))
        rs = np.random.RandomState(3)
        t = np.arange(700)
        y1 = np.sin(2 * np.pi * t / 23.0) + 0.5 * rs.randn(700)
        y2 = np.sin(2 * np.pi * t / 23.0 + 0.6) + 0.5 * rs.randn(700)
        lo, hi, WCT, aWCT, mean, freq = D.wct_scale_sharded(y1, y2, 1.0, dj=0.25, s0=2.0, J=27, engine=eng,
                                                            comm=D.TorchComm(dist))
        q.put((rank, lo, hi, WCT, aWCT, mean))
        eng.close()
    finally:
        dist.destroy_process_group()


def test_wct_scale_sharded_gloo():
    """World size 3: every rank's coherence slab (block + halo computed, interior kept) equals the
    rows of the single-process result bit for bit, and the gathered mean coherence is complete."""
    pytest.importorskip("torch")
    import torch.multiprocessing as mp
    from pycwt_b200 import build as _build, _engine, distributed as D
    lib = _build.build_emulation(os.path.join(ROOT, "tests", "_emu"))
    eng = _engine.Engine(0, lib_path=lib)
    rs = np.random.RandomState(3)
    t = np.arange(700)
    y1 = np.sin(2 * np.pi * t / 23.0) + 0.5 * rs.randn(700)
    y2 = np.sin(2 * np.pi * t / 23.0 + 0.6) + 0.5 * rs.randn(700)
    lo, hi, W1, A1, mean1, _ = D.wct_scale_sharded(y1, y2, 1.0, dj=0.25, s0=2.0, J=27, engine=eng)
    eng.close()
    assert (lo, hi) == (0, 28) and W1.shape == (28, 700)
    # the single-process result is the oracle's coherence
    from oracle import cwt_oracle as orc
    Wr, Ar = orc.wct(y1, y2, 1.0, dj=0.25, s0=2.0, J=27, sig=False, wavelet=orc.Morlet(6))[:2]
    assert np.abs(W1 - Wr).max() < 1e-10 and np.abs(np.angle(np.exp(1j * (A1 - Ar)))).max() < 1e-8
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_wct_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {g[0]: g for g in (q.get(timeout=300) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for r in range(world):
        _, lo, hi, W, A, mean = got[r]
        assert (lo, hi) == D.shard_range(28, r, world)
        assert np.array_equal(W, W1[lo:hi]) and np.array_equal(A, A1[lo:hi])
        assert np.allclose(mean, mean1, rtol=0, atol=0)
        covered += list(range(lo, hi))
    assert covered == list(range(28))
