"""SURVEY 8d config 5 on N GPUs of one box (torchrun): independent float32 channels are
block-partitioned over the ranks (no data-path collective), each rank transforms its block with
the batched engine path and only the reduced [channels, scales] spectra are all-gathered over
NCCL.  Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29533 profiles/config5_multi_gpu.py [channels_total]
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from pycwt_b200 import _engine, distributed as D

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nch_total = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n0, S = 2 ** 16, 128
sj = 2.0 * 2 ** (np.arange(S) / 8.0)
lo, hi = D.shard_range(nch_total, rank, world)
X = np.random.RandomState(1000 + rank).randn(hi - lo, n0).astype(np.float32)   # this rank's channels
eng = _engine.Engine(local)
dev = torch.device("cuda", local)
power, _ = eng.cwt_batch(X[:64], 1.0, sj, 0, 6.0, precision=1)                 # warm-up (same chunk shape: buffers allocated)
if world > 1:
    dist.barrier(device_ids=[local])
torch.cuda.synchronize()
t0 = time.perf_counter()
power, _ = eng.cwt_batch(X, 1.0, sj, 0, 6.0, precision=1, want_power=True)     # H2D + kernels + reduce
full = D.gather_rows(power, nch_total, dist if world > 1 else None, dev)
torch.cuda.synchronize()
dt = D.max_over_ranks(time.perf_counter() - t0, dist if world > 1 else None, dev)
assert full.shape == (nch_total, S)
if rank == 0:
    from oracle import cwt_oracle as orc
    Wr = orc.cwt(X[0].astype(np.float64), 1.0, wavelet=orc.Morlet(6), freqs=1 / (orc.Morlet(6).flambda() * sj))[0]
    err = np.abs(full[0] - (np.abs(Wr) ** 2).mean(axis=1)).max() / (np.abs(Wr) ** 2).mean(axis=1).max()
    pts = nch_total * S * n0
    print("config5 %d GPUs: %d channels x %d scales x 2^16 fp32 (host float32 in, [ch,scales] spectra "
          "gathered over NCCL): %.3f s -> %.3e scale-points/s; channel 0 spectrum vs oracle rel.err %.1e"
          % (world, nch_total, S, dt, pts / dt, err))
if world > 1:
    dist.barrier(device_ids=[local]); dist.destroy_process_group()
