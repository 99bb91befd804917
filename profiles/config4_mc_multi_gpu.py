"""SURVEY 8e row 3 on N GPUs of one box (torchrun): the Monte-Carlo surrogate pairs of the
coherence significance test (config 4: s0=2, dj=1/12, J=144, 200 pairs of length 49152) are
block-partitioned over the ranks; each rank accumulates its [145, 1000] int64 histograms on its
GPU and ONE NCCL all-reduce (1.16 MB) sums them.  The result is independent of the world size
(pair i always comes from RNG stream (seed, i)); rank 0 prints a checksum to compare runs.
Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29534 profiles/config4_mc_multi_gpu.py [mc_count]
"""
import hashlib, os, sys, time
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
mc = int(sys.argv[1]) if len(sys.argv) > 1 else 200
eng = _engine.Engine(local)
dev = torch.device("cuda", local)
args = (0.3, 0.25, 1.0, 1 / 12, 2.0, 144)
d = dist if world > 1 else None
D.wct_significance_sharded(*args, mc_count=2 * world, seed=1, engine=eng, comm=D.TorchComm(d, dev))   # warm-up
if world > 1:
    dist.barrier(device_ids=[local])
torch.cuda.synchronize()
t0 = time.perf_counter()
sig = D.wct_significance_sharded(*args, significance_level=0.95, wavelet='morlet', mc_count=mc, seed=7,
                                 engine=eng, comm=D.TorchComm(d, dev))
torch.cuda.synchronize()
dt = D.max_over_ranks(time.perf_counter() - t0, d, dev)
if rank == 0:
    ok = sig[np.isfinite(sig)]
    print("config4 Monte-Carlo significance, %d GPUs: %d surrogate pairs (N=49152, 145 scales) in %.3f s "
          "(%.2f ms per pair incl. host RNG); sig95 range %.3f..%.3f, sha1 %s"
          % (world, mc, dt, 1e3 * dt / mc, ok.min(), ok.max(), hashlib.sha1(sig.tobytes()).hexdigest()[:12]))
if world > 1:
    dist.barrier(device_ids=[local]); dist.destroy_process_group()
