"""SURVEY 8e row 4 on N GPUs of one box (torchrun is only the launcher): the deterministic part of
the wavelet coherence at config-4 size (two N = 2^18 series, 145 scales, Morlet) with the scales
block-partitioned over the ranks.  Every rank computes its block plus the boxcar halo (7 rows
below, 6 above at dj = 1/12), keeps the interior, and the per-scale mean coherence is all-gathered
through the engine's C ABI (NCCL).  Each rank also runs the whole ladder alone and checks that its
slab is bit-identical to those rows.
Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29535 profiles/config4_wct_sharded.py
"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads as wl                                  # noqa: E402
import pycwt_b200 as pycwt                              # noqa: E402
from pycwt_b200 import _engine, distributed as D        # noqa: E402

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
eng = _engine.Engine(local)
comm = D.NcclComm(eng, rank, world)
c = wl.C4
y1, y2 = wl.config4_signals()
args = dict(dj=c["dj"], s0=c["s0"], J=c["J"], wavelet=pycwt.Morlet(c["f0"]), engine=eng)
for _ in range(3):     # warm-up: buffers, pinned result arrays, the communicator's first collective
    out = D.wct_scale_sharded(y1, y2, c["dt"], comm=comm, **args)
comm.max(0.0)          # barrier
t0 = time.perf_counter()
lo, hi, WCT, aWCT, mean, freq = D.wct_scale_sharded(y1, y2, c["dt"], comm=comm, **args)
dt = comm.max(time.perf_counter() - t0)
k_ms = comm.max(eng.last_kernel_ms())
# the same ladder on this GPU alone
for _ in range(3):     # (the full-ladder result arrays are larger: page-locked once, then pooled)
    flo, fhi, Wf, Af, meanf, _ = D.wct_scale_sharded(y1, y2, c["dt"], comm=None, **args)
t0 = time.perf_counter()
D.wct_scale_sharded(y1, y2, c["dt"], comm=None, **args)
dt1 = time.perf_counter() - t0
k1 = eng.last_kernel_ms()
same = bool(np.array_equal(WCT, Wf[lo:hi]) and np.array_equal(aWCT, Af[lo:hi]) and np.array_equal(mean, meanf))
bad = comm.max(0.0 if same else 1.0)
if rank == 0:
    print("config4 wct(sig=False), scales over %d GPUs (block + halo): %.1f ms per call, kernels %.2f ms "
          "(max over ranks); one GPU alone: %.1f ms, kernels %.2f ms; slabs bit-identical to the single-GPU rows "
          "on every rank: %s; mean coherence sha1 %s"
          % (world, 1e3 * dt, k_ms, 1e3 * dt1, k1, bad == 0.0, hashlib.sha1(mean.tobytes()).hexdigest()[:12]))
comm.close()
