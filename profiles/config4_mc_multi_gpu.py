"""SURVEY 8e row 3 on N GPUs of one box (torchrun is only the launcher): the Monte-Carlo surrogate
pairs of the coherence significance test (config 4: s0=2, dj=1/12, J=144, 200 pairs of length
49152) are block-partitioned over the ranks; each rank accumulates its [145, 1000] int64
histograms on its GPU and ONE NCCL all-reduce (1.16 MB, through the engine's C ABI -- no torch
collective) sums them.  The result is independent of the world size (pair i always comes from
the stream keyed by (seed, i)); rank 0 prints a checksum to compare runs.
Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29534 profiles/config4_mc_multi_gpu.py [mc_count]
"""
import hashlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycwt_b200 import _engine, distributed as D

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
mc = int(sys.argv[1]) if len(sys.argv) > 1 else 200
eng = _engine.Engine(local)
comm = D.NcclComm(eng, rank, world)          # NCCL id through the rendezvous file
args = (0.3, 0.25, 1.0, 1 / 12, 2.0, 144)
for device_rng in (False, True):
    D.wct_significance_sharded(*args, mc_count=2 * world, seed=1, engine=eng, comm=comm, device_rng=device_rng)
    comm.max(0.0)                            # barrier
    t0 = time.perf_counter()
    sig = D.wct_significance_sharded(*args, significance_level=0.95, wavelet='morlet', mc_count=mc, seed=7,
                                     engine=eng, comm=comm, device_rng=device_rng)
    dt = comm.max(time.perf_counter() - t0)
    if rank == 0:
        ok = sig[np.isfinite(sig)]
        print("config4 Monte-Carlo significance, %d GPUs, %s surrogates: %d pairs (N=49152, 145 scales) in %.3f s "
              "(%.2f ms per pair); sig95 range %.3f..%.3f, sha1 %s"
              % (world, "device-RNG (Philox)" if device_rng else "host-RNG", mc, dt, 1e3 * dt / mc, ok.min(), ok.max(),
                 hashlib.sha1(sig.tobytes()).hexdigest()[:12]))
comm.close()
