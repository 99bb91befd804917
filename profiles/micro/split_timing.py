"""Where does the config-2 step go?  Device-timed step of the full scale ladder and of its parts
(dense rows, pruned two-kernel rows, expansion rows) run as separate transforms: if the parts add
up to the whole the step is throughput-bound and stream overlap has nothing left to give.

    python profiles/micro/split_timing.py [lib.so ...]     (default: the in-tree build)

Every library given is timed in turn (A/B of compile-time variants built by
profiles/micro/build_variant.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import workloads as wl          # noqa: E402
from pycwt_b200 import _engine  # noqa: E402
from bench import pin_to_gpu_numa_node  # noqa: E402


def timed(eng, dsig, n, dt, sj, f0, iters=20):
    eng.cwt_dev(dsig, 0, n, dt, sj, _engine.MORLET, f0, _engine.F64)
    eng.bench_last(3)
    return eng.bench_last(iters)


def main():
    pin_to_gpu_numa_node(0)   # like bench.py: launch latency depends on the CPU node
    libs = sys.argv[1:] or [None]
    c = wl.C2
    sj = wl.config2_scales()
    x = wl.config2_signal()
    parts = [("full 0..255", slice(0, 256)), ("dense 0..23", slice(0, 24)), ("band 24..71", slice(24, 72)),
             ("exact 0..71", slice(0, 72)), ("expand 72..255", slice(72, 256)),
             ("expand w12 117..255", slice(117, 256))]
    if os.environ.get("SPLIT_PARTS"):   # e.g. SPLIT_PARTS=full,exact
        keep = os.environ["SPLIT_PARTS"].split(",")
        parts = [p for p in parts if p[0].split()[0] in keep]
    for lib in libs:
        eng = _engine.Engine(0, lib_path=lib)
        dsig = eng.dev_alloc(x.nbytes)
        eng.h2d(dsig, x)
        print("== %s" % (lib or "in-tree build"))
        for name, sl in parts:
            ms = timed(eng, dsig, c["n"], c["dt"], sj[sl], c["f0"])
            rows = len(sj[sl])
            print("  %-22s %7.4f ms  %6.2f us/row  launches %d" % (name, ms, 1e3 * ms / rows, eng.last_launch_count()))
        eng.dev_free(dsig)
        eng.close()


if __name__ == "__main__":
    main()
