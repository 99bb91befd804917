"""Device-timed step of every BASELINE.json configuration (kernels only, inputs resident), for
sweeps of engine switches through environment variables:

    CWTB_GEN_BAND=0 python profiles/micro/config_kernels.py [2,3,4,5] [--prof]

--prof adds the serialised per-kernel table of each."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import workloads as wl          # noqa: E402
from pycwt_b200 import _engine  # noqa: E402
from bench import pin_to_gpu_numa_node  # noqa: E402


def show(eng, name, ms, prof):
    print("%-28s %8.4f ms  launches %d" % (name, ms, eng.last_launch_count()))
    if prof:
        for k in sorted(eng.profile_last(), key=lambda k: -k["ms"]):
            print("      %-46s %3d x  %7.4f ms  rows %d" % (k["name"], k["launches"], k["ms"], k["rows"]))


def main():
    pin_to_gpu_numa_node(0)   # like bench.py: launch latency depends on the CPU node
    which = [a for a in sys.argv[1:] if not a.startswith("--")]
    which = which[0].split(",") if which else ["2", "3", "4", "5"]
    prof = "--prof" in sys.argv
    eng = _engine.Engine(0)
    if "2" in which:
        c = wl.C2
        x = wl.config2_signal()
        d = eng.dev_alloc(x.nbytes)
        eng.h2d(d, x)
        eng.cwt_dev(d, 0, c["n"], c["dt"], wl.config2_scales(), _engine.MORLET, c["f0"], _engine.F64)
        eng.bench_last(3)
        show(eng, "config 2", eng.bench_last(20), prof)
        eng.dev_free(d)
    if "3" in which:
        c = wl.C3
        x = wl.config3_signal()
        d = eng.dev_alloc(x.nbytes)
        eng.h2d(d, x)
        for fam, code in (("paul", _engine.PAUL), ("dog", _engine.DOG)):
            p = c[fam]
            sj = wl.geometric_scales(p["s0"], p["dj"], p["J"])
            eng.cwt_dev(d, 1, c["n"], c["dt"], sj, code, float(p["m"]), _engine.F32)
            eng.bench_last(3)
            show(eng, "config 3 " + fam, eng.bench_last(40), prof)
        eng.dev_free(d)
    if "4" in which:
        import pycwt_b200 as pycwt
        c = wl.C4
        y1, y2 = wl.config4_signals()
        m = pycwt.Morlet(c["f0"])
        deng = pycwt.default_engine()
        for name, fn in (("config 4 xwt", lambda: pycwt.xwt(y1, y2, c["dt"], c["dj"], c["s0"], c["J"], wavelet=m)),
                         ("config 4 wct", lambda: pycwt.wct(y1, y2, c["dt"], c["dj"], c["s0"], c["J"], sig=False,
                                                            wavelet=m))):
            ks = []
            for _ in range(4):
                fn()
                ks.append(deng.last_kernel_ms())
            print("%-28s %8.4f ms  (kernels of one call, best of 4; launches %d)"
                  % (name, min(ks), deng.last_launch_count()))
            if prof:
                deng.profile_begin()
                fn()
                for k in sorted(deng.profile_end(), key=lambda k: -k["ms"])[:16]:
                    print("      %-46s %3d x  %7.4f ms  rows %d" % (k["name"], k["launches"], k["ms"], k["rows"]))
        deng.trim()
    if "5" in which:
        c = wl.C5
        sj = wl.geometric_scales(c["s0"], c["dj"], c["J"])
        chunk = 256
        X = wl.config5_channels(0, chunk)
        d = eng.dev_alloc(chunk * c["n"] * 4)
        eng.h2d(d, X)
        eng.cwt_batch_dev(d, chunk, c["n"], c["dt"], sj, _engine.MORLET, c["f0"], _engine.F32)
        eng.bench_last(2)
        show(eng, "config 5 (256 channels)", eng.bench_last(5), prof)
        eng.dev_free(d)
    eng.close()


if __name__ == "__main__":
    main()
