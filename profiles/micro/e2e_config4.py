"""Where does the wall time of pycwt_b200.xwt / wct at config-4 size go (host in, host out)?"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import workloads as wl      # noqa: E402
import pycwt_b200 as pycwt  # noqa: E402

c = wl.C4
y1, y2 = wl.config4_signals()
m = pycwt.Morlet(c["f0"])
eng = pycwt.default_engine()
for name, fn in (("xwt", lambda: pycwt.xwt(y1, y2, c["dt"], c["dj"], c["s0"], c["J"], wavelet=m)),
                 ("wct", lambda: pycwt.wct(y1, y2, c["dt"], c["dj"], c["s0"], c["J"], sig=False, wavelet=m))):
    for _ in range(3):
        r = fn()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        r = fn()
        ts.append(1e3 * (time.perf_counter() - t0))
    print("== %s: %s ms per call; kernels %.2f ms" % (name, ", ".join("%.1f" % t for t in ts), eng.last_kernel_ms()))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        r = fn()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
