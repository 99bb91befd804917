"""Cost of the un-padded mode (Bluestein) next to the padded mode, for the record."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pycwt_b200 as pycwt
eng = pycwt.default_engine()
for n in (100000, 1000000):
    t = np.arange(n) / n
    x = np.sin(2 * np.pi * (50 * t + (n / 8) * t ** 2))
    sj = 2.0 * 2 ** (np.arange(64) / 4.0)
    for pad in (True, False):
        eng.set_padding(pad)
        for _ in range(3):
            eng.cwt(x, 1.0, sj, 0, 6.0, fetch=False); eng.sync()
        t0 = time.perf_counter()
        for _ in range(5):
            eng.cwt(x, 1.0, sj, 0, 6.0, fetch=False)
        eng.sync()
        dt_ = (time.perf_counter() - t0) / 5
        print("n0=%d, 64 scales, %s: %.2f ms per transform (host input, W resident) -> %.2e scale-points/s"
              % (n, "padded to 2^k" if pad else "un-padded (Bluestein)", 1e3 * dt_, 64 * n / dt_))
eng.set_padding(True)
