"""Where does the time of one pycwt_b200.cwt() call at the north-star size go?"""
import os, sys, time, ctypes, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pycwt_b200 as pycwt
eng = pycwt.default_engine()
N = 2 ** 20
t = np.arange(N) / N
x = np.sin(2 * np.pi * (50 * t + (N / 8) * t ** 2))
sj = 2.0 * 2 ** (np.arange(256) / 16.0)
for _ in range(3):
    W, *_ = pycwt.cwt(x, 1.0, 1 / 16, 2.0, 255, pycwt.Morlet(6))
del W, _   # a held result keeps its 4.3 GB pinned buffer out of the pool (two circulate below)
def tm(f, n=5):
    f(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    return (time.perf_counter() - t0) / n * 1e3
print("full pycwt.cwt            %.1f ms" % tm(lambda: pycwt.cwt(x, 1.0, 1 / 16, 2.0, 255, pycwt.Morlet(6))))
print("engine.cwt fetch=False    %.1f ms" % tm(lambda: eng.cwt(x, 1.0, sj, 0, 6.0, fetch=False)))
print("engine.get_w (4.29 GB)    %.1f ms" % tm(lambda: eng.get_w(256, N)))
print("signal_fft                %.1f ms" % tm(lambda: eng.signal_fft()))
P = ctypes.c_void_p
for gb in (0.5, 1, 2, 4):
    n = int(gb * (1 << 30))
    arr = eng.result_array((n // 16,), np.complex128)
    d = eng.dev_alloc(n)
    ms = tm(lambda: eng.lib.cwtb_memcpy_d2h(eng.h, arr.ctypes.data_as(P), d, n), 3)
    print("raw D2H %.1f GiB: %.1f ms = %.1f GB/s" % (gb, ms, n / ms / 1e6))
    eng.dev_free(d); del arr
