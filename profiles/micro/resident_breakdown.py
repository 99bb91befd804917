"""Where does the time of the device-resident path at the north-star size go?"""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pycwt_b200 as pycwt
N = 2 ** 20
t = np.arange(N) / N
x = np.sin(2 * np.pi * (50 * t + (N / 8) * t ** 2))
def tm(f, n=5):
    f(); f(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    return (time.perf_counter() - t0) / n * 1e3
mk = lambda: pycwt.cwt_resident(x, 1.0, 1 / 16, 2.0, 255, pycwt.Morlet(6))
print("cwt_resident             %.2f ms" % tm(mk))
r = mk()
print("global_power             %.2f ms" % tm(r.global_power))
print("global_power inside coi  %.2f ms" % tm(lambda: r.global_power(inside_coi=True)))
print("scale_avg_power          %.2f ms" % tm(lambda: r.scale_avg_power(16.0, 64.0)))
print("icwt                     %.2f ms" % tm(r.icwt))
print("coi (host, first use)    %.2f ms" % tm(lambda: setattr(r, '_coi', None) or r.coi))
