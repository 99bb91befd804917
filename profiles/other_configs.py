"""Timings of SURVEY 8d configs 3-5 (single GPU slices), for the record; not bench lines."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pycwt_b200 as pycwt
eng = pycwt.default_engine()


def chirp(n, ph=0.0):
    t = np.arange(n) / n
    return np.sin(2 * np.pi * (50 * t + (n / 8) * t ** 2) + ph)


def kernel_time(x, sj, fam, par, prec, is32):
    d = eng.dev_alloc(x.nbytes)
    eng.h2d(d, x)
    eng.cwt_dev(d, is32, x.size, 1.0, sj, fam, par, prec)
    ms = eng.bench_last(20)
    eng.dev_free(d)
    return ms

# config 3: Paul(4) / DOG(2), N = 2^18, 128 scales, fp32
n = 2 ** 18
x32 = chirp(n).astype(np.float32)
for name, fam, par, s0, dj in (("paul4", 1, 4.0, 1.4324, 1 / 18), ("dog2", 2, 2.0, 0.5033, 1 / 8)):
    sj = s0 * 2 ** (np.arange(128) * dj)
    ms = kernel_time(x32, sj, fam, par, 1, 1)
    print("config3 %s fp32 N=2^18 S=128: %.3f ms kernels -> %.3e scale-points/s, %.0f GB/s algorithmic"
          % (name, ms, 128 * n / ms * 1e3, 128 * n * 8 / ms / 1e6))
# config 3 end to end through the reference API (float32 host input, complex128 host result)
os.environ["CWTB_PRECISION"] = "fp32"
for name, mother, kw in (("paul4", pycwt.Paul(4), dict(s0=1.4324, dj=1 / 18, J=127)),
                         ("dog2", pycwt.DOG(2), dict(s0=0.5033, dj=1 / 8, J=127))):
    for _ in range(3):
        t0 = time.perf_counter(); W = pycwt.cwt(x32, 1.0, wavelet=mother, **kw)[0]; t_e = time.perf_counter() - t0
    print("config3 %s e2e pycwt.cwt (fp32 engine, complex128 out, %d MB D2H): %.1f ms -> %.3e scale-points/s"
          % (name, W.nbytes >> 20, 1e3 * t_e, W.size / t_e))
del os.environ["CWTB_PRECISION"]
# device-resident products of config 2: nothing of size S x N crosses PCIe
N2 = 2 ** 20
x2 = chirp(N2)
for _ in range(3):
    t0 = time.perf_counter()
    r = pycwt.cwt_resident(x2, 1.0, 1 / 16, 2.0, 255, pycwt.Morlet(6))
    g = r.global_power(); sa = r.scale_avg_power(16.0, 64.0); iw = r.icwt()
    t_r = time.perf_counter() - t0
print("config2 resident: cwt + global power + scale-averaged power + icwt, host in / O(S+N) out: %.1f ms" % (1e3 * t_r))
# config 4: xwt + wct of two N=2^18 series, 145 scales, fp64 (deterministic part) + MC rate
rs = np.random.RandomState(0)
y1 = chirp(n) + 0.5 * rs.randn(n)
y2 = chirp(n, 0.7) + 0.5 * rs.randn(n)
sj = 2.0 * 2 ** (np.arange(145) / 12.0)
for _ in range(5):   # the first iterations allocate device buffers and the pinned result pool
    t0 = time.perf_counter(); W12 = eng.xwt(y1, y2, 1.0, sj, 0, 6.0); t_x = time.perf_counter() - t0
    t0 = time.perf_counter(); WCT, aWCT = eng.wct(y1, y2, 1.0, 1 / 12, sj, 0, 6.0, 14); t_w = time.perf_counter() - t0
print("config4 xwt (host in/out) %.3f s, wct(sig=False) %.3f s  [reference: 9.7 s / 31.7 s]" % (t_x, t_w))
nmc = 65536
noise = rs.randn(8, 2, 49152)
mask = np.ones((145, 49152), dtype=np.uint8)
hist = np.zeros((145, 1000), dtype=np.int64)
eng.wct_mc(noise, 1.0, 1 / 12, sj, 0, 6.0, 14, mask, 144, 1000, hist)
t0 = time.perf_counter(); eng.wct_mc(noise, 1.0, 1 / 12, sj, 0, 6.0, 14, mask, 144, 1000, hist); t_mc = (time.perf_counter() - t0) / 8
print("config4 wct_significance Monte-Carlo: %.4f s per surrogate pair (N=49152, 145 scales) -> 200 pairs %.1f s [reference ~23 s/pair]" % (t_mc, 200 * t_mc))
# config 5 slice: 64 channels of N = 2^16, 128 scales, fp32, one launch set
nch = 64
X = np.random.RandomState(1).randn(nch, 2 ** 16).astype(np.float32)
sj = 2.0 * 2 ** (np.arange(128) / 8.0)
d = eng.dev_alloc(X.nbytes); eng.h2d(d, X)
eng.cwt_batch_dev(d, nch, 2 ** 16, 1.0, sj, 0, 6.0, precision=1)
ms = eng.bench_last(10)
for k in sorted(eng.profile_last(), key=lambda k: -k["ms"])[:12]:
    print("   %-40s launches %2d  ms %.4f  rows %5d  us/row %.3f" % (k["name"][:40], k["launches"], k["ms"], k["rows"], 1e3 * k["ms"] / max(k["rows"], 1)))
print("config5 slice: %d channels x 128 scales x 2^16, fp32: %.3f ms -> %.3e scale-points/s, %.0f GB/s algorithmic (8192 ch on 8 GPUs = 16 such chunks per GPU)"
      % (nch, ms, nch * 128 * 65536 / ms * 1e3, nch * 128 * 65536 * 8 / ms / 1e6))
