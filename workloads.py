"""The BASELINE.json configurations as input generators, shared by bench.py and
the full-size parity tests.  Pure NumPy; nothing here touches the engine or the oracle."""
import numpy as np

# ---- config 2: Morlet CWT, synthetic chirp N = 2^20, 256 scales, fp64 --------------------------
C2 = dict(n=2 ** 20, dt=1.0, s0=2.0, dj=1.0 / 16, J=255, f0=6.0)
# ---- config 3: Paul(4) and DOG(2), N = 2^18, 128 scales, fp32 ----------------------------------
C3 = dict(n=2 ** 18, dt=1.0,
          paul=dict(m=4, s0=1.4324, dj=1.0 / 18, J=127),
          dog=dict(m=2, s0=0.5033, dj=1.0 / 8, J=127))
# ---- config 4: xwt + wct of two N = 2^18 series, Morlet, 200 surrogates ------------------------
C4 = dict(n=2 ** 18, dt=1.0, s0=2.0, dj=1.0 / 12, J=144, f0=6.0, mc_count=200)
# ---- config 5: 8192 channels of N = 2^16, 128 scales, fp32, 1024 channels per GPU --------------
C5 = dict(n=2 ** 16, dt=1.0, s0=2.0, dj=1.0 / 8, J=127, f0=6.0, channels=8192, per_gpu=1024)


def chirp(n, phase=0.0):
    """Linear chirp 50 -> n/4 cycles per record."""
    t = np.arange(n) / n
    return np.sin(2 * np.pi * (50 * t + (n / 8) * t ** 2) + phase)


def geometric_scales(s0, dj, J):
    return s0 * 2 ** (np.arange(0, J + 1) * dj)


def config2_signal(rank=0):
    return chirp(C2["n"], phase=0.1 * rank)


def config2_scales():
    return geometric_scales(C2["s0"], C2["dj"], C2["J"])


def config3_signal():
    return chirp(C3["n"]).astype(np.float32)


def config4_signals(n=None):
    n = n or C4["n"]
    rs = np.random.RandomState(0)
    y1 = chirp(n) + 0.5 * rs.randn(n)
    y2 = chirp(n, phase=0.7) + 0.5 * rs.randn(n)
    return y1, y2


def config5_channels(first, count, n=None):
    """Channels [first, first + count) of the 8192 x 2^16 float32 matrix.  Every channel has its
    own RandomState so that any slice can be generated without the rest (the matrix of SURVEY 8d
    is `RandomState(1).randn(8192, 2**16)`; per-channel seeding keeps shards reproducible)."""
    n = n or C5["n"]
    X = np.empty((count, n), dtype=np.float32)
    for i in range(count):
        X[i] = np.random.RandomState(1000 + first + i).randn(n).astype(np.float32)
    return X
