"""CPU oracle for the CWT hot path -- TEST INFRASTRUCTURE ONLY.

This module is a numpy/scipy restatement of the algorithm regeirk/pycwt uses on
its `cwt / icwt / xwt / wct / wct_significance` path.  It exists so that the
CUDA engine in `pycwt_b200/` can be checked against an independent CPU
implementation on a box where `/root/reference` does not exist.

Rules:
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
    `--impl reference` legs may import this file;
  * the product package `pycwt_b200` never imports it and has no CPU fallback.

Parity pinning: the reference's own test-suite holds no golden vectors
(SURVEY.md section 4), so this oracle is pinned against *outputs of the
reference itself*, generated in the build container by
`tests/golden/make_golden.py` (which imports `/root/reference/pycwt`) and
committed as `tests/golden/*.npz`.  `tests/test_oracle_golden.py` checks every
function below against those fixtures.

Each function cites the reference lines it restates (paths relative to
/root/reference/).  The FFT itself is third-party in the reference
(scipy.fftpack -> scipy.fft / DUCC, unpinned version; scipy 1.18.1 here); any
exact FFT agrees to ~1e-15*log2(N), we use scipy.fft.
"""

import numpy as np
import scipy.fft as _sfft
from scipy.special import gamma as _gamma
from scipy.stats import chi2 as _chi2

__all__ = ["Morlet", "Paul", "DOG", "MexicanHat", "cwt", "icwt", "xwt", "wct",
           "smooth", "wct_significance", "ar1", "ar1_spectrum", "rednoise",
           "rect", "resolve_wavelet", "next_pow2"]


def next_pow2(n):
    """Padding policy of the scipy branch, pycwt/helpers.py:27-30."""
    return int(2 ** np.ceil(np.log2(n)))


# Transform-length policy, pycwt/helpers.py:7-30.  True: the scipy branch (zero-pad to the next
# power of two, :27-30) -- the behaviour of the reference in this image.  False: the policy of
# the pyfftw branch (:15-19, `kwargs['n'] = len(signal)  # do not pad`).
PAD_NEXT_POW2 = True


def transform_length(n):
    return next_pow2(n) if PAD_NEXT_POW2 else int(n)


# --------------------------------------------------------------------------
# Mother wavelets (pycwt/mothers.py)
# --------------------------------------------------------------------------
class _Mother(object):
    cdelta = gamma = deltaj0 = -1


class Morlet(_Mother):
    """pycwt/mothers.py:13-104."""
    name = "Morlet"
    dofmin = 2

    def __init__(self, f0=6):
        self.f0 = f0
        if f0 == 6:  # TC98 table 2, mothers.py:52-55
            self.cdelta, self.gamma, self.deltaj0 = 0.776, 2.32, 0.60

    def psi_ft(self, f):  # mothers.py:26-28 (two-sided, no Heaviside)
        return np.pi ** -0.25 * np.exp(-0.5 * (f - self.f0) ** 2)

    def psi(self, t):  # mothers.py:30-32
        return np.pi ** -0.25 * np.exp(1j * self.f0 * t - t ** 2 / 2)

    def flambda(self):  # mothers.py:34-36
        return 4 * np.pi / (self.f0 + np.sqrt(2 + self.f0 ** 2))

    def coi(self):  # mothers.py:38-40
        return 1.0 / np.sqrt(2)

    def smooth(self, W, dt, dj, scales):
        return smooth(W, dt, dj, scales, self.deltaj0)


class Paul(_Mother):
    """pycwt/mothers.py:107-155."""
    name = "Paul"
    dofmin = 2

    def __init__(self, m=4):
        self.m = m
        if m == 4:  # mothers.py:148-151
            self.cdelta, self.gamma, self.deltaj0 = 1.132, 1.17, 1.50

    def psi_ft(self, f):  # mothers.py:118-122; inf*0 -> NaN for f < -709.78
        m = self.m
        fact = float(np.prod(range(2, 2 * m)))  # (2m-1)!
        return 2 ** m / np.sqrt(m * fact) * f ** m * np.exp(-f) * (f > 0)

    def psi(self, t):  # mothers.py:124-128 (note prod(range(2, m-1)) quirk)
        m = self.m
        return (2 ** m * 1j ** m * np.prod(range(2, m - 1)) /
                np.sqrt(np.pi * np.prod(range(2, 2 * m + 1))) *
                (1 - 1j * t) ** (-(m + 1)))

    def flambda(self):  # mothers.py:130-132
        return 4 * np.pi / (2 * self.m + 1)

    def coi(self):  # mothers.py:134-136
        return np.sqrt(2)


class DOG(_Mother):
    """pycwt/mothers.py:158-222."""
    name = "DOG"
    dofmin = 1

    def __init__(self, m=2):
        self.m = m
        if m == 2:  # mothers.py:211-218
            self.cdelta, self.gamma, self.deltaj0 = 3.541, 1.43, 1.40
        elif m == 6:
            self.cdelta, self.gamma, self.deltaj0 = 1.966, 1.37, 0.97

    def psi_ft(self, f):  # mothers.py:170-173; unary minus binds after **
        m = self.m
        return -(1j ** m) / np.sqrt(_gamma(m + 0.5)) * f ** m * np.exp(-0.5 * f ** 2)

    def psi(self, t):  # mothers.py:175-191 (probabilists' Hermite polynomial)
        from numpy.polynomial import hermite_e as _He
        m = self.m
        he = _He.hermeval(t, [0] * m + [1])
        return (-1) ** (m + 1) * he * np.exp(-t ** 2 / 2) / np.sqrt(_gamma(m + 0.5))

    def flambda(self):  # mothers.py:193-195
        return 2 * np.pi / np.sqrt(self.m + 0.5)

    def coi(self):  # mothers.py:197-199
        return 1 / np.sqrt(2)


class MexicanHat(DOG):
    """pycwt/mothers.py:225-233."""

    def __init__(self):
        DOG.__init__(self, 2)
        self.name = "Mexican Hat"


def resolve_wavelet(w):
    """pycwt/wavelet.py:650-663."""
    table = {"morlet": Morlet, "paul": Paul, "dog": DOG, "mexicanhat": MexicanHat}
    if isinstance(w, str):
        return table[w]()
    return w


# --------------------------------------------------------------------------
# cwt (pycwt/wavelet.py:13-124)
# --------------------------------------------------------------------------
def cwt(signal, dt, dj=1 / 12, s0=-1, J=-1, wavelet="morlet", freqs=None,
        workers=1):
    mother = resolve_wavelet(wavelet)
    n0 = len(signal)
    lam = mother.flambda()
    if freqs is None:  # wavelet.py:75-85
        if s0 == -1:
            s0 = 2 * dt / lam
        if J == -1:
            J = int(np.round(np.log2(n0 * dt / s0) / dj))
        sj = s0 * 2 ** (np.arange(0, J + 1) * dj)
        freqs = 1 / (lam * sj)
    else:  # wavelet.py:86-88
        sj = 1 / (lam * freqs)

    npad = transform_length(n0)  # helpers.py:15-19 / 27-30
    spec = _sfft.fft(np.asarray(signal), n=npad)  # wavelet.py:91
    omega = 2 * np.pi * _sfft.fftfreq(npad, dt)  # wavelet.py:94
    col = sj[:, None]
    # wavelet.py:102-104
    filt = (col * omega[1] * npad) ** 0.5 * np.conjugate(mother.psi_ft(col * omega))
    W = _sfft.ifft(spec * filt, axis=1, workers=workers)  # wavelet.py:105-106

    keep = ~np.isnan(W).all(axis=1)  # wavelet.py:111-115
    if keep.any():
        sj, freqs, W = sj[keep], freqs[keep], W[keep, :]

    tri = n0 / 2 - np.abs(np.arange(0, n0) - (n0 - 1) / 2)  # wavelet.py:120-121
    coi = lam * mother.coi() * dt * tri
    return (W[:, :n0], sj, freqs, coi, spec[1:npad // 2] / npad ** 0.5,
            omega[1:npad // 2] / (2 * np.pi))


# --------------------------------------------------------------------------
# icwt (pycwt/wavelet.py:127-171)
# --------------------------------------------------------------------------
def icwt(W, sj, dt, dj=1 / 12, wavelet="morlet"):
    mother = resolve_wavelet(wavelet)
    a, b = W.shape
    c = sj.size
    if a == c:
        smat = np.broadcast_to(sj[:, None], (a, b))
    elif b == c:
        smat = np.broadcast_to(sj[None, :], (a, b))
    else:
        raise Warning("Input array dimensions do not match.")
    coef = dj * np.sqrt(dt) / (mother.cdelta * mother.psi(0))
    return coef * (np.real(W) / np.sqrt(smat)).sum(axis=0)


# --------------------------------------------------------------------------
# helpers (pycwt/helpers.py)
# --------------------------------------------------------------------------
def ar1(x):
    """Allen & Smith lag-1 estimate, pycwt/helpers.py:43-104."""
    x = np.asarray(x)
    N = x.size
    x = x - x.mean()
    c0 = x.dot(x) / N
    c1 = x[:-1].dot(x[1:]) / (N - 1)
    B = -c1 * N - c0 * N ** 2 - 2 * c0 + 2 * c1 - c1 * N ** 2 + c0 * N
    A = c0 * N ** 2
    C = N * (c0 + c1 * N - c1)
    D = B ** 2 - 4 * A * C
    if D <= 0:
        raise Warning("Cannot place an upperbound on the unbiased AR(1). "
                      "Series is too short or trend is to large.")
    g = (-B - D ** 0.5) / (2 * A)
    mu2 = -1 / N + (2 / N ** 2) * ((N - g ** N) / (1 - g) -
                                   g * (1 - g ** (N - 1)) / (1 - g) ** 2)
    c0t = c0 / (1 - mu2)
    a = ((1 - g ** 2) * c0t) ** 0.5
    return g, a, mu2


def ar1_spectrum(freqs, ar1=0.0):
    """pycwt/helpers.py:107-143."""
    freqs = np.asarray(freqs)
    return (1 - ar1 ** 2) / np.abs(1 - ar1 * np.exp(-2j * np.pi * freqs)) ** 2


def rednoise(N, g, a=1.0, rng=None):
    """pycwt/helpers.py:146-173.

    The reference filters an (N+tau, 1) array along its last (length-1) axis, so
    the 'red' noise it returns is the white-noise draw itself with the first
    tau samples discarded.  Reproduced as-is (parity target).
    `rng` defaults to numpy's global RNG like the reference.
    """
    randn = np.random.randn if rng is None else rng.randn
    tau = int(np.ceil(-2 / np.log(np.abs(g))))
    draw = randn(N + tau, 1) * a
    return draw[tau:].flatten()


def rect(k, normalize=False):
    """Boxcar with half-weight end taps, pycwt/helpers.py:176-191."""
    win = np.zeros(k)
    win[0] = win[-1] = 0.5
    win[1:-1] = 1
    if normalize:
        win /= win.sum()
    return win


# --------------------------------------------------------------------------
# Morlet.smooth (pycwt/mothers.py:61-104)
# --------------------------------------------------------------------------
def smooth(W, dt, dj, scales, deltaj0=0.60):
    m, n = W.shape
    npad = transform_length(n)
    k2 = (2 * np.pi * _sfft.fftfreq(npad)) ** 2  # mothers.py:83-84
    snorm = scales / dt
    F = np.exp(-0.5 * (snorm[:, None] ** 2) * k2)  # mothers.py:89
    T = _sfft.ifft(F * _sfft.fft(W, n=npad, axis=1), axis=1)[:, :n]  # :90-93
    if np.isreal(W).all():  # :95-96
        T = T.real
    win = rect(int(np.round(deltaj0 / dj * 2)), normalize=True)  # :100-101
    # scipy.signal.convolve2d(T, win[:, None], 'same') with zero fill (:102):
    # out[i] = sum_q T[q] * win[i + (K-1)//2 - q]
    K = win.size
    off = (K - 1) // 2
    out = np.zeros_like(T)
    for t in range(K):
        sh = t - off  # out[i] += win[t] * T[i - sh]
        if sh >= 0:
            if sh < m:
                out[sh:, :] += win[t] * T[:m - sh, :]
        else:
            if -sh < m:
                out[:m + sh, :] += win[t] * T[-sh:, :]
    return out


# --------------------------------------------------------------------------
# xwt (pycwt/wavelet.py:316-419)
# --------------------------------------------------------------------------
def _normalise(y, normalize):
    y = np.asarray(y)
    std = y.std()
    yn = (y - y.mean()) / std if normalize else y
    return y, yn, std


def xwt(y1, y2, dt, dj=1 / 12, s0=-1, J=-1, significance_level=0.95,
        wavelet="morlet", normalize=True):
    mother = resolve_wavelet(wavelet)
    y1, y1n, std1 = _normalise(y1, normalize)
    y2, y2n, std2 = _normalise(y2, normalize)
    W1, sj, freq, coi, _, _ = cwt(y1n, dt, dj=dj, s0=s0, J=J, wavelet=mother)
    W2, sj, freq, coi, _, _ = cwt(y2n, dt, dj=dj, s0=s0, J=J, wavelet=mother)
    W12 = W1 * W2.conj()  # wavelet.py:399
    if normalize:  # wavelet.py:408-409
        std1 = std2 = 1.0
    a1, _, _ = ar1(y1)
    a2, _, _ = ar1(y2)
    Pk1 = ar1_spectrum(freq * dt, a1)
    Pk2 = ar1_spectrum(freq * dt, a2)
    dof = mother.dofmin
    ppf = _chi2.ppf(significance_level, dof)
    signif = std1 * std2 * (Pk1 * Pk2) ** 0.5 * ppf / dof  # wavelet.py:416
    return W12, coi, freq, signif


# --------------------------------------------------------------------------
# wct (pycwt/wavelet.py:422-528)
# --------------------------------------------------------------------------
def wct(y1, y2, dt, dj=1 / 12, s0=-1, J=-1, sig=True, significance_level=0.95,
        wavelet="morlet", normalize=True, **kwargs):
    mother = resolve_wavelet(wavelet)
    if s0 == -1:
        s0 = 2 * dt / mother.flambda()
    if J == -1:
        J = int(np.round(np.log2(y1.size * dt / s0) / dj))
    y1, y1n, _ = _normalise(y1, normalize)
    y2, y2n, _ = _normalise(y2, normalize)
    W1, sj, freq, coi, _, _ = cwt(y1n, dt, dj=dj, s0=s0, J=J, wavelet=mother)
    W2, sj, freq, coi, _, _ = cwt(y2n, dt, dj=dj, s0=s0, J=J, wavelet=mother)
    inv_s = 1.0 / sj[:, None]
    S1 = mother.smooth(np.abs(W1) ** 2 * inv_s, dt, dj, sj)  # wavelet.py:506
    S2 = mother.smooth(np.abs(W2) ** 2 * inv_s, dt, dj, sj)  # :507
    W12 = W1 * W2.conj()  # :510
    S12 = mother.smooth(W12 * inv_s, dt, dj, sj)  # :512
    WCT = np.abs(S12) ** 2 / (S1 * S2)  # :513
    aWCT = np.angle(W12)  # :514 (un-smoothed)
    if sig:
        a1, _, _ = ar1(y1)
        a2, _, _ = ar1(y2)
        sig = wct_significance(a1, a2, dt=dt, dj=dj, s0=s0, J=J,
                               significance_level=significance_level,
                               wavelet=mother, **kwargs)
    else:
        sig = np.asarray([0])
    return WCT, aWCT, coi, freq, sig


# --------------------------------------------------------------------------
# wct_significance (pycwt/wavelet.py:531-647), cache handling omitted: the
# oracle always recomputes.  RNG consumption order matches the reference: one
# set-up draw, then (noise1, noise2) per iteration, from numpy's global RNG
# unless `rng` is given.
# --------------------------------------------------------------------------
def wct_significance(al1, al2, dt, dj, s0, J, significance_level=0.95,
                     wavelet="morlet", mc_count=300, progress=False,
                     cache=False, rng=None, return_hist=False):
    mother = resolve_wavelet(wavelet)
    ms = s0 * (2 ** (J * dj)) / dt  # wavelet.py:592-593
    N = int(np.ceil(ms * 6))
    noise1 = rednoise(N, al1, 1, rng)
    nW1, sj, freq, coi, _, _ = cwt(noise1, dt=dt, dj=dj, s0=s0, J=J, wavelet=mother)
    period = 1.0 / freq[:, None]
    outside = period <= coi[None, :]  # wavelet.py:598-600
    has = outside.any(axis=1)
    maxscale = np.nonzero(has)[0][-1]
    sig95 = np.zeros(J + 1)
    sig95[has] = np.nan
    nbins = 1000
    hist = np.zeros((J + 1, nbins), dtype=np.int64)
    inv_s = 1.0 / sj[:, None]
    for _ in range(mc_count):
        noise1 = rednoise(N, al1, 1, rng)
        noise2 = rednoise(N, al2, 1, rng)
        nW1 = cwt(noise1, dt=dt, dj=dj, s0=s0, J=J, wavelet=mother)[0]
        nW2 = cwt(noise2, dt=dt, dj=dj, s0=s0, J=J, wavelet=mother)[0]
        nW12 = nW1 * nW2.conj()
        S1 = mother.smooth(np.abs(nW1) ** 2 * inv_s, dt, dj, sj)
        S2 = mother.smooth(np.abs(nW2) ** 2 * inv_s, dt, dj, sj)
        S12 = mother.smooth(nW12 * inv_s, dt, dj, sj)
        R2 = np.abs(S12) ** 2 / (S1 * S2)
        # wavelet.py:627-630, vectorised: floor(R2*nbins) for valid points of
        # rows s < maxscale (a value of exactly nbins would raise IndexError in
        # the reference; it cannot occur for |S12|^2 <= S1*S2 up to rounding).
        for s in range(maxscale):
            idx = np.floor(R2[s, outside[s]] * nbins).astype(np.int64)
            hist[s] += np.bincount(idx, minlength=nbins)[:nbins]
    sig95 = percentile_from_hist(hist, sig95, maxscale, significance_level)
    if return_hist:
        return sig95, hist
    return sig95


def percentile_from_hist(hist, sig95, maxscale, level):
    """pycwt/wavelet.py:634-640."""
    nbins = hist.shape[1]
    centres = (np.arange(nbins) + 0.5) / nbins
    for s in range(maxscale):
        sel = hist[s] != 0
        P = hist[s, sel].astype(float).cumsum()
        P = (P - 0.5) / P[-1]
        sig95[s] = np.interp(level, P, centres[sel])
    return sig95


# --------------------------------------------------------------------------
# Smoothing operator for wavelets the reference has none for.
# NOT a restatement of reference code (mothers.py:107-222 define no `smooth`): an independent
# NumPy statement of the general definition (Torrence & Webster 1999; TC98 sec. 6) that the
# package offers behind `mothers.enable_generic_smoothing()`, so that the GPU path has a checker.
# --------------------------------------------------------------------------
def smooth_generic(W, dt, dj, scales, mother):
    """Time: circular convolution (length = transform_length(n)) with |psi0(t/s)| normalised to unit
    sum; scale: the boxcar of `smooth` with width 2*deltaj0/dj (same alignment as mothers.py:100-102)."""
    from scipy.signal import convolve2d
    m, n = W.shape
    npad = transform_length(n)
    idx = np.arange(npad)
    t = dt * np.where(idx <= npad // 2, idx, idx - npad)
    T = np.zeros((m, n), dtype=complex)
    for j in range(m):
        k = np.abs(mother.psi(t / scales[j]))
        k = k / k.sum()
        x = np.zeros(npad, dtype=complex)
        x[:n] = W[j]
        # direct O(n^2) would be the purest statement; the FFT of the SAME sampled kernel is exact
        # to rounding and keeps the test fast
        T[j] = _sfft.ifft(_sfft.fft(x) * _sfft.fft(k))[:n]
    if np.isreal(W).all():
        T = T.real
    wsize = int(np.round(mother.deltaj0 / dj * 2))
    win = rect(wsize, normalize=True)
    return convolve2d(T, win[:, np.newaxis], "same")
