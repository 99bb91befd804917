"""Mother-wavelet objects with the interface of pycwt/mothers.py (reference
file:line cited per method).  They are parameter carriers for the CUDA engine plus the
small closed-form host formulas user code calls directly (flambda, coi, psi, psi_ft)."""
import numpy as np
from numpy.polynomial import hermite_e as _hermite_e
from scipy.special import gamma as _gamma

from . import _engine


#: Opt-in: a smoothing operator for Paul and DOG.  The reference defines
#: `smooth` for Morlet only (mothers.py:61-104; `wct` with Paul/DOG raises AttributeError there,
#: and here while this is off).  When on, Paul/DOG objects expose `smooth` with the general
#: definition of Torrence & Webster (1999) / TC98 sec. 6 that Morlet's Gaussian is the special
#: case of: in time a filter given by the absolute value of the wavelet function at each scale,
#: normalised to unit weight; in scale a boxcar of width deltaj0 (TC98 table 2).
_GENERIC_SMOOTHING = False


def enable_generic_smoothing(on=True):
    """Give Paul and DOG a `smooth` method (see `_GENERIC_SMOOTHING`); returns the old setting."""
    global _GENERIC_SMOOTHING
    old, _GENERIC_SMOOTHING = _GENERIC_SMOOTHING, bool(on)
    return old


def time_filter_table(wavelet, scales, dt, npad):
    """Real frequency responses [S, npad] of the time smoothing |psi0(t/s)| / sum: the DFT of the
    wavelet modulus sampled on the circular grid of `npad` points (for Morlet this is the sampled
    counterpart of the Gaussian exp(-0.5 (s/dt)^2 k^2) of mothers.py:83-91)."""
    n = np.arange(npad)
    t = dt * np.where(n <= npad // 2, n, n - npad)
    out = np.empty((len(scales), npad))
    for j, s in enumerate(np.asarray(scales, dtype=float)):
        k = np.abs(wavelet.psi(t / s))
        out[j] = np.fft.fft(k / k.sum()).real
    return out


class _Base(object):
    #: engine family id and the attribute holding its parameter
    _family = None
    cdelta = -1
    gamma = -1
    deltaj0 = -1

    @property
    def smooth(self):
        """Only with `enable_generic_smoothing()`: see `_GENERIC_SMOOTHING`."""
        if not _GENERIC_SMOOTHING:
            raise AttributeError("'{}' object has no attribute 'smooth'".format(type(self).__name__))
        return self._smooth_generic

    def _smooth_generic(self, W, dt, dj, scales):
        from .wavelet import _smooth_device
        return _smooth_device(W, dt, dj, scales, self.deltaj0, wavelet=self)

    def _engine_spec(self):
        """(family, param) if the engine evaluates this wavelet analytically, else None."""
        return None

    def sup(self):
        """Wavelet support defined by the e-folding time (reference mothers.py:42-44
        divides by the bound method and raises TypeError; here it is evaluated)."""
        return 1.0 / self.coi()


class Morlet(_Base):
    """Morlet wavelet, angular wavenumber f0 (reference mothers.py:13-104)."""

    def __init__(self, f0=6):
        self._set_f0(f0)
        self.name = 'Morlet'

    def psi_ft(self, f):
        # mothers.py:26-28: two-sided, no Heaviside step
        return np.pi ** -0.25 * np.exp(-0.5 * (f - self.f0) ** 2)

    def psi(self, t):
        # mothers.py:30-32
        return np.pi ** -0.25 * np.exp(1j * self.f0 * t - t ** 2 / 2)

    def flambda(self):
        # mothers.py:34-36
        return (4 * np.pi) / (self.f0 + np.sqrt(2 + self.f0 ** 2))

    def coi(self):
        # mothers.py:38-40
        return 1. / np.sqrt(2)

    def _set_f0(self, f0):
        # Torrence & Compo (1998) table 2; mothers.py:46-59
        self.f0 = f0
        self.dofmin = 2
        if self.f0 == 6:
            self.cdelta, self.gamma, self.deltaj0 = 0.776, 2.32, 0.60
        else:
            self.cdelta = self.gamma = self.deltaj0 = -1

    def _engine_spec(self):
        if type(self).psi_ft is not Morlet.psi_ft:
            return None
        return _engine.MORLET, float(self.f0)

    def smooth(self, W, dt, dj, scales):
        """Coherence smoothing operator (mothers.py:61-104): Gaussian in time per scale
        (FFT, zero-padded to the next power of two) then a boxcar of width
        2*deltaj0/dj with half-weight end taps along the scale axis.  Runs on the GPU."""
        from .wavelet import _smooth_device
        return _smooth_device(W, dt, dj, scales, self.deltaj0)


class Paul(_Base):
    """Paul wavelet of order m (reference mothers.py:107-155)."""

    def __init__(self, m=4):
        self._set_m(m)
        self.name = 'Paul'

    def psi_ft(self, f):
        # mothers.py:118-122; for f < -709.78 exp overflows and inf*0 gives NaN
        m = self.m
        return (2 ** m / np.sqrt(m * np.prod(range(2, 2 * m))) *
                f ** m * np.exp(-f) * (f > 0))

    def psi(self, t):
        # mothers.py:124-128 (keeps the reference's prod(range(2, m-1)) factor)
        m = self.m
        return (2 ** m * 1j ** m * np.prod(range(2, m - 1)) /
                np.sqrt(np.pi * np.prod(range(2, 2 * m + 1))) *
                (1 - 1j * t) ** (-(m + 1)))

    def flambda(self):
        # mothers.py:130-132
        return 4 * np.pi / (2 * self.m + 1)

    def coi(self):
        # mothers.py:134-136
        return np.sqrt(2)

    def _set_m(self, m):
        # mothers.py:142-155
        self.m = m
        self.dofmin = 2
        if self.m == 4:
            self.cdelta, self.gamma, self.deltaj0 = 1.132, 1.17, 1.50
        else:
            self.cdelta = self.gamma = self.deltaj0 = -1

    def _engine_spec(self):
        if type(self).psi_ft is not Paul.psi_ft or int(self.m) != self.m or not 1 <= self.m <= 64:
            return None
        return _engine.PAUL, float(self.m)


class DOG(_Base):
    """m-th derivative of a Gaussian (reference mothers.py:158-222)."""

    def __init__(self, m=2):
        self._set_m(m)
        self.name = 'DOG'

    def psi_ft(self, f):
        # mothers.py:170-173; note -(1j**m): the minus applies after the power
        return (- 1j ** self.m / np.sqrt(_gamma(self.m + 0.5)) * f ** self.m *
                np.exp(- 0.5 * f ** 2))

    def psi(self, t):
        # mothers.py:175-191: probabilists' Hermite polynomial He_m
        he = _hermite_e.hermeval(t, [0] * int(self.m) + [1])
        return ((-1) ** (self.m + 1) * he * np.exp(-t ** 2 / 2) /
                np.sqrt(_gamma(self.m + 0.5)))

    def flambda(self):
        # mothers.py:193-195
        return 2 * np.pi / np.sqrt(self.m + 0.5)

    def coi(self):
        # mothers.py:197-199
        return 1 / np.sqrt(2)

    def _set_m(self, m):
        # mothers.py:205-222
        self.m = m
        self.dofmin = 1
        if self.m == 2:
            self.cdelta, self.gamma, self.deltaj0 = 3.541, 1.43, 1.40
        elif self.m == 6:
            self.cdelta, self.gamma, self.deltaj0 = 1.966, 1.37, 0.97
        else:
            self.cdelta = self.gamma = self.deltaj0 = -1

    def _engine_spec(self):
        if type(self).psi_ft is not DOG.psi_ft or int(self.m) != self.m or not 1 <= self.m <= 64:
            return None
        return _engine.DOG, float(self.m)


class MexicanHat(DOG):
    """DOG with m = 2 (reference mothers.py:225-233)."""

    def __init__(self):
        self.name = 'Mexican Hat'
        self._set_m(2)
