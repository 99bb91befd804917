"""Host-side helpers with the interface of pycwt/helpers.py.  O(N) or O(S) glue that
stays in NumPy (8f rank 1); the FFT-heavy work is in the engine."""
from os import environ, makedirs
from os.path import exists, expanduser

import numpy as np
import scipy.fftpack as fft  # the reference's backend module under the reference's name (helpers.py:22)

# True: the reference's scipy branch (helpers.py:22-30), transforms zero-padded to the next power
# of two.  False: the policy of its pyfftw branch (helpers.py:15-19), transforms at the signal's
# own length -- what a user with pyfftw installed gets from the stock package.  (In the
# reference the switch is "is pyfftw importable"; here it is this flag / CWTB_FFT_PAD=0.)
_FFT_NEXT_POW2 = environ.get('CWTB_FFT_PAD', '1') != '0'


def set_fft_padding(pad_to_next_pow2):
    """Choose the transform-length policy (see `_FFT_NEXT_POW2`).  Un-padded transforms run in
    fp64 (Bluestein's algorithm on the engine's power-of-two kernels)."""
    global _FFT_NEXT_POW2
    _FFT_NEXT_POW2 = bool(pad_to_next_pow2)


def fft_kwargs(signal, **kwargs):
    """Transform length for `signal` (helpers.py:15-19 and 27-30).  Other keyword arguments
    are dropped, as in the reference's scipy branch."""
    if _FFT_NEXT_POW2:
        return {'n': int(2 ** np.ceil(np.log2(len(signal))))}
    return {'n': len(signal)}


def find(condition):
    """Indices where ravel(condition) is true (helpers.py:37-40)."""
    res, = np.nonzero(np.ravel(condition))
    return res


def ar1(x):
    """Allen & Smith (1996) estimate of the lag-1 autocorrelation of a series, with the
    finite-sample bias correction used by the reference (helpers.py:43-104).

    Returns ``(g, a, mu2)``: lag-one autocorrelation, white-noise amplitude of the AR(1)
    model ``x_t - <x> = g (x_{t-1} - <x>) + a z_t`` and the (normalised) squared mean of a
    finite AR(1) segment.  Raises ``Warning`` when no upper bound can be placed on g
    (series too short or dominated by a trend), like the reference."""
    series = np.asarray(x)
    n = series.size
    dev = series - series.mean()
    if n < 4096:
        cov0 = np.dot(dev, dev) / n                     # lag-0 covariance
        cov1 = np.dot(dev[:-1], dev[1:]) / (n - 1)      # lag-1 covariance
    else:
        # long series: the threaded BLAS dot costs ~30 ms of thread start-up per call on a
        # many-core host (more than the GPU transform it precedes); a plain pairwise sum does not
        cov0 = float(np.sum(dev * dev)) / n
        cov1 = float(np.sum(dev[:-1] * dev[1:])) / (n - 1)

    # unbiased estimate: smaller root of  qa*g^2 + qb*g + qc = 0
    qa = cov0 * n ** 2
    qb = -cov1 * n - cov0 * n ** 2 - 2 * cov0 + 2 * cov1 - cov1 * n ** 2 + cov0 * n
    qc = n * (cov0 + cov1 * n - cov1)
    disc = qb ** 2 - 4 * qa * qc
    if not disc > 0:
        raise Warning('Cannot place an upperbound on the unbiased AR(1). '
                      'Series is too short or trend is to large.')
    g = (-qb - disc ** 0.5) / (2 * qa)

    # Allen & Smith footnote 4: expected squared mean of a length-n AR(1) segment
    mu2 = (2 / n ** 2) * ((n - g ** n) / (1 - g) - g * (1 - g ** (n - 1)) / (1 - g) ** 2) - 1 / n
    a = np.sqrt((1 - g ** 2) * cov0 / (1 - mu2))
    return g, a, mu2


def ar1_spectrum(freqs, ar1=0.):
    """Theoretical AR(1) power spectrum (helpers.py:107-143)."""
    freqs = np.asarray(freqs)
    return (1 - ar1 ** 2) / np.abs(1 - ar1 * np.exp(-2 * np.pi * 1j * freqs)) ** 2


def rednoise(N, g, a=1.):
    """Surrogate generator of the reference (helpers.py:146-173).

    The reference runs `lfilter` along the length-1 axis of an (N+tau, 1) array, which
    is the identity; its output is therefore the white-noise draw itself with the first
    tau = ceil(-2/ln|g|) samples dropped.  Reproduced exactly (same RNG consumption) so
    that Monte-Carlo significance levels match the reference draw for draw."""
    if g == 0:
        # the reference calls the non-existent np.randn here (AttributeError)
        yr = np.random.randn(N, 1) * a
    else:
        tau = int(np.ceil(-2 / np.log(np.abs(g))))
        yr = (np.random.randn(N + tau, 1) * a)[tau:]
    return yr.flatten()


def rect(x, normalize=False):
    """Boxcar window whose two end taps carry half weight (helpers.py:176-191).  `x` is the
    length (int/float), a shape list, or an array whose shape is used."""
    if isinstance(x, (np.ndarray, np.ma.core.MaskedArray)):
        shape = x.shape
    elif type(x) in (list, dict):
        shape = x
    else:
        shape = [x, ]
    win = np.ones(shape)
    win[0] = 0.5
    win[-1] = 0.5          # a single-tap window ends up as [0.5], like the reference
    if normalize:
        win /= win.sum()
    return win


def boxpdf(x):
    """Maps the data onto its empirical percentiles (helpers.py:194-225; the reference
    raises NameError on its last line, this version returns the intended result)."""
    x = np.asarray(x)
    n = x.size
    i = np.argsort(x)
    d = (np.diff(x[i]) != 0)
    j = find(np.concatenate([d, [True]]))
    X = x[i][j]
    j = np.concatenate([[0], j + 1])
    Y = 0.5 * (j[0:-1] + j[1:]) / n
    bX = np.interp(x, X, Y)
    return bX, X, Y


def get_cache_dir():
    """Location of the significance cache (helpers.py:228-236)."""
    cache_dir = '{}/.cache/pycwt/'.format(expanduser('~'))
    if not exists(cache_dir):
        makedirs(cache_dir)
    return cache_dir
