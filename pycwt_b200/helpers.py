"""Host-side helpers with the interface of pycwt/helpers.py.  O(N) or O(S) glue that
stays in NumPy (8f rank 1); the FFT-heavy work is in the engine."""
from os import makedirs
from os.path import exists, expanduser

import numpy as np
import scipy.fft as fft  # the reference exposes its FFT backend module under this name

_FFT_NEXT_POW2 = True


def fft_kwargs(signal, **kwargs):
    """Padding policy of the reference's scipy branch (helpers.py:27-30): transform
    length is the next power of two.  Other keyword arguments are dropped, as there."""
    if _FFT_NEXT_POW2:
        return {'n': int(2 ** np.ceil(np.log2(len(signal))))}


def find(condition):
    """Indices where ravel(condition) is true (helpers.py:37-40)."""
    res, = np.nonzero(np.ravel(condition))
    return res


def ar1(x):
    """Allen & Smith (1996) lag-1 autocorrelation estimate (helpers.py:43-104).

    Returns (g, a, mu2): lag-one autocorrelation, noise amplitude, and the squared mean
    of a finite AR(1) segment normalised by the process variance."""
    x = np.asarray(x)
    N = x.size
    x = x - x.mean()
    c0 = x.transpose().dot(x) / N
    c1 = x[0:N - 1].transpose().dot(x[1:N]) / (N - 1)
    B = -c1 * N - c0 * N ** 2 - 2 * c0 + 2 * c1 - c1 * N ** 2 + c0 * N
    A = c0 * N ** 2
    C = N * (c0 + c1 * N - c1)
    D = B ** 2 - 4 * A * C
    if D > 0:
        g = (-B - D ** 0.5) / (2 * A)
    else:
        raise Warning('Cannot place an upperbound on the unbiased AR(1). '
                      'Series is too short or trend is to large.')
    mu2 = -1 / N + (2 / N ** 2) * ((N - g ** N) / (1 - g) -
                                   g * (1 - g ** (N - 1)) / (1 - g) ** 2)
    c0t = c0 / (1 - mu2)
    a = ((1 - g ** 2) * c0t) ** 0.5
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
    """Boxcar window with half-weight end taps (helpers.py:176-191)."""
    if type(x) in [int, float]:
        shape = [x, ]
    elif type(x) in [list, dict]:
        shape = x
    elif type(x) in [np.ndarray, np.ma.core.MaskedArray]:
        shape = x.shape
    X = np.zeros(shape)
    X[0] = X[-1] = 0.5
    X[1:-1] = 1
    if normalize:
        X /= X.sum()
    return X


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
