"""GPU parity tests of the rows around the cwt kernel pair: icwt, xwt, wct, Morlet.smooth,
Monte-Carlo coherence significance, derived products, batched channels.  Fixtures come from
the real reference (tests/golden/make_golden.py); seeded cases are checked against the CPU
oracle.  fp64 tolerance 1e-10 relative to max|ref| (north_star); achieved ~1e-14."""
import numpy as np
import pytest

from conftest import load_golden, relerr
from oracle import cwt_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def pycwt():
    import pycwt_b200
    return pycwt_b200


def noisy_chirps(n, seed=0):
    rs = np.random.RandomState(seed)
    t = np.arange(n) / n
    ph = 2 * np.pi * (50 * t + (n / 8) * t ** 2)
    return np.sin(ph) + 0.5 * rs.randn(n), np.sin(ph + 0.7) + 0.5 * rs.randn(n)


def test_icwt_golden(pycwt):
    for name in ("nino3_morlet_tutorial", "nino3_dog_default", "nino3_paul_default",
                 "chirp4000_morlet"):
        g = load_golden(name)
        if "iW" not in g.files:
            continue
        kw = {str(k): float(v) for k, v in zip(g["kw_keys"], g["kw_vals"])}
        if "J" in kw:
            kw["J"] = int(kw["J"])
        mother = {"morlet": pycwt.Morlet, "paul": pycwt.Paul, "dog": pycwt.DOG}[str(g["wavelet"])](int(g["param"]))
        W, sj, *_ = pycwt.cwt(g["x"], float(g["dt"]), wavelet=mother, **kw)
        iW = pycwt.icwt(W, sj, float(g["dt"]), kw.get("dj", 1 / 12), mother)
        assert iW.dtype == g["iW"].dtype
        assert relerr(iW, g["iW"]) < TOL
        # transposed input is accepted like in the reference (still reduces axis 0)
        if W.shape[0] != W.shape[1]:
            ref_t = orc.icwt(W.T.copy(), sj, float(g["dt"]), kw.get("dj", 1 / 12),
                             {"morlet": orc.Morlet, "paul": orc.Paul, "dog": orc.DOG}[str(g["wavelet"])](int(g["param"])))
            assert relerr(pycwt.icwt(W.T.copy(), sj, float(g["dt"]), kw.get("dj", 1 / 12), mother), ref_t) < TOL
    with pytest.raises(Warning):
        pycwt.icwt(np.zeros((3, 5), complex), np.ones(4), 1.0)


def test_icwt_resident_and_power(pycwt):
    g = load_golden("nino3_morlet_tutorial")
    W, sj, *_ = pycwt.cwt(g["x"], 0.25, 0.25, 0.5, 28, pycwt.Morlet(6))
    eng = pycwt.default_engine()
    red = eng.icwt_sum()
    assert relerr(red, (W.real / np.sqrt(sj)[:, None]).sum(axis=0)) < 1e-13
    assert relerr(eng.global_power(len(sj)), (np.abs(W) ** 2).mean(axis=1)) < 1e-13
    assert relerr(eng.power(len(sj), W.shape[1]), np.abs(W) ** 2) < 1e-13


def test_xwt_golden(pycwt):
    g = load_golden("ao_baltic_xwt_wct")
    y1, y2, dt = g["y1"], g["y2"], float(g["dt"])
    W12, coi, freq, signif = pycwt.xwt(y1, y2, dt, dj=1 / 12, significance_level=0.8646,
                                       wavelet=pycwt.Morlet(6))
    assert W12.shape == (76, 147) and W12.dtype == np.complex128
    assert relerr(W12, g["W12"]) < TOL
    np.testing.assert_allclose(signif, g["signif"], rtol=1e-12)
    np.testing.assert_array_equal(coi, g["coi"])
    np.testing.assert_array_equal(freq, g["freq"])
    W12n, _, _, sn = pycwt.xwt(list(y1), list(y2), dt, dj=1 / 12, wavelet="morlet", normalize=False)
    assert relerr(W12n, g["W12_nonorm"]) < TOL
    np.testing.assert_allclose(sn, g["signif_nonorm"], rtol=1e-12)


def test_wct_golden(pycwt):
    g = load_golden("ao_baltic_xwt_wct")
    WCT, aWCT, coi, freq, sig = pycwt.wct(g["y1"], g["y2"], float(g["dt"]), dj=1 / 12, sig=False,
                                          wavelet="morlet")
    assert relerr(WCT, g["WCT"]) < TOL
    assert relerr(aWCT, g["aWCT"]) < TOL
    assert sig.shape == (1,) and sig[0] == 0
    with pytest.raises(AttributeError):   # Paul / DOG have no smoothing operator (reference too)
        pycwt.wct(g["y1"], g["y2"], float(g["dt"]), sig=False, wavelet="paul")


def test_smooth_golden(pycwt):
    g = load_golden("smooth_cases")
    m = pycwt.Morlet(6)
    Sr = m.smooth(g["Wr"], 1.0, 0.25, g["sj"])
    Sc = m.smooth(g["Wc"], 1.0, 0.25, g["sj"])
    assert Sr.dtype == np.float64 and Sc.dtype == np.complex128
    assert relerr(Sr, g["Sr"]) < TOL and relerr(Sc, g["Sc"]) < TOL


def test_wct_significance_seeded_exact(pycwt):
    """Same seeded numpy RNG stream as the reference -> same surrogates -> the device
    histograms must reproduce the reference's significance levels."""
    g = load_golden("wct_significance_seed1234")
    np.random.seed(int(g["seed"]))
    sig = pycwt.wct_significance(float(g["al1"]), float(g["al2"]), dt=float(g["dt"]),
                                 dj=float(g["dj"]), s0=float(g["s0"]), J=int(g["J"]),
                                 significance_level=float(g["level"]), wavelet="morlet",
                                 mc_count=int(g["mc_count"]), progress=False, cache=False)
    assert relerr(sig, g["sig95"]) < 1e-12
    g = load_golden("wct_sig_seed99")
    np.random.seed(99)
    WCT, aWCT, coi, freq, sig = pycwt.wct(g["y1"], g["y2"], 1.0, dj=0.5, s0=2.0, J=8, sig=True,
                                          wavelet=pycwt.Morlet(6), mc_count=4, progress=False,
                                          cache=False)
    assert relerr(WCT, g["WCT"]) < TOL
    assert relerr(sig, g["sig"]) < 1e-12


def test_wct_mc_histogram_vs_oracle(pycwt):
    """Host-supplied surrogates: device histogram == oracle histogram, bin for bin
    (bins are integers; a handful of samples may sit within rounding of a bin edge)."""
    rs = np.random.RandomState(5)
    dt, dj, s0, J = 1.0, 0.25, 2.0, 24
    sig_ref, hist_ref = orc.wct_significance(0.3, 0.2, dt, dj, s0, J, mc_count=3, rng=rs,
                                             return_hist=True)
    rs = np.random.RandomState(5)
    np.random.set_state(rs.get_state())
    sig = pycwt.wct_significance(0.3, 0.2, dt, dj, s0, J, wavelet="morlet", mc_count=3,
                                 progress=False, cache=False)
    assert relerr(sig, sig_ref) < 1e-9


def test_xwt_wct_medium_two_kernel_paths(pycwt):
    a, b = noisy_chirps(6000, 0)
    W12, *_ = pycwt.xwt(a, b, 1.0, dj=1 / 4, s0=2.0, J=40, wavelet="morlet")
    W12r, *_ = orc.xwt(a, b, 1.0, dj=1 / 4, s0=2.0, J=40)
    assert relerr(W12, W12r) < TOL
    WCT, aWCT, *_ = pycwt.wct(a, b, 1.0, dj=1 / 4, s0=2.0, J=40, sig=False)
    WCTr, aWCTr, *_ = orc.wct(a, b, 1.0, dj=1 / 4, s0=2.0, J=40, sig=False)
    assert relerr(WCT, WCTr) < TOL
    # the angle is discontinuous at +-pi: compare on the unit circle
    assert np.abs(np.exp(1j * aWCT) - np.exp(1j * aWCTr)).max() < 1e-9


def test_xwt_wct_config4_properties(pycwt):
    """SURVEY 8d config 4 size (two N=2^18 series, 145 scales): size-independent checks.
    xwt(y, y) is real and equals |W|^2; 0 <= WCT <= 1; WCT(y, y) == 1."""
    n = 2 ** 18
    a, b = noisy_chirps(n, 0)
    eng = pycwt.default_engine()
    sj = 2.0 * 2 ** (np.arange(0, 145, 12) / 12.0)   # 13 of the 145 scales, all classes
    W12 = eng.xwt(a, a, 1.0, sj, 0, 6.0)
    Wa = eng.cwt(a, 1.0, sj, 0, 6.0)
    assert np.abs(W12.imag).max() <= 1e-12 * np.abs(W12.real).max()
    assert relerr(W12.real, np.abs(Wa) ** 2) < 1e-13
    WCT, aWCT = eng.wct(a, b, 1.0, 1 / 12, sj, 0, 6.0, 14)
    assert np.isfinite(WCT).all() and WCT.min() >= 0 and WCT.max() <= 1 + 1e-9
    WCTs, _ = eng.wct(a, a, 1.0, 1 / 12, sj, 0, 6.0, 14)
    assert np.abs(WCTs - 1).max() < 1e-9
    # spot parity of 2 rows of W12 against the oracle (CPU inverse FFTs of full length)
    om = 2 * np.pi * np.fft.fftfreq(n, 1.0)
    rows = [0, 12]
    filt = np.sqrt(sj[rows, None] * om[1] * n) * orc.Morlet(6).psi_ft(sj[rows, None] * om)
    Wa_r = np.fft.ifft(np.fft.fft(a) * filt, axis=1)
    Wb_r = np.fft.ifft(np.fft.fft(b) * filt, axis=1)
    W12ab = eng.xwt(a, b, 1.0, sj, 0, 6.0)
    assert relerr(W12ab[rows], Wa_r * Wb_r.conj()) < TOL


def test_cwt_batch_channels(pycwt):
    """Config 5 flavour: independent float32 channels, fp32 engine, power spectra + W.  All
    channels of a chunk share every kernel launch (one descriptor row per channel x scale)."""
    rs = np.random.RandomState(1)
    X = rs.randn(6, 4096).astype(np.float32)
    sj = 2.0 * 2 ** (np.arange(0, 40) / 4.0)
    eng = pycwt.default_engine()
    power, W = eng.cwt_batch(X, 1.0, sj, 0, 6.0, precision=1, want_w=True)
    for ch in (0, 3, 5):
        Wr = orc.cwt(X[ch].astype(np.float64), 1.0, wavelet=orc.Morlet(6),
                     freqs=1 / (orc.Morlet(6).flambda() * sj))[0]
        assert relerr(W[ch], Wr) < 1e-5
        assert relerr(power[ch], (np.abs(Wr) ** 2).mean(axis=1)) < 1e-5


def test_cwt_batch_fp64_two_kernel_sizes_and_device_variant(pycwt):
    """Batched fp64 channels long enough for the direct and two-kernel classes; the batched
    result must equal the per-channel transform bit for bit (same kernels, same order)."""
    rs = np.random.RandomState(2)
    n0, nch = 2 ** 15, 3
    X = rs.randn(nch, n0)
    sj = 2.0 * 2 ** (np.arange(0, 26) / 2.0)
    eng = pycwt.default_engine()
    power, W = eng.cwt_batch(X, 1.0, sj, 0, 6.0, precision=0, want_w=True)
    for ch in range(nch):
        Wc = eng.cwt(X[ch], 1.0, sj, 0, 6.0)
        assert np.array_equal(W[ch], Wc)
    Wr = orc.cwt(X[1], 1.0, wavelet=orc.Morlet(6), freqs=1 / (orc.Morlet(6).flambda() * sj))[0]
    assert relerr(W[1], Wr) < TOL
    # device-resident entry point
    d = eng.dev_alloc(X.nbytes)
    eng.h2d(d, X)
    p2 = eng.cwt_batch_dev(d, nch, n0, 1.0, sj, 0, 6.0, precision=0, want_power=True)
    eng.dev_free(d)
    assert relerr(p2, power) < 1e-13
    W2 = eng.get_w(nch * len(sj), n0).reshape(nch, len(sj), n0)
    assert np.array_equal(W2, W)


def test_sharded_wct_significance_matches_oracle(pycwt):
    """The sharded Monte-Carlo driver (here with one rank) fed to the oracle draw for
    draw: identical significance levels, and identical to the sum of two disjoint shards."""
    from pycwt_b200 import distributed as D
    args = (0.2, 0.1, 1.0, 0.5, 2.0, 10)
    sig = D.wct_significance_sharded(*args, significance_level=0.95, wavelet='morlet', mc_count=6, seed=5)

    class Replay(object):
        """rng stand-in: hands the oracle the same white-noise columns, in its call order"""
        def __init__(self, N):
            tau = lambda al: int(np.ceil(-2 / np.log(abs(al))))
            self.q = [np.zeros(N + tau(0.2))]          # the reference's discarded set-up draw
            for i in range(6):
                rs = np.random.RandomState([5, i])
                self.q += [rs.randn(N + tau(0.2)), rs.randn(N + tau(0.1))]

        def randn(self, n, one):
            a = self.q.pop(0)
            assert a.size == n and one == 1
            return a.reshape(n, 1)

    N = int(np.ceil(2.0 * 2 ** (10 * 0.5) / 1.0 * 6))
    ref = orc.wct_significance(*args, significance_level=0.95, wavelet="morlet", mc_count=6, rng=Replay(N))
    assert np.array_equal(np.isnan(sig), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.abs(sig[ok] - ref[ok]).max() < 1e-12


def test_seeded_monte_carlo_device_rng(pycwt):
    """cwtb_wct_mc_seeded on the GPU: the same checks as on the emulation build, plus the public
    `wct_significance(..., seed=)` entry."""
    from test_emu_kernels import check_seeded_monte_carlo
    check_seeded_monte_carlo(pycwt.default_engine())
    a = pycwt.wct_significance(0.2, 0.4, 1.0, 0.5, 2.0, 8, mc_count=30, progress=False, cache=False, seed=1)
    b = pycwt.wct_significance(0.2, 0.4, 1.0, 0.5, 2.0, 8, mc_count=30, progress=False, cache=False, seed=1)
    c = pycwt.wct_significance(0.2, 0.4, 1.0, 0.5, 2.0, 8, mc_count=30, progress=False, cache=False, seed=2)
    assert np.array_equal(a, b, equal_nan=True) and not np.array_equal(a, c, equal_nan=True)


def test_generic_smoothing_for_paul_and_dog_gpu(pycwt):
    from test_emu_kernels import check_generic_smoothing
    check_generic_smoothing(pycwt)


def test_batch_pipeline_equals_synchronous_chunks(monkeypatch):
    """cwtb_cwt_batch overlaps the input copy of chunk k+1 with the kernels of chunk k and reads the
    spectra back once (engine.cu: cwt_batch_pipelined); the spectra must equal the ones of the
    chunk-after-chunk path (same kernels: differences only from the order of the row-sum atomics)."""
    from pycwt_b200 import _engine
    rs = np.random.RandomState(8)
    X = rs.randn(21, 4096).astype(np.float32)
    sj = 2.0 * 2 ** (np.arange(24) / 3.0)
    monkeypatch.setenv("CWTB_BATCH_MB", "4")           # 5 channels per chunk: four full chunks and a short last one
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("CWTB_BATCH_PIPELINE", mode)
        eng = _engine.Engine(0)
        try:
            out[mode], _ = eng.cwt_batch(X, 1.0, sj, _engine.MORLET, 6.0, _engine.F32, want_power=True)
            out[mode] = np.array(out[mode])
        finally:
            eng.close()
    assert out["1"].shape == (21, 24)
    assert np.allclose(out["1"], out["0"], rtol=1e-12, atol=0)
    ref = np.stack([(np.abs(orc.cwt(x.astype(np.float64), 1.0, wavelet=orc.Morlet(6),
                                    freqs=1 / (orc.Morlet(6).flambda() * sj))[0]) ** 2).mean(axis=1) for x in X[:3]])
    assert np.allclose(out["1"][:3], ref, rtol=2e-5)


def test_wct_scale_sharded_single_process_equals_wct(pycwt):
    """distributed.wct_scale_sharded without a communicator is the whole ladder: the same arrays as wct()."""
    from pycwt_b200 import distributed as D
    rs = np.random.RandomState(9)
    t = np.arange(5000)
    y1 = np.sin(2 * np.pi * t / 40.0) + rs.randn(5000)
    y2 = np.sin(2 * np.pi * t / 40.0 + 1.0) + rs.randn(5000)
    W, A, coi, freq, _ = pycwt.wct(y1, y2, 1.0, dj=1 / 6, s0=2.0, J=40, sig=False)
    lo, hi, Ws, As, mean, freqs = D.wct_scale_sharded(y1, y2, 1.0, dj=1 / 6, s0=2.0, J=40)
    assert (lo, hi) == (0, 41) and np.array_equal(Ws, W) and np.array_equal(As, A)
    assert np.array_equal(freqs, freq) and np.allclose(mean, W.mean(axis=1), rtol=0, atol=0)
