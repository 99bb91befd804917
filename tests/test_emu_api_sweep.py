"""Public-API sweeps on the host-emulation build (CPU; see tests/test_emu_kernels.py for what
that build is): every return value of `cwt` for tiny and awkward lengths under both
transform-length policies, unordered custom frequencies, list / float32 input, both `icwt`
orientations, special signals and the reference's exception types -- against the oracle."""
import warnings

import numpy as np
import pytest
from scipy.signal import lfilter

from conftest import ROOT
from oracle import cwt_oracle as orc


@pytest.fixture(scope="module")
def api():
    import os
    import pycwt_b200 as pycwt
    from pycwt_b200 import build as _build, _engine
    eng = _engine.Engine(0, lib_path=_build.build_emulation(os.path.join(ROOT, "tests", "_emu")))
    saved = _engine.default_engine
    _engine.default_engine = lambda *a, **k: eng
    yield pycwt
    _engine.default_engine = saved
    eng.close()


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("pad", [True, False])
def test_every_output_for_tiny_and_awkward_lengths(api, pad):
    from pycwt_b200 import helpers
    rs = np.random.RandomState(123)
    helpers.set_fft_padding(pad)
    orc.PAD_NEXT_POW2 = pad
    try:
        for n0 in list(range(2, 41)) + [63, 64, 65, 100, 127, 128, 129]:
            x = rs.randn(n0)
            for mo, mr in ((api.Morlet(6), orc.Morlet(6)), (api.Paul(4), orc.Paul(4)),
                           (api.DOG(2), orc.DOG(2)), (api.DOG(3), orc.DOG(3))):
                with np.errstate(all="ignore"), warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    try:
                        r = orc.cwt(x, 0.5, dj=0.5, wavelet=mr)
                    except Exception:
                        with pytest.raises(Exception):
                            api.cwt(x, 0.5, dj=0.5, wavelet=mo)
                        continue
                    g = api.cwt(x, 0.5, dj=0.5, wavelet=mo)
                tag = (pad, n0, type(mo).__name__)
                assert r[0].shape == g[0].shape, tag
                fin = np.isfinite(r[0])
                assert np.array_equal(fin, np.isfinite(g[0])), tag
                if fin.any():
                    assert rel(g[0][fin], r[0][fin]) < 1e-10, tag
                for k in (1, 2, 3, 5):      # sj, freqs, coi, fftfreqs: bit-identical host outputs
                    assert np.array_equal(r[k], g[k]), (tag, k)
                if r[4].size:
                    assert np.abs(r[4] - g[4]).max() <= 1e-12 * max(np.abs(r[4]).max(), 1e-300), tag
    finally:
        helpers.set_fft_padding(True)
        orc.PAD_NEXT_POW2 = True


def test_api_corners(api):
    warnings.filterwarnings("ignore")
    rs = np.random.RandomState(7)
    for it in range(12):
        n = int(2 ** rs.uniform(5, 13))
        dt = float(10 ** rs.uniform(-1, 1))
        x = rs.randn(n).cumsum()
        fam = rs.randint(3)
        mo, mr = [(api.Morlet(6), orc.Morlet(6)), (api.Paul(4), orc.Paul(4)), (api.DOG(2), orc.DOG(2))][fam]
        name = ["morlet", "paul", "dog"][fam]
        fr = rs.uniform(2 / (n * dt), 0.4 / dt, size=rs.randint(1, 30))     # unordered, non-geometric
        g, r = api.cwt(x, dt, wavelet=mo, freqs=fr), orc.cwt(x, dt, wavelet=mr, freqs=fr)
        assert g[0].shape == r[0].shape and rel(g[0], r[0]) < 1e-10
        g2, r2 = api.cwt(list(x), dt, dj=0.5, wavelet=name), orc.cwt(list(x), dt, dj=0.5, wavelet=name)
        assert rel(g2[0], r2[0]) < 1e-10
        if mr.cdelta != -1:
            assert rel(api.icwt(g2[0], g2[1], dt, 0.5, name), orc.icwt(r2[0], r2[1], dt, 0.5, name)) < 1e-10
            if g2[0].shape[0] != g2[0].shape[1]:     # (N, S) input: the reference still sums axis 0
                a = api.icwt(g2[0].T.copy(), g2[1], dt, 0.5, name)
                b = orc.icwt(r2[0].T.copy(), r2[1], dt, 0.5, name)
                assert a.shape == b.shape and rel(a, b) < 1e-10
        xs = lfilter([1], [1, -0.6], rs.randn(n))
        y2 = np.roll(xs, 5) + rs.randn(n)
        a = api.xwt(xs, y2, dt, dj=0.5, wavelet=name, normalize=False)
        b = orc.xwt(xs, y2, dt, dj=0.5, wavelet=name, normalize=False)
        assert rel(a[0], b[0]) < 1e-10 and np.allclose(a[3], b[3], rtol=1e-12)
        x32 = x.astype(np.float32)
        assert rel(api.cwt(x32, dt, dj=0.5, wavelet=name)[0],
                   orc.cwt(x32.astype(np.float64), dt, dj=0.5, wavelet=name)[0]) < 1e-10


def test_special_signals_and_exception_types(api):
    for x in (np.zeros(100), np.ones(257), np.r_[np.zeros(50), 1e300, np.zeros(50)],
              np.r_[1.0, np.inf, np.zeros(30)], np.r_[np.nan, np.ones(20)]):
        with np.errstate(all="ignore"):
            g, r = api.cwt(x, 1.0, dj=0.5)[0], orc.cwt(x, 1.0, dj=0.5)[0]
        assert np.array_equal(np.isfinite(g), np.isfinite(r))
        f = np.isfinite(r)
        if f.any() and np.abs(r[f]).max() > 0:
            assert rel(g[f], r[f]) < 1e-10
    with pytest.raises(KeyError):
        api.cwt(np.ones(64), 1.0, wavelet="nope")
    with pytest.raises(AttributeError):
        api.wct(np.ones(64), np.ones(64), 1.0, wavelet="paul", sig=False)
    with pytest.raises(Warning):
        api.icwt(np.ones((5, 7)), np.ones(3), 1.0)


def test_resident_handle_random(api):
    """Derived products of the device-resident handle for random lengths, families and both
    length policies (incl. odd un-padded lengths where Paul rows are dropped) vs the oracle."""
    from pycwt_b200 import helpers
    warnings.filterwarnings("ignore")
    rs = np.random.RandomState(17)
    checked = 0
    try:
        for it in range(40):
            pad = bool(rs.rand() < 0.5)
            helpers.set_fft_padding(pad)
            orc.PAD_NEXT_POW2 = pad
            n = int(2 ** rs.uniform(3, 12))
            dt = float(10 ** rs.uniform(-1, 1))
            x = rs.randn(n)
            fam = rs.randint(3)
            mo, mr = [(api.Morlet(6), orc.Morlet(6)), (api.Paul(4), orc.Paul(4)), (api.DOG(2), orc.DOG(2))][fam]
            dj = float(rs.choice([0.5, 0.25, 0.125]))
            try:
                with np.errstate(all="ignore"):
                    W, sj, fr, coi, _, _ = orc.cwt(x, dt, dj=dj, wavelet=mr)
            except Exception:
                continue
            if W.size == 0 or not np.isfinite(W).all():
                continue
            r = api.cwt_resident(x, dt, dj, wavelet=mo)
            tag = (it, pad, n, fam, dj)
            P = np.abs(W) ** 2
            assert r.shape == W.shape, tag
            assert np.array_equal(r.scales, sj) and np.array_equal(r.coi, coi), tag
            assert rel(r.power(), P) < 1e-10 and rel(r.power(rectify=True), P / sj[:, None]) < 1e-10, tag
            assert rel(r.global_power(), P.mean(axis=1)) < 1e-10, tag
            per = 1 / fr
            inside = per[:, None] <= coi[None, :]
            g = r.global_power(inside_coi=True)
            with np.errstate(all="ignore"):
                ref = np.where(inside.any(axis=1), (P * inside).sum(axis=1) / inside.sum(axis=1), np.nan)
            assert np.array_equal(np.isnan(g), np.isnan(ref)), tag
            ok = ~np.isnan(ref)
            if ok.any():
                assert rel(g[ok], ref[ok]) < 1e-10, tag
            lo, hi = np.percentile(per, [20, 70])
            sel = (per >= lo) & (per < hi)
            want = 2.0 * dj * dt / mr.cdelta * (P / sj[:, None])[sel].sum(axis=0)
            assert np.abs(r.scale_avg_power(lo, hi, variance=2.0) - want).max() <= 1e-10 * max(np.abs(want).max(), 1e-300), tag
            assert rel(r.icwt(), orc.icwt(W, sj, dt, dj, mr)) < 1e-10, tag
            assert rel(r.wave(), W) < 1e-10, tag
            checked += 1
    finally:
        helpers.set_fft_padding(True)
        orc.PAD_NEXT_POW2 = True
    assert checked > 25


class _Duck(object):
    """A mother wavelet the engine does not recognise: forces the host-table path."""
    def __init__(self, m):
        self.m = m

    def __getattr__(self, k):
        if k == '_engine_spec':
            raise AttributeError(k)
        return getattr(self.m, k)


def test_table_path_and_fp32_engine_random(api, monkeypatch):
    from pycwt_b200 import helpers
    warnings.filterwarnings("ignore")
    rs = np.random.RandomState(5)
    try:
        for it in range(40):
            pad = bool(rs.rand() < 0.6)
            helpers.set_fft_padding(pad)
            orc.PAD_NEXT_POW2 = pad
            n = int(2 ** rs.uniform(1.2, 11))
            dt = float(10 ** rs.uniform(-1, 1))
            x = rs.randn(n)
            fam = rs.randint(3)
            mo, mr = [(api.Morlet(6), orc.Morlet(6)), (api.Paul(4), orc.Paul(4)), (api.DOG(2), orc.DOG(2))][fam]
            try:
                with np.errstate(all="ignore"):
                    r = orc.cwt(x, dt, dj=0.5, wavelet=mr)
            except Exception:
                continue
            if r[0].size == 0:
                continue
            fin = np.isfinite(r[0])
            tag = (it, pad, n, fam)
            g = api.cwt(x, dt, dj=0.5, wavelet=_Duck(mo))
            assert g[0].shape == r[0].shape, tag
            if fin.any():
                assert rel(g[0][fin], r[0][fin]) < 1e-10, tag
            x32 = x.astype(np.float32)
            r32 = orc.cwt(x32.astype(np.float64), dt, dj=0.5, wavelet=mr)
            f2 = np.isfinite(r32[0])
            monkeypatch.setenv("CWTB_PRECISION", "fp32")
            try:
                g32 = api.cwt(x32, dt, dj=0.5, wavelet=mo)
                gd = api.cwt(x32, dt, dj=0.5, wavelet=_Duck(mo))
            finally:
                monkeypatch.delenv("CWTB_PRECISION")
            assert g32[0].shape == r32[0].shape and g32[0].dtype == np.complex128, tag
            if f2.any() and np.abs(r32[0][f2]).max() > 1e-30:
                assert rel(g32[0][f2], r32[0][f2]) < 3e-5, tag
                assert rel(gd[0][f2], r32[0][f2]) < 3e-5, tag
    finally:
        helpers.set_fft_padding(True)
        orc.PAD_NEXT_POW2 = True


def test_coherence_api_random(api):
    from pycwt_b200 import helpers
    warnings.filterwarnings("ignore")
    rs = np.random.RandomState(77)
    try:
        for it in range(16):
            pad = bool(rs.rand() < 0.5)
            helpers.set_fft_padding(pad)
            orc.PAD_NEXT_POW2 = pad
            n = int(2 ** rs.uniform(5, 11))
            dt = float(10 ** rs.uniform(-1, 1))
            y1 = lfilter([1], [1, -0.5], rs.randn(n))
            y2 = np.roll(y1, 2) + 0.7 * rs.randn(n)
            dj = float(rs.choice([0.5, 0.25, 1 / 6, 1 / 12]))
            kw = dict(dj=dj, sig=False, wavelet='morlet', normalize=bool(rs.rand() < 0.7))
            if rs.rand() < 0.5:
                kw.update(s0=2 * dt, J=int(rs.randint(3, int(np.log2(n) / dj))))
            b = orc.wct(y1, y2, dt, **kw)
            a = api.wct(y1, y2, dt, **kw)
            tag = (it, pad, n, dj)
            assert a[0].shape == b[0].shape, tag
            assert np.abs(a[0] - b[0]).max() < 1e-8, tag
            assert np.abs(np.exp(1j * a[1]) - np.exp(1j * b[1])).max() < 1e-7, tag
            assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]), tag
    finally:
        helpers.set_fft_padding(True)
        orc.PAD_NEXT_POW2 = True


def test_concurrent_cwt_calls_on_the_default_engine(api):
    """Several Python threads transform signals of different lengths on the SHARED engine: every
    result is the one of its own call (the engine lock spans policy, transform, fetch and spectrum;
    without it the resident job of one call was replaced under another's fetch)."""
    import threading
    rs = np.random.RandomState(11)
    sigs = [rs.randn(n) for n in (4096, 300, 20000, 1000)]
    want = [orc.cwt(x, 1.0, dj=0.5, wavelet=orc.Morlet(6)) for x in sigs]
    errs = []

    def work(i):
        try:
            for _ in range(6):
                g = api.cwt(sigs[i], 1.0, dj=0.5, wavelet=api.Morlet(6))
                assert g[0].shape == want[i][0].shape
                assert rel(g[0], want[i][0]) < 1e-10 and rel(g[4], want[i][4]) < 1e-12
        except Exception as exc:      # surfaced in the main thread
            errs.append((i, repr(exc)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(sigs))]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errs, errs


def test_complex_signal_and_unusable_frequencies(api):
    """A complex signal goes through the reference's FFT unchanged (linearity): W(re) + i W(im);
    custom frequencies of 0 (s = inf) or < 0 (NaN norm) give all-NaN rows, which the reference
    drops from W, sj and freqs (wavelet.py:111-115)."""
    rs = np.random.RandomState(4)
    z = rs.randn(300) + 1j * rs.randn(300)
    for mo, mr in ((api.Morlet(6), orc.Morlet(6)), (api.DOG(3), orc.DOG(3))):
        r = orc.cwt(z, 0.5, dj=0.5, wavelet=mr)
        g = api.cwt(z, 0.5, dj=0.5, wavelet=mo)
        assert g[0].shape == r[0].shape and rel(g[0], r[0]) < 1e-10 and rel(g[4], r[4]) < 1e-12
    x = rs.randn(500)
    for fr in ([0.5, 0.1, 0.0, 0.01], [0.5, 0.1, -0.05, 0.01]):
        fr = np.array(fr)
        for mo, mr in ((api.Morlet(6), orc.Morlet(6)), (api.Paul(4), orc.Paul(4)), (api.DOG(2), orc.DOG(2))):
            with np.errstate(all="ignore"), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r = orc.cwt(x, 1.0, wavelet=mr, freqs=fr)
                g = api.cwt(x, 1.0, wavelet=mo, freqs=fr)
            assert r[0].shape == (3, 500) and g[0].shape == (3, 500)
            assert np.array_equal(r[1], g[1]) and np.array_equal(r[2], g[2])
            assert rel(g[0], r[0]) < 1e-10


def test_resident_rejects_wavelets_the_engine_cannot_evaluate(api):
    class MyMorlet(api.Morlet):
        def psi_ft(self, f):
            return super().psi_ft(f) * 1.0

    x = np.random.RandomState(1).randn(256)
    for w in (MyMorlet(6), api.DOG(2.5)):
        with pytest.raises(TypeError, match="analytic families"):
            api.cwt_resident(x, 1.0, wavelet=w)
    # the plain call still works for them (host-evaluated response table)
    g = api.cwt(x, 1.0, dj=0.5, wavelet=MyMorlet(6))
    r = orc.cwt(x, 1.0, dj=0.5, wavelet=orc.Morlet(6))
    assert rel(g[0], r[0]) < 1e-10


def test_pinned_pool_is_bounded_and_trimmable(api):
    from pycwt_b200 import _engine
    eng = _engine.default_engine()
    old = eng.POOL_MAX_BYTES
    eng.POOL_MAX_BYTES = 3 << 20
    try:
        arrs = [eng.result_array((1 << 17,), np.complex128) for _ in range(4)]     # 2 MiB each, pinned
        assert eng._outstanding >= 4
        del arrs
        assert eng._pool_bytes <= eng.POOL_MAX_BYTES      # least recently released ones were retired
        eng.trim()
        assert eng._pool_bytes == 0 and not eng._pool and not eng._dead
    finally:
        eng.POOL_MAX_BYTES = old
