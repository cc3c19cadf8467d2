"""The oracle's canonical transcendental definitions (oracle/gsmo_math.h) must stay close to the
true functions -- they stand in for MSL built-ins whose bits are unspecified."""
import numpy as np


def _ulp_err(got, ref64):
    ref32 = ref64.astype(np.float32)
    ulp = np.spacing(np.abs(ref32)).astype(np.float64)
    return np.abs(got.astype(np.float64) - ref64) / np.maximum(ulp, 1e-45)


def test_sincos(oracle):
    x = np.linspace(0.0, np.pi, 200_001).astype(np.float32)
    s, c = oracle.probe_sincos(x)
    assert np.max(np.abs(s - np.sin(x.astype(np.float64)))) < 2.5e-7
    assert np.max(np.abs(c - np.cos(x.astype(np.float64)))) < 2.5e-7
    x = np.linspace(-20.0, 20.0, 100_001).astype(np.float32)
    s, c = oracle.probe_sincos(x)
    assert np.max(np.abs(s - np.sin(x.astype(np.float64)))) < 1e-6
    assert np.max(np.abs(c - np.cos(x.astype(np.float64)))) < 1e-6
    # every packed theta (GaussianShared.h:442-444)
    th = (np.arange(65536, dtype=np.float32) * np.float32(np.float32(np.pi) / np.float32(65535.0)))
    s, c = oracle.probe_sincos(th)
    assert np.max(np.abs(s * s + c * c - 1.0)) < 5e-7


def test_log(oracle):
    x = np.exp(np.linspace(np.log(1e-6), np.log(1e6), 200_001)).astype(np.float32)
    y = oracle.probe_log(x)
    assert np.max(_ulp_err(y, np.log(x.astype(np.float64)))[np.abs(np.log(x)) > 0.1]) < 2.0
    assert np.max(np.abs(y - np.log(x.astype(np.float64)))) < 2e-6
    assert oracle.probe_log(np.array([1.0], np.float32))[0] == 0.0


def test_atan2(oracle):
    a = np.linspace(-np.pi, np.pi, 100_001)
    for r in (1.0, 1e-3, 1e3):
        y, x = (r * np.sin(a)).astype(np.float32), (r * np.cos(a)).astype(np.float32)
        got = oracle.probe_atan2(y, x)
        ref = np.arctan2(y.astype(np.float64), x.astype(np.float64))
        assert np.max(np.abs(got - ref)) < 5e-7
    assert oracle.probe_atan2(np.array([0.0, 1.0, -1.0], np.float32), np.array([0.0, 0.0, 0.0], np.float32)).tolist() \
        == [0.0, np.float32(np.pi / 2), np.float32(-np.pi / 2)]


def test_powr_srgb_range(oracle):
    c = np.linspace(0.04045, 1.0, 100_001).astype(np.float32)
    x = ((c + np.float32(0.055)) / np.float32(1.055)).astype(np.float32)
    got = oracle.probe_powr(x, 2.4)
    ref = x.astype(np.float64) ** np.float64(np.float32(2.4))
    assert np.max(np.abs(got - ref) / ref) < 1e-6


def test_hexp_exhaustive(oracle):
    bits = np.arange(65536, dtype=np.uint16)
    got = oracle.probe_hexp(bits)
    x = bits.view(np.float16).astype(np.float64)
    finite = np.isfinite(x)
    with np.errstate(over="ignore"):
        ref = np.exp(x)
    gotf = got.view(np.float16).astype(np.float64)
    # NaN -> NaN
    assert np.all(np.isnan(gotf[np.isnan(x)]))
    # faithful: within one half-ulp-of-result of the true value, overflow/underflow where it must
    ref16 = ref[finite].astype(np.float16)
    lo = np.nextafter(ref16, np.float16(-np.inf)).astype(np.float64)
    hi = np.nextafter(ref16, np.float16(np.inf)).astype(np.float64)
    g = gotf[finite]
    assert np.all((g >= lo) & (g <= hi))
    # and almost always the correctly rounded one
    assert np.mean(g == ref16.astype(np.float64)) > 0.999
    # exact anchors
    assert got[0x0000] == 0x3C00 and got[0x8000] == 0x3C00  # exp(+-0) = 1
    assert got[0xFC00] == 0x0000 and got[0x7C00] == 0x7C00  # exp(-inf)=0, exp(inf)=inf


def test_half_conversion(oracle):
    x = np.random.default_rng(0).normal(0, 100, 100_000).astype(np.float32)
    x = np.concatenate([x, np.array([-1e10, 1e10, 65504, 65520, 65519.9, 6e-8, 3e-8, 2.9e-8, 0.0, -0.0], np.float32)])
    assert np.array_equal(oracle.probe_f2h(x), x.astype(np.float16).view(np.uint16))
