"""Survival-time utilities and lookup tables: Rpexp, Rpexp_gamma
(inst/include/microsimulation.h:527-704), NumericInterpolate and Table (inst/include/rcpp_table.h:52-324).
The reference prints no expected values for them; the oracle's restatement is checked against an independent numpy
rendering of the dev script's R implementation rpexp2 (test/test_microsimulation.R:268-295) and against closed
forms, the device functions (compiled for the host) against the oracle, and the GPU against the oracle."""
import ctypes as C

import numpy as np
import pytest

H3, T3 = np.array([1.0, 2.0, 3.0]), np.array([0.0, 1.0, 2.0])


def rpexp2_numpy(e, rate, t, start):
    """rpexp2 of the dev script with the exponential draws e = -log(u) given: H = c(0, cumsum(head(rate,-1)*diff(t)));
    Hstart = H(start); i = findInterval(e + Hstart, H); t[i] + (e + Hstart - H[i]) / rate[i]."""
    H = np.r_[0.0, np.cumsum(rate[:-1] * np.diff(t))]
    i0 = np.maximum(np.searchsorted(t, start, side="right") - 1, 0)
    Hs = np.where(start == 0.0, 0.0, H[i0] + (start - t[i0]) * rate[i0])
    v = e + Hs
    i = np.maximum(np.searchsorted(H, v, side="right") - 1, 0)
    return t[i] + (v - H[i]) / rate[i]


def test_oracle_rpexp_against_the_dev_scripts_r_version(oracle):
    rs = np.random.RandomState(1)
    u = rs.uniform(1e-9, 1, 5000)
    for rate, t, start in ((H3, T3, 0.0), (H3, T3 + 1, 1.7), (np.full(11, 0.2), np.arange(11.0), 0.0), (H3, T3, 4.0)):
        got = oracle.rpexp(0, rate, t, u, start)
        np.testing.assert_allclose(got, rpexp2_numpy(-np.log(u), rate, t, np.full(u.size, start)), rtol=1e-13)
        assert (got >= start).all()
    # constant hazard: an exponential; S(rand(u)) = u in general
    np.testing.assert_allclose(oracle.rpexp(0, [0.5], [0.0], u), -np.log(u) / 0.5, rtol=1e-14)
    t = oracle.rpexp(0, H3, T3, u)
    np.testing.assert_allclose(1 - oracle.rpexp(3, H3, T3, t), u, rtol=1e-10, atol=1e-15)
    # zero hazards: a leading zero is skipped (the dev script's "allows for rate = 0"); a zero tail cannot be inverted
    assert (oracle.rpexp(0, [0.0, 1.0, 2.0], T3, u) >= 1.0).all()
    with pytest.raises(RuntimeError):
        oracle.rpexp(0, [0.01, 0.0], [0.0, 1.0], [1e-9])


def test_oracle_rpexp_gamma_and_interpolate_closed_forms(oracle):
    rs = np.random.RandomState(2)
    u, Z = rs.uniform(1e-6, 1, 4000), rs.gamma(2.0, 0.5, 4000)
    mu, t, theta = np.array([0.01, 0.03, 0.08]), np.array([0.0, 40.0, 70.0]), 0.7
    x = oracle.rpexp_gamma(0, theta, mu, t, u, Z)
    np.testing.assert_allclose(np.exp(-oracle.rpexp_gamma(1, theta, mu, t, x, Z)), u, rtol=1e-9)  # S(rand(u, Z) | Z) = u
    xs, ys = np.array([0.0, 1.0, 3.0, 4.0]), np.array([2.0, 4.0, 5.0, 9.0])
    v = np.array([-1.0, 0.0, 0.5, 1.0, 2.0, 3.5, 4.0, 6.0])
    np.testing.assert_allclose(oracle.interpolate(0, xs, ys, v), [2, 2, 3, 4, 4.5, 7, 9, 17])
    np.testing.assert_allclose(oracle.interpolate(1, xs, ys, v), [2, 2, 2, 2, 4, 5, 5, 9])
    np.testing.assert_allclose(oracle.interpolate(2, xs, ys, oracle.interpolate(0, xs, ys, v[2:])), v[2:])
    np.testing.assert_allclose(oracle.interpolate(3, xs, -ys, -oracle.interpolate(0, xs, ys, v[2:])), v[2:])
    rows = np.array([[50, 0, 1.0], [50, 1, 2.0], [60, 0, 3.0], [70, 1, 4.0]])
    keys = np.array([[40, 0], [55, 1], [65, 0.5], [65, 1], [99, 1], [70, 0]])
    np.testing.assert_allclose(oracle.table_lookup(rows, keys), [1, 2, 3, 0, 4, 0])  # (60,1) and (70,0) are not rows: 0


def _emu(fn, *args):
    assert fn(*args) == 0


def test_device_lookups_on_host_vs_oracle(oracle, emu):
    rs = np.random.RandomState(3)
    n = 4000
    dp = C.POINTER(C.c_double)
    ptr = lambda a: a.ctypes.data_as(dp)
    u, start = rs.uniform(1e-9, 1, n), rs.uniform(0, 3, n)
    h, t = np.array([0.0, 0.3, 0.1, 0.9, 2.0]), np.array([0.0, 0.5, 1.0, 2.5, 4.0])
    out = np.empty(n)
    for mode, b in ((0, start), (1, None), (2, None), (3, None), (4, None)):
        a = u if mode == 0 else rs.uniform(0, 6, n)
        _emu(emu.emu_rpexp, C.c_int(mode), C.c_int(h.size), ptr(h), ptr(t), C.c_int64(n), ptr(a), None if b is None else ptr(b), ptr(out))
        assert np.array_equal(out, oracle.rpexp(mode, h, t, a, b)), mode
    Z = rs.gamma(2.0, 0.5, n)
    mu, tg, theta = np.array([0.01, 0.03, 0.08, 0.2]), np.array([0.0, 40.0, 70.0, 85.0]), 1.3
    for mode in (0, 1, 2, 3):
        a = u if mode == 0 else rs.uniform(0, 100, n)
        st = rs.uniform(0, 50, n) if mode == 0 else None
        _emu(emu.emu_rpexp_gamma, C.c_int(mode), C.c_int(mu.size), C.c_double(theta), ptr(mu), ptr(tg), C.c_int64(n), ptr(a), ptr(Z),
             None if st is None else ptr(st), ptr(out))
        np.testing.assert_array_equal(out, oracle.rpexp_gamma(mode, theta, mu, tg, a, Z, st))
    xs = np.sort(rs.uniform(0, 10, 12))
    ys = np.cumsum(rs.uniform(0.1, 2, 12))
    for mode, y in ((0, ys), (1, ys), (2, ys), (3, -ys)):
        v = rs.uniform(-2, 12, n) if mode < 2 else rs.uniform(y.min() - 1, y.max() + 1, n)
        v[:12] = xs if mode < 2 else y  # exactly on the points
        _emu(emu.emu_interpolate, C.c_int(mode), C.c_int(12), ptr(xs), ptr(np.ascontiguousarray(y)), C.c_int64(n), ptr(v), ptr(out))
        np.testing.assert_array_equal(out, oracle.interpolate(mode, xs, y, v))


@pytest.mark.gpu
def test_lookups_on_gpu_vs_oracle(msim, oracle):
    from microsimulation_b200.lookups import NumericInterpolate, Rpexp, RpexpGamma, Table
    rs = np.random.RandomState(4)
    n = 300000
    u, start = rs.uniform(1e-9, 1, n), rs.uniform(0, 3, n)
    h, t = np.array([0.0, 0.3, 0.1, 0.9, 2.0]), np.array([0.0, 0.5, 1.0, 2.5, 4.0])
    R = Rpexp(h, t)
    np.testing.assert_allclose(R.rand(u, start), oracle.rpexp(0, h, t, u, start), rtol=1e-14)  # log(): device vs glibc
    s = rs.uniform(0, 6, n)
    assert np.array_equal(R.H(s), oracle.rpexp(1, h, t, s)) and np.array_equal(R.hazard(s), oracle.rpexp(2, h, t, s))
    np.testing.assert_allclose(R.p(s), oracle.rpexp(3, h, t, s), rtol=1e-13, atol=1e-16)
    np.testing.assert_allclose(R.d(s), oracle.rpexp(4, h, t, s), rtol=1e-13)
    with pytest.raises(msim.MsimError):
        Rpexp([0.01, 0.0], [0.0, 1.0]).rand([1e-9])
    Z = rs.gamma(2.0, 0.5, n)
    mu, tg, theta = np.array([0.01, 0.03, 0.08, 0.2]), np.array([0.0, 40.0, 70.0, 85.0]), 1.3
    G = RpexpGamma(theta, mu, tg)
    st = rs.uniform(0, 50, n)
    np.testing.assert_allclose(G.rand(u, Z, st), oracle.rpexp_gamma(0, theta, mu, tg, u, Z, st), rtol=1e-12)
    a = rs.uniform(0, 100, n)
    np.testing.assert_allclose(G.H(a, Z), oracle.rpexp_gamma(1, theta, mu, tg, a, Z), rtol=1e-13)
    np.testing.assert_allclose(G.hazard(a, Z), oracle.rpexp_gamma(2, theta, mu, tg, a, Z), rtol=1e-13)
    np.testing.assert_allclose(G.p(a, Z), oracle.rpexp_gamma(3, theta, mu, tg, a, Z), rtol=1e-12, atol=1e-16)
    xs = np.sort(rs.uniform(0, 10, 12))
    ys = np.cumsum(rs.uniform(0.1, 2, 12))
    F = NumericInterpolate(xs, ys)
    v = np.r_[xs, rs.uniform(-2, 12, n)]
    assert np.array_equal(F.approx(v), oracle.interpolate(0, xs, ys, v)) and np.array_equal(F(v), oracle.interpolate(1, xs, ys, v))
    w = np.r_[ys, rs.uniform(ys.min() - 1, ys.max() + 1, n)]
    assert np.array_equal(F.invert(w), oracle.interpolate(2, xs, ys, w))
    assert np.array_equal(NumericInterpolate(xs, -ys).invert_decreasing(-w), oracle.interpolate(3, xs, -ys, -w))
    # a 3-key table: age x year x grade, a third of the combinations missing
    ages, years, grades = np.arange(40, 90, 5.0), np.arange(1990, 2021, 10.0), np.array([0.0, 1.0, 2.0])
    rows = np.array([[a_, y_, g_, rs.rand()] for a_ in ages for y_ in years for g_ in grades if rs.rand() > 0.33])
    keys = np.c_[rs.uniform(30, 100, n), rs.uniform(1980, 2030, n), rs.randint(0, 3, n)]
    assert np.array_equal(Table(rows)(keys), oracle.table_lookup(rows, keys))
