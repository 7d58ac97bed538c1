"""Generalised survival model (src/gsm.cpp, src/splines.cpp, inst/include/c_optim.h): PARITY UNPINNED in the reference
(no expected value anywhere in the tree), so the oracle's restatement is cross-checked against scipy and through
S(randU(u)) = u, and the device functions (compiled for the host) against the oracle."""
import ctypes as C

import numpy as np
import pytest

from gsm_fixture import make_model


def test_oracle_spline_basis_matches_scipy(oracle):
    from scipy.interpolate import BSpline
    model, q_const, knots, B = make_model(oracle, two_terms=False)
    m = model.cstruct()
    t = np.r_[[B[0]] * 4, knots, [B[1]] * 4]
    x = np.r_[np.linspace(B[0], B[1], 101), knots]
    got = oracle.spline_eval(m, 0, x, der=0, natural=False, width=6)
    for der in (0, 1, 2):
        got = oracle.spline_eval(m, 0, x[:-len(knots)][1:-1], der=der, natural=False, width=6)
        for j in range(6):  # bs::eval drops the first B-spline (no intercept): column j is B_{j+1}
            c = np.zeros(7)
            c[j + 1] = 1.0
            want = BSpline(t, c, 3).derivative(der)(x[:-len(knots)][1:-1]) if der else BSpline(t, c, 3)(x[:-len(knots)][1:-1])
            np.testing.assert_allclose(got[:, j], want, rtol=1e-10, atol=1e-10)
    # natural basis: second derivative vanishes at the boundary knots, linear outside them
    d2 = oracle.spline_eval(m, 0, B, der=2, natural=True, width=4)
    assert np.abs(d2).max() < 1e-9
    out = np.array([B[0] - 1.0, B[0] - 0.5, B[1] + 0.7, B[1] + 1.4])
    v = oracle.spline_eval(m, 0, out, der=0, natural=True, width=4)
    sl = oracle.spline_eval(m, 0, B, der=1, natural=True, width=4)
    at = oracle.spline_eval(m, 0, B, der=0, natural=True, width=4)
    np.testing.assert_allclose(v[:2], at[0] + np.outer(out[:2] - B[0], sl[0]), rtol=1e-12)
    np.testing.assert_allclose(v[2:], at[1] + np.outer(out[2:] - B[1], sl[1]), rtol=1e-12)
    # eta = etap + ns . gamma equals the scipy spline with coefficients q^T gamma
    y = np.linspace(B[0], B[1], 57)
    coef = np.r_[0.0, q_const.T @ model.terms[0]["gamma"]]
    np.testing.assert_allclose(oracle.gsm_eta(m, y), model.etap[0] + BSpline(t, coef, 3)(y), rtol=1e-11)


def test_oracle_inversion_solves_survival_equation(oracle):
    model, _, _, _ = make_model(oracle)
    m = model.cstruct()
    rs = np.random.RandomState(7)
    u = rs.uniform(0.001, 0.999, 300)
    for idx in (0, 1, 2):
        index = np.full(u.size, idx, dtype=np.int32)
        t = oracle.gsm_randU(m, u, index=index)
        S = np.exp(-np.exp(oracle.gsm_eta(m, np.log(t), index)))  # linkinv(eta(log t))
        # randU solves on the t < t0 branch only (operator() switches to eta0 - target0 at t0, gsm.cpp:29-31)
        inside = (t > model.tmin / 10 * 1.0001) & (t < model.t0 * 0.999)
        assert inside.sum() > 200
        np.testing.assert_allclose(S[inside], u[inside], rtol=2e-7)
        # left truncation: S(t)/S(tentry) = u
        te = np.full(u.size, 2.0)
        t2 = oracle.gsm_randU(m, u, tentry=te, index=index)
        S2 = np.exp(-np.exp(oracle.gsm_eta(m, np.log(t2), index))) / np.exp(-np.exp(oracle.gsm_eta(m, np.log(te), index)))
        ok = t2 < model.t0 * 0.999
        assert (t2 >= 2.0 - 1e-7).all()
        np.testing.assert_allclose(S2[ok], u[ok], rtol=2e-7)


def _emu_call(emu, model, mode, u, tentry=None, index=None, scale=10.0):
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty_like(u)
    m = model.cstruct()
    dp = C.POINTER(C.c_double)
    assert emu.emu_gsm(C.byref(m), C.c_int(mode), C.c_int64(u.size), u.ctypes.data_as(dp),
                       None if tentry is None else np.ascontiguousarray(tentry).ctypes.data_as(dp),
                       None if index is None else np.ascontiguousarray(index, dtype=np.int32).ctypes.data_as(C.POINTER(C.c_int32)),
                       C.c_double(scale), out.ctypes.data_as(dp)) == 0
    return out


@pytest.mark.parametrize("intercept", [0, 1])
def test_device_functions_on_host_vs_oracle(oracle, emu, intercept):
    """gsm.cuh compiled for the host against the restated reference: the term expanded into per-interval cubics on the
    host (gsm_term_build) against the reference's basis + projection evaluation — eta to rounding, roots to the
    solver's tolerance (Brent stops within 1e-8 on the log-time scale; a last-bit difference in a function value can
    end a search one step earlier or later)."""
    model, _, _, B = make_model(oracle, intercept=intercept)
    m = model.cstruct()
    rs = np.random.RandomState(11)
    n = 2000
    u = rs.uniform(1e-6, 1 - 1e-6, n)
    index = rs.randint(0, 3, n).astype(np.int32)
    y = rs.uniform(B[0] - 2, B[1] + 2, n)
    y = np.r_[y, B, B[0] - 1e-9, B[1] + 1e-9, np.asarray(model.terms[0]["knots"], dtype=float)]  # at and around every breakpoint
    index_y = np.r_[index, np.zeros(len(y) - n, dtype=np.int32)].astype(np.int32)
    np.testing.assert_allclose(_emu_call(emu, model, 2, y, index=index_y), oracle.gsm_eta(m, y, index_y), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np.log(_emu_call(emu, model, 0, u, index=index)), np.log(oracle.gsm_randU(m, u, index=index)), atol=2e-8)
    te = rs.uniform(0.2, 10, n)
    np.testing.assert_allclose(np.log(_emu_call(emu, model, 0, u, tentry=te, index=index)),
                               np.log(oracle.gsm_randU(m, u, tentry=te, index=index)), atol=2e-8)
    np.testing.assert_allclose(np.log(_emu_call(emu, model, 1, u, index=index)), np.log(oracle.gsm_randU0(m, u, index=index)), atol=2e-8)


@pytest.mark.gpu
def test_gsm_on_gpu_vs_oracle(msim, oracle):
    model, _, _, B = make_model(oracle)
    m = model.cstruct()
    rs = np.random.RandomState(5)
    n = 200000
    u = rs.uniform(1e-6, 1 - 1e-6, n)
    index = rs.randint(0, 3, n).astype(np.int32)
    y = rs.uniform(B[0] - 2, B[1] + 2, n)
    np.testing.assert_allclose(model.eta(y, index), oracle.gsm_eta(m, y, index), rtol=1e-12, atol=1e-12)
    # Brent stops within tol = 1e-8 of the root on the log-time scale; device libm may end a step earlier or later
    t_gpu, t_ref = model.randU(u, index=index), oracle.gsm_randU(m, u, index=index)
    np.testing.assert_allclose(np.log(t_gpu), np.log(t_ref), atol=2e-8)
    te = rs.uniform(0.2, 10, n)
    np.testing.assert_allclose(np.log(model.randU(u, tentry=te, index=index)), np.log(oracle.gsm_randU(m, u, tentry=te, index=index)),
                               atol=2e-8)
    np.testing.assert_allclose(np.log(model.randU0(u, index=index)), np.log(oracle.gsm_randU0(m, u, index=index)), atol=2e-8)
    # the property the inversion exists for: S(t) = u
    S = np.exp(-np.exp(model.eta(np.log(t_gpu), index)))
    inside = (t_gpu > model.tmin / 10 * 1.0001) & (t_gpu < model.t0 * 0.999)
    assert inside.sum() > n // 2
    np.testing.assert_allclose(S[inside], u[inside], rtol=2e-7)
