"""csrc/fastmath.cuh (exp, log and pow of the person-r hot path) on the host — the device runs the same source with
the same explicit fma — against mpmath at 60 digits and against glibc, in ulp.  The bar: each function within 1.05 ulp of
the true value (measured maxima: exp 1.005, log 1.0003, pow 1.05), so that an event time stays within 2 ulp of the reference's glibc value (numerical contract)."""
import ctypes as C

import numpy as np
import pytest

mpmath = pytest.importorskip("mpmath")


def _call(emu, mode, x, c=0.0):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    dp = C.POINTER(C.c_double)
    emu.emu_fastmath.restype = None
    emu.emu_fastmath(C.c_int(mode), C.c_int64(x.size), x.ctypes.data_as(dp), C.c_double(c), out.ctypes.data_as(dp))
    return out


def _ulp_err_vs_true(got, true_mp):
    """|got - true| in units of the last place of the true value"""
    worst = 0.0
    for g, t in zip(got, true_mp):
        tf = float(t)
        ulp = np.spacing(abs(tf))
        worst = max(worst, abs(float((mpmath.mpf(float(g)) - t) / mpmath.mpf(float(ulp)))))
    return worst


def _ulp_diff(a, b):
    return np.max(np.abs(a - b) / np.spacing(np.abs(b)))


# person-r's Weibull exponents: 1/exp(2.3525), 1/exp(1.0353), 1/exp(1.4404) (person-r.cc:100, 128-173)
EXPONENTS = [1.0 / np.exp(2.3525), 1.0 / np.exp(1.0353), 1.0 / np.exp(1.4404)]


def test_exp_against_mpmath_and_glibc(emu):
    mpmath.mp.dps = 60
    rng = np.random.RandomState(1)
    x = np.concatenate([-rng.exponential(2.0, 20000), rng.uniform(-700, 700, 5000), [0.0, -0.0, -1e-300, 1e-17, -700.0, 700.0],
                        np.log(2) * (rng.randint(-1000, 1000, 2000) + 0.5) * (1 + rng.uniform(-1e-15, 1e-15, 2000))])
    got = _call(emu, 0, x)
    assert _ulp_err_vs_true(got, [mpmath.exp(mpmath.mpf(float(v))) for v in x]) < 1.05
    big = np.concatenate([-rng.exponential(3.0, 2000000), rng.uniform(-100, 5, 1000000)])
    assert _ulp_diff(_call(emu, 0, big), np.exp(big)) <= 1.0
    assert _call(emu, 0, np.array([-701.0, -1e6, -np.inf]))[0] == 0.0
    assert np.isinf(_call(emu, 0, np.array([701.0]))[0])


def test_log_against_mpmath_and_glibc(emu):
    mpmath.mp.dps = 60
    rng = np.random.RandomState(2)
    u = rng.uniform(0, 1, 20000)
    x = np.concatenate([u, 1 - 2.0 ** -rng.uniform(1, 33, 4000), rng.exponential(1.0, 5000) + 1e-12, 2.0 ** rng.uniform(-40, 40, 3000),
                        [1.0, 0.5, 2.0, np.sqrt(2), np.sqrt(0.5), np.nextafter(1.0, 0), np.nextafter(1.0, 2), 2.0 ** -32]])
    got = _call(emu, 1, x)
    assert _ulp_err_vs_true(got, [mpmath.log(mpmath.mpf(float(v))) for v in x]) < 1.05
    big = np.concatenate([rng.uniform(0, 1, 3000000), rng.exponential(1.0, 1000000) + 1e-300])
    assert _ulp_diff(_call(emu, 1, big), np.log(big)) <= 1.0


@pytest.mark.parametrize("c", EXPONENTS)
def test_pow_against_mpmath_and_glibc(emu, c):
    """x^c for x = -log(u), u uniform: the onset and dwell-time Weibulls"""
    mpmath.mp.dps = 60
    rng = np.random.RandomState(3)
    x = np.concatenate([-np.log(rng.uniform(0, 1, 20000)), 2.0 ** rng.uniform(-34, 5, 5000), [1.0, 2.0 ** -32, 22.18, np.nextafter(1.0, 0)]])
    got = _call(emu, 2, x, c)
    cm = mpmath.mpf(float(c))
    assert _ulp_err_vs_true(got, [mpmath.power(mpmath.mpf(float(v)), cm) for v in x]) < 1.05
    big = -np.log(rng.uniform(0, 1, 3000000))
    assert _ulp_diff(_call(emu, 2, big, c), np.power(big, c)) <= 1.0


def test_weibull_times_within_two_ulp_of_glibc(emu):
    """scale * pow(-log(u), 1/shape) as the reference evaluates it (nmath rweibull, glibc) against the device's
    arithmetic (fm_log, fm_pow): the event-time bar of the numerical contract"""
    rng = np.random.RandomState(4)
    u = rng.uniform(0, 1, 3000000)
    for c, scale in zip(EXPONENTS, (64.0218, 19.8617, 16.3863)):
        ref = scale * np.power(-np.log(u), c)
        got = scale * _call(emu, 2, -_call(emu, 1, u), c)
        assert _ulp_diff(got, ref) <= 2.0
