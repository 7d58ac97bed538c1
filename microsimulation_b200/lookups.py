"""Survival-time utilities and lookup tables of the reference, evaluated element-wise on the GPU through the C ABI
(include/msim.h): Rpexp, Rpexp_gamma (inst/include/microsimulation.h:527-704), NumericInterpolate and Table
(inst/include/rcpp_table.h:52-324)."""
import ctypes as C

import numpy as np

from ._lib import check, load


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


class Rpexp:
    """Piecewise constant hazards `h[i]` from time `t[i]` on; rand(u, from) inverts the survival function."""
    MODES = {"rand": 0, "H": 1, "h": 2, "p": 3, "d": 4}

    def __init__(self, h, t):
        self.h, self.t = _d(h), _d(t)

    def _run(self, mode, a, b=None):
        a = _d(a)
        b = None if b is None else _d(np.broadcast_to(b, a.shape))
        out = np.empty_like(a)
        check(load().msim_rpexp(self.MODES[mode], self.h.size, _p(self.h), _p(self.t), a.size, _p(a), _p(b), _p(out)))
        return out

    def rand(self, u, start=0.0):
        return self._run("rand", u, start)

    def H(self, s):
        return self._run("H", s)

    def hazard(self, s):
        return self._run("h", s)

    def p(self, s):
        return self._run("p", s)

    def d(self, s):
        return self._run("d", s)


class RpexpGamma:
    MODES = {"rand": 0, "H": 1, "h": 2, "p": 3}

    def __init__(self, theta, mu, t):
        self.theta, self.mu, self.t = float(theta), _d(mu), _d(t)

    def _run(self, mode, a, Z, start=None):
        a = _d(a)
        Z = _d(np.broadcast_to(Z, a.shape))
        start = None if start is None else _d(np.broadcast_to(start, a.shape))
        out = np.empty_like(a)
        check(load().msim_rpexp_gamma(self.MODES[mode], self.mu.size, self.theta, _p(self.mu), _p(self.t), a.size, _p(a), _p(Z),
                                      _p(start), _p(out)))
        return out

    def rand(self, u, Z, start=0.0):
        return self._run("rand", u, Z, start)

    def H(self, a, Z):
        return self._run("H", a, Z)

    def hazard(self, a, Z):
        return self._run("h", a, Z)

    def p(self, a, Z):
        return self._run("p", a, Z)


class NumericInterpolate:
    MODES = {"approx": 0, "step": 1, "invert": 2, "invert_decreasing": 3}

    def __init__(self, x, y):
        self.x, self.y = _d(x), _d(y)

    def _run(self, mode, v):
        v = _d(v)
        out = np.empty_like(v)
        check(load().msim_interpolate(self.MODES[mode], self.x.size, _p(self.x), _p(self.y), v.size, _p(v), _p(out)))
        return out

    def approx(self, x):
        return self._run("approx", x)

    def __call__(self, x):
        return self._run("step", x)

    def invert(self, y):
        return self._run("invert", y)

    def invert_decreasing(self, y):
        return self._run("invert_decreasing", y)


class Table:
    """Table<keys..., Outcome> from the rows of a data frame: `rows` is n_rows x (n_axes + 1), outcome last."""

    def __init__(self, rows):
        self.rows = _d(rows)
        self.n_axes = self.rows.shape[1] - 1

    def __call__(self, keys):
        keys = _d(keys).reshape(-1, self.n_axes)
        out = np.empty(keys.shape[0])
        check(load().msim_table_lookup(self.n_axes, self.rows.shape[0], _p(self.rows), keys.shape[0], _p(keys), _p(out)))
        return out
