"""Generalised survival models: the host-side mirror of the reference's `gsm` class (src/gsm.cpp, inst/include/gsm.h).

`GsmModel` holds the fields of the list that `gsm::gsm(Rcpp::List)` reads (gsm.cpp:60-92) — what rstpm2's
`gsm_design()` hands to the reference — and `randU / randU0 / eta` run the inversions on the GPU through the C ABI
(include/msim.h: msim_gsm_randU, msim_gsm_randU0, msim_gsm_eta)."""
import ctypes as C

import numpy as np

from ._lib import check, load


class GsmTermC(C.Structure):
    _fields_ = [("n_gamma", C.c_int32), ("n_knots", C.c_int32), ("intercept", C.c_int32), ("cure", C.c_int32),
                ("q_rows", C.c_int32), ("q_cols", C.c_int32), ("gamma", C.POINTER(C.c_double)),
                ("knots", C.POINTER(C.c_double)), ("q_const", C.POINTER(C.c_double)), ("Boundary_knots", C.c_double * 2),
                ("x", C.POINTER(C.c_double))]


class GsmC(C.Structure):
    _fields_ = [("link", C.c_int32), ("log_time", C.c_int32), ("n_index", C.c_int32), ("n_terms", C.c_int32),
                ("tmin", C.c_double), ("tmax", C.c_double), ("inflate", C.c_double), ("t0", C.c_double),
                ("etap", C.POINTER(C.c_double)), ("etap0", C.POINTER(C.c_double)), ("terms", C.POINTER(GsmTermC))]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class GsmModel:
    """link_name, tmin, tmax, inflate, etap, etap0, terms=[dict(gamma, knots, Boundary_knots, intercept, q_const, cure, x)],
    log_time, t0 — the list elements gsm.cpp:60-92 reads."""

    def __init__(self, tmin, tmax, etap, terms, etap0=None, inflate=1.0, log_time=True, t0=np.inf, link_name="PH"):
        if link_name != "PH":
            raise ValueError('only link "PH" exists in the reference')
        self.etap = np.ascontiguousarray(etap, dtype=np.float64)
        self.etap0 = np.ascontiguousarray(self.etap if etap0 is None else etap0, dtype=np.float64)
        self.tmin, self.tmax, self.inflate, self.log_time, self.t0 = float(tmin), float(tmax), float(inflate), bool(log_time), float(t0)
        self.terms = []
        for t in terms:
            self.terms.append(dict(gamma=np.ascontiguousarray(t["gamma"], dtype=np.float64),
                                   knots=np.ascontiguousarray(t["knots"], dtype=np.float64),
                                   Boundary_knots=[float(v) for v in t["Boundary_knots"]],
                                   intercept=int(t.get("intercept", 0)), cure=int(t.get("cure", 0)),
                                   q_const=np.ascontiguousarray(t["q_const"], dtype=np.float64),
                                   x=np.ascontiguousarray(t["x"], dtype=np.float64)))
        self._keep = None

    def cstruct(self):
        arr = (GsmTermC * max(len(self.terms), 1))()
        for i, t in enumerate(self.terms):
            arr[i] = GsmTermC(len(t["gamma"]), len(t["knots"]), t["intercept"], t["cure"], t["q_const"].shape[0],
                              t["q_const"].shape[1], _dp(t["gamma"]), _dp(t["knots"]), _dp(t["q_const"]),
                              (C.c_double * 2)(*t["Boundary_knots"]), _dp(t["x"]))
        m = GsmC(0, int(self.log_time), len(self.etap), len(self.terms), self.tmin, self.tmax, self.inflate, self.t0,
                 _dp(self.etap), _dp(self.etap0), arr)
        self._keep = (arr, m)
        return m

    # ---- on the GPU
    def _run(self, fn, u, extra, index, scale):
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.empty_like(u)
        idx = None if index is None else np.ascontiguousarray(index, dtype=np.int32)
        m = self.cstruct()
        args = [C.byref(m), C.c_int64(u.size), _dp(u)] + extra
        args.append(idx.ctypes.data_as(C.POINTER(C.c_int32)) if idx is not None else None)
        if scale is not None:
            args.append(C.c_double(scale))
        args.append(_dp(out))
        check(fn(*args))
        return out

    def randU(self, u, tentry=None, index=None, scale=10.0):
        te = None if tentry is None else np.ascontiguousarray(tentry, dtype=np.float64)
        return self._run(load().msim_gsm_randU, u, [None if te is None else _dp(te)], index, scale)

    def randU0(self, u, index=None, scale=10.0):
        return self._run(load().msim_gsm_randU0, u, [], index, scale)

    def eta(self, y, index=None):
        return self._run(load().msim_gsm_eta, y, [], index, None)
