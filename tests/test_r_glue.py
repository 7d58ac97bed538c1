"""The R glue (glue/msim_glue.c) driven the way R's .Call / .C would drive it.

R is not in the image, so the glue is compiled against tests/r_shim — a small stand-in for the R API with R's names
and semantics (typed vectors with attributes, a global environment holding .Random.seed, non-returning Rf_error) —
and called by symbol with the argument counts the reference registers (src/init.c:32-62).  The results are read back
as R would see them (names, classes, integer vs double columns) and compared with the oracle.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_report_close, assert_rowlog_close

# what src/init.c:32-62 registers and the glue replaces: (R-level name, C symbol, number of arguments)
DOT_C = [("r_get_user_random_seed", "r_get_user_random_seed", 1), ("r_next_rng_substream", "r_next_rng_substream", 0),
         ("r_rng_advance_substream", "r_rng_advance_substream", 2), ("r_set_user_random_seed", "r_set_user_random_seed", 1),
         ("r_create_current_stream", "r_create_current_stream", 0), ("r_remove_current_stream", "r_remove_current_stream", 0)]
DOT_CALL = [("callCalibrationSimulation", "callCalibrationSimulation", 1), (".callIllnessDeath", "callIllnessDeath", 1),
            (".callPersonSimulation", "callPersonSimulation", 2), (".callSimplePerson", "callSimplePerson", 1),
            (".callSimplePerson2", "callSimplePerson2", 1)]
INTSXP, REALSXP, STRSXP, VECSXP = 13, 14, 16, 19


class R:
    """A few lines of R's evaluator over the shim."""

    def __init__(self, lib):
        self.lib = lib
        for f in ("shim_int", "shim_real", "shim_list", "shim_get_global", "shim_elt", "shim_attr", "shim_dot_call", "shim_data"):
            getattr(lib, f).restype = C.c_void_p
        for f in ("shim_string", "shim_last_error", "shim_last_stderr"):
            getattr(lib, f).restype = C.c_char_p
        lib.shim_length.restype = C.c_int64
        lib.user_unif_rand.restype = C.POINTER(C.c_double)
        lib.shim_reset()

    # R objects in
    def integer(self, v):
        a = np.ascontiguousarray(v, dtype=np.int32)
        return C.c_void_p(self.lib.shim_int(a.ctypes.data_as(C.POINTER(C.c_int)), C.c_int64(a.size)))

    def double(self, v):
        a = np.ascontiguousarray(v, dtype=np.float64)
        return C.c_void_p(self.lib.shim_real(a.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(a.size)))

    def list(self, **kw):
        names = (C.c_char_p * len(kw))(*[k.encode() for k in kw])
        elts = (C.c_void_p * len(kw))(*[v.value for v in kw.values()])
        return C.c_void_p(self.lib.shim_list(len(kw), names, elts))

    def set_global(self, name, value):
        self.lib.shim_set_global(name.encode(), value)

    def get_global(self, name):
        return self.to_python(C.c_void_p(self.lib.shim_get_global(name.encode())))

    # R objects out: vectors -> numpy arrays, named lists -> dicts (order kept), plus the attributes R would show
    def to_python(self, x):
        L = self.lib
        t, n = L.shim_typeof(x), L.shim_length(x)
        if t == REALSXP:
            return np.ctypeslib.as_array(C.cast(L.shim_data(x), C.POINTER(C.c_double)), (n,)).copy() if n else np.empty(0)
        if t == INTSXP:
            return np.ctypeslib.as_array(C.cast(L.shim_data(x), C.POINTER(C.c_int)), (n,)).copy() if n else np.empty(0, np.int32)
        if t == STRSXP:
            return [L.shim_string(x, C.c_int64(i)).decode() for i in range(n)]
        assert t == VECSXP, t
        names = L.shim_attr(x, b"names")
        elts = [self.to_python(C.c_void_p(L.shim_elt(x, C.c_int64(i)))) for i in range(n)]
        if not names:
            assert n == 0
            return {}
        out = dict(zip(self.to_python(C.c_void_p(names)), elts))
        assert len(out) == n
        return out

    def attr(self, x, name):
        a = self.lib.shim_attr(x, name.encode())
        return self.to_python(C.c_void_p(a)) if a else None

    def call(self, symbol, *args, raw=False):
        """.Call(symbol, ...)"""
        nargs = dict((s, n) for _, s, n in DOT_CALL)[symbol]
        assert len(args) == nargs, "R checks the registered argument count"
        fn = C.cast(getattr(self.lib, symbol), C.c_void_p)
        a = (C.c_void_p * max(1, len(args)))(*[x.value for x in args])
        out = self.lib.shim_dot_call(fn, len(args), a)
        if not out:
            raise RuntimeError(self.lib.shim_last_error().decode())
        assert self.lib.shim_protect_depth() == 0, "unbalanced PROTECT/UNPROTECT"
        return C.c_void_p(out) if raw else self.to_python(C.c_void_p(out))

    def dot_c(self, symbol, *arrays):
        """.C(symbol, ...): arrays are modified in place, as R copies them back"""
        nargs = dict((s, n) for _, s, n in DOT_C)[symbol]
        assert len(arrays) == nargs
        fn = C.cast(getattr(self.lib, symbol), C.c_void_p)
        a = (C.c_void_p * max(1, len(arrays)))(*[x.ctypes.data for x in arrays])
        if self.lib.shim_dot_c(fn, len(arrays), a):
            raise RuntimeError(self.lib.shim_last_error().decode())

    def set_seed_mersenne(self, state625):
        """RNGkind("Mersenne-Twister"); set.seed(.) leaves .Random.seed = c(10403L, position, mt[624])"""
        self.set_global(".Random.seed", self.integer(np.concatenate([[10403], np.asarray(state625, np.uint32).view(np.int32)])))


@pytest.fixture(scope="module")
def r():
    import microsimulation_b200  # noqa: F401  (builds nothing; makes sure the package directory is where we link)
    so = os.path.join(ROOT, "build", "libglue_test.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    shim = os.path.join(ROOT, "tests", "r_shim")
    cmd = ["gcc", "-O1", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", "-I" + os.path.join(shim, "include"),
           "-I" + os.path.join(ROOT, "include"), "-o", so, os.path.join(ROOT, "glue", "msim_glue.c"),
           os.path.join(shim, "rshim.c"), "-L" + os.path.join(ROOT, "microsimulation_b200"), "-lmsim_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "microsimulation_b200"), "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    rr = R(C.CDLL(so))
    yield rr
    rr.lib.shim_reset()


def test_glue_defines_every_symbol_init_c_registers(r):
    for _, sym, _ in DOT_C + DOT_CALL:
        assert hasattr(r.lib, sym), sym
    assert hasattr(r.lib, "user_unif_rand")  # resolved by name under RNGkind("user")


def test_dot_c_rng_table(r, oracle):
    """set.user.Random.seed / user.Random.seed / next.user.Random.substream / advance.substream through .C, and
    runif() through the user_unif_rand hook (R/rcpp_hello_world.R:104-152, src/microsimulation.cc:33-85)."""
    r.dot_c("r_create_current_stream")
    seed = np.full(6, 12345.0)
    r.dot_c("r_set_user_random_seed", seed)
    out = np.zeros(6)
    r.dot_c("r_get_user_random_seed", out)
    assert np.array_equal(out, seed)
    want = oracle.rng_uniforms(12345, 0, 2, 6)
    got = np.array([r.lib.user_unif_rand()[0] for _ in range(6)])
    assert np.array_equal(got, want[0])
    r.dot_c("r_next_rng_substream")
    got = np.array([r.lib.user_unif_rand()[0] for _ in range(6)])
    assert np.array_equal(got, want[1])
    s = np.full(6, 12345.0)
    r.dot_c("r_rng_advance_substream", s, np.array([1000], dtype=np.int32))
    assert list(s) == oracle.rng_advance_substream(12345, 1000)
    with pytest.raises(RuntimeError, match="illegal MRG32k3a seed"):
        r.dot_c("r_set_user_random_seed", np.array([0.0, 0, 0, 1, 1, 1]))
    r.dot_c("r_remove_current_stream")
    assert not r.lib.user_unif_rand()  # NULL, and the reference's message
    assert b"No stream created yet" in r.lib.shim_last_stderr()
    r.dot_c("r_create_current_stream")


def test_argument_errors_become_r_errors(r):
    with pytest.raises(RuntimeError, match="no element 'n'"):
        r.call("callPersonSimulation", r.integer([12345] * 6), r.list(m=r.integer([5])))
    with pytest.raises(RuntimeError, match="six integers"):
        r.call("callPersonSimulation", r.double([12345.0] * 6), r.list(n=r.integer([5])))
    with pytest.raises(RuntimeError, match="runpar"):
        r.call("callCalibrationSimulation", r.list(n=r.integer([5]), runpar=r.integer([1, 2, 3])))


def test_no_device_is_an_r_error_not_a_fallback(r):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="GPU engine"):
        r.call("callPersonSimulation", r.integer([12345] * 6), r.list(n=r.integer([5])))


@pytest.mark.gpu
def test_call_person_simulation(r, oracle):
    out = r.call("callPersonSimulation", r.integer([12345] * 6), r.list(n=r.integer([20000])))
    assert list(out) == ["endTime", "event", "id", "startTime", "state"]  # std::map key order (person-r.cc:224)
    assert all(v.dtype == np.float64 for v in out.values())
    assert_rowlog_close(out, oracle.callPersonSimulation(20000))
    other = r.call("callPersonSimulation", r.integer([1, 2, 3, 4, 5, 6]), r.list(n=r.integer([300])))
    assert_rowlog_close(other, oracle.callPersonSimulation(300, seed=[1, 2, 3, 4, 5, 6]))
    with pytest.raises(RuntimeError, match="seed"):
        r.call("callPersonSimulation", r.integer([0, 0, 0, 1, 2, 3]), r.list(n=r.integer([5])))


@pytest.mark.gpu
def test_mersenne_twister_models_read_and_write_random_seed(r, oracle, msim):
    """callSimplePerson: .Random.seed goes in, the advanced state comes out (Rcpp::RNGScope), so a second call goes on
    where the first stopped."""
    with pytest.raises(RuntimeError, match="Mersenne-Twister"):
        r.lib.shim_reset()
        r.call("callSimplePerson", r.list(n=r.integer([10])))
    msim.set_seed(12345)
    r.set_seed_mersenne(msim.mt_state())
    a = r.call("callSimplePerson", r.list(n=r.integer([3000])))
    assert list(a) == ["endtime", "event", "id", "startTime", "state"]  # simple-example.cc:45 spells it so
    want = oracle.callSimplePerson(5000)
    first = want["id"] < 3000
    assert_rowlog_close(a, {k: v[first] for k, v in want.items()}, end_col="endtime")
    seed_after = r.get_global(".Random.seed")
    assert seed_after[0] == 10403 and 1 <= seed_after[1] <= 624
    b = r.call("callSimplePerson", r.list(n=r.integer([2000])))  # ids start at 0 again; the draws go on
    rest = {k: v[~first] for k, v in want.items()}
    rest["id"] = rest["id"] - 3000
    assert_rowlog_close(b, rest, end_col="endtime")


def _as_report(out):
    for tb in out.values():
        for k in ("Key", "number", "event"):
            if k in tb:
                assert tb[k].dtype == np.int32, k   # Rcpp wraps short / int as integer vectors
        assert tb["age"].dtype == np.float64
    return {tb: {k: np.asarray(v, dtype=np.float64) for k, v in cols.items()} for tb, cols in out.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("symbol", ["callIllnessDeath", "callSimplePerson2"])
def test_event_report_models_return_data_frames(r, oracle, msim, symbol):
    msim.set_seed(12345)
    r.set_seed_mersenne(msim.mt_state())
    n = 20000
    if symbol == "callIllnessDeath":
        parms = r.list(n=r.integer([n]), cure=r.double([0.1]), zsd=r.double([0.0]))
        want = oracle.callIllnessDeath(n, 0.1, 0.0)
    else:
        parms = r.list(n=r.integer([n]))
        want = oracle.callSimplePerson2(n)
    raw = r.call(symbol, parms, raw=True)
    out = r.to_python(raw)
    assert list(out) == ["pt", "ut", "events", "prev"]
    assert list(out["events"]) == ["Key", "event", "age", "number"] and list(out["ut"]) == ["Key", "age", "utility"]
    for i, tb in enumerate(out):
        df = C.c_void_p(r.lib.shim_elt(raw, C.c_int64(i)))
        assert r.attr(df, "class") == ["data.frame"]
        rn = r.attr(df, "row.names")
        assert rn[0] == -2**31 and rn[1] == -len(out[tb]["age"])  # compact row names c(NA, -nrow)
    assert_report_close(_as_report(out), want)
    # n = 0: EventReport::wrap() returns list() when no event was reported (microsimulation.h:976)
    parms0 = r.list(n=r.integer([0]), cure=r.double([0.1]), zsd=r.double([0.0]))
    assert r.call(symbol, parms0) == {}


@pytest.mark.gpu
def test_call_calibration_simulation(r, oracle):
    """callCalibrationPerson's inner .Call (R/rcpp_hello_world.R, callCalibrationPerson): the package seed comes from
    set.user.Random.seed; the list has the stages in std::map order and only those ever counted."""
    r.dot_c("r_create_current_stream")
    r.dot_c("r_set_user_random_seed", np.full(6, 12345.0))
    runpar = [4, 0.5, 0.05, 10, 3, 0.5]
    out = r.call("callCalibrationSimulation", r.list(n=r.integer([5000]), runpar=r.double(runpar)))
    want = oracle.callCalibrationPerson(12345, 5000, runpar)   # the same named list, as a dict
    assert list(out) == sorted(want)
    for k in want:
        if k == "TimeAtRisk":
            np.testing.assert_allclose(out[k], want[k], rtol=1e-9)
        else:
            assert np.array_equal(out[k], want[k]), k
    # a tiny run never reaches the later stages: the reference's list then lacks those names
    few = r.call("callCalibrationSimulation", r.list(n=r.integer([2]), runpar=r.double(runpar)))
    want_few = oracle.callCalibrationPerson(12345, 2, runpar)
    assert list(few) == sorted(want_few) and len(few) < 5
