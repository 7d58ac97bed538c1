"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of the CPU oracle (oracle/libmsim_oracle.so, and the
`oracle/_ref` build that links the reference's own ssim.cc / RngStream.cpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module; the product package
(microsimulation_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"


class Table(C.Structure):
    _fields_ = [("nrow", C.c_int64), ("ncol", C.c_int32), ("colnames", C.POINTER(C.c_char_p)),
                ("data", C.POINTER(C.c_double))]

    def to_dict(self):
        out = {}
        for j in range(self.ncol):
            name = self.colnames[j].decode()
            if self.nrow:
                col = np.ctypeslib.as_array(self.data, shape=(self.ncol * self.nrow,))[j * self.nrow:(j + 1) * self.nrow]
                out[name] = col.copy()
            else:
                out[name] = np.zeros(0)
        return out


class Report(C.Structure):
    _fields_ = [(k, Table) for k in ("pt", "ut", "events", "prev", "costs", "indiv")]

    def to_dict(self):
        return {k: getattr(self, k).to_dict() for k in ("pt", "ut", "events", "prev", "costs", "indiv")}


class Calibration(C.Structure):
    _fields_ = [("occupancy", C.c_double * 40), ("present", C.c_int32 * 4), ("time_at_risk", C.c_double * 4),
                ("n_time_at_risk", C.c_int32)]

    def to_dict(self):
        names = ["DiseaseFree", "Precursor", "PreClinical", "Clinical"]
        out = {"TimeAtRisk": np.array(self.time_at_risk[:self.n_time_at_risk])}
        for s, nm in enumerate(names):
            if self.present[s]:
                out[nm] = np.array(self.occupancy[s * 10:(s + 1) * 10])
        return out


class PersonShard(C.Structure):
    _fields_ = [("seed", C.c_double * 6), ("first_person", C.c_int64), ("n", C.c_int64), ("outputs", C.c_uint32),
                ("discount_rate", C.c_double)]


def build(flavour="port", quiet=True):
    """Compile the oracle (make) if needed; returns the .so path or None."""
    if flavour == "port":
        path = os.path.join(HERE, "libmsim_oracle.so")
        target = "all"
    else:
        path = os.path.join(HERE, "_ref", "libmsim_oracle_ref.so")
        target = "ref"
        if not os.path.isdir(REF_ROOT):
            return path if os.path.exists(path) else None
    r = subprocess.run(["make", "-C", HERE, target], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return path


SICK_SICKER_KEYS = ["r_HS1", "r_HD", "r_S1H", "r_S1S2", "r_S1D", "r_S2D", "c_H", "c_S1", "c_S2", "c_Trt", "u_H",
                    "u_S1", "u_S2", "u_Trt", "discountRate", "partitionBy", "Trt"]


class Oracle:
    def __init__(self, flavour="port"):
        path = build(flavour)
        if path is None or not os.path.exists(path):
            raise FileNotFoundError("oracle flavour %r is not built" % flavour)
        self.lib = L = C.CDLL(path)
        L.oracle_flavour.restype = C.c_char_p
        self.flavour = L.oracle_flavour().decode()
        L.oracle_events_handled.restype = C.c_uint64
        L.oracle_draws_taken.restype = C.c_uint64
        L.oracle_qnorm.restype = C.c_double
        L.oracle_qnorm.argtypes = [C.c_double]
        L.oracle_gammafn.restype = C.c_double
        L.oracle_gammafn.argtypes = [C.c_double]
        L.oracle_discountedInterval.restype = C.c_double
        L.oracle_discountedInterval.argtypes = [C.c_double] * 4
        L.oracle_tick_tock.restype = C.c_int64
        L.oracle_doc_illness_death.restype = C.c_int64
        L.oracle_r_create_current_stream()

    # ---- RNG
    def set_user_random_seed(self, seed):
        s = (C.c_double * 6)(*[float(x) for x in _seed6(seed)])
        self.lib.oracle_r_set_user_random_seed(s)

    def get_user_random_seed(self):
        s = (C.c_double * 6)()
        self.lib.oracle_r_get_user_random_seed(s)
        return list(s)

    def next_rng_substream(self):
        self.lib.oracle_r_next_rng_substream()

    def rng_advance_substream(self, seed, n):
        s = (C.c_double * 6)(*[float(x) for x in _seed6(seed)])
        nn = C.c_int(n)
        self.lib.oracle_r_rng_advance_substream(s, C.byref(nn))
        return list(s)

    def set_seed(self, seed):
        self.lib.oracle_mt_set_seed(C.c_uint32(seed))

    def rng_uniforms(self, seed, first_substream, n_substreams, draws):
        s = (C.c_double * 6)(*[float(x) for x in _seed6(seed)])
        out = np.empty(n_substreams * draws)
        rc = self.lib.oracle_rng_uniforms(s, C.c_int64(first_substream), C.c_int64(n_substreams), C.c_int32(draws),
                                          out.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        return out.reshape(n_substreams, draws)

    def mt_uniforms(self, n):
        out = np.empty(n)
        self.lib.oracle_mt_uniforms(C.c_int64(n), out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def sample(self, which, a, b, n):
        idx = {"runif": 0, "rweibull": 1, "rexp": 2, "rnorm": 3, "exp_rand": 4, "norm_rand": 5}[which]
        out = np.empty(n)
        rc = self.lib.oracle_sample(idx, C.c_double(a), C.c_double(b), C.c_int64(n),
                                    out.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        return out

    def qnorm(self, p):
        return self.lib.oracle_qnorm(p)

    def gammafn(self, x):
        return self.lib.oracle_gammafn(x)

    def discountedInterval(self, y, start, end, rate):
        return self.lib.oracle_discountedInterval(y, start, end, rate)

    def counters(self):
        return {"events": self.lib.oracle_events_handled(), "draws": self.lib.oracle_draws_taken()}

    def reset_counters(self):
        self.lib.oracle_reset_counters()

    # ---- models
    def _table_call(self, fn, *args):
        t = Table()
        rc = fn(*args, C.byref(t))
        assert rc == 0
        d = t.to_dict()
        self.lib.oracle_table_free(C.byref(t))
        return d

    def _report_call(self, fn, *args):
        r = Report()
        rc = fn(*args, C.byref(r))
        assert rc == 0
        d = r.to_dict()
        self.lib.oracle_report_free(C.byref(r))
        return d

    def callPersonSimulation(self, n=20, seed=(12345,) * 6):
        s = (C.c_int32 * 6)(*[int(x) for x in _seed6(seed)])
        return self._table_call(self.lib.oracle_callPersonSimulation, s, C.c_int32(n))

    def person_run(self, seed, first_person, n, rowlog=True, report=False, discount_rate=0.0):
        sh = PersonShard((C.c_double * 6)(*[float(x) for x in _seed6(seed)]), first_person, n, 0, discount_rate)
        t, r = Table(), Report()
        counts = (C.c_double * 9)()
        rc = self.lib.oracle_person_run(C.byref(sh), C.byref(t) if rowlog else None, counts,
                                        C.byref(r) if report else None)
        assert rc == 0
        out = {"counts": np.array(counts[:])}
        if rowlog:
            out["rowlog"] = t.to_dict()
            self.lib.oracle_table_free(C.byref(t))
        if report:
            out["report"] = r.to_dict()
            self.lib.oracle_report_free(C.byref(r))
        return out

    def person_run_options(self, seed, first_person, n, discount_rate, start, finish, delta, max_time=1.0e100,
                           start_report_age=0.0, output_utilities=True):
        """person-r with an EventReport built from setPartition(start, finish, delta, maxTime), startReportAge and
        outputUtilities, individualReset after every person: (report tables, counts, wrap_indiv, wrap_means)."""
        class Opt(C.Structure):
            _fields_ = [("start", C.c_double), ("finish", C.c_double), ("delta", C.c_double), ("max_time", C.c_double),
                        ("start_report_age", C.c_double), ("output_utilities", C.c_int32)]
        sh = PersonShard((C.c_double * 6)(*[float(x) for x in _seed6(seed)]), first_person, n, 0, discount_rate)
        o = Opt(start, finish, delta, max_time, start_report_age, int(bool(output_utilities)))
        r = Report()
        counts = (C.c_double * 9)()
        indiv = np.empty(max(n, 1))
        means = (C.c_double * 7)()
        rc = self.lib.oracle_person_run_options(C.byref(sh), C.byref(o), counts, C.byref(r),
                                                indiv.ctypes.data_as(C.POINTER(C.c_double)), means)
        assert rc == 0
        rep = r.to_dict()
        self.lib.oracle_report_free(C.byref(r))
        return rep, np.array(counts[:]), indiv[:n], dict(zip(("n", "mean", "var", "sd", "se", "sum", "sumsq"), means[:]))

    def callSimplePerson(self, n=10, seed=12345):
        self.set_seed(seed)
        return self._table_call(self.lib.oracle_callSimplePerson, C.c_int32(n))

    def callSimplePerson2(self, n=10, seed=12345):
        self.set_seed(seed)
        return self._report_call(self.lib.oracle_callSimplePerson2, C.c_int32(n))

    def callIllnessDeath(self, n=10, cure=0.1, zsd=0.0, seed=12345):
        self.set_seed(seed)
        return self._report_call(self.lib.oracle_callIllnessDeath, C.c_int32(n), C.c_double(cure), C.c_double(zsd))

    def callCalibrationPerson(self, seed=12345, n=500, runpar=(4, 0.5, 0.05, 10, 3, 0.5)):
        """R/calibSim.R:10-40 with mc.cores=1."""
        self.set_user_random_seed(seed)
        par = (C.c_double * 6)(*[float(x) for x in runpar])
        c = Calibration()
        rc = self.lib.oracle_callCalibrationSimulation(C.c_int32(n), par, C.byref(c))
        assert rc == 0
        return c.to_dict()

    def calibration_sweep(self, seed, first_person, n, runpars):
        runpars = np.ascontiguousarray(runpars, dtype=np.float64)
        ns = runpars.shape[0]
        arr = (Calibration * ns)()
        s = (C.c_double * 6)(*[float(x) for x in _seed6(seed)])
        rc = self.lib.oracle_calibration_sweep(s, C.c_int64(first_person), C.c_int64(n), C.c_int32(ns),
                                               runpars.ctypes.data_as(C.POINTER(C.c_double)), arr)
        assert rc == 0
        return [arr[i].to_dict() for i in range(ns)]

    def sick_sicker(self, n, param, indivp=True, substreams=False, first_person=0, seed=12345, trace=False):
        if not substreams:
            self.set_seed(seed)
        par = (C.c_double * 17)(*[float(param[k]) for k in SICK_SICKER_KEYS])
        r = Report()
        cap = 5 * 64 * max(n, 1) if trace else 0
        tr = np.zeros(max(cap, 1))
        tlen = C.c_int64(0)
        rc = self.lib.oracle_sick_sicker(C.c_int64(n), par, C.c_int(int(indivp)), C.c_int(int(substreams)),
                                         C.c_int64(first_person), C.byref(r),
                                         tr.ctypes.data_as(C.POINTER(C.c_double)) if trace else None,
                                         C.c_int64(cap), C.byref(tlen))
        assert rc == 0
        d = r.to_dict()
        self.lib.oracle_report_free(C.byref(r))
        if trace:
            d["trace"] = tr[:tlen.value].reshape(-1, 5)
        return d

    def cancer_management(self, n, odShape=8.0, discountRate=0.03, partitionBy=50.0, indivp=True, first_person=0):
        """simulations() of test/test_microsimulation.R:782-865 / 1006-1093 (SummaryReport + addPointCost), person i
        on substream first_person + i + 1 of a stream made from the package seed."""
        r = Report()
        rc = self.lib.oracle_cancer_management(C.c_int64(n), C.c_double(odShape), C.c_double(discountRate), C.c_double(partitionBy),
                                               C.c_int(int(indivp)), C.c_int64(first_person), C.byref(r))
        assert rc == 0
        d = r.to_dict()
        self.lib.oracle_report_free(C.byref(r))
        return d

    def screening_skeleton(self, inputs, n, first_person=0, discount_rate=0.0):
        """The screening-model skeleton (examples/plugins/screening_skeleton.h) on the oracle's building blocks:
        (report tables, values n x 4: age at death, age at diagnosis, screen-detected, grade)."""
        r = Report()
        values = np.zeros((max(n, 1), 4))
        rc = self.lib.oracle_screening_skeleton(C.byref(inputs), C.c_int64(first_person), C.c_int64(n), C.c_double(discount_rate),
                                                C.byref(r), values.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        d = r.to_dict()
        self.lib.oracle_report_free(C.byref(r))
        return d, values[:n]

    # ---- generalised survival model (gsm_lite); `model` is a ctypes msim_gsm (include/msim.h), passed by reference
    def _gsm(self, fn, model, u, extra, index, scale):
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.empty_like(u)
        idx = None if index is None else np.ascontiguousarray(index, dtype=np.int32)
        args = [C.byref(model), C.c_int64(u.size), u.ctypes.data_as(C.POINTER(C.c_double))] + extra
        args.append(idx.ctypes.data_as(C.POINTER(C.c_int32)) if idx is not None else None)
        if scale is not None:
            args.append(C.c_double(scale))
        args.append(out.ctypes.data_as(C.POINTER(C.c_double)))
        assert fn(*args) == 0
        return out

    def gsm_randU(self, model, u, tentry=None, index=None, scale=10.0):
        te = None if tentry is None else np.ascontiguousarray(tentry, dtype=np.float64)
        return self._gsm(self.lib.oracle_gsm_randU, model, u, [None if te is None else te.ctypes.data_as(C.POINTER(C.c_double))],
                         index, scale)

    def gsm_randU0(self, model, u, index=None, scale=10.0):
        return self._gsm(self.lib.oracle_gsm_randU0, model, u, [], index, scale)

    def gsm_eta(self, model, y, index=None):
        return self._gsm(self.lib.oracle_gsm_eta, model, y, [], index, None)

    def spline_eval(self, model, term, x, der=0, natural=True, width=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros((x.size, width))
        fn = self.lib.oracle_ns_eval if natural else self.lib.oracle_bs_eval
        assert fn(C.byref(model), C.c_int(term), C.c_int64(x.size), x.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(der),
                  out.ctypes.data_as(C.POINTER(C.c_double))) == 0
        return out

    # ---- survival-time utilities and lookup tables (lookups_lite.h); modes as in include/msim.h
    def rpexp(self, mode, h, t, a, b=None):
        h, t, a = (np.ascontiguousarray(v, dtype=np.float64) for v in (h, t, a))
        b = None if b is None else np.ascontiguousarray(np.broadcast_to(b, a.shape), dtype=np.float64)
        out = np.empty_like(a)
        dp = C.POINTER(C.c_double)
        rc = self.lib.oracle_rpexp(C.c_int(mode), C.c_int(h.size), h.ctypes.data_as(dp), t.ctypes.data_as(dp), C.c_int64(a.size),
                                   a.ctypes.data_as(dp), None if b is None else b.ctypes.data_as(dp), out.ctypes.data_as(dp))
        if rc:
            raise RuntimeError("Rpexp.rand() failed due to zero hazards")
        return out

    def rpexp_gamma(self, mode, theta, mu, t, a, Z, start=None):
        mu, t, a = (np.ascontiguousarray(v, dtype=np.float64) for v in (mu, t, a))
        Z = np.ascontiguousarray(np.broadcast_to(Z, a.shape), dtype=np.float64)
        start = None if start is None else np.ascontiguousarray(np.broadcast_to(start, a.shape), dtype=np.float64)
        out = np.empty_like(a)
        dp = C.POINTER(C.c_double)
        assert self.lib.oracle_rpexp_gamma(C.c_int(mode), C.c_int(mu.size), C.c_double(theta), mu.ctypes.data_as(dp),
                                           t.ctypes.data_as(dp), C.c_int64(a.size), a.ctypes.data_as(dp), Z.ctypes.data_as(dp),
                                           None if start is None else start.ctypes.data_as(dp), out.ctypes.data_as(dp)) == 0
        return out

    def interpolate(self, mode, x, y, v):
        x, y, v = (np.ascontiguousarray(w, dtype=np.float64) for w in (x, y, v))
        out = np.empty_like(v)
        dp = C.POINTER(C.c_double)
        assert self.lib.oracle_interpolate(C.c_int(mode), C.c_int(x.size), x.ctypes.data_as(dp), y.ctypes.data_as(dp),
                                           C.c_int64(v.size), v.ctypes.data_as(dp), out.ctypes.data_as(dp)) == 0
        return out

    def table_lookup(self, rows, keys):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        k = rows.shape[1] - 1
        keys = np.ascontiguousarray(keys, dtype=np.float64).reshape(-1, k)
        out = np.empty(keys.shape[0])
        dp = C.POINTER(C.c_double)
        assert self.lib.oracle_table_lookup(C.c_int(k), C.c_int64(rows.shape[0]), rows.ctypes.data_as(dp), C.c_int64(keys.shape[0]),
                                            keys.ctypes.data_as(dp), out.ctypes.data_as(dp)) == 0
        return out

    def tick_tock(self):
        buf = np.zeros(4096)
        n = self.lib.oracle_tick_tock(buf.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(buf.size))
        return buf[:n].copy()

    def doc_illness_death(self, n=5):
        buf = np.zeros(4096)
        qc = np.zeros(2 * n)
        m = self.lib.oracle_doc_illness_death(C.c_int(n), buf.ctypes.data_as(C.POINTER(C.c_double)),
                                              C.c_int64(buf.size), qc.ctypes.data_as(C.POINTER(C.c_double)))
        return buf[:m].copy(), qc.reshape(n, 2)

    def priority_demo(self):
        o = (C.c_double * 4)()
        n = self.lib.oracle_priority_demo(o)
        return list(o)[:n]


def _seed6(seed):
    """set.user.Random.seed's recycling rule (R/rcpp_hello_world.R:104-110)."""
    if np.isscalar(seed):
        return [seed] * 6
    seed = list(seed)
    if len(seed) == 7:
        seed = seed[1:]
    assert len(seed) == 6
    return seed


def sick_sicker_param(Trt=False):
    """README.org:148-187.  Rates are logm(Pmat) entries (scipy)."""
    return dict(r_HS1=0.28424764950662157, r_HD=0.0034485118628178875, r_S1H=0.9474921650220719,
                r_S1S2=0.17396002318490003, r_S1D=0.017327099626200958, r_S2D=0.05012541823544285,
                c_H=2000.0, c_S1=4000.0, c_S2=15000.0, c_Trt=12000.0, u_H=1.0, u_S1=0.75, u_S2=0.5, u_Trt=0.95,
                discountRate=0.03, partitionBy=1.0, Trt=int(Trt))
