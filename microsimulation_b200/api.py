"""Host-side mirror of the reference's R wrappers, over the C ABI.

Same names, arguments and defaults as the R functions a user of the reference
calls (R/rcpp_hello_world.R:104-152, 300-390; R/calibSim.R:10-40), so parity
tests read like the reference's own tests (tests/testthat/*.R).  Everything
here is thin: seeds in, tables out; all simulation work happens in
libmsim_b200.so on the GPU.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (MSIM_OUT_COUNTS, MSIM_OUT_INDIV, MSIM_OUT_REPORT, MSIM_OUT_ROWLOG, Calibration, DenseDims,
                   PersonDeviceResult, PersonShard, Report, ReportOptions, Table, check, load)

_stream_created = False


def _ensure_stream():
    """.onLoad of the R package calls r_create_current_stream (R/zzz.R)."""
    global _stream_created
    if not _stream_created:
        load().msim_r_create_current_stream()
        _stream_created = True


def _seed6(seed):
    """set.user.Random.seed's recycling rule (R/rcpp_hello_world.R:104-110)."""
    if np.isscalar(seed):
        seed = [seed] * 6
    seed = [float(x) + (2.0 ** 32 if x < 0 else 0.0) for x in seed]  # unsigned()
    if len(seed) == 7:
        seed = seed[1:]
    if len(seed) != 6:
        raise ValueError("seed must have length 1, 6 or 7")
    return seed


def init(device=0):
    check(load().msim_init(device))


def launch_count():
    return int(load().msim_launch_count())


def set_stream(cuda_stream=None):
    """Enqueue on `cuda_stream` (an integer cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream); None = own."""
    check(load().msim_set_stream(C.c_void_p(cuda_stream or 0), C.c_int(1 if cuda_stream is None else 0)))


def set_person_schedule(which):
    """0 / "staged" (default) or 1 / "generic": which GPU schedule runs person-r, and how the calibration and Sick-Sicker
    models keep their pending actions: 0 = the calibration sweep computes each person's draws once for all parameter
    sets and needs no queue (csrc/calib_shared.cuh), Sick-Sicker uses an ordered array; 1 = the binary heap;
    3 / "queue" = as 0, but the calibration sweep goes through the event loop with an ordered array.  Same results
    either way.  2 / "repeat" is 0 with the repeat paths forced (tests)."""
    which = {"staged": 0, "generic": 1, "repeat": 2, "queue": 3}.get(which, which)
    check(load().msim_set_person_schedule(C.c_int(int(which))))


def set_mt_parallel(on=True):
    """Parallel (jump-ahead) generation of R's Mersenne-Twister stream on / off; returns False if no table is loaded."""
    return load().msim_set_mt_parallel(C.c_int(int(bool(on)))) == 0


def set_rowlog_plan(rows_per_person=1.25, extra_rows_per_person=0.26):
    check(load().msim_set_rowlog_plan(rows_per_person, extra_rows_per_person))


def set_wave_share(share=0.0):
    """Staged person-r schedule: the first wave of blocks' share of the persons (0: adapted from the finishing times
    of earlier launches, the default).  Scheduling only."""
    check(load().msim_set_wave_share(share))


def last_wave_share():
    v = C.c_double(0.0)
    check(load().msim_last_wave_share(C.byref(v)))
    return v.value


# ---- RNG utilities (R/rcpp_hello_world.R:104-152)

def set_user_random_seed(seed):
    _ensure_stream()
    s = (C.c_double * 6)(*_seed6(seed))
    check(load().msim_r_set_user_random_seed(s))
    return list(s)


def user_random_seed():
    _ensure_stream()
    s = (C.c_double * 6)()
    check(load().msim_r_get_user_random_seed(s))
    return [407] + [int(x) for x in s]


def next_user_random_substream():
    _ensure_stream()
    check(load().msim_r_next_rng_substream())


def user_unif_rand(n=1):
    """n draws of the hook R calls under RNGkind("user") (src/microsimulation.cc:77-85): runif(n) there."""
    _ensure_stream()
    L = load()
    out = np.empty(int(n))
    for i in range(int(n)):
        p = L.msim_user_unif_rand()
        if not p:
            raise RuntimeError(L.msim_last_error().decode())
        out[i] = p[0]
    return out


def advance_substream(seed, n):
    s = (C.c_double * 6)(*_seed6(seed))
    nn = C.c_int(int(n))
    check(load().msim_r_rng_advance_substream(s, C.byref(nn)))
    return list(s)


def set_seed(seed=12345):
    """RNGkind("Mersenne-Twister"); set.seed(seed)."""
    check(load().msim_mt_set_seed(C.c_uint32(seed)))


def mt_state():
    st = (C.c_uint32 * 625)()
    check(load().msim_mt_get_state(st))
    return np.array(st[:], dtype=np.uint32)


def set_mt_state(state):
    st = (C.c_uint32 * 625)(*[int(x) for x in state])
    check(load().msim_mt_set_state(st))


def rng_uniforms(seed, first_substream, n_substreams, draws):
    s = (C.c_double * 6)(*_seed6(seed))
    out = np.empty(n_substreams * draws)
    check(load().msim_rng_uniforms(s, C.c_int64(first_substream), C.c_int64(n_substreams), C.c_int32(draws),
                                   out.ctypes.data_as(C.POINTER(C.c_double))))
    return out.reshape(n_substreams, draws)


def mt_uniforms(n):
    out = np.empty(n)
    check(load().msim_mt_uniforms(C.c_int64(n), out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def discountedInterval(y, start, end, discountRate):
    """R/rcpp_hello_world.R:392-400 / microsimulation.h:821-824, element-wise on the GPU."""
    y, a, b = (np.ascontiguousarray(np.broadcast_to(v, np.broadcast(y, start, end).shape), dtype=np.float64).ravel()
               for v in (y, start, end))
    out = np.empty_like(y)
    dp = C.POINTER(C.c_double)
    check(load().msim_discounted(0, y.size, y.ctypes.data_as(dp), a.ctypes.data_as(dp), b.ctypes.data_as(dp), float(discountRate),
                                 out.ctypes.data_as(dp)))
    return out


def discountedPoint(y, time, discountRate):
    y, a = (np.ascontiguousarray(np.broadcast_to(v, np.broadcast(y, time).shape), dtype=np.float64).ravel() for v in (y, time))
    out = np.empty_like(y)
    dp = C.POINTER(C.c_double)
    check(load().msim_discounted(1, y.size, y.ctypes.data_as(dp), a.ctypes.data_as(dp), None, float(discountRate),
                                 out.ctypes.data_as(dp)))
    return out


# ---- the model wrappers

def _table(fn, *args):
    t = Table()
    check(fn(*args, C.byref(t)))
    d = t.to_dict()
    load().msim_table_free(C.byref(t))
    return d


def _report(fn, *args):
    r = Report()
    check(fn(*args, C.byref(r)))
    d = r.to_dict()
    load().msim_report_free(C.byref(r))
    return d


def callPersonSimulation(n=20, seed=(12345,) * 6):
    """R/rcpp_hello_world.R:300-315 -> src/person-r.cc:195-226."""
    s6 = _seed6(list(seed) if not np.isscalar(seed) else seed)
    set_user_random_seed(s6)
    ints = (C.c_int32 * 6)(*[int(x) for x in s6])
    return _table(load().msim_callPersonSimulation, ints, C.c_int32(n))


def callSimplePerson(n=10):
    """R/rcpp_hello_world.R:325-336 -> src/simple-example.cc:72-83."""
    set_seed(12345)
    return _table(load().msim_callSimplePerson, C.c_int32(n))


def callSimplePerson2(n=10):
    """R/rcpp_hello_world.R:346-364 -> src/simple-example2.cc:62-83."""
    set_seed(12345)
    return _report(load().msim_callSimplePerson2, C.c_int32(n))


def callIllnessDeath(n=10, cure=0.1, zsd=0.0):
    """R/rcpp_hello_world.R:374-390 -> src/illness-death.cpp:73-94."""
    set_seed(12345)
    return _report(load().msim_callIllnessDeath, C.c_int32(n), C.c_double(cure), C.c_double(zsd))


def next_rng_stream(seed):
    """parallel::nextRNGStream on the six seed words (R/calibSim.R:21)."""
    s = (C.c_double * 6)(*_seed6(seed))
    check(load().msim_next_rng_stream(s))
    return [int(x) for x in s]


def calibration_chunks(n, mc_cores):
    """Sizes of split(1:n, rep(1:mc.cores, length.out = n)) (R/calibSim.R:17-18)."""
    return [len(range(i, n, mc_cores)) for i in range(mc_cores)]


def callCalibrationPerson(seed=12345, n=500, runpar=(4, 0.5, 0.05, 10, 3, 0.5), mc_cores=1):
    """R/calibSim.R:10-40.  As in the reference, mc.cores chunks each run on a stream of their own (the first on the
    seed's, chunk i on the stream after chunk i-1's), so the result depends on mc_cores; to use several GPUs WITHOUT
    changing the random numbers, shard with distributed.calibration_sweep_sharded instead."""
    set_user_random_seed(seed)
    current = user_random_seed()[1:]
    par = (C.c_double * 6)(*[float(x) for x in runpar])
    states = ["DiseaseFree", "Precursor", "PreClinical", "Clinical"]
    mat, tar, raw = np.zeros((10, 4)), None, []
    for i, n_chunk in enumerate(calibration_chunks(int(n), int(mc_cores))):
        if i > 0:
            current = next_rng_stream(current)
        set_user_random_seed(current)
        c = Calibration()
        check(load().msim_callCalibrationSimulation(C.c_int32(n_chunk), par, C.byref(c)))
        d = c.to_dict()
        raw.append(d)
        for j, nm in enumerate(states):
            if nm in d:
                mat[:, j] += d[nm]
        # Reduce(... u$TimeAtRisk + z$TimeAtRisk): R recycles a shorter vector; chunks of any size reach all four
        t = d["TimeAtRisk"]
        tar = t.copy() if tar is None else (tar + t if len(tar) == len(t) else np.resize(tar, max(len(tar), len(t))) + np.resize(t, max(len(tar), len(t))))
    return {"StateOccupancy": mat, "TimeAtRisk": tar, "raw": raw[0] if mc_cores == 1 else raw}


SICK_SICKER_KEYS = ["r_HS1", "r_HD", "r_S1H", "r_S1S2", "r_S1D", "r_S2D", "c_H", "c_S1", "c_S2", "c_Trt", "u_H", "u_S1",
                    "u_S2", "u_Trt", "discountRate", "partitionBy", "Trt"]


def sick_sicker_param(Trt=False):
    """The README's parameter list (README.org:148-187); rates are the logm(Pmat) entries."""
    return dict(r_HS1=0.28424764950662157, r_HD=0.0034485118628178875, r_S1H=0.9474921650220719,
                r_S1S2=0.17396002318490003, r_S1D=0.017327099626200958, r_S2D=0.05012541823544285,
                c_H=2000.0, c_S1=4000.0, c_S2=15000.0, c_Trt=12000.0, u_H=1.0, u_S1=0.75, u_S2=0.5, u_Trt=0.95,
                discountRate=0.03, partitionBy=1.0, Trt=int(Trt))


def callSim(n, param, indivp=True, substreams=False, first_person=0, seed=12345):
    """README.org:289-301 callSim(n, param, indivp): the Sick-Sicker model with a SummaryReport.  As in the README's
    runs, set.seed(seed) precedes the call (substreams=False); substreams=True uses one RngStream substream per
    person of a stream made from the package seed (set_user_random_seed)."""
    if substreams:
        _ensure_stream()
    else:
        set_seed(seed)
    par = (C.c_double * 17)(*[float(param[k]) for k in SICK_SICKER_KEYS])
    return _report(load().msim_callSim_SickSicker, C.c_int32(n), par, C.c_int(int(indivp)), C.c_int(int(substreams)),
                   C.c_int64(first_person))


def cancer_management(n, odShape=8.0, discountRate=0.03, partitionBy=50.0, indivp=True, first_person=0):
    """simulations(n, discountRate, odShape, partitionBy) of test/test_microsimulation.R:1006-1093: SummaryReport with
    point costs (addPointCost); one substream per person of a stream made from the package seed."""
    _ensure_stream()
    return _report(load().msim_callSim_CancerManagement, C.c_int32(n), C.c_double(odShape), C.c_double(discountRate),
                   C.c_double(partitionBy), C.c_int(int(indivp)), C.c_int64(first_person))


def cost_report(state, time, cost, discountRate=0.0, startReportAge=0.0, partition=(0.0, 100.0, 1.0), n_state=8):
    """CostReport(discountRate, startReportAge); setPartition(*partition); add(state[i], time[i], cost[i]) for all i
    (microsimulation.h:1198-1296).  Returns (wrap() as a dict of columns, the individual's running total)."""
    state = np.ascontiguousarray(state, dtype=np.int32)
    time = np.ascontiguousarray(time, dtype=np.float64)
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    t = Table()
    cur = C.c_double()
    dp = C.POINTER(C.c_double)
    check(load().msim_cost_report(float(discountRate), float(startReportAge), float(partition[0]), float(partition[1]),
                                  float(partition[2]), int(n_state), state.size, state.ctypes.data_as(C.POINTER(C.c_int32)),
                                  time.ctypes.data_as(dp), cost.ctypes.data_as(dp), C.byref(t), C.byref(cur)))
    d = t.to_dict()
    load().msim_table_free(C.byref(t))
    return d, cur.value


def calibration_sweep(seed, runpars, n, first_person=0):
    """n_sets parameter vectors x persons [first_person, first_person+n), one seed (common random numbers)."""
    runpars = np.ascontiguousarray(runpars, dtype=np.float64).reshape(-1, 6)
    ns = runpars.shape[0]
    arr = (Calibration * ns)()
    s = (C.c_double * 6)(*_seed6(seed))
    check(load().msim_calibration_sweep(s, C.c_int64(first_person), C.c_int64(n), C.c_int32(ns),
                                        runpars.ctypes.data_as(C.POINTER(C.c_double)), arr))
    return [arr[i].to_dict() for i in range(ns)]


def calibration_sweep_device(seed, runpars, n, first_person=0, set_lo=0, set_hi=None):
    """Sets [set_lo, set_hi) of `runpars` on persons [first_person, first_person + n), results left on the device as
    one vector of 48 doubles per set in which every cell is a sum.  Returns (device pointer, number of doubles)."""
    runpars = np.ascontiguousarray(runpars, dtype=np.float64).reshape(-1, 6)
    ns = runpars.shape[0]
    s = (C.c_double * 6)(*_seed6(seed))
    p = C.POINTER(C.c_double)()
    nd = C.c_int64()
    check(load().msim_calibration_sweep_device(s, C.c_int64(first_person), C.c_int64(n), C.c_int32(ns),
                                               runpars.ctypes.data_as(C.POINTER(C.c_double)), C.c_int32(set_lo),
                                               C.c_int32(ns if set_hi is None else set_hi), C.byref(p), C.byref(nd)))
    return C.cast(p, C.c_void_p).value, nd.value


def calibration_fetch(n_doubles):
    out = np.empty(n_doubles)
    check(load().msim_calibration_fetch(out.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(n_doubles)))
    return out


def calibration_unpack(vector, as_dicts=True):
    """The (merged) vector of a sweep -> the reference's per-set lists (pure host code)."""
    vector = np.ascontiguousarray(vector, dtype=np.float64)
    ns = vector.size // 48
    arr = (Calibration * ns)()
    check(load().msim_calibration_unpack(vector.ctypes.data_as(C.POINTER(C.c_double)), C.c_int32(ns), arr))
    return [arr[i].to_dict() for i in range(ns)] if as_dicts else arr


# ---- sharded / device-resident person-r runs (multi-GPU, benchmarks)

def person_shard(seed, first_person, n, outputs, discount_rate=0.0):
    return PersonShard((C.c_double * 6)(*_seed6(seed)), int(first_person), int(n), int(outputs), float(discount_rate))


def person_run_device(seed, first_person, n, outputs, discount_rate=0.0):
    """Runs a shard and leaves results on the device; returns the result struct."""
    sh = person_shard(seed, first_person, n, outputs, discount_rate)
    res = PersonDeviceResult()
    check(load().msim_person_run_device(C.byref(sh), C.byref(res)))
    return res


def person_enqueue(seed, first_person, n, outputs, discount_rate=0.0):
    sh = person_shard(seed, first_person, n, outputs, discount_rate)
    res = PersonDeviceResult()
    check(load().msim_person_enqueue(C.byref(sh), C.byref(res)))
    return res


def sync():
    check(load().msim_sync())


def last_kernel_ms():
    ms = C.c_float()
    check(load().msim_last_kernel_ms(C.byref(ms)))
    return ms.value


def timer_start():
    check(load().msim_timer_start())


def timer_stop():
    ms = C.c_float()
    check(load().msim_timer_stop(C.byref(ms)))
    return ms.value


def person_fetch_rowlog(copy=True):
    t = Table()
    check(load().msim_person_fetch_rowlog(C.byref(t)))
    d = t.to_dict(copy=copy)
    if copy:
        load().msim_table_free(C.byref(t))
        return d
    return d, t  # caller frees with free_table(t)


def free_table(t):
    load().msim_table_free(C.byref(t))


def person_fetch_counts():
    c = (C.c_double * 9)()
    check(load().msim_person_fetch_counts(c))
    return np.array(c[:])


def dense_fetch(n_doubles):
    out = np.empty(n_doubles)
    check(load().msim_dense_fetch(out.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(n_doubles)))
    return out


def dense_device_ptr():
    p = C.POINTER(C.c_double)()
    n = C.c_int64()
    check(load().msim_dense_device_ptr(C.byref(p), C.byref(n)))
    return C.cast(p, C.c_void_p).value, n.value


# ---- merge across the GPUs of a node inside the person-r launch (include/msim.h, msim_peer_*)

PEER_HANDLE_BYTES = 64


def peer_export(world):
    """This rank's mailbox: returns its CUDA IPC handle (bytes) for the other ranks."""
    h = (C.c_ubyte * PEER_HANDLE_BYTES)()
    check(load().msim_peer_export(C.c_int32(int(world)), h))
    return bytes(h)


def peer_import(rank, world, handles):
    """`handles`: every rank's peer_export(), in rank order."""
    blob = b"".join(handles)
    if len(blob) != PEER_HANDLE_BYTES * int(world):
        raise ValueError("one %d-byte handle per rank" % PEER_HANDLE_BYTES)
    buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
    check(load().msim_peer_import(C.c_int32(int(rank)), C.c_int32(int(world)), buf))


def peer_merge(mode=1):
    """0 / False: off; 1 / True: every report run ends with the sum over all ranks in its dense vector; 2: the sum of
    run k is formed during run k + 1 (nobody waits for a slower rank) or by peer_flush(), in peer_result_ptr()'s vector."""
    check(load().msim_peer_merge(C.c_int(int(mode))))


def peer_flush():
    check(load().msim_peer_flush())


def peer_result_ptr():
    p = C.POINTER(C.c_double)()
    n = C.c_int64()
    check(load().msim_peer_result_ptr(C.byref(p), C.byref(n)))
    return C.cast(p, C.c_void_p).value, n.value


def peer_close():
    check(load().msim_peer_close())


def dense_to_report(dense, dims, discount_rate=0.0):
    dense = np.ascontiguousarray(dense, dtype=np.float64)
    r = Report()
    check(load().msim_dense_to_report(dense.ctypes.data_as(C.POINTER(C.c_double)), C.byref(dims),
                                      C.c_double(discount_rate), C.byref(r)))
    d = r.to_dict()
    load().msim_report_free(C.byref(r))
    return d


def set_report_options(start=None, finish=100.0, delta=1.0, max_time=1.0e100, start_report_age=0.0, output_utilities=True):
    """The EventReport attached to person-r (and plugin) runs: setPartition(start, finish, delta, maxTime), the
    constructor's startReportAge and outputUtilities (microsimulation.h:873-904).  No arguments: the default (the built
    models' partition {0, 1, ..., 100, 1e6}, startReportAge 0, utilities on)."""
    if start is None:
        check(load().msim_set_report_options(None))
        return
    o = ReportOptions(float(start), float(finish), float(delta), float(max_time), float(start_report_age), int(bool(output_utilities)))
    check(load().msim_set_report_options(C.byref(o)))


def person_fetch_indiv(n):
    """wrap_indiv() and wrap_means() after a run with MSIM_OUT_REPORT | MSIM_OUT_INDIV: (vector of n totals, NaN = NA;
    dict n / mean / var / sd / se / sum / sumsq)."""
    out = np.empty(int(n))
    means = (C.c_double * 7)()
    check(load().msim_person_fetch_indiv(out.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(int(n)), means))
    return out, dict(zip(("n", "mean", "var", "sd", "se", "sum", "sumsq"), means[:]))


def person_report(seed, first_person, n, discount_rate=0.0, indiv=False):
    """Single-GPU person-r run with the EventReport attached; returns (report tables, counts) and, with indiv, the
    individuals' discounted totals and their means as well."""
    outs = MSIM_OUT_REPORT | MSIM_OUT_COUNTS | (MSIM_OUT_INDIV if indiv else 0)
    res = person_run_device(seed, first_person, n, outs, discount_rate)
    dense = dense_fetch(res.dims.n_doubles)
    rep, counts = dense_to_report(dense, res.dims, discount_rate), person_fetch_counts()
    if indiv:
        return (rep, counts) + person_fetch_indiv(n)
    return rep, counts
