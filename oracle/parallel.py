"""ORACLE — TEST INFRASTRUCTURE ONLY.

The CPU oracle on several host cores, the way the reference's own wrappers use
several cores (R/calibSim.R:16-39 `mclapply` over person chunks): forked
processes over contiguous person ranges — the engine is static-global
(src/ssim.cc:35-79), so processes, not threads.  Unlike the R wrappers, every
chunk jumps to its first person's substream, so the merged result is the
single-process result (integer tables exactly, FP sums to rounding).

Only tests/ and bench.py's CPU legs import this (never the product).
"""
import multiprocessing as mp
import os
import time

import numpy as np

_ORACLES = {}
_ROWS = None  # set by the parent before forking: the row log every worker checks its chunk against


def flavour_available():
    from . import pyoracle
    return "ref" if pyoracle.build("ref") else "port"


def _oracle(flavour):
    if flavour not in _ORACLES:  # one library handle per worker process
        from .pyoracle import Oracle
        _ORACLES[flavour] = Oracle(flavour)
    return _ORACLES[flavour]


def _worker(args):
    flavour, seed, first, n, want_rows, want_report, rate, time_rtol = args
    o = _oracle(flavour)
    t0 = time.perf_counter()
    r = o.person_run(seed, first, n, rowlog=want_rows, report=want_report, discount_rate=rate)
    dt = time.perf_counter() - t0
    out = {"seconds": dt, "counts": r["counts"], "report": r.get("report"), "n_rows": None, "rows_ok": None, "rows_msg": ""}
    if want_rows:
        mine = r["rowlog"]
        out["n_rows"] = len(mine["id"])
        if _ROWS is not None:  # compare this chunk with the rows the parent holds for the same persons
            ids = _ROWS["id"]
            lo, hi = np.searchsorted(ids, first, "left"), np.searchsorted(ids, first + n, "left")
            ok, msg = (hi - lo) == len(mine["id"]), ""
            if not ok:
                msg = "persons [%d, %d): %d rows against %d" % (first, first + n, hi - lo, len(mine["id"]))
            else:
                for k in ("id", "state", "event"):
                    if not np.array_equal(_ROWS[k][lo:hi], mine[k]):
                        ok, msg = False, "column %s differs in persons [%d, %d)" % (k, first, first + n)
                        break
                if ok:
                    worst = 0.0
                    for k in ("startTime", "endTime"):
                        a, b = _ROWS[k][lo:hi], mine[k]
                        with np.errstate(divide="ignore", invalid="ignore"):
                            rel = np.where(b != 0, np.abs(a - b) / np.abs(b), np.abs(a - b))
                        worst = max(worst, float(rel.max()) if len(rel) else 0.0)
                    out["max_time_rel_err"] = worst
                    if worst > time_rtol:
                        ok, msg = False, "event times differ by %.3g relative in persons [%d, %d)" % (worst, first, first + n)
            out["rows_ok"], out["rows_msg"] = bool(ok), msg
    return out


def _calib_worker(args):
    flavour, seed, first, n, runpars = args
    o = _oracle(flavour)
    t0 = time.perf_counter()
    r = o.calibration_sweep(seed, first, n, runpars)
    return time.perf_counter() - t0, r


def calibration_parallel(seed, first, n, runpars, cores=None, flavour=None):
    """The calibration model (src/calibperson-r.cc) for every parameter vector of `runpars` on persons
    [first, first + n), the sets dealt out over `cores` forked processes.  Returns the per-set lists in order, the
    wall time of the map, and the number of person simulations."""
    cores = cores or os.cpu_count() or 1
    flavour = flavour or flavour_available()
    runpars = np.ascontiguousarray(runpars, dtype=np.float64).reshape(-1, 6)
    groups = [list(range(k, runpars.shape[0], cores)) for k in range(cores)]
    groups = [g for g in groups if g]
    with mp.get_context("fork").Pool(len(groups)) as pool:
        pool.map(_calib_worker, [(flavour, seed, 0, 10, runpars[:1])] * len(groups))  # load the library
        t0 = time.perf_counter()
        res = pool.map(_calib_worker, [(flavour, seed, first, n, runpars[g]) for g in groups])
        wall = time.perf_counter() - t0
    out = [None] * runpars.shape[0]
    for g, (_, lists) in zip(groups, res):
        for idx, d in zip(g, lists):
            out[idx] = d
    return {"flavour": flavour, "cores": len(groups), "seconds": wall, "person_simulations": n * runpars.shape[0], "sets": out}


_KEYS = {"pt": ("Key", "age"), "ut": ("Key", "age"), "events": ("Key", "event", "age"), "prev": ("Key", "age")}
_VALUE = {"pt": "pt", "ut": "utility", "events": "number", "prev": "number"}


def merge_reports(reports):
    """Sum of sparse EventReport tables: union of the key tuples, values added, rows in std::map order
    (microsimulation.h:974-986 sorts by Key, [event,] age)."""
    out = {}
    for tb, keys in _KEYS.items():
        val = _VALUE[tb]
        acc = {}
        for r in reports:
            t = r[tb]
            cols = [t[k] for k in keys]
            for i in range(len(t[val])):
                kt = tuple(float(c[i]) for c in cols)
                acc[kt] = acc.get(kt, 0.0) + float(t[val][i])
        order = sorted(acc)
        out[tb] = {k: np.array([kt[j] for kt in order]) for j, k in enumerate(keys)}
        out[tb][val] = np.array([acc[kt] for kt in order])
    return out


def person_run_parallel(seed, first, n, cores=None, flavour=None, rows_to_check=None, report=True, discount_rate=0.03,
                        time_rtol=4e-15, pool=None):
    """Persons [first, first + n) on `cores` forked processes.  Returns merged counts, the merged report (if asked
    for), the wall time of the map, and — when `rows_to_check` (a row log dict sorted by id) is given — whether
    every chunk's rows equal the corresponding rows of it (ids, states, events exactly; times to time_rtol)."""
    global _ROWS
    cores = cores or os.cpu_count() or 1
    flavour = flavour or flavour_available()
    per = (n + cores - 1) // cores
    chunks = [(first + k * per, max(0, min(per, n - k * per))) for k in range(cores)]
    chunks = [c for c in chunks if c[1] > 0]
    _ROWS = rows_to_check
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(cores)  # forked after _ROWS is set: the children see it copy-on-write
    try:
        pool.map(_worker, [(flavour, seed, 0, 100, False, False, 0.0, time_rtol)] * cores)  # load the library, page in
        t0 = time.perf_counter()
        res = pool.map(_worker, [(flavour, seed, f, c, rows_to_check is not None, report, discount_rate, time_rtol)
                                 for f, c in chunks])
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.close()
            pool.join()
        _ROWS = None
    # seconds = wall time of the map (simulation + merging the workers' results + any row comparison);
    # sim_seconds = the slowest worker's simulation alone, which is what a throughput figure should use
    out = {"flavour": flavour, "cores": len(chunks), "seconds": wall, "sim_seconds": max(r["seconds"] for r in res),
           "persons": n, "counts": np.sum([r["counts"] for r in res], axis=0)}
    if report:
        out["report"] = merge_reports([r["report"] for r in res])
    if rows_to_check is not None:
        out["n_rows"] = int(sum(r["n_rows"] for r in res))
        out["rows_ok"] = all(r["rows_ok"] for r in res)
        out["rows_msg"] = "; ".join(r["rows_msg"] for r in res if r["rows_msg"])
        out["max_time_rel_err"] = max([r.get("max_time_rel_err", 0.0) for r in res] + [0.0])
    return out


def report_equal(a, b, rtol=1e-9):
    """events / prev tables bit-equal (keys and numbers), pt / ut keys bit-equal and sums within rtol.  Returns
    (ok, largest relative difference of an FP sum, message)."""
    worst = 0.0
    for tb, keys in _KEYS.items():
        for k in keys:
            if a[tb][k].shape != b[tb][k].shape or not np.array_equal(a[tb][k], b[tb][k]):
                return False, worst, "%s: key column %s differs" % (tb, k)
        val = _VALUE[tb]
        if tb in ("events", "prev"):
            if not np.array_equal(a[tb][val], b[tb][val]):
                return False, worst, "%s: counts differ" % tb
        else:
            x, y = a[tb][val], b[tb][val]
            with np.errstate(divide="ignore", invalid="ignore"):
                rel = np.where(y != 0, np.abs(x - y) / np.abs(y), np.abs(x - y))
            if len(rel):
                worst = max(worst, float(rel.max()))
            if worst > rtol:
                return False, worst, "%s: sums differ by %.3g relative" % (tb, worst)
    return True, worst, ""
