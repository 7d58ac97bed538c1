#!/usr/bin/env python
"""bench.py — simulated person-lifetimes per second, FHCRC prostate natural-history model
(the reference's src/person-r.cc), n = 1e8 persons.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--persons P]
                  [--workload report|rowlog|calib]

A "step" is one pass of the hot path over one cohort: every rank simulates its contiguous range of persons
(person i draws from substream i+1 of the package-seed stream) and leaves its results in HBM.

  --workload report (default, BASELINE configs[3] as SURVEY.md 8(d) makes it concrete: the in-tree natural-history
      model at n = 1e8 with an EventReport — person-time, events, prevalence by state x age, utilities discounted
      at 3% — and per-kind counts attached to every handled event; ranks merge with ONE NCCL all-reduce)
  --workload rowlog (BASELINE configs[2]: callPersonSimulation's own output, the life-history row log, 5 f64
      columns in person order; rows stay sharded)
  --workload calib (BASELINE configs[4]: calibperson-r.cc, 256 parameter sets x 1e6 persons sharded over the ranks —
      sets first, persons second — merged by one all-reduce; total work fixed)

The JSON line of the default run carries all of it: the headline is the report workload with P = 1e8 persons PER
GPU ("scaling": "weak", as in round 1, so numbers compare); "strong" is the same workload with 1e8 persons IN
TOTAL (BASELINE.json's metric: "1e8 persons at 1/2/4/8 B200"); "other_workload" is the row log; "calib" the sweep.

  value  persons/s over all ranks, kernels only, results resident in HBM, device-timed (CUDA events on the
         stream the library enqueues on), max over ranks.
  e2e    the same through the C-ABI calls a user makes, results in HOST memory (report: the dense accumulators
         copied back and turned into the reference's sparse tables; rowlog: every row copied to page-locked host
         memory), H2D of the arguments and D2H of the results inside the timed region.
  roofline / cpu_baseline / merge_check / clocks: see DESIGN.md "Measurement".

--impl reference times the reference's own CPU path for the same model and the same outputs (oracle/_ref: the
reference's ssim.cc + RngStream.cpp + heap.h compiled where they lie, plus the restated Rcpp layer; falls back to
the portable oracle where oracle/_ref is absent) on all host cores, on a bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = (12345,) * 6
METRIC = "simulated person-lifetimes/sec (FHCRC person-r.cc natural-history model, 1e8 persons per GPU)"
UNIT = "persons/s"
BYTES_PER_ROW = 40        # 5 f64 columns
TOTAL_STRONG = 10 ** 8    # BASELINE.json: 1e8 persons at 1/2/4/8 B200
N_SETS, CALIB_PERSONS = 256, 10 ** 6  # BASELINE configs[4]


def calib_lattice():
    """256 parameter vectors: base c(4, 0.5, 0.05, 10, 3, 0.5) (R/calibSim.R:10), every component on a +-20% lattice
    (4 levels for the first four components, the last two fixed): deterministic, no generator involved."""
    import numpy as np
    base = np.array([4, 0.5, 0.05, 10, 3, 0.5])
    levels = np.array([0.8, 0.9333333333333333, 1.0666666666666667, 1.2])
    out = []
    for a in levels:
        for b in levels:
            for c in levels:
                for d in levels:
                    out.append(base * np.array([a, b, c, d, 1.0, 1.0]))
    return np.array(out)


def workload_name(w):
    if w == "report":
        return ("FHCRC natural-history model (person-r.cc), BASELINE configs[3] as SURVEY 8(d) states it: EventReport "
                "(pt/ut/events/prev by state x age, utilities discounted 3%) + per-kind counts on every event")
    if w == "calib":
        return ("calibration sweep (calibperson-r.cc), BASELINE configs[4]: %d parameter sets x %d persons, sets then "
                "persons sharded over the ranks, one all-reduce" % (N_SETS, CALIB_PERSONS))
    return "FHCRC natural-history model (person-r.cc), BASELINE configs[2]: callPersonSimulation life-history row log"


# --------------------------------------------------------------------------- CPU reference arm

def cpu_person(sample_total, workload, cores=None, gpu_rows=None, pool=None):
    """The reference's CPU path on all host cores over contiguous person ranges (oracle/parallel.py).  persons/s is
    taken over the slowest worker's simulation time (merging and checking are not the reference's work)."""
    from oracle import parallel
    r = parallel.person_run_parallel(SEED, 0, sample_total, cores=cores, rows_to_check=gpu_rows,
                                     report=(workload == "report"), discount_rate=0.03 if workload == "report" else 0.0,
                                     pool=pool)
    kind = "reference" if r["flavour"] == "ref" else "port"
    out = {"value": sample_total / r["sim_seconds"], "unit": UNIT, "cores": r["cores"], "kind": kind, "persons": sample_total,
           "seconds": r["sim_seconds"], "event_counts": [int(x) for x in r["counts"][:8]],
           "sample": "persons [0, %d) over %d forked processes (contiguous ranges, each jumping to its first substream), %s, "
                     "seed 12345x6; the full workload is %s" % (
                         sample_total, r["cores"],
                         "EventReport (3 percent discounting) + counts" if workload == "report" else "row log built in memory",
                         "1e8 persons: same model, outputs and seed, the sample is a prefix of it (rate metric)")}
    return out, r


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 measures the host
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    if args.workload == "calib":
        from oracle import parallel
        pars = calib_lattice()[:max(cores, 1) * 2]
        vals, last = [], None
        for i in range(args.warmup + args.steps):
            last = parallel.calibration_parallel(SEED, 0, args.cpu_sample or 250000, pars, cores=cores)
            if i >= args.warmup:
                vals.append(last["person_simulations"] / last["seconds"])
        v = sum(vals) / len(vals)
        base = {"value": v, "unit": "person-simulations/s", "cores": last["cores"], "kind": "reference" if last["flavour"] == "ref" else "port",
                "sample": "%d parameter sets x %d persons over %d processes" % (len(pars), args.cpu_sample or 250000, last["cores"])}
        emit({"impl": "reference", "metric": "person-simulations/sec (calibration sweep)", "value": v, "unit": base["unit"],
              "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": last["person_simulations"] / v * 1e3,
              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
              "config": {"workload": workload_name("calib")}, "cpu_baseline": base,
              "e2e": {"value": v, "unit": base["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    sample = (args.cpu_sample or (10 ** 6 if args.workload == "report" else 5 * 10 ** 6)) * cores
    vals, base = [], None
    with mp.get_context("fork").Pool(cores) as pool:  # one pool for the whole run
        for i in range(args.warmup + args.steps):
            base, _ = cpu_person(sample, args.workload, cores=cores, pool=pool)
            if i >= args.warmup:
                vals.append(base["value"])
    v = sum(vals) / len(vals)
    base["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sample / v * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload) + "; CPU sample of %d persons per step" % sample,
                       "seed": "12345x6"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------- clocks

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        # under load = the upper half of the samples (idle samples before/after the timed region drag the median)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- our arm

PEER_MERGE = False  # the merge across ranks happens inside the simulation launch (microsimulation_b200.distributed.enable_peer_merge)


def merge_check(P, first, outputs, rate, world):
    """One step with the merge taken apart: every rank's accumulator vector is gathered BEFORE any merge, and the
    merged vector must equal their sum (integer cells exactly: counters are integer-valued doubles; FP sums to 1e-12)
    and account for every person exactly once (each dies once: toDeath or toPCDeath).  Checked for the NCCL
    all-reduce and, where it is in use, for the merge inside the launch — which must also leave the SAME bits on
    every rank (it adds in rank order)."""
    import torch
    import torch.distributed as dist
    import microsimulation_b200.api as api
    from microsimulation_b200.distributed import _CudaBuffer
    if PEER_MERGE:
        api.peer_flush()
        api.peer_merge(0)
    res = api.person_run_device(SEED, first, P, outputs, rate)
    ptr, n = api.dense_device_ptr()
    t = torch.as_tensor(_CudaBuffer(ptr, n), device="cuda")
    mine = t.clone()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    dist.all_reduce(t)
    total = torch.stack(parts).sum(0)
    n_fp = 2 * res.dims.n_state * res.dims.n_bin if n > 9 else 0  # pt and ut cells lead the vector (csrc/report.cuh)

    def against_total(v):
        ints_equal = bool(torch.equal(v[n_fp:-1], total[n_fp:-1]))
        fp = torch.cat([v[:n_fp], v[-1:]]), torch.cat([total[:n_fp], total[-1:]])
        fp_diff = float(((fp[0] - fp[1]).abs() / fp[1].abs().clamp_min(1e-300)).max()) if fp[0].numel() else 0.0
        return ints_equal, fp_diff, float(v[-9] + v[-8])

    ints_equal, fp_diff, deaths = against_total(t)
    out = {"allreduced_equals_sum_of_ranks_int_cells": ints_equal, "fp_cells_max_rel_diff": fp_diff,
           "deaths_equal_persons": deaths == float(world) * P, "persons": world * P, "ranks": world,
           "ok": ints_equal and fp_diff <= 1e-12 and deaths == float(world) * P}
    if PEER_MERGE:
        api.peer_merge(2)
        api.person_enqueue(SEED, first, P, outputs, rate)  # two steps: the first one's sum is formed inside the second launch,
        api.person_enqueue(SEED, first, P, outputs, rate)  # the second one's by the flush
        api.sync()
        ptr, n = api.peer_result_ptr()
        first_step = torch.as_tensor(_CudaBuffer(ptr, n), device="cuda").clone()
        api.peer_flush()
        api.sync()
        v = torch.as_tensor(_CudaBuffer(ptr, n), device="cuda").clone()
        i2, f2, d2 = against_total(v)
        every = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(every, v)
        same_bits = all(bool(torch.equal(e.view(torch.int64), v.view(torch.int64))) for e in every)
        # the two steps are the same simulation: integer cells equal, FP sums to rounding (within a block the adds into a
        # cell arrive in any order)
        same_steps = bool(torch.equal(first_step[n_fp:-1], v[n_fp:-1])) and bool(torch.allclose(first_step, v, rtol=1e-12, atol=0))
        out["in_launch_merge"] = {"equals_sum_of_ranks_int_cells": i2, "fp_cells_max_rel_diff": f2, "deaths_equal_persons": d2 == float(world) * P,
                                  "identical_bits_on_all_ranks": same_bits, "sum_formed_in_next_launch_equals_flushed_sum": same_steps}
        out["ok"] = out["ok"] and i2 and f2 <= 1e-12 and d2 == float(world) * P and same_bits and same_steps
    return out


def measure(args, workload, steps, P, rank, local, world, clocks_wanted, with_e2e=True):
    """One person-r workload on this rank's GPU: device-timed value, e2e, launches, the top kernel's time."""
    import torch
    import torch.distributed as dist
    import microsimulation_b200 as m
    import microsimulation_b200.api as api
    first = rank * P
    report_mode = workload == "report"
    outputs = (api.MSIM_OUT_REPORT | api.MSIM_OUT_COUNTS) if report_mode else api.MSIM_OUT_ROWLOG
    rate = 0.03 if report_mode else 0.0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def merge():
        if world > 1 and report_mode and not PEER_MERGE:  # the single collective: accumulator vector (report cells + counts)
            from microsimulation_b200.distributed import _CudaBuffer
            ptr, n = api.dense_device_ptr()
            dist.all_reduce(torch.as_tensor(_CudaBuffer(ptr, n), device="cuda"))

    def finish():
        # with the merge inside the launches, a launch adds up what the ranks sent in the launch before it (nobody waits
        # for a slower rank): the last step of a sequence needs one small launch more
        if world > 1 and report_mode and PEER_MERGE:
            api.peer_flush()

    # ---- warm-up (>= 3 steps)
    for _ in range(max(args.warmup, 3)):
        api.person_run_device(SEED, first, P, outputs, rate)
        merge()
    finish()
    rows = api.person_run_device(SEED, first, P, outputs, rate).n_rows
    kernel_ms_one = api.last_kernel_ms()  # CUDA events around this workload's kernels, one step

    # ---- timed region: K steps, device-timed with CUDA events on the stream everything is enqueued on
    # (the library runs on torch's current stream — api.set_stream — so the events also cover the all-reduce)
    sampler = ClockSampler(local)
    if clocks_wanted:
        sampler.start()
        time.sleep(0.25)
    launches0 = m.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        api.person_enqueue(SEED, first, P, outputs, rate)
        merge()
    finish()
    ev1.record()
    barrier()
    api.sync()
    dev_ms = ev0.elapsed_time(ev1)
    launches = m.launch_count() - launches0
    clocks = sampler.stop() if clocks_wanted else None
    t = torch.tensor([dev_ms / steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms = float(t[0])
    out = {"value": world * P / (step_ms * 1e-3), "ms_per_step": step_ms, "rows": int(rows), "kernel_ms": kernel_ms_one,
           "launches": int(launches), "clocks": clocks, "persons_per_gpu": P,
           "wave_share": api.last_wave_share()}  # rank 0's: the favoured blocks' share of the persons in the timed launches
    if world > 1 and report_mode:
        out["merge_check"] = merge_check(P, first, outputs, rate, world)
    if not with_e2e:
        return out

    # ---- e2e: the calls a user makes, results in HOST memory
    e2e_steps = max(1, min(steps, args.e2e_steps))

    def e2e_step():
        res = api.person_run_device(SEED, first, P, outputs, rate)
        if report_mode:
            merge()
            finish()
            if world > 1 and PEER_MERGE:
                from microsimulation_b200.distributed import _CudaBuffer
                ptr, nd = api.peer_result_ptr()
                dense = torch.as_tensor(_CudaBuffer(ptr, nd), device="cuda").cpu().numpy()  # D2H of the merged accumulators
            else:
                dense = api.dense_fetch(res.dims.n_doubles)   # D2H of the accumulators
            api.dense_to_report(dense, res.dims, rate)     # the reference's sparse tables (host)
            return dense.nbytes
        d, tab = api.person_fetch_rowlog(copy=False)       # D2H of every row into page-locked host memory
        nb = sum(v.nbytes for v in d.values())
        api.free_table(tab)  # returns the pinned buffer to the library's cache
        return nb

    e2e_step()  # first call pins the host buffer
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(e2e_steps):
        d2h = e2e_step()
    barrier()
    t = torch.tensor([(time.perf_counter() - t0) * 1e3 / e2e_steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[0])
    out["e2e"] = {"value": world * P / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 6 * 8 + 3 * 8 + 4 + 8,
                  "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": e2e_ms}
    return out


def measure_calib(args, steps, rank, world, with_cpu):
    """BASELINE configs[4]: 256 sets x 1e6 persons over the ranks.  Whole calls (the sweep checks a flag on the host
    after its kernel, so a call is synchronous): wall time per sweep, max over ranks, with the result vector merged on
    the device and unpacked into the reference's per-set lists on rank 0."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import microsimulation_b200 as m
    import microsimulation_b200.api as api
    from microsimulation_b200 import distributed as D
    pars = calib_lattice()

    def sweep(unpack):
        return D.calibration_sweep_sharded(SEED, pars, CALIB_PERSONS, unpack=unpack)

    # the same sweep through the event loop (every (person, set) pair simulated on its own: ordered event array),
    # which is what the default kernel — draws shared between the sets, no queue — replaces; reported beside it
    api.set_person_schedule("queue")
    try:
        for _ in range(2):
            sweep(False)
        event_loop_ms = api.last_kernel_ms()
    finally:
        api.set_person_schedule("staged")
    for _ in range(2):
        sweep(False)
    kernel_ms = api.last_kernel_ms()
    launches0 = m.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        vec = sweep(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([(time.perf_counter() - t0) * 1e3 / steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    launches = m.launch_count() - launches0
    lists = sweep(True)  # + D2H of the 98 KB vector and the per-set lists; the first call sets up the host side
    e2e_steps = max(1, min(steps, args.e2e_steps))
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        lists = sweep(True)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    gs, gp = D.calibration_grid(world, N_SETS)
    out = {"workload": workload_name("calib"), "value": N_SETS * CALIB_PERSONS / (ms * 1e-3), "unit": "person-simulations/s",
           "ms_per_step": ms, "kernel_ms_rank0": kernel_ms, "kernel": "calib_shared_kernel (csrc/calib_shared.cuh)",
           "event_loop_kernel_ms_rank0": event_loop_ms, "scaling": "strong", "gpu_launches": int(launches),
           "split": {"set_groups": gs, "person_groups": gp},
           "e2e": {"value": N_SETS * CALIB_PERSONS / (e2e_ms * 1e-3), "unit": "person-simulations/s", "ms_per_step": e2e_ms,
                   "steps": e2e_steps, "h2d_bytes_per_step": N_SETS * 48, "d2h_bytes_per_step": N_SETS * 48 * 8},
           "persons_counted_per_set": float(np.sum([lists[0].get(k, np.zeros(10))[0] for k in ("DiseaseFree", "Precursor", "PreClinical", "Clinical")]))}
    if with_cpu and rank == 0:
        from oracle import parallel
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_sample or 250000
        sub = pars[:2 * cores]
        ref = parallel.calibration_parallel(SEED, 0, n_cpu, sub, cores=cores)
        got = api.calibration_sweep(SEED, sub, n_cpu)
        same = True
        for g_, r_ in zip(got, ref["sets"]):
            same = same and set(g_) == set(r_)
            for k in r_:
                if k not in g_:
                    same = False
                elif k == "TimeAtRisk":
                    same = same and bool(np.allclose(g_[k], r_[k], rtol=1e-9, atol=0))
                else:
                    same = same and bool(np.array_equal(g_[k], r_[k]))
        out["cpu_baseline"] = {"value": ref["person_simulations"] / ref["seconds"], "unit": "person-simulations/s", "cores": ref["cores"],
                               "kind": "reference" if ref["flavour"] == "ref" else "port",
                               "sample": "%d parameter sets x %d persons over %d processes" % (len(sub), n_cpu, ref["cores"]),
                               "tables_equal_gpu": same}
    return out


def roofline_of(workload, meas, P):
    """The dominant kernel (person_staged_kernel) against the peak it is bound by: INSTRUCTION ISSUE (DESIGN.md,
    Roofline).  achieved = warp instructions the kernel executes per person (ncu capture of THIS
    build, profiles/traffic.json, keyed by the hash of the sources) x persons / the live CUDA-event time;
    peak = 148 SMs x 4 schedulers x max SM clock; frac = the divergence-free fraction SURVEY 8(d) defines
    (x threads per warp instruction / 32), frac_executed beside it.  The HBM figures the schema also asks for are
    given under "hbm"; they are tiny by construction (the report leaves 78 KB)."""
    from microsimulation_b200 import build as _build
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks \
        else (6650.0, "fallback (B200_PROFILING.md)")
    sm_mhz = peaks.get("sm_max_mhz", 1965.0)
    report_mode = workload == "report"
    kernel_s = meas["kernel_ms"] * 1e-3
    # algorithmic bytes per launch: row log = 40 B per row out; report = the accumulator vector
    alg_bytes = meas["rows"] * BYTES_PER_ROW if not report_mode else 8 * (4 * 8 * 102 + 8 * 8 * 102 + 9)
    peak_issue = 148 * 4 * sm_mhz * 1e6
    r = {"bound": "issue", "achieved": None, "peak": peak_issue / 1e9, "unit": "G warp-inst/s", "frac": None, "traffic": None,
         "kernel": "person_staged_kernel" + ("" if report_mode else " (+ rowpass_* for the row log: see hbm)"),
         "kernel_ms": meas["kernel_ms"], "peak_source": "148 SMs x 4 schedulers x %.0f MHz (MEASURED_PEAKS.json sm_max_mhz)" % sm_mhz,
         "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / kernel_s / 1e9, "peak_gbs": hbm_peak,
                 "frac": alg_bytes / kernel_s / 1e9 / hbm_peak, "peak_source": hbm_src}}
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        allp = json.load(open(tr_path))
        tr = allp.get("staged_report" if report_mode else "staged_rowlog")
        now_hash = _build.source_hash()
        wi = tr["warp_inst_per_person"]
        lanes = tr["smsp__thread_inst_executed_per_inst_executed.ratio"]
        ach = wi * P / kernel_s
        r["achieved"] = ach / 1e9
        r["frac_executed"] = ach / peak_issue
        r["frac"] = ach / peak_issue * lanes / 32.0
        r["warp_inst_per_person"] = wi
        r["threads_per_warp_inst"] = lanes
        r["ncu_issue_active_pct"] = tr.get("smsp__issue_active.avg.pct_of_peak_sustained_active")
        r["profile"] = {"file": "profiles/traffic.json", "captured_on_build": allp.get("build_hash"), "this_build": now_hash,
                        "matches_this_build": allp.get("build_hash") == now_hash, "captured_persons": tr.get("persons")}
        if tr.get("persons") == P and tr.get("dram_bytes") is not None:
            r["traffic"] = tr["dram_bytes"]  # dram__bytes_read + write of one launch of the same size
        else:
            r["traffic_note"] = "no ncu capture at %d persons per launch (DRAM traffic is not extrapolated across sizes)" % P
        floor = allp.get("instruction_floor")
        if floor:
            r["floor"] = dict(floor, frac_of_executed=floor["warp_inst_per_person"] / wi)
    except Exception as e:  # a missing profile must not break the bench
        r["profile_error"] = str(e)
    if not report_mode:
        r["bound"] = "issue (simulation kernel); the row passes are HBM bound"
    return r


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import microsimulation_b200 as m
    import microsimulation_b200.api as api
    m.init(local)
    api.set_stream(torch.cuda.current_stream().cuda_stream)  # kernels and the all-reduce on ONE stream
    global PEER_MERGE
    if world > 1 and not args.nccl_merge:
        from microsimulation_b200 import distributed as D
        PEER_MERGE = D.enable_peer_merge(mode=2)  # report runs are merged inside the launches (NVLink peer memory)
    P = args.persons
    few = max(3, min(args.steps, 5))
    if args.workload == "calib":
        c = measure_calib(args, max(3, min(args.steps, 10)), rank, world, not args.no_cpu_baseline)
        if rank == 0:
            emit({"metric": "person-simulations/sec (calibration sweep, calibperson-r.cc, %d sets x %d persons)" % (N_SETS, CALIB_PERSONS),
                  "value": c["value"], "unit": c["unit"], "n_gpus": world, "steps": args.steps, "warmup": 2, "ms_per_step": c["ms_per_step"],
                  "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                  "config": {"workload": c["workload"], "seed": "12345x6", "split": c["split"]}, "e2e": c["e2e"],
                  "gpu_launches": c["gpu_launches"], "cpu_baseline": c.get("cpu_baseline"), "kernel_ms_rank0": c["kernel_ms_rank0"]})
        if world > 1:
            dist.destroy_process_group()
        return
    main = measure(args, args.workload, args.steps, P, rank, local, world, clocks_wanted=(rank == 0))
    other_name = "rowlog" if args.workload == "report" else "report"
    other = None if args.no_other_workload else measure(args, other_name, few, P, rank, local, world, False)
    # strong scaling of the headline workload: 1e8 persons in total (at one GPU that IS the headline run)
    strong = None
    if world > 1 and not args.no_strong and TOTAL_STRONG % world == 0:
        strong = measure(args, args.workload, args.steps, TOTAL_STRONG // world, rank, local, world, False)
        if PEER_MERGE and args.workload == "report":
            # the same steps with the merge as a separate NCCL all-reduce after every launch, for comparison
            api.peer_flush()
            api.peer_merge(0)
            PEER_MERGE = False
            try:
                strong["nccl_ms_per_step"] = measure(args, args.workload, few, TOTAL_STRONG // world, rank, local, world, False)["ms_per_step"]
            finally:
                api.peer_merge(2)
                PEER_MERGE = True
    if PEER_MERGE:
        api.peer_flush()
        api.peer_merge(0)  # what follows runs on rank 0 alone (CPU baseline checks) or does not involve person-r
    calib = None
    if not args.no_calib:
        calib = measure_calib(args, 3, rank, world, with_cpu=not args.no_cpu_baseline)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    report_mode = args.workload == "report"
    cpu = None
    if not args.no_cpu_baseline:
        import numpy as np
        from oracle import parallel
        cores = os.cpu_count() or 1
        sample = (args.cpu_sample or (10 ** 6 if report_mode else 5 * 10 ** 6)) * cores
        # the persons the CPU reference simulates, on the GPU: every cell of the report (or every row) is compared
        if report_mode:
            cpu, ref = cpu_person(sample, "report", cores=cores)
            rep, counts = api.person_report(SEED, 0, sample, 0.03)
            ok, worst, msg = parallel.report_equal(rep, ref["report"], rtol=1e-9)
            cpu["event_counts_equal_gpu"] = [int(x) for x in counts[:8]] == cpu["event_counts"]
            cpu["report_equal_gpu"] = bool(ok)
            cpu["report_cells_compared"] = int(sum(len(rep[tb][k]) for tb, k in (("pt", "pt"), ("ut", "utility"), ("events", "number"), ("prev", "number"))))
            cpu["report_fp_max_rel_diff"] = worst
            cpu["report_check"] = msg or "events and prev tables bit-equal, pt and ut within 1e-9"
        else:
            api.person_run_device(SEED, 0, sample, api.MSIM_OUT_ROWLOG | api.MSIM_OUT_COUNTS)
            rows = api.person_fetch_rowlog()
            counts = api.person_fetch_counts()
            cpu, ref = cpu_person(sample, "rowlog", cores=cores, gpu_rows=rows)
            cpu["event_counts_equal_gpu"] = [int(x) for x in counts[:8]] == cpu["event_counts"]
            cpu["rows_equal_gpu"] = bool(ref["rows_ok"])
            cpu["rows_compared"] = int(ref["n_rows"])
            cpu["rows_max_time_rel_diff"] = ref["max_time_rel_err"]
    line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload) + "; %d persons per GPU per step" % P,
                       "seed": "12345x6", "persons_per_gpu": P, "rows_per_step_rank0": main["rows"],
                       "wave_share_rank0": main["wave_share"],
                       "merge": (("inside the simulation launches: the block that finishes the fold sends the accumulator vector to the other "
                                  "ranks over NVLink peer memory; the next launch (or, after the last step, one small launch) adds the "
                                  "ranks' vectors in rank order" if PEER_MERGE else
                                  "one NCCL all-reduce(sum, f64) of the accumulator vector per step") if report_mode and world > 1
                                 else ("none (one GPU)" if world == 1 else "none: row logs stay sharded")),
                       "l2": ("no input: each person is generated from (seed, index); outputs are KB-sized tables, nothing to flush"
                              if report_mode else
                              "no input reuse; each step writes a %.1f GB row log (>> 126 MB L2)" % (main["rows"] * 40 / 1e9))},
            "clocks": main["clocks"], "e2e": main["e2e"], "gpu_launches": main["launches"],
            "roofline": roofline_of(args.workload, main, P), "cpu_baseline": cpu}
    if "merge_check" in main:
        line["merge_check"] = main["merge_check"]
    if world == 1 and P == TOTAL_STRONG:
        line["strong"] = {"persons_total": TOTAL_STRONG, "persons_per_gpu": P, "value": main["value"], "ms_per_step": main["ms_per_step"],
                          "e2e": main["e2e"], "note": "one GPU: the headline run is the strong-scaling base"}
    elif strong:
        line["strong"] = {"persons_total": TOTAL_STRONG, "persons_per_gpu": strong["persons_per_gpu"], "value": strong["value"],
                          "unit": UNIT, "ms_per_step": strong["ms_per_step"], "kernel_ms_rank0": strong["kernel_ms"], "e2e": strong["e2e"],
                          "gpu_launches": strong["launches"], "merge_check": strong.get("merge_check"),
                          "nccl_ms_per_step": strong.get("nccl_ms_per_step"),
                          "note": "BASELINE.json's metric: 1e8 persons in total over the ranks; per step = one launch that simulates and "
                                  "merges (nccl_ms_per_step: the same with the merge as a separate NCCL all-reduce)" if PEER_MERGE else
                                  "BASELINE.json's metric: 1e8 persons in total over the ranks; per step = kernel + one all-reduce"}
    if other:
        line["other_workload"] = {"workload": workload_name(other_name), "value": other["value"], "unit": UNIT,
                                  "ms_per_step": other["ms_per_step"], "e2e": other["e2e"], "gpu_launches": other["launches"],
                                  "rows_per_step_rank0": other["rows"], "roofline": roofline_of(other_name, other, P)}
    if calib:
        line["calib"] = calib
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """stdout carries ONE line, the JSON: whatever libraries print there (NCCL's version banner, ...) goes to
    stderr instead.  The JSON is written to the saved descriptor by emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        sys.stdout.buffer.write(data)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--persons", type=int, default=10 ** 8, help="persons per GPU per step (the weak-scaling headline)")
    ap.add_argument("--workload", default="report", choices=["rowlog", "report", "calib"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=0, help="persons per host process for the CPU legs (0: 1e6 report / 5e6 rowlog / 2.5e5 calib)")
    ap.add_argument("--no-other-workload", action="store_true", help="measure only the chosen workload")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    ap.add_argument("--nccl-merge", action="store_true", help="merge with an NCCL all-reduce after every launch instead of inside the launch")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (1e8 persons in total) at N > 1")
    ap.add_argument("--no-calib", action="store_true", help="skip the calibration-sweep leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
