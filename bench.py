#!/usr/bin/env python
"""bench.py — simulated person-lifetimes per second, FHCRC prostate natural-history model
(the reference's src/person-r.cc), n = 1e8 persons per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--persons P] [--workload report|rowlog]

A "step" is one pass of the hot path over one cohort: every rank simulates its contiguous range of persons
(rank r: persons [r*P, (r+1)*P); person i draws from substream i+1 of the package-seed stream) and leaves its
results in HBM.  Weak scaling: P persons per GPU.

  --workload report (default, BASELINE configs[3] as SURVEY.md 8(d) makes it concrete: the in-tree natural-history
      model at n = 1e8 with an EventReport — person-time, events, prevalence by state x age, utilities discounted
      at 3% — and per-kind counts attached to every handled event; ranks merge with ONE NCCL all-reduce)
  --workload rowlog (BASELINE configs[2]: callPersonSimulation's own output, the life-history row log, 5 f64
      columns in person order; rows stay sharded)
  The JSON line is for the chosen workload and carries the other one's numbers under "other_workload".

  value  persons/s over all ranks, kernels only, results resident in HBM, device-timed (CUDA events on the
         library's stream), max over ranks.
  e2e    the same through the C-ABI calls a user makes, results in HOST memory (report: the dense accumulators
         copied back and turned into the reference's sparse tables; rowlog: every row copied to page-locked host
         memory), H2D of the arguments and D2H of the results inside the timed region.
  roofline / cpu_baseline / clocks: see DESIGN.md "Measurement".

--impl reference times the reference's own CPU path for the same model and the same outputs (oracle/_ref: the
reference's ssim.cc + RngStream.cpp + heap.h compiled where they lie, plus the restated Rcpp layer; falls back to
the portable oracle where oracle/_ref is absent) on all host cores, on a bounded sample.
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
ROWS_PER_PERSON = 1.2316  # oracle: 1.232 events (rows) per person at seed 12345
BYTES_PER_ROW = 40        # 5 f64 columns


# --------------------------------------------------------------------------- CPU reference arm

_ORACLES = {}


def _cpu_worker(args):
    flavour, first, n, workload = args
    if flavour not in _ORACLES:  # one library handle per worker process
        from oracle.pyoracle import Oracle
        _ORACLES[flavour] = Oracle(flavour)
    o = _ORACLES[flavour]
    t0 = time.perf_counter()
    if workload == "report":  # the same outputs as the GPU step: EventReport (3% discounting) + counts
        r = o.person_run(SEED, first, n, rowlog=False, report=True, discount_rate=0.03)
    else:  # builds the row log, as callPersonSimulation does
        r = o.person_run(SEED, first, n, rowlog=True)
    dt = time.perf_counter() - t0
    return dt, [int(x) for x in r["counts"][:8]]


def cpu_reference(sample_per_core, workload, cores=None, repeats=1):
    """The reference's CPU path on `cores` forked processes over contiguous person ranges (what its
    mclapply wrappers do; the engine is static-global, so processes, not threads).  Returns persons/s."""
    import multiprocessing as mp
    from oracle import pyoracle
    flavour = "ref" if pyoracle.build("ref") else "port"
    pyoracle.build("port")
    cores = cores or os.cpu_count() or 1
    best = None
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(flavour, 0, 1000, workload)] * cores)  # warm: load the library, page in
        for _ in range(repeats):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(flavour, k * sample_per_core, sample_per_core, workload) for k in range(cores)])
            wall = time.perf_counter() - t0
            v = cores * sample_per_core / wall
            best = v if best is None else max(best, v)
    kind = "reference" if flavour == "ref" else "port"
    counts = [sum(r[1][k] for r in res) for k in range(8)]  # per event kind, persons [0, cores * sample_per_core)
    return {"value": best, "unit": UNIT, "cores": cores, "kind": kind, "persons": cores * sample_per_core, "event_counts": counts,
            "sample": "%d persons per process x %d processes (contiguous person ranges, each jumping to its first substream), %s, seed 12345x6"
                      % (sample_per_core, cores, "EventReport (3 percent discounting) + counts" if workload == "report" else "row log built in memory")}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 measures the host
    sample = args.cpu_sample or (10 ** 6 if args.workload == "report" else 5 * 10 ** 6)
    t_steps = []
    base = None
    for i in range(args.warmup + args.steps):
        r = cpu_reference(sample, args.workload)
        base = r
        if i >= args.warmup:
            t_steps.append(r["value"])
    v = sum(t_steps) / len(t_steps)
    base["value"] = v
    persons_per_step = base["cores"] * sample
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": persons_per_step / v * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload) + "; CPU sample of %d persons per step" % persons_per_step,
                       "seed": "12345x6"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_name(w):
    if w == "report":
        return ("FHCRC natural-history model (person-r.cc), BASELINE configs[3] as SURVEY 8(d) states it: EventReport "
                "(pt/ut/events/prev by state x age, utilities discounted 3%) + per-kind counts on every event")
    return "FHCRC natural-history model (person-r.cc), BASELINE configs[2]: callPersonSimulation life-history row log"


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

def measure(args, workload, steps, rank, local, world, clocks_wanted):
    """One workload on this rank's GPU: device-timed value, e2e, launches, the top kernel's time."""
    import torch
    import torch.distributed as dist
    import microsimulation_b200 as m
    import microsimulation_b200.api as api
    P = args.persons
    first = rank * P
    report_mode = workload == "report"
    outputs = (api.MSIM_OUT_REPORT | api.MSIM_OUT_COUNTS) if report_mode else api.MSIM_OUT_ROWLOG
    rate = 0.03 if report_mode else 0.0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def merge():
        if world > 1 and report_mode:  # the single collective: accumulator vector (report cells + counts)
            from microsimulation_b200.distributed import _CudaBuffer
            ptr, n = api.dense_device_ptr()
            dist.all_reduce(torch.as_tensor(_CudaBuffer(ptr, n), device="cuda"))

    # ---- warm-up (>= 3 steps)
    for _ in range(max(args.warmup, 3)):
        api.person_run_device(SEED, first, P, outputs, rate)
        merge()
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

    # ---- e2e: the calls a user makes, results in HOST memory
    e2e_steps = max(1, min(steps, args.e2e_steps))

    def e2e_step():
        res = api.person_run_device(SEED, first, P, outputs, rate)
        if report_mode:
            merge()
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
    return {"value": world * P / (step_ms * 1e-3), "ms_per_step": step_ms, "rows": int(rows), "kernel_ms": kernel_ms_one,
            "launches": int(launches), "clocks": clocks,
            "e2e": {"value": world * P / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 6 * 8 + 3 * 8 + 4 + 8,
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": e2e_ms}}


def roofline_of(workload, meas, P):
    """The dominant kernel (person_staged_kernel) against the measured peaks.  The schema's bound is HBM or
    tensor; this kernel is neither — it is instruction-issue bound (Roofline) — so the HBM figures are
    given as asked and the issue-slot figures, which are the ones that mean something here, beside them."""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks \
        else (6650.0, "fallback (B200_PROFILING.md)")
    sm_mhz = peaks.get("sm_max_mhz", 1965.0)
    report_mode = workload == "report"
    # algorithmic bytes per launch: row log = 40 B per row out (+ 10 B per person and 24 B per
    # overflow row staged and read back); report = the accumulator vector, flushed once per block
    alg_bytes = meas["rows"] * BYTES_PER_ROW if not report_mode else 8 * (4 * 8 * 102 + 8 * 8 * 102 + 9)
    achieved = alg_bytes / (meas["kernel_ms"] * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
         "kernel": "person_staged_kernel" + ("" if report_mode else " + rowpass_*"), "kernel_ms": meas["kernel_ms"],
         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
         "note": "instruction-issue bound, not HBM bound: see 'issue' (warp instructions per person from the ncu "
                 "capture under profiles/, times persons, over the live CUDA-event time, against 148 SMs x 4 "
                 "schedulers x max SM clock)"}
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        try:
            tr = json.load(open(tr_path)).get("staged_report" if report_mode else "staged_rowlog")
            if tr:
                if tr.get("dram_bytes_per_person") is not None:
                    r["traffic"] = tr["dram_bytes_per_person"] * P
                wi = tr["warp_inst_per_person"] * P
                peak_issue = 148 * 4 * sm_mhz * 1e6
                ach = wi / (meas["kernel_ms"] * 1e-3)
                r["issue"] = {"achieved": ach / 1e9, "peak": peak_issue / 1e9, "unit": "G warp-inst/s", "frac": ach / peak_issue,
                              "warp_inst_per_person": tr["warp_inst_per_person"],
                              "ncu_issue_active_pct": tr.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                              "ncu_threads_per_warp_inst": tr.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
                              "source": "profiles/traffic.json (ncu --set full of the same kernel)"}
                lanes = tr.get("smsp__thread_inst_executed_per_inst_executed.ratio")
                if lanes:
                    # against the work of warps with zero divergence (thread instructions / 32)
                    r["issue"]["lane_efficiency"] = lanes / 32.0
                    r["issue"]["frac_divergence_free"] = r["issue"]["frac"] * lanes / 32.0
        except Exception as e:  # a missing profile must not break the bench
            r["issue_error"] = str(e)
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
    P = args.persons
    main = measure(args, args.workload, args.steps, rank, local, world, clocks_wanted=(rank == 0))
    other_name = "rowlog" if args.workload == "report" else "report"
    other = None if args.no_other_workload else measure(args, other_name, max(3, min(args.steps, 5)), rank, local, world, False)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    report_mode = args.workload == "report"
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_reference(args.cpu_sample or (10 ** 6 if report_mode else 5 * 10 ** 6), args.workload)
        # the persons the CPU reference just simulated, on the GPU: the per-kind event counts must be identical
        import microsimulation_b200.api as api
        api.person_run_device(SEED, 0, cpu["persons"], api.MSIM_OUT_COUNTS)
        gpu_counts = [int(x) for x in api.person_fetch_counts()[:8]]
        cpu["event_counts_equal_gpu"] = gpu_counts == cpu["event_counts"]
    line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload) + "; %d persons per GPU per step" % P,
                       "seed": "12345x6", "persons_per_gpu": P, "rows_per_step_rank0": main["rows"],
                       "merge": "one NCCL all-reduce(sum, f64) of the accumulator vector per step" if report_mode and world > 1
                                else ("none (one GPU)" if world == 1 else "none: row logs stay sharded"),
                       "l2": ("no input: each person is generated from (seed, index); outputs are KB-sized tables, nothing to flush"
                              if report_mode else
                              "no input reuse; each step writes a %.1f GB row log (>> 126 MB L2)" % (main["rows"] * 40 / 1e9))},
            "clocks": main["clocks"], "e2e": main["e2e"], "gpu_launches": main["launches"],
            "roofline": roofline_of(args.workload, main, P), "cpu_baseline": cpu}
    if other:
        line["other_workload"] = {"workload": workload_name(other_name), "value": other["value"], "unit": UNIT,
                                  "ms_per_step": other["ms_per_step"], "e2e": other["e2e"], "gpu_launches": other["launches"],
                                  "rows_per_step_rank0": other["rows"], "roofline": roofline_of(other_name, other, P)}
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
    ap.add_argument("--persons", type=int, default=10 ** 8, help="persons per GPU per step")
    ap.add_argument("--workload", default="report", choices=["rowlog", "report"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=0, help="persons per host process for the CPU legs (0: 1e6 report / 5e6 rowlog)")
    ap.add_argument("--no-other-workload", action="store_true", help="measure only the chosen workload")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
