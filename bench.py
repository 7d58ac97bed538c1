#!/usr/bin/env python
"""bench.py — simulated person-lifetimes per second, FHCRC prostate natural-history model
(the reference's src/person-r.cc through callPersonSimulation), n = 1e8 persons per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--persons P] [--workload rowlog|report]

A "step" is one pass of the hot path over one cohort: every rank simulates its contiguous range of
persons (rank r: persons [r*P, (r+1)*P), substream i+1 of the package-seed stream for person i) and
leaves the life-history row log (5 f64 columns, person order) in HBM.  Weak scaling: P persons per GPU.

  value  persons/s over all ranks, kernels only, results resident in HBM, device-timed (CUDA events on
         the library's stream), max over ranks.
  e2e    the same through the C-ABI call a user makes (seed and n in, the row log in page-locked host
         memory out: H2D of the arguments and D2H of every row inside the timed region).
  roofline / cpu_baseline / clocks: see DESIGN.md "Measurement".

--impl reference times the reference's own CPU path for the same model (oracle/_ref: the reference's
ssim.cc + RngStream.cpp + heap.h compiled where they lie, plus the restated Rcpp layer; falls back to
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
    flavour, first, n = args
    if flavour not in _ORACLES:  # one library handle per worker process
        from oracle.pyoracle import Oracle
        _ORACLES[flavour] = Oracle(flavour)
    o = _ORACLES[flavour]
    t0 = time.perf_counter()
    r = o.person_run(SEED, first, n, rowlog=True)  # builds the row log, as callPersonSimulation does
    dt = time.perf_counter() - t0
    return dt, int(r["counts"][:8].sum())


def cpu_reference(sample_per_core, cores=None, repeats=1):
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
        pool.map(_cpu_worker, [(flavour, 0, 1000)] * cores)  # warm: load the library, page in
        for _ in range(repeats):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(flavour, k * sample_per_core, sample_per_core) for k in range(cores)])
            wall = time.perf_counter() - t0
            v = cores * sample_per_core / wall
            best = v if best is None else max(best, v)
    kind = "reference" if flavour == "ref" else "port"
    return {"value": best, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d persons per process x %d processes, row log built in memory, seed 12345x6" % (sample_per_core, cores)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 measures the host
    sample = args.cpu_sample
    t_steps = []
    base = None
    for i in range(args.warmup + args.steps):
        r = cpu_reference(sample)
        base = r
        if i >= args.warmup:
            t_steps.append(r["value"])
    v = sum(t_steps) / len(t_steps)
    base["value"] = v
    persons_per_step = base["cores"] * sample
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": persons_per_step / v * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "person-r.cc callPersonSimulation, row log; CPU sample of %d persons per step" % persons_per_step,
                       "seed": "12345x6"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


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
    P = args.persons
    first = rank * P
    report_mode = args.workload == "report"
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
            t = torch.as_tensor(_CudaBuffer(ptr, n), device="cuda")
            dist.all_reduce(t)

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        api.person_run_device(SEED, first, P, outputs, rate)
        merge()
    rows = api.person_run_device(SEED, first, P, outputs, rate).n_rows
    kernel_ms_one = api.last_kernel_ms()

    # ---- timed region: K steps, device-timed on the library's stream
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = m.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    if report_mode and world > 1:
        dev_ms = 0.0
        for _ in range(args.steps):  # the all-reduce runs on torch's stream: time step by step
            api.timer_start()
            api.person_enqueue(SEED, first, P, outputs, rate)
            dev_ms += api.timer_stop()
            merge()
        torch.cuda.synchronize()
        api.sync()
    else:
        api.timer_start()
        for _ in range(args.steps):
            api.person_enqueue(SEED, first, P, outputs, rate)
        dev_ms = api.timer_stop()
        api.sync()
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = m.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = (wall_ms if (report_mode and world > 1) else dev_ms) / args.steps
    t = torch.tensor([step_ms, float(launches)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms = float(t[0])
    value = world * P / (step_ms * 1e-3)

    # ---- e2e: the call a user makes — rows end up in page-locked HOST memory
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def e2e_step():
        api.person_run_device(SEED, first, P, outputs, rate)
        if report_mode:
            merge()
            dense = api.dense_fetch(api.dense_device_ptr()[1])
            return dense.nbytes
        d, tab = api.person_fetch_rowlog(copy=False)
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
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * P / (float(t[0]) * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (cohort_kernel<PersonTraits>), measured live
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = (rows * BYTES_PER_ROW + ((P + 255) // 256) * 16) if not report_mode else 8 * 14840 * 2
    achieved = alg_bytes / (kernel_ms_one * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "cohort_kernel<PersonTraits,false>", "kernel_ms": kernel_ms_one,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "note": "the kernel is instruction-issue bound, not HBM bound (DESIGN.md); see profiles/ for issue-slot utilisation"}
    ncu_traffic = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(ncu_traffic):
        try:
            tr = json.load(open(ncu_traffic))
            roofline["traffic"] = tr.get("dram_bytes_per_person", 0) * P
            roofline["issue"] = tr.get("issue")
        except Exception:
            pass

    cpu = cpu_reference(args.cpu_sample) if not args.no_cpu_baseline else None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "person-r.cc callPersonSimulation (BASELINE configs[2]/[3] model), %s, %d persons per GPU per step"
                                   % ("EventReport tables + counts, one NCCL all-reduce" if report_mode else "life-history row log", P),
                       "seed": "12345x6", "persons_per_gpu": P, "rows_per_step_rank0": int(rows),
                       "l2": "no input reuse; each step writes a %.1f GB row log (>> 126 MB L2)" % (rows * 40 / 1e9)
                             if not report_mode else "tables are KB-sized; nothing to flush"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 6 * 8 + 3 * 8, "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "ms_per_step": float(t[0])},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--persons", type=int, default=10 ** 8, help="persons per GPU per step")
    ap.add_argument("--workload", default="report", choices=["rowlog", "report"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=5 * 10 ** 6, help="persons per host process for the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
