#!/usr/bin/env python
"""The instruction FLOOR of the person-r path: W = J + D (c_rng + c_samp) + E c_evt + B c_bin.

  ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none -k regex:microbench \
      --csv --log-file gpurun_out/inst_floor.csv python scripts/inst_floor.py run
  python scripts/inst_floor.py parse gpurun_out/inst_floor.csv profiles/traffic.json [build_hash]

`run` launches every building block of csrc/microbench.cuh once (full, converged warps; the product's own device
functions with the product's flags).  `parse` turns ncu's per-launch instruction counts into weights — thread
instructions per execution of the block, i.e. what one person costs when its warp does not diverge: (block - empty
loop) / (threads x repetitions) — multiplies them with the path's per-person counts and stores the floor, in warp
instructions per person (/ 32), under "instruction_floor" in profiles/traffic.json.

Per-person counts of person-r.cc at seed 12345 (SURVEY.md section 4 "oracle-derived KATs" and 8(d); n = 1e6):
  J = 1 substream jump;  D = 3.383 uniforms;  E = 1.231988 handled events (= report intervals B);
  Weibull draws (log + pow): 0.2241 onset (person-r.cc:100) + (105038 + 38786 + 20628)/1e6 progression = 0.388552;
  exp_rand: 1 (R::rexp(80), person-r.cc:102), 1.6932 uniforms each on average (Ahrens-Dieter: 1 + 0.3069 (1 + 1.2588)),
  so D - 1.6932 = 1.6898 uniforms are drawn outside exp_rand.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OPS = ["empty", "jump", "u01", "weibull", "exp_rand", "heap", "report", "exp"]
N_THREADS, REPS = 148 * 2048, 64
COUNTS = {"jump": 1.0, "u01": 3.383 - 1.6932, "exp_rand": 1.0, "weibull": 0.388552, "heap": 1.231988, "report": 1.231988}


def run():
    import ctypes as C
    import microsimulation_b200 as m
    m.init(0)
    L = m.load()
    ms = C.c_float()
    for op, name in enumerate(OPS):
        for _ in range(2):
            rc = L.msim_microbench(C.c_int(op), C.c_int64(N_THREADS), C.c_int(REPS), C.byref(ms))
            assert rc == 0, (name, rc)
        print("%-9s %.3f ms" % (name, ms.value))


def parse(csv_path, json_path, build_hash=None):
    rows = [r for r in csv.reader(open(csv_path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    iN, iM, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    per = {}
    for r in rows[1:]:
        name = r[iN]
        if "microbench_kernel" not in name:
            continue
        op = int(name.split("microbench_kernel<")[1].split(">")[0].replace("(int)", "").strip())
        per.setdefault(op, {})[r[iM]] = float(r[iV].replace(",", ""))  # the last launch of each block wins
    base = per[0]["smsp__thread_inst_executed.sum"]
    weights = {}
    for op, name in enumerate(OPS):
        if op == 0 or op not in per:
            continue
        weights[name] = (per[op]["smsp__thread_inst_executed.sum"] - base) / (N_THREADS * REPS)
    thread_inst = sum(COUNTS[k] * weights[k] for k in COUNTS)
    out = {"definition": "SURVEY.md 8(d): thread instructions one person needs when its warp never diverges, / 32",
           "weights_thread_inst_per_execution": weights, "per_person_counts": COUNTS,
           "thread_inst_per_person": thread_inst, "warp_inst_per_person": thread_inst / 32.0,
           "terms_warp_inst_per_person": {k: COUNTS[k] * weights[k] / 32.0 for k in COUNTS},
           "measured_with": "ncu smsp__thread_inst_executed.sum on csrc/microbench.cuh, %d threads x %d repetitions per block" % (N_THREADS, REPS)}
    try:
        cur = json.load(open(json_path))
    except Exception:
        cur = {}
    cur["instruction_floor"] = out
    if build_hash:
        cur["build_hash"] = build_hash
    with open(json_path, "w") as f:
        json.dump(cur, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run()
    else:
        parse(*sys.argv[2:5])
