"""Times the staged person-r kernel for each library variant in build/variants (development aid)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r)
import microsimulation_b200 as m, microsimulation_b200.api as api
m.init(0)
n = 10**8
api.set_person_schedule(int(os.environ.get("MSIM_SCHED", "0")))
for outputs, nm in ((4, "counts"), (2 | 4, "report+counts"), (1, "rowlog")):
    best = 1e9
    for _ in range(4):
        api.person_run_device(12345, 0, n, outputs, float(os.environ.get("MSIM_RATE", "0.03")))
        best = min(best, api.last_kernel_ms())
    print("%%-28s %%-14s %%.3f ms  %%.3e persons/s" %% (os.path.basename(os.environ.get("MSIM_B200_LIB", "default")) + " s" + os.environ.get("MSIM_SCHED", "0"), nm, best, n / best * 1e3), flush=True)
''' % ROOT
libs = [None] + sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so")))
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["MSIM_B200_LIB"] = lib
    for sched in os.environ.get("MSIM_SCHEDS", "0").split(","):
        env["MSIM_SCHED"] = sched
        subprocess.run([sys.executable, "-c", code], env=env)
