"""Times the row-log step (and its final pass alone) for the default library and each variant in build/variants (development aid)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r)
import microsimulation_b200 as m, microsimulation_b200.api as api
m.init(0)
n = 10**8
best = 1e9
for _ in range(5):
    res = api.person_run_device(12345, 0, n, 1, 0.0)
    best = min(best, api.last_kernel_ms())
print("%%-28s rowlog step %%.3f ms  rows %%d" %% (os.path.basename(os.environ.get("MSIM_B200_LIB", "default")), best, res.n_rows), flush=True)
''' % ROOT
for lib in [None] + sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so"))):
    env = dict(os.environ)
    if lib:
        env["MSIM_B200_LIB"] = lib
    subprocess.run([sys.executable, "-c", code], env=env)
