"""Times the Sick-Sicker kernels for each library variant in build/variants (development aid)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r)
import microsimulation_b200 as m, microsimulation_b200.api as api
m.init(0)
par = api.sick_sicker_param(False)
for n, sub in ((10**5, False), (10**6, True)):
    best = 1e9
    for _ in range(3):
        if sub:
            api.set_user_random_seed(12345)
        r = api.callSim(n, par, substreams=sub)
        best = min(best, api.last_kernel_ms())
    print("%%-30s %%s n=%%.0e kernels %%.2f ms" %% (os.path.basename(os.environ.get("MSIM_B200_LIB", "default")), "substreams" if sub else "R stream  ", n, best), flush=True)
''' % ROOT
for lib in [None] + sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*sick*.so"))):
    env = dict(os.environ)
    if lib:
        env["MSIM_B200_LIB"] = lib
    subprocess.run([sys.executable, "-c", code], env=env)
