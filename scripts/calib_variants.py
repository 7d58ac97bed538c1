"""Times the calibration sweep for the default library and each variant in build/variants (development aid)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r)
import numpy as np
import microsimulation_b200 as m, microsimulation_b200.api as api
m.init(0)
rs = np.random.RandomState(2024)
base = np.array([4, 0.5, 0.05, 10, 3, 0.5])
for n_sets, n in ((256, 10**6), (32, 10**6), (4, 10**6)):
    pars = base * (1 + 0.2 * (2 * rs.rand(n_sets, 6) - 1))
    best = 1e9
    for _ in range(5):
        api.calibration_sweep(12345, pars, n)
        best = min(best, api.last_kernel_ms())
    print("%%-28s %%4d sets x %%.0e: %%.3f ms  %%.3e person-simulations/s" %% (os.path.basename(os.environ.get("MSIM_B200_LIB", "default")), n_sets, n, best, n_sets * n / best * 1e3), flush=True)
''' % ROOT
for lib in [None] + sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so"))):
    env = dict(os.environ)
    if lib:
        env["MSIM_B200_LIB"] = lib
    subprocess.run([sys.executable, "-c", code], env=env)
