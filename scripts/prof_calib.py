"""A few launches of the calibration sweep for ncu (development aid): python scripts/prof_calib.py [SETS] [PERSONS]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import microsimulation_b200 as m, microsimulation_b200.api as api
sets, n = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(float(sys.argv[2])) if len(sys.argv) > 2 else 10**5
m.init(0)
pars = np.tile(np.array([4, 0.5, 0.05, 10, 3, 0.5]), (sets, 1)) * (1 + 0.2 * (np.random.RandomState(1).rand(sets, 6) - 0.5))
for _ in range(3):
    m.calibration_sweep(12345, pars, n)
print("ok", api.last_kernel_ms(), "ms", sets * n / api.last_kernel_ms() * 1e3, "person-sims/s")
