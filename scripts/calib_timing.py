#!/usr/bin/env python
"""Calibration sweep (BASELINE configs[4]: 256 parameter sets x 1e6 persons) under each kernel: shared draws without a
queue (default), event loop with an ordered array, event loop with the binary heap; and a few other shapes."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import microsimulation_b200.api as api  # noqa: E402

api.init(0)
rs = np.random.RandomState(2024)
base = np.array([4, 0.5, 0.05, 10, 3, 0.5])
shapes = [(256, 10 ** 6), (4, 10 ** 6), (1, 10 ** 7), (1024, 10 ** 5), (40, 10 ** 6)]
for sched in ("staged", "queue", "generic"):
    api.set_person_schedule(sched)
    for n_sets, n in shapes if sched == "staged" else shapes[:2]:
        pars = base * (1 + 0.2 * (2 * rs.rand(n_sets, 6) - 1))
        api.calibration_sweep(12345, pars[: min(4, n_sets)], 10 ** 4)
        best, wall = 1e9, 1e9
        for _ in range(3):
            t0 = time.time()
            api.calibration_sweep(12345, pars, n)
            wall = min(wall, time.time() - t0)
            best = min(best, api.last_kernel_ms())
        print("%-8s %4d sets x %.0e persons: kernels %8.3f ms  whole call %7.1f ms  %.3e person-simulations/s"
              % (sched, n_sets, n, best, wall * 1e3, n_sets * n / best * 1e3), flush=True)
api.set_person_schedule("staged")
