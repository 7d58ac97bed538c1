"""Kernel timings of the main paths (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import microsimulation_b200 as m
import microsimulation_b200.api as api
m.init(0)
for sched in ("staged", "generic"):
  api.set_person_schedule(sched)
  for n in (10**7, 10**8):
    for outputs, nm in ((4, "counts"), (1, "rowlog"), (2 | 4, "report+counts"), (1 | 2 | 4, "all")):
        best = 1e9
        for rep_ in range(3):
            res = api.person_run_device(12345, 0, n, outputs, 0.03)
            best = min(best, api.last_kernel_ms())
        print("%s person n=%g %-14s %.3f ms  %.3e persons/s rows=%d" % (sched, n, nm, best, n / best * 1e3, res.n_rows), flush=True)
api.set_person_schedule("staged")
t0 = time.time(); g = m.callPersonSimulation(10**7); t1 = time.time()
t0 = time.time(); g = m.callPersonSimulation(10**7); t1 = time.time()
print("e2e callPersonSimulation(1e7) warm: %.3f s rows %d" % (t1 - t0, len(g["id"])))
for n in (10**5, 10**6):
    m.callIllnessDeath(n); t0 = time.time(); m.callIllnessDeath(n); t1 = time.time(); print("illness n=%g: %.4f s pipeline %.3f ms" % (n, t1 - t0, api.last_kernel_ms()))
    m.callSimplePerson(n); t0 = time.time(); m.callSimplePerson(n); t1 = time.time(); print("simple n=%g: %.4f s pipeline %.3f ms" % (n, t1 - t0, api.last_kernel_ms()))
    m.callSimplePerson2(n); t0 = time.time(); m.callSimplePerson2(n); t1 = time.time(); print("simple2 n=%g: %.4f s pipeline %.3f ms" % (n, t1 - t0, api.last_kernel_ms()))
pars = np.tile(np.array([4, 0.5, 0.05, 10, 3, 0.5]), (256, 1)) * (1 + 0.2 * (np.random.RandomState(1).rand(256, 6) - 0.5))
m.calibration_sweep(12345, pars[:4], 10**5)
t0 = time.time(); out = m.calibration_sweep(12345, pars, 10**6); t1 = time.time()
print("calib sweep 256x1e6: %.3f s kernel %.3f ms  %.3e person-sims/s" % (t1 - t0, api.last_kernel_ms(), 256e6 / api.last_kernel_ms() * 1e3))
