"""Times the Sick-Sicker model (SummaryReport) in both RNG regimes and both event-queue kinds (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import microsimulation_b200 as m, microsimulation_b200.api as api
m.init(0)
par = api.sick_sicker_param(False)
for sched in ("staged", "generic"):
    m.set_person_schedule(sched)
    for n, sub in ((10**5, False), (10**6, True)):
        best, wall = 1e9, 1e9
        for _ in range(3):
            if sub:
                api.set_user_random_seed(12345)
            t0 = time.perf_counter()
            r = api.callSim(n, par, substreams=sub)
            wall = min(wall, time.perf_counter() - t0)
            best = min(best, api.last_kernel_ms())
        print("%-8s %s n=%.0e  kernels %.2f ms  whole call %.1f ms  mean cost %.1f" % (sched, "substreams" if sub else "R stream  ", n, best, wall * 1e3, r["indiv"]["costs"].mean()), flush=True)
