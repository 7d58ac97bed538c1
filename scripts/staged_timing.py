"""Kernel timings of the person-r staged schedule in every output mode (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import microsimulation_b200 as m
import microsimulation_b200.api as api
m.init(0)
sizes = [int(float(x)) for x in sys.argv[1:]] or [10**7, 10**8]
for n in sizes:
    for outputs, nm in ((4, "counts"), (1, "rowlog"), (2 | 4, "report+counts"), (1 | 2 | 4, "all")):
        best = 1e9
        for rep_ in range(4):
            res = api.person_run_device(12345, 0, n, outputs, 0.03)
            best = min(best, api.last_kernel_ms())
        print("staged person n=%g %-14s %.3f ms  %.3e persons/s rows=%d" % (n, nm, best, n / best * 1e3, res.n_rows), flush=True)
