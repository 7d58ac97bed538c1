"""Development aid: kernel times of nine successive person-r runs per output mode at 1e8 persons, for the library named by
MSIM_B200_LIB (default: the in-tree build) — variants are compared within ONE gpurun call: boxes differ by a few per cent."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import microsimulation_b200 as m, microsimulation_b200.api as api
m.init(0)
n = 10**8
for outputs, nm in ((6, "report+counts"), (4, "counts"), (1, "rowlog")):
    ms = []
    for _ in range(9):
        api.person_run_device(12345, 0, n, outputs, 0.03)
        ms.append(api.last_kernel_ms())
    print("%-28s %-14s %s" % (os.path.basename(os.environ.get("MSIM_B200_LIB", "default")), nm, " ".join("%.3f" % v for v in ms)), flush=True)
