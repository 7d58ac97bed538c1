"""Development aid: the staged person-r kernel at fixed and adapted shares of the first wave of blocks (msim_set_wave_share)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import microsimulation_b200 as m, microsimulation_b200.api as api
m.init(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**8
for outputs, nm in ((2 | 4, "report+counts"), (4, "counts"), (1, "rowlog")):
    for share in (0.5, 0.52, 0.535, 0.55, 0.0):
        api.set_wave_share(share)
        ms, sh = [], []
        for _ in range(7 if share == 0.0 else 4):
            api.person_run_device(12345, 0, n, outputs, 0.03)
            ms.append(api.last_kernel_ms())
            sh.append(api.last_wave_share())
        print("%-14s n=%g share %-8s ms %s  used %s" % (nm, n, share or "adaptive", " ".join("%.3f" % v for v in ms), " ".join("%.4f" % v for v in sh)), flush=True)
api.set_wave_share(0.0)
