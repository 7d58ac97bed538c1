"""Development aid: where a person-r launch's time goes (the kernel's fixed part).  Needs the library built with
-DMSIM_TRACE (build/variants/libmsim_trace.so, see scripts/build_variants.sh for the flags); the kernel then prints, per
launch, the earliest and latest block's time stamps at entry / start of the loop / persons done / counters flushed /
fold done, in ns after the first block's entry."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MSIM_B200_LIB"] = os.path.join(ROOT, "build", "variants", "libmsim_trace.so")
import microsimulation_b200 as m, microsimulation_b200.api as api
m.init(0)
for n in (1000, 12500000, 100000000, 100000000, 100000000):
    for rep in range(2):
        api.person_run_device(12345, 0, n, 6, 0.03)
        print("n=%d rep=%d kernel_ms %.4f" % (n, rep, api.last_kernel_ms()), flush=True)
