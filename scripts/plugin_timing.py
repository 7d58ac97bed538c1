"""Times the example plugin model (examples/plugins/illness_death_doc.cu) in both RNG regimes (development aid)."""
import os, struct, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import microsimulation_b200 as m
from microsimulation_b200 import plugin
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
p = plugin.Plugin(plugin.build(os.path.join(ROOT, "examples", "plugins", "illness_death_doc.cu")), "illness_death_doc")
consts = struct.pack("5d", 30.0, 60.0, 5.0, 10.0, 0.03)
for n in (10**6, 10**7):
    best = 1e9
    for _ in range(3):
        v, _ = p.run(n, consts, seed=[12345] * 6)
        best = min(best, p.last_kernel_ms)
    print("substreams  n=%.0e  %.3f ms  %.3e persons/s  mean QALY %.4f" % (n, best, n / best * 1e3, v["QALY"].mean()), flush=True)
m.init(0)
for n in (10**5, 10**6):
    m.set_seed(12345)
    st = m.mt_state().copy()
    import time
    t0 = time.perf_counter()
    v, _ = p.run(n, consts, regime=plugin.MT_STREAM, mt_state=st, mean_draws=9)
    dt = time.perf_counter() - t0
    print("MT stream   n=%.0e  %.1f ms wall (whole call)  mean QALY %.4f" % (n, dt * 1e3, v["QALY"].mean()), flush=True)
