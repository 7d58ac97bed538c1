"""Quick GPU parity + timing probe (development aid; the real checks are tests/ -m gpu)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import microsimulation_b200 as m
from oracle.pyoracle import Oracle

o = Oracle("port")
m.init(0)

def cmp(name, a, b, exact=True):
    if a.shape != b.shape:
        print(name, "SHAPE", a.shape, b.shape); return
    if np.array_equal(a, b):
        print(name, "exact"); return
    rel = np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
    print(name, "maxrel %.3e" % rel, "n_diff", int((a != b).sum()), "of", a.size)

u = m.rng_uniforms(12345, 5, 100000, 8); cmp("mrg uniforms", u, o.rng_uniforms(12345, 5, 100000, 8))
m.set_seed(12345); o.set_seed(12345); cmp("mt uniforms", m.mt_uniforms(100000), o.mt_uniforms(100000))

for n in (20, 1000, 200000):
    g = m.callPersonSimulation(n); r = o.callPersonSimulation(n)
    for k in r: cmp("person n=%d %s" % (n, k), g[k], r[k])
n = 200000
rep, cnt = m.person_report(12345, 0, n, 0.03)
ref = o.person_run(12345, 0, n, rowlog=False, report=True, discount_rate=0.03)
cmp("person counts", cnt[:8], ref["counts"][:8]); print("sum end", cnt[8], ref["counts"][8])
for tb in ("pt", "ut", "events", "prev"):
    for k in rep[tb]: cmp("person report %s.%s" % (tb, k), rep[tb][k], ref["report"][tb][k])

for n in (10, 1000, 100000):
    g = m.callSimplePerson(n); r = o.callSimplePerson(n)
    for k in r: cmp("simple n=%d %s" % (n, k), g[k], r[k])
    g = m.callSimplePerson2(n); r = o.callSimplePerson2(n)
    for tb in ("pt", "ut", "events", "prev"):
        for k in g[tb]: cmp("simple2 n=%d %s.%s" % (n, tb, k), g[tb][k], r[tb][k])
    for cure, zsd in ((0.1, 0.0), (0.2, 0.0), (0.1, 0.5)):
        g = m.callIllnessDeath(n, cure, zsd); r = o.callIllnessDeath(n, cure, zsd)
        for tb in ("pt", "ut", "events", "prev"):
            for k in g[tb]: cmp("illness n=%d cure=%g zsd=%g %s.%s" % (n, cure, zsd, tb, k), g[tb][k], r[tb][k])
    cmp("mt state after", m.mt_uniforms(8), o.mt_uniforms(8))

g = m.callCalibrationPerson(10, 500); r = o.callCalibrationPerson(10, 500)
print(g["raw"]); print(r)
g = m.callCalibrationPerson(12345, 100000)["raw"]; r = o.callCalibrationPerson(12345, 100000)
for k in r: cmp("calib " + k, g[k], r[k])

# timing
import microsimulation_b200.api as api
for n in (10**6, 10**7, 10**8):
    for outputs, nm in ((4, "counts"), (1, "rowlog"), (2 | 4, "report+counts"), (1 | 2 | 4, "all")):
        best = 1e9
        for rep_ in range(3):
            res = api.person_run_device(12345, 0, n, outputs, 0.03)
            best = min(best, api.last_kernel_ms())
        print("person n=%g %-14s kernel %.3f ms  %.3e persons/s rows=%d" % (n, nm, best, n / best * 1e3, res.n_rows))
t0 = time.time(); g = m.callPersonSimulation(10**7); t1 = time.time()
print("e2e callPersonSimulation(1e7): %.3f s rows %d" % (t1 - t0, len(g["id"])))
t0 = time.time(); g = m.callPersonSimulation(10**7); t1 = time.time()
print("e2e callPersonSimulation(1e7) again: %.3f s rows %d" % (t1 - t0, len(g["id"])))
for n in (10**5, 10**6):
    t0 = time.time(); m.callIllnessDeath(n); t1 = time.time(); print("illness n=%g: %.4f s kernel-pipeline %.3f ms" % (n, t1 - t0, api.last_kernel_ms()))
    t0 = time.time(); m.callSimplePerson(n); t1 = time.time(); print("simple n=%g: %.4f s kernel-pipeline %.3f ms" % (n, t1 - t0, api.last_kernel_ms()))
import numpy as np
pars = np.tile(np.array([4, 0.5, 0.05, 10, 3, 0.5]), (256, 1)) * (1 + 0.2 * (np.random.RandomState(1).rand(256, 6) - 0.5))
t0 = time.time(); out = m.calibration_sweep(12345, pars, 10**6); t1 = time.time()
print("calib sweep 256x1e6: %.3f s kernel %.3f ms" % (t1 - t0, api.last_kernel_ms()))
print("launches", m.launch_count())
