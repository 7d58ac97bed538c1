"""Large shards (development aid): counts must add up and ids must stay exact well beyond the bench size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import microsimulation_b200 as m, microsimulation_b200.api as api
m.init(0)
for n in (2 * 10**9, 4 * 10**9):
    t0 = time.time()
    rep, counts = m.person_report(12345, 0, n, 0.03)
    dt = time.time() - t0
    ev = rep["events"]
    by_kind = np.bincount(ev["event"].astype(int), weights=ev["number"], minlength=8)
    assert np.array_equal(by_kind, counts[:8]) and by_kind[0] + by_kind[1] == n, (by_kind, counts)
    at0 = rep["prev"]["number"][(rep["prev"]["Key"] == 0) & (rep["prev"]["age"] == 0)]
    assert at0[0] == n
    print("report n=%.0e ok: %.3f s wall, kernel %.1f ms, %.3e persons/s" % (n, dt, api.last_kernel_ms(), n / api.last_kernel_ms() * 1e3), flush=True)
try:
    api.person_run_device(12345, 0, 2**32, api.MSIM_OUT_COUNTS)
    print("2^32 persons accepted?!")
except m.MsimError as e:
    print("2^32 persons refused as documented:", e)
n = 5 * 10**8
res = api.person_run_device(12345, 3 * 10**9, n, api.MSIM_OUT_ROWLOG | api.MSIM_OUT_COUNTS)
c = api.person_fetch_counts()
print("rowlog n=%.0e first=3e9: rows %d (counts %d) kernel %.1f ms" % (n, res.n_rows, int(c[:8].sum()), api.last_kernel_ms()), flush=True)
assert res.n_rows == int(c[:8].sum())
g = api.person_fetch_rowlog()
ids = g["id"]
assert ids[0] == 3e9 and ids[-1] == 3e9 + n - 1
d = np.diff(ids)
assert ((d == 0) | (d == 1)).all()
print("rowlog ids exact and in person order; host table %.1f GB" % (5 * 8 * len(ids) / 1e9))
