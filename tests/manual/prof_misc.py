"""The secondary kernels once or twice each, for ncu (development aid; under tests/ because the gsm fixture is built
with the oracle's spline code): gsm inversion, survival-time lookups, Sick-Sicker, the documented plugin model."""
import os, struct, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import microsimulation_b200 as m, microsimulation_b200.api as api
from microsimulation_b200 import lookups, plugin
from oracle.pyoracle import Oracle
from gsm_fixture import make_model

m.init(0)
model = make_model(Oracle("port"))[0]
rs = np.random.RandomState(1)
n = 10**7
u = rs.uniform(0.02, 0.98, n)
idx = rs.randint(0, 3, n).astype(np.int32)
for _ in range(2):
    model.randU(u, index=idx)
print("gsm randU       n=%.0e  %.3f ms" % (n, api.last_kernel_ms()), flush=True)
h, t = np.array([0.01, 0.02, 0.05, 0.1, 0.3]), np.array([0.0, 40.0, 60.0, 75.0, 90.0])
pexp = lookups.Rpexp(h, t)
for _ in range(2):
    pexp.rand(u)
print("rpexp rand      n=%.0e  %.3f ms" % (n, api.last_kernel_ms()), flush=True)
api.set_user_random_seed(12345)
for _ in range(2):
    api.callSim(10**6, api.sick_sicker_param(False), substreams=True)
print("sick substreams n=1e6   %.3f ms" % api.last_kernel_ms(), flush=True)
p = plugin.Plugin(plugin.build(os.path.join(ROOT, "examples", "plugins", "illness_death_doc.cu")), "illness_death_doc")
for _ in range(2):
    p.run(10**7, struct.pack("5d", 30.0, 60.0, 5.0, 10.0, 0.03), seed=[12345] * 6)
print("plugin doc      n=1e7   %.3f ms" % p.last_kernel_ms, flush=True)
