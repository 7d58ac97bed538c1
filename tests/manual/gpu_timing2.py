"""Timings of the widened paths (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import microsimulation_b200 as m, microsimulation_b200.api as api
m.init(0)
for n in (10**5, 10**6):
    for sub in (False, True):
        if sub: api.set_user_random_seed(12345)
        api.callSim(1000, api.sick_sicker_param(), substreams=sub)
        t0 = time.time(); r = api.callSim(n, api.sick_sicker_param(), substreams=sub); t1 = time.time()
        print("sick-sicker n=%g substreams=%s: %.3f s (device part %.2f ms) mean cost %.1f" % (n, sub, t1 - t0, api.last_kernel_ms(), r["indiv"]["costs"].mean()), flush=True)
from oracle.pyoracle import Oracle
from gsm_fixture import make_model
model = make_model(Oracle("port"))[0]
u = np.random.RandomState(1).uniform(1e-6, 1 - 1e-6, 10**7)
model.randU(u[:1000])
t0 = time.time(); model.randU(u); t1 = time.time()
print("gsm randU n=1e7: %.3f s (kernel %.2f ms = %.2e inversions/s)" % (t1 - t0, api.last_kernel_ms(), 1e7 / api.last_kernel_ms() * 1e3))
