"""Sick-Sicker from R's stream, a few calls, for ncu launch lists (development aid): python scripts/prof_sick.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import microsimulation_b200 as m, microsimulation_b200.api as api
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**5
m.init(0)
par = api.sick_sicker_param(False)
for _ in range(3):
    api.callSim(n, par)
print("ok %.3f ms" % api.last_kernel_ms())
