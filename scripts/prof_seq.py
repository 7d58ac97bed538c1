"""Sequential-stream models for ncu launch lists (development aid): python scripts/prof_seq.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import microsimulation_b200 as m, microsimulation_b200.api as api
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**6
m.init(0)
for _ in range(2):
    m.callIllnessDeath(n); a = api.last_kernel_ms()
    m.callSimplePerson2(n); b = api.last_kernel_ms()
    m.callSimplePerson(n); c = api.last_kernel_ms()
print("ok illness %.3f simple2 %.3f simple %.3f ms" % (a, b, c))
