"""A few launches of the person-r path for ncu (development aid): python scripts/prof_person.py OUTPUTS PERSONS [SCHEDULE]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import microsimulation_b200 as m, microsimulation_b200.api as api
outputs, n = int(sys.argv[1]), int(float(sys.argv[2]))
m.init(0)
api.set_person_schedule(int(sys.argv[3]) if len(sys.argv) > 3 else 0)
for _ in range(4):
    api.person_run_device(12345, 0, n, outputs, 0.03)
print("ok", api.last_kernel_ms(), "ms")
