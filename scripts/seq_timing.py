"""Times the sequential-stream models (configs 1 and 2) for each library variant in build/variants (development aid)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r)
import microsimulation_b200 as m, microsimulation_b200.api as api
m.init(0)
for name, fn in (("simple-example 1e6", lambda: m.callSimplePerson(10**6)), ("illness-death 1e5", lambda: m.callIllnessDeath(10**5)),
                 ("illness-death 1e6", lambda: m.callIllnessDeath(10**6))):
    best = 1e9
    for _ in range(4):
        fn()
        best = min(best, api.last_kernel_ms())
    print("%%-24s %%-20s %%.3f ms" %% (os.path.basename(os.environ.get("MSIM_B200_LIB", "default")), name, best), flush=True)
''' % ROOT
for lib in [None] + sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so"))):
    env = dict(os.environ)
    if lib:
        env["MSIM_B200_LIB"] = lib
    subprocess.run([sys.executable, "-c", code], env=env)
