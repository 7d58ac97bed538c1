"""The reference's CPU path (oracle/_ref where built, else the port) timed on this host for the BASELINE configs the
bench does not cover: one process, as the reference runs them (development aid; prints a small table)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle
flavour = "ref" if pyoracle.build("ref") else "port"
o = pyoracle.Oracle(flavour)


def timed(label, persons, fn):
    fn_small = None
    t0 = time.perf_counter()
    fn()
    dt = time.perf_counter() - t0
    print("%-52s %9d persons  %8.3f s  %.3e persons/s  (%s, 1 core)" % (label, persons, dt, persons / dt, flavour), flush=True)


timed("config 1: illness-death, n = 1e5", 10**5, lambda: o.callIllnessDeath(10**5, 0.1, 0.0))
timed("config 2: simple-example row log, n = 1e6", 10**6, lambda: o.callSimplePerson(10**6))
timed("config 3: person-r row log, n = 1e6 (of 1e7)", 10**6, lambda: o.callPersonSimulation(10**6))
timed("config 5: calibration, 1 set x 1e6 (of 256 sets)", 10**6, lambda: o.callCalibrationPerson(12345, 10**6))
par = pyoracle.sick_sicker_param(False)
timed("Sick-Sicker (README example), n = 1e5", 10**5, lambda: o.sick_sicker(10**5, par))
timed("documented user model (plugin example), n = 2e5", 2 * 10**5, lambda: o.doc_illness_death(2 * 10**5))
