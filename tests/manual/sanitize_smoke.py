"""Small instances of every kernel, for compute-sanitizer memcheck (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import microsimulation_b200 as m, microsimulation_b200.api as api
m.init(0)
for sched in ("staged", "generic"):
    api.set_person_schedule(sched)
    for outputs in (1, 2, 4, 7):
        api.person_run_device(12345, 5, 40001, outputs, 0.03)
    m.callPersonSimulation(3000)
api.set_person_schedule("staged")
m.callIllnessDeath(5000); m.callSimplePerson(5000); m.callSimplePerson2(5000)
m.callCalibrationPerson(10, 3000)
api.callSim(800, api.sick_sicker_param()); api.set_user_random_seed(12345); api.callSim(800, api.sick_sicker_param(), substreams=True)
m.rng_uniforms(12345, 0, 1000, 4); m.set_seed(1); m.mt_uniforms(5000)
api.discountedInterval(np.ones(100), np.zeros(100), np.ones(100), 0.03)
from microsimulation_b200.lookups import Rpexp, RpexpGamma, NumericInterpolate, Table
Rpexp([1, 2, 3], [0, 1, 2]).rand(np.linspace(0.01, 0.99, 500)); RpexpGamma(0.5, [0.01, 0.02], [0, 50]).rand(np.linspace(0.01, 0.99, 500), 1.0)
NumericInterpolate([0, 1, 2], [0, 2, 3]).approx(np.linspace(-1, 3, 500)); Table(np.array([[1, 1, 5.0], [2, 1, 6.0]]))(np.array([[1.5, 1], [3, 2]]))
from oracle.pyoracle import Oracle
from gsm_fixture import make_model
make_model(Oracle("port"))[0].randU(np.linspace(0.01, 0.99, 2000))
print("sanitize smoke ok, launches", m.launch_count())
