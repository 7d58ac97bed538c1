"""Inputs of the screening-model skeleton (examples/plugins/screening_skeleton.h) for the tests: the scalar parameters
of the reference's PSA natural-history sketch (test/test_microsimulation.R:1949-1956), look-up tables taken from the
fhcrcData fixture (tests/golden/fhcrc_tables.npz), and the tests' generalised survival model as the sojourn-time
model."""
import ctypes as C
import os
import struct

import numpy as np

from conftest import GOLDEN
from gsm_fixture import make_model
from microsimulation_b200 import fhcrc_data
from microsimulation_b200.gsm import GsmC


class ScreeningInputs(C.Structure):
    _fields_ = [("g0", C.c_double), ("mubeta0", C.c_double), ("sebeta0", C.c_double), ("mubeta1", C.c_double),
                ("sebeta1", C.c_double), ("mubeta2", C.c_double * 2), ("sebeta2", C.c_double * 2), ("tau2", C.c_double),
                ("psa_threshold", C.c_double), ("compliance", C.c_double), ("cohort", C.c_double), ("screen", C.c_int32),
                ("n_mort", C.c_int32), ("mort_age", C.POINTER(C.c_double)), ("mort_rate", C.POINTER(C.c_double)),
                ("n_biopsy", C.c_int32), ("biopsy_rows", C.POINTER(C.c_double)),
                ("n_sens", C.c_int32), ("sens_year", C.POINTER(C.c_double)), ("sens_value", C.POINTER(C.c_double)),
                ("n_surv", C.c_int32), ("surv_time", C.POINTER(C.c_double)), ("surv_value", C.POINTER(C.c_double)),
                ("gsm", C.POINTER(GsmC))]


def _tables():
    z = np.load(os.path.join(GOLDEN, "fhcrc_tables.npz"))
    out = {}
    for k in z.files:
        name, col = k.split("/", 1)
        out.setdefault(name, {})[col] = z[k]
    return out


class Screening:
    """Keeps the arrays alive and builds both forms of the inputs: the C struct (oracle, plugin setup) and the scalar
    parameter blob a plugin call carries (scr::Consts: 14 doubles)."""

    def __init__(self, oracle, screen=2, cohort=1950, compliance=0.75, psa_threshold=3.0):
        t = _tables()
        dp = C.POINTER(C.c_double)
        self.ages, self.rates = fhcrc_data.mortality_hazards(t, cohort)
        bf = t["biopsy_frequency"]
        rows = []
        for psa_from, probs in zip(bf["PSA.beg"], zip(*[bf[c] for c in ("X55.59", "X60.64", "X65.69", "X70.74", "X75.79")])):
            for age_from, p in zip((55.0, 60.0, 65.0, 70.0, 75.0), probs):
                rows.append((psa_from, age_from, p))
        self.biopsy = np.ascontiguousarray(rows, dtype=np.float64)
        self.sens_year = np.ascontiguousarray(t["biopsy_sensitivity"]["Year"], dtype=np.float64)
        self.sens_value = np.ascontiguousarray(t["biopsy_sensitivity"]["Sensitivity"], dtype=np.float64)
        sd = t["survival_dist"]
        sel = sd["Grade"] == 0
        self.surv_time = np.ascontiguousarray(sd["Time"][sel], dtype=np.float64)
        self.surv_value = np.ascontiguousarray(sd["Survival"][sel], dtype=np.float64)
        self.model = make_model(oracle)[0]
        self.gsm_c = self.model.cstruct()
        s = ScreeningInputs()
        s.g0, s.mubeta0, s.sebeta0, s.mubeta1, s.sebeta1 = 0.0005, -1.609, 0.2384, 0.04463, 0.0430
        s.mubeta2 = (C.c_double * 2)(0.0397, 0.1678)
        s.sebeta2 = (C.c_double * 2)(0.0913, 0.3968)
        s.tau2 = 0.0829
        s.psa_threshold, s.compliance, s.cohort, s.screen = psa_threshold, compliance, float(cohort), int(screen)
        s.n_mort, s.mort_age, s.mort_rate = len(self.ages), self.ages.ctypes.data_as(dp), self.rates.ctypes.data_as(dp)
        s.n_biopsy, s.biopsy_rows = len(self.biopsy), self.biopsy.ctypes.data_as(dp)
        s.n_sens, s.sens_year, s.sens_value = len(self.sens_year), self.sens_year.ctypes.data_as(dp), self.sens_value.ctypes.data_as(dp)
        s.n_surv, s.surv_time, s.surv_value = len(self.surv_time), self.surv_time.ctypes.data_as(dp), self.surv_value.ctypes.data_as(dp)
        s.gsm = C.pointer(self.gsm_c)
        self.inputs = s

    def consts(self):
        s = self.inputs
        return struct.pack("14d", s.g0, s.mubeta0, s.sebeta0, s.mubeta1, s.sebeta1, s.mubeta2[0], s.mubeta2[1], s.sebeta2[0],
                           s.sebeta2[1], s.tau2, s.psa_threshold, s.compliance, s.cohort, float(s.screen))
