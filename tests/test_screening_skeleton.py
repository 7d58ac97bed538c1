"""The screening-model skeleton: Rpexp + Table + NumericInterpolate + gsm + EventReport together in one model
(examples/plugins/screening_skeleton.{h,cu}).  PARITY UNPINNED: the FHCRC screening model is not in
the reference tree, the skeleton only strings together what the tree keeps of it, and nothing in the reference prints a
result for it.  What is checked: the oracle's restatement behaves sensibly (CPU), and the plugin on the GPU equals it."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, assert_table_close
from screening_fixture import Screening


def test_oracle_skeleton_behaves(oracle):
    n = 4000
    oracle.set_user_random_seed(12345)
    none = Screening(oracle, screen=0)
    rep0, v0 = oracle.screening_skeleton(none.inputs, n)
    oracle.set_user_random_seed(12345)
    two = Screening(oracle, screen=2)
    rep2, v2 = oracle.screening_skeleton(two.inputs, n)
    assert np.all(v0[:, 2] == 0) and v2[:, 2].sum() > 20          # screen-detected diagnoses only with screening
    assert np.array_equal(v0[:, 3], v2[:, 3])                      # common random numbers: the same natural history
    assert 60 < v0[:, 0].mean() < 90 and v0[:, 0].max() < 130      # ages at death from the 1950 cohort's mortality
    dx0, dx2 = v0[:, 1][v0[:, 1] > 0], v2[:, 1][v2[:, 1] > 0]
    assert len(dx2) >= len(dx0) > 50 and dx2.mean() < dx0.mean()   # screening finds cancers, and finds them earlier
    ev = rep2["events"]
    assert ev["number"][ev["event"] == 4].sum() > n                # toScreen events were reported (ages 50..70)
    assert set(np.unique(ev["age"][ev["event"] == 4])) <= set(np.arange(50.0, 71.0, 2.0))


@pytest.mark.gpu
@pytest.mark.parametrize("screen,first", [(0, 0), (2, 0), (4, 1000)])
def test_skeleton_plugin_on_the_gpu_vs_oracle(msim, oracle, screen, first):
    from microsimulation_b200 import plugin
    lib = plugin.build(os.path.join(ROOT, "examples", "plugins", "screening_skeleton.cu"))
    p = plugin.Plugin(lib, "screening_skeleton")
    s = Screening(oracle, screen=screen)
    p.lib.screening_skeleton_setup.argtypes = [C.c_void_p, C.c_int]
    assert p.lib.screening_skeleton_setup(C.byref(s.inputs), 0) == 0, p._err().decode()
    n = 30000
    values, report = p.run(n, s.consts(), seed=[12345] * 6, first_person=first, outputs=plugin.OUT_VALUES | plugin.OUT_REPORT,
                           n_state=3, n_event=7, discount_rate=0.03)
    oracle.set_user_random_seed(12345)
    ref_report, ref = oracle.screening_skeleton(s.inputs, n, first_person=first, discount_rate=0.03)
    assert np.array_equal(values["grade"], ref[:, 3])
    assert np.array_equal(values["screen_dx"], ref[:, 2])
    assert np.array_equal(values["age_dx"] > 0, ref[:, 1] > 0)
    # the sojourn time is a Brent root (tolerance 1e-8 on the log-time scale): times agree to that, not to the last bit
    np.testing.assert_allclose(values["age_death"], ref[:, 0], rtol=1e-6)
    np.testing.assert_allclose(values["age_dx"], ref[:, 1], rtol=1e-6)
    for tb in ("events", "prev"):
        for k in ref_report[tb]:
            assert np.array_equal(report[tb][k], ref_report[tb][k]), (tb, k)
    assert_table_close(report["pt"], ref_report["pt"], 1e-7)
    assert_table_close(report["ut"], ref_report["ut"], 1e-7)
