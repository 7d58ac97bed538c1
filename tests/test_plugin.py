"""Out-of-tree models through csrc/plugin.cuh (the device mirror of the reference's cProcess plugin mechanism).

examples/plugins/illness_death_doc.cu is the user model the reference documents (inst/doc/examples.org:254-312,
compiled there with Rcpp::sourceCpp); its printed results are the known answers.  examples/plugins/
illness_death_builtin.cu exports the package's own illness-death model through the same mechanism, so the plugin
path's EventReport can be compared with msim_callIllnessDeath.
"""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from conftest import ROOT, assert_report_close

PLUGINS = os.path.join(ROOT, "examples", "plugins")
DOC_CONSTS = struct.pack("5d", 30.0, 60.0, 5.0, 10.0, 0.03)


@pytest.fixture(scope="module")
def doc_plugin():
    from microsimulation_b200 import plugin
    return plugin.Plugin(plugin.build(os.path.join(PLUGINS, "illness_death_doc.cu")), "illness_death_doc")


@pytest.fixture(scope="module")
def builtin_plugin():
    from microsimulation_b200 import plugin
    return plugin.Plugin(plugin.build(os.path.join(PLUGINS, "illness_death_builtin.cu")), "illness_death_builtin")


def test_plugin_library_exports(doc_plugin, builtin_plugin):
    assert doc_plugin.n_values == 2 and doc_plugin.consts_bytes == len(DOC_CONSTS)
    assert builtin_plugin.n_values == 0 and builtin_plugin.consts_bytes == 32
    for sym in ("_run", "_free", "_last_error", "_n_values", "_consts_bytes"):
        assert hasattr(doc_plugin.lib, "illness_death_doc" + sym)
    # nothing else leaks out of the library (-fvisibility=hidden): it cannot clash with libmsim_b200.so
    assert not hasattr(doc_plugin.lib, "msim_init")


def test_plugin_argument_errors_need_no_device(doc_plugin):
    from microsimulation_b200 import MsimError
    with pytest.raises(MsimError) as e:
        doc_plugin.run(5, DOC_CONSTS[:-8])
    assert e.value.code == -3 and "consts_bytes" in str(e.value)
    with pytest.raises(MsimError) as e:
        doc_plugin.run(-1, DOC_CONSTS)
    assert e.value.code == -3
    with pytest.raises(MsimError) as e:
        doc_plugin.run(5, DOC_CONSTS, regime=7)
    assert e.value.code == -3


def test_plugin_has_no_cpu_path(doc_plugin):
    import torch
    from microsimulation_b200 import MsimError
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(MsimError) as e:
        doc_plugin.run(5, DOC_CONSTS)
    assert e.value.code == -1  # MSIM_ERR_NO_DEVICE


def _seeded_state(msim, seed=12345):
    msim.set_seed(seed)
    return msim.mt_state().copy()


@pytest.mark.gpu
def test_documented_model_reproduces_the_documents_results(doc_plugin, msim, kats, oracle):
    from microsimulation_b200 import plugin
    st = _seeded_state(msim)
    values, _ = doc_plugin.run(5, DOC_CONSTS, regime=plugin.MT_STREAM, mt_state=st, mean_draws=9)
    want = np.array(kats["doc_illness_death_qaly_cost"])  # as printed by the reference: 6 and 3 decimals
    assert np.allclose(values["QALY"], want[:, 0], atol=5.1e-7, rtol=0)
    assert np.allclose(values["cumulative_costs"], want[:, 1], atol=5.1e-4, rtol=0)
    oracle.set_seed(12345)
    _, qc = oracle.doc_illness_death(5)
    np.testing.assert_allclose(values["QALY"], qc[:, 0], rtol=1e-12)
    np.testing.assert_allclose(values["cumulative_costs"], qc[:, 1], rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 33, 1500])
def test_documented_model_vs_oracle_and_generator_state(doc_plugin, msim, oracle, n):
    from microsimulation_b200 import plugin
    st = _seeded_state(msim)
    values, _ = doc_plugin.run(n, DOC_CONSTS, regime=plugin.MT_STREAM, mt_state=st, mean_draws=9)
    oracle.set_seed(12345)
    _, qc = oracle.doc_illness_death(n)
    assert values["QALY"].shape == (n,)
    np.testing.assert_allclose(values["QALY"], qc[:, 0], rtol=1e-12)
    np.testing.assert_allclose(values["cumulative_costs"], qc[:, 1], rtol=1e-12)
    # R's generator is left where the reference leaves it: the next uniforms agree
    msim.set_mt_state(st)
    assert np.array_equal(msim.mt_uniforms(5), oracle.mt_uniforms(5))


@pytest.mark.gpu
def test_documented_model_substream_regime_shards(doc_plugin):
    whole, _ = doc_plugin.run(5000, DOC_CONSTS, seed=[12345] * 6)
    a, _ = doc_plugin.run(2000, DOC_CONSTS, seed=[12345] * 6)
    b, _ = doc_plugin.run(3000, DOC_CONSTS, seed=[12345] * 6, first_person=2000)
    for k in whole:
        assert np.array_equal(whole[k], np.concatenate([a[k], b[k]])), k   # person i <-> substream i+1, wherever it runs
    other, _ = doc_plugin.run(5000, DOC_CONSTS, seed=[1, 2, 3, 4, 5, 6])
    assert not np.array_equal(other["QALY"], whole["QALY"])
    # undiscounted life expectancy from healthy is 60 years (onset only restarts the clocks); QALYs at 3% are far below
    assert 10 < whole["QALY"].mean() < 25 and abs(whole["QALY"].mean() - other["QALY"].mean()) < 1.0
    assert np.all(whole["cumulative_costs"] > 0)


@pytest.mark.gpu
def test_speed_test_model_of_the_reference():
    """callSpeedTest's VerySimple process (src/person-r.cc:234-255): two events each, nothing else."""
    import struct
    from microsimulation_b200 import plugin
    p = plugin.Plugin(plugin.build(os.path.join(PLUGINS, "very_simple.cu")), "very_simple")
    values, _ = p.run(10**6, struct.pack("i", 0), seed=[12345] * 6)
    assert values["events"].shape == (10**6,) and np.all(values["events"] == 2.0)
    # (no timing claim: with two constant event times and an empty handler the compiler folds the whole life away)


@pytest.mark.gpu
def test_heap_capacity_overflow_is_an_error_code():
    from microsimulation_b200 import MsimError, plugin
    lib = plugin.build(os.path.join(PLUGINS, "illness_death_doc.cu"), out=os.path.join(ROOT, "build", "libillness_death_doc_cap3.so"),
                       defines=["DOC_HEAP_CAP=3"])
    small = plugin.Plugin(lib, "illness_death_doc")
    with pytest.raises(MsimError) as e:
        small.run(2000, DOC_CONSTS, seed=[12345] * 6)
    assert e.value.code == -4  # MSIM_ERR_HEAP: two onsets leave four actions pending


@pytest.mark.gpu
@pytest.mark.parametrize("n,cure", [(20000, 0.1), (3000, 0.2)])
def test_builtin_model_as_plugin_matches_the_builtin_entry_point(builtin_plugin, msim, n, cure):
    from microsimulation_b200 import plugin
    consts = C.create_string_buffer(32)
    builtin_plugin.lib.illness_death_builtin_consts(C.c_double(cure), C.c_double(0.0), consts)
    st = _seeded_state(msim)
    _, rep = builtin_plugin.run(n, consts.raw, regime=plugin.MT_STREAM, mt_state=st, max_draws=5, mean_draws=3.6,
                                outputs=plugin.OUT_REPORT, n_state=2, n_event=3)
    msim.set_seed(12345)
    want = msim.callIllnessDeath(n, cure, 0.0)
    assert_report_close(rep, want, rtol=1e-12)
    assert np.array_equal(st, msim.mt_state())   # both leave the generator after the last person's last draw
