"""The portable oracle (restated engine + integer RngStream) against the build that
links the reference's OWN src/ssim.cc, src/RngStream.cpp and inst/include/heap.h
(oracle/_ref), and both against the committed fixtures made from oracle/_ref."""
import ctypes as C

import numpy as np
import pytest

from oracle.pyoracle import sick_sicker_param


def _eq(a, b):
    assert set(a) == set(b)
    for k in a:
        if isinstance(a[k], dict):
            _eq(a[k], b[k])
        else:
            assert np.array_equal(a[k], b[k]), k


def test_flavours(oracle, oracle_ref):
    assert oracle.flavour == "port" and oracle_ref.flavour == "reference"


def test_rngstream_draw_for_draw(oracle, oracle_ref):
    for seed, first in ((12345, 0), (12345, 99999990), ([1, 2, 3, 4, 5, 6], 17), (4294944442, 3)):
        assert np.array_equal(oracle.rng_uniforms(seed, first, 500, 16), oracle_ref.rng_uniforms(seed, first, 500, 16))
    for n in (1, 2, 1000, 99999999, -5):
        assert oracle.rng_advance_substream(12345, n) == oracle_ref.rng_advance_substream(12345, n)
    for e, c in ((3, 0), (-3, 2), (51, -1), (0, 123456)):
        a = (C.c_double * 6)(*[12345.0] * 6)
        b = (C.c_double * 6)(*[12345.0] * 6)
        oracle.lib.oracle_rng_advance_e(a, C.c_int(e), C.c_int(c))
        oracle_ref.lib.oracle_rng_advance_e(b, C.c_int(e), C.c_int(c))
        assert list(a) == list(b)


def test_models_identical(oracle, oracle_ref):
    _eq(oracle.person_run(12345, 0, 20000, rowlog=True, report=True, discount_rate=0.03),
        oracle_ref.person_run(12345, 0, 20000, rowlog=True, report=True, discount_rate=0.03))
    _eq(oracle.person_run(987654321, 4321, 3000), oracle_ref.person_run(987654321, 4321, 3000))
    _eq(oracle.callSimplePerson(5000), oracle_ref.callSimplePerson(5000))
    _eq(oracle.callSimplePerson2(5000), oracle_ref.callSimplePerson2(5000))
    _eq(oracle.callIllnessDeath(5000, 0.1, 0.0), oracle_ref.callIllnessDeath(5000, 0.1, 0.0))
    _eq(oracle.callIllnessDeath(5000, 0.3, 0.7), oracle_ref.callIllnessDeath(5000, 0.3, 0.7))
    _eq(oracle.callCalibrationPerson(10, 5000), oracle_ref.callCalibrationPerson(10, 5000))
    _eq(oracle.sick_sicker(500, sick_sicker_param(True)), oracle_ref.sick_sicker(500, sick_sicker_param(True)))
    _eq(oracle.sick_sicker(500, sick_sicker_param(False), substreams=True, first_person=7),
        oracle_ref.sick_sicker(500, sick_sicker_param(False), substreams=True, first_person=7))
    assert oracle.priority_demo() == oracle_ref.priority_demo()


def test_shard_starts_anywhere(oracle):
    """Person i is a pure function of (seed, i): a run that starts at person k equals the tail of a full run."""
    full = oracle.person_run(12345, 0, 3000)["rowlog"]
    tail = oracle.person_run(12345, 1000, 2000)["rowlog"]
    sel = full["id"] >= 1000
    for k in full:
        assert np.array_equal(full[k][sel], tail[k])


def test_port_matches_committed_fixtures(oracle, golden):
    """Fixtures were generated from oracle/_ref (tests/golden/make_golden.py); they travel to the GPU box."""
    assert np.array_equal(oracle.rng_uniforms(12345, 0, 64, 8), golden["mrg_uniforms_seed12345_sub0_64x8"])
    assert np.array_equal(oracle.rng_uniforms(12345, 99999990, 8, 4), golden["mrg_uniforms_seed12345_sub99999990_8x4"])
    oracle.set_seed(12345)
    assert np.array_equal(oracle.mt_uniforms(2000), golden["mt_uniforms_12345_2000"])
    r = oracle.person_run(12345, 0, 2000, rowlog=True, report=True, discount_rate=0.03)
    for k, v in r["rowlog"].items():
        assert np.array_equal(v, golden["person_n2000_" + k])
    assert np.array_equal(r["counts"], golden["person_n2000_counts"])
    for tb in ("pt", "ut", "events", "prev"):
        for k, v in r["report"][tb].items():
            assert np.array_equal(v, golden["person_n2000_report_%s_%s" % (tb, k)])
    r = oracle.callSimplePerson(3000)
    for k, v in r.items():
        assert np.array_equal(v, golden["simple_n3000_" + k])
    for name, rep in (("simple2_n20000", oracle.callSimplePerson2(20000)),
                      ("illness_n20000_cure01", oracle.callIllnessDeath(20000, 0.1, 0.0)),
                      ("illness_n20000_cure02_zsd05", oracle.callIllnessDeath(20000, 0.2, 0.5))):
        for tb in ("pt", "ut", "events", "prev"):
            for k, v in rep[tb].items():
                assert np.array_equal(v, golden["%s_%s_%s" % (name, tb, k)]), (name, tb, k)
    c = oracle.callCalibrationPerson(12345, 20000)
    for k, v in c.items():
        assert np.array_equal(v, golden["calib_seed12345_n20000_" + k])


def test_reference_engine_cannot_hold_two_processes(tmp_path):
    """SURVEY 8(f) 4 (multi-process simulations, cProcess::send): the reference's current engine cannot run them — the
    second create_process compares two A_Init actions whose events are null (src/ssim.cc:60-62, 108-113) and
    dereferences null.  So there is no reference behaviour to be in parity with; DESIGN.md section 8."""
    import os
    import subprocess
    from oracle import pyoracle
    if not os.path.isdir(pyoracle.REF_ROOT):
        pytest.skip("the reference tree is not on this machine")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "two")
    r = subprocess.run(["g++", "-O0", "-std=gnu++17", "-Wno-deprecated", "-I" + os.path.join(root, "oracle", "shim"),
                        "-I" + os.path.join(pyoracle.REF_ROOT, "inst", "include"), "-o", exe,
                        os.path.join(root, "tests", "manual", "probe_two_processes.cc"),
                        os.path.join(pyoracle.REF_ROOT, "src", "ssim.cc")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert "one created" in run.stdout and "two created" not in run.stdout
    assert run.returncode != 0   # killed by SIGSEGV
