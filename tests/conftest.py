import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle("port")


@pytest.fixture(scope="session")
def oracle_ref():
    """The oracle linked against the reference's own ssim.cc / RngStream.cpp (oracle/_ref)."""
    from oracle.pyoracle import Oracle, build
    if build("ref") is None:
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    return Oracle("ref")


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(GOLDEN, "doc_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "oracle_ref.npz")))


@pytest.fixture(scope="session")
def emu():
    """Host emulation of the device engine (tests/host_emulation)."""
    d = os.path.join(ROOT, "tests", "host_emulation")
    so = os.path.join(d, "libmsim_emu.so")
    srcs = [os.path.join(d, "emu.cc")] + [os.path.join(ROOT, "microsimulation_b200", "csrc", f)
                                          for f in os.listdir(os.path.join(ROOT, "microsimulation_b200", "csrc"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        r = subprocess.run(["g++", "-O2", "-g", "-fPIC", "-std=gnu++17", "-ffp-contract=off", "-shared", "-o", so,
                            os.path.join(d, "emu.cc")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    return C.CDLL(so)


@pytest.fixture(scope="session")
def msim():
    """The product: microsimulation_b200 over libmsim_b200.so on cuda:0.  No fallback."""
    import microsimulation_b200 as m
    m.init(0)
    return m


def assert_table_close(a, b, rtol=1e-9, int_cols=("Key", "age", "event", "number")):
    """Report tables: key and count columns bit-exact, FP sums within rtol (north_star: 1e-9)."""
    assert set(a.keys()) == set(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape, (k, a[k].shape, b[k].shape)
        if k in int_cols:
            assert np.array_equal(a[k], b[k]), k
        else:
            np.testing.assert_allclose(a[k], b[k], rtol=rtol, atol=0, err_msg=k)


def assert_report_close(a, b, rtol=1e-9):
    for tb in ("pt", "ut", "events", "prev"):
        assert_table_close(a[tb], b[tb], rtol)


def assert_rowlog_close(a, b, time_rtol=4e-15, end_col="endTime"):
    """Row logs: id / state / event bit-exact; event times to a few ulp (device libm vs glibc)."""
    for k in ("id", "state", "event"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("startTime", end_col):
        np.testing.assert_allclose(a[k], b[k], rtol=time_rtol, atol=0, err_msg=k)
