"""data/fhcrcData.rda (the FHCRC model's ten look-up tables): read without R
(microsimulation_b200/rdata.py), frozen as tests/golden/fhcrc_tables.npz, and looked up row by row through the oracle's
Table (CPU) and the device's (GPU).  The reference's C++ never reads the file, so nothing pins a value beyond
"Table(key of row i) = outcome of row i" and the floor rule between keys (rcpp_table.h:136-324)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from microsimulation_b200 import fhcrc_data

RDA = "/root/reference/data/fhcrcData.rda"


@pytest.fixture(scope="module")
def tables():
    z = np.load(os.path.join(GOLDEN, "fhcrc_tables.npz"))
    out = {}
    for k in z.files:
        name, col = k.split("/", 1)
        out.setdefault(name, {})[col] = z[k]
    return out


@pytest.mark.skipif(not os.path.exists(RDA), reason="the reference tree is not on this machine")
def test_rda_reader_reproduces_the_fixture(tables):
    got = fhcrc_data.load(RDA)
    assert sorted(got) == sorted(tables) == sorted(fhcrc_data.TABLES)
    for name, tab in tables.items():
        assert list(got[name]) == list(tab)
        for col, v in tab.items():
            assert np.array_equal(np.asarray(got[name][col], dtype=np.float64), v), (name, col)
    assert got["all_cause_mortality"]["BirthCohort"].dtype == np.int32  # integer columns stay integers
    assert len(got["all_cause_mortality"]["Age"]) == 12120 and len(got["pradt"]["ADT"]) == 8208


def test_rda_reader_rejects_what_it_does_not_understand(tmp_path):
    from microsimulation_b200 import rdata
    p = tmp_path / "x.rda"
    p.write_bytes(b"not an R file")
    with pytest.raises(ValueError):
        rdata.read_rda(str(p))
    import gzip
    p.write_bytes(gzip.compress(b"RDX2\nA\n"))  # ASCII serialisation: not supported
    with pytest.raises(ValueError):
        rdata.read_rda(str(p))


def _queries(tab, keys, rs):
    """every row's own key, and points between / beyond the keys"""
    exact = np.stack([tab[k] for k in keys], axis=1)
    lo, hi = exact.min(0), exact.max(0)
    between = rs.uniform(lo - 1.0, hi + 2.0, size=(3000, len(keys)))
    return exact, between


def test_every_row_of_every_table_through_the_oracle(oracle, tables):
    rs = np.random.RandomState(3)
    for name, spec in fhcrc_data.TABLES.items():
        tab = tables[name]
        for outcome in spec["outcomes"]:
            rows = fhcrc_data.table_rows(tab, spec["keys"], outcome)
            exact, _ = _queries(tab, spec["keys"], rs)
            assert np.array_equal(oracle.table_lookup(rows, exact), tab[outcome]), (name, outcome)


@pytest.mark.gpu
def test_every_row_of_every_table_on_the_gpu(msim, oracle, tables):
    from microsimulation_b200 import lookups
    rs = np.random.RandomState(4)
    n_lookups = 0
    for name, spec in fhcrc_data.TABLES.items():
        tab = tables[name]
        for outcome in spec["outcomes"]:
            rows = fhcrc_data.table_rows(tab, spec["keys"], outcome)
            T = lookups.Table(rows)
            exact, between = _queries(tab, spec["keys"], rs)
            assert np.array_equal(T(exact), tab[outcome]), (name, outcome)           # Table(key of row i) = outcome of row i
            assert np.array_equal(T(between), oracle.table_lookup(rows, between)), (name, outcome)  # the floor rule
            n_lookups += len(exact) + len(between)
    assert n_lookups > 60000
    # prob_grade7 as a NumericInterpolate, all_cause_mortality of one cohort as an Rpexp: the other two device look-ups
    g7 = tables["prob_grade7"]
    x = rs.uniform(-0.1, 0.7, 5000)
    f = lookups.NumericInterpolate(g7["slope"], g7["Pr.grade.7."])
    np.testing.assert_allclose(f.approx(x), oracle.interpolate(0, g7["slope"], g7["Pr.grade.7."], x), rtol=1e-14)
    ages, rates = fhcrc_data.mortality_hazards(tables, 1950)
    assert len(ages) > 90 and ages[0] == 0
    u = rs.uniform(1e-9, 1, 20000)
    r = lookups.Rpexp(rates, ages)
    np.testing.assert_allclose(r.rand(u), oracle.rpexp(0, rates, ages, u), rtol=1e-13)
    np.testing.assert_allclose(r.rand(u, 50.0), oracle.rpexp(0, rates, ages, u, 50.0), rtol=1e-13)
