"""The oracle against everything the reference itself prints for this path
(SURVEY.md section 4 / 8c): doc outputs, dev-script asserts, RngStream KATs."""
import numpy as np
import pytest

from oracle.pyoracle import sick_sicker_param


def test_rngstream_kats(oracle, kats):
    import ctypes as C
    k = kats["rngstream_first_uniforms"]  # src/rngstream-example.cpp:9-13
    u = oracle.rng_uniforms(12345, 0, 2, 1).ravel()
    assert round(u[0], 6) == k["stream1"]
    assert round(u[1], 6) == k["stream1_substream1"]
    # stream 2 starts one stream jump (2^127 steps = 2^51 substreams of 2^76) after stream 1:
    # RngStream::AdvanceSubstream(e = 51, c = 0)
    g = (C.c_double * 6)(*[12345.0] * 6)
    assert oracle.lib.oracle_rng_advance_e(g, C.c_int(51), C.c_int(0)) == 0
    assert round(oracle.rng_uniforms(list(g), 0, 1, 1).ravel()[0], 6) == k["stream2"]


def test_advance_substream_kat(oracle):
    # substream seed of person 999 (1000 jumps from 12345 x 6)
    assert oracle.rng_advance_substream(12345, 1000) == [3009716804.0, 2079495440.0, 3691030853.0, 1985753873.0,
                                                         2695694265.0, 3749022466.0]


def test_tick_tock(oracle, kats):  # inst/doc/examples.org:162-180
    oracle.set_seed(12345)
    for want in kats["tick_tock"]:
        got = oracle.tick_tock()
        assert [round(x, 6) for x in got] == want


def test_doc_illness_death(oracle, kats):  # inst/doc/examples.org:341-360
    oracle.set_seed(12345)
    t, qc = oracle.doc_illness_death(5)
    assert [round(x, 6) for x in t] == kats["doc_illness_death_times"]
    for (q, c), (wq, wc) in zip(qc, kats["doc_illness_death_qaly_cost"]):
        assert round(q, 6) == wq and round(c, 3) == wc


def test_priority_order(oracle):  # test/test_microsimulation.R:206-242: priorities 1,0,-1,2 at t=10
    assert oracle.priority_demo() == [2.0, 1.0, 0.0, 3.0]


def test_simple_person_doc(oracle, kats):  # inst/doc/examples.org:535-546
    r = oracle.callSimplePerson(5)
    assert [round(x, 2) for x in r["endtime"]] == kats["simple_person_5"]
    assert list(r["id"]) == [0, 0, 1, 2, 3, 3, 4]


def test_simple_person2_readme(oracle, kats):  # README.org:95-107
    k = kats["simple2_events_1e5"]
    ev = oracle.callSimplePerson2(100000)["events"]
    num = ev["number"]
    assert len(num) == k["rows"]
    assert [int((ev["event"] == e).sum()) for e in range(3)] == k["by_event"]
    assert num.max() == k["max"]
    assert round(num.mean(), 1) == k["mean"]
    assert np.median(num) == k["median"]
    assert round(float(np.percentile(num, 25)), 1) == k["q1"] and round(float(np.percentile(num, 75)), 1) == k["q3"]
    assert round(ev["age"].mean(), 2) == k["age_mean"]
    # (the README's state column reads "Healthy:338": the R wrapper's factor conversion of the short
    # Key column, not a property of the C++ tables -- toCancerDeath rows are necessarily in state Cancer)
    assert set(ev["Key"][ev["event"] == 2]) == {1.0}


def test_illness_death_readme(oracle, kats):  # README.org:114-124
    k = kats["illness_prev_1e5_cure02"]
    pv = oracle.callIllnessDeath(100000, cure=0.2)["prev"]
    num = pv["number"]
    assert len(num) == k["rows"] and num.max() == k["max"]
    assert round(num.mean()) == k["mean"]
    assert round(float(np.median(num)) + 1e-9) == k["median"]  # R prints 8095.5 as 8096
    assert round(float(np.percentile(num, 25))) == k["q1"] and round(float(np.percentile(num, 75))) == k["q3"]


def test_dev_script_asserts(oracle, kats):  # test/test_microsimulation.R:1602-1612
    r = oracle.callCalibrationPerson(10, 500)
    assert list(r["DiseaseFree"][:2]) == kats["calibration_seed10_n500_DiseaseFree_first2"]
    le = oracle.callIllnessDeath(10)["pt"]["pt"].sum() / 10
    assert abs(le - kats["illness_death_n10_LE"]) < 1e-3


def test_sick_sicker_trace(oracle, kats):  # README.org:398-434
    tr = oracle.sick_sicker(5, sick_sicker_param(False), trace=True)["trace"]
    want = np.array(kats["sick_sicker_trace"])
    assert tr.shape == want.shape
    assert np.array_equal(tr[:, :3], want[:, :3])
    assert np.array_equal(np.round(tr[:, 3:], 6), want[:, 3:])
    # the treatment arm prints the same trace (README.org:435-470)
    tr2 = oracle.sick_sicker(5, sick_sicker_param(True), trace=True)["trace"]
    assert np.array_equal(tr2, tr)


def test_sick_sicker_costs_qalys(oracle, kats):  # README.org:333-337 (SummaryReport, discounting, costs)
    k = kats["sick_sicker_n1000"]
    for arm, trt in (("no_treatment", False), ("treatment", True)):
        r = oracle.sick_sicker(1000, sick_sicker_param(trt))
        assert round(r["indiv"]["costs"].mean(), 1) == k[arm][0]
        assert round(r["indiv"]["utilities"].mean(), 3) == k[arm][1]


def test_report_structure(oracle):  # tests/testthat/test-callIllnessDeath.R:9-49: "List of 4": pt, ut, events, prev
    for r in (oracle.callIllnessDeath(10), oracle.callSimplePerson2(10)):
        assert list(r["pt"].keys()) == ["Key", "age", "pt"]
        assert list(r["ut"].keys()) == ["Key", "age", "utility"]
        assert list(r["events"].keys()) == ["Key", "event", "age", "number"]
        assert list(r["prev"].keys()) == ["Key", "age", "number"]
    assert oracle.callIllnessDeath(0)["pt"]["pt"].size == 0  # empty list when nothing happened


def test_person_counts_survey(oracle):  # SURVEY.md section 4 extra KATs (oracle-derived)
    r = oracle.person_run(12345, 0, 5, rowlog=True)
    rl = r["rowlog"]
    deaths = rl["endTime"][rl["event"] == 0]
    assert [round(x, 6) for x in deaths] == [9.407128, 5.758767, 136.614857, 60.238907, 154.553957]


def test_nmath_pieces(oracle):
    import math
    from scipy import stats
    ps = np.concatenate([np.linspace(1e-12, 1 - 1e-12, 2001), [1e-300, 1e-20, 0.5, 0.075, 0.925]])
    q = np.array([oracle.qnorm(float(p)) for p in ps])
    np.testing.assert_allclose(q, stats.norm.ppf(ps), rtol=2e-15, atol=1e-15)
    for x in (1.25, 4.0 / 3.0, 0.5, 2.5, 7.0):
        assert abs(oracle.gammafn(x) - math.gamma(x)) / math.gamma(x) < 1e-15
    assert oracle.discountedInterval(1.0, 0.0, 1.0, 0.0) == 1.0
    assert abs(oracle.discountedInterval(2.0, 1.0, 3.0, 0.03) - 2 * (1.03 ** -1 - 1.03 ** -3) / math.log(1.03)) < 1e-14
    # rnorm(mu, 0) draws nothing; exp_rand / norm_rand draw counts
    oracle.set_seed(1)
    oracle.reset_counters()
    oracle.sample("rnorm", 3.0, 0.0, 10)
    assert oracle.counters()["draws"] == 0
    oracle.sample("norm_rand", 0, 0, 10)
    assert oracle.counters()["draws"] == 20
    x = oracle.sample("rexp", 80.0, 0, 200000)
    assert abs(x.mean() - 80.0) < 0.6  # rexp's argument is the MEAN (person-r.cc:102)
