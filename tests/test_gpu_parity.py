"""Parity tests proper: the CUDA path, called through the C ABI (microsimulation_b200 ->
libmsim_b200.so), against the CPU oracle on the same seeds, against the committed
fixtures made from oracle/_ref, and — at BASELINE.json's full sizes — through
size-independent properties.

Bars (north_star): RNG streams, event kinds, ids, states and every integer table
bit-exact; FP64 report sums within 1e-9 relative; event TIMES within a few ulp
(device libm vs glibc for log/pow/exp; DESIGN.md "libm parity")."""
import numpy as np
import pytest

from conftest import assert_report_close, assert_rowlog_close, assert_table_close

pytestmark = pytest.mark.gpu

TIME_RTOL = 4e-15  # <= ~18 ulp guard; measured max is 2 ulp (4.4e-16)


@pytest.fixture(params=["staged", "generic"])
def pmsim(msim, request):
    """The product with person-r on either GPU schedule (csrc/person_staged.cuh / the generic cohort kernel)."""
    import microsimulation_b200.api as api
    api.set_person_schedule(request.param)
    yield msim
    api.set_person_schedule("staged")


# ------------------------------------------------------------------ RNG (K1)

@pytest.mark.parametrize("seed,first,n,draws", [(12345, 0, 100000, 8), (12345, 99999000, 5000, 4),
                                                ([1, 2, 3, 4, 5, 6], 17, 70000, 3), (4294944442, 0, 1000, 16)])
def test_mrg32k3a_substreams_bit_exact(msim, oracle, seed, first, n, draws):
    assert np.array_equal(msim.rng_uniforms(seed, first, n, draws), oracle.rng_uniforms(seed, first, n, draws))


def test_mrg32k3a_golden(msim, golden):
    assert np.array_equal(msim.rng_uniforms(12345, 0, 64, 8), golden["mrg_uniforms_seed12345_sub0_64x8"])
    assert np.array_equal(msim.rng_uniforms(12345, 99999990, 8, 4), golden["mrg_uniforms_seed12345_sub99999990_8x4"])
    assert np.array_equal(np.array(msim.advance_substream(12345, 1000)), golden["advance_substream_12345_1000"])


def test_mersenne_twister_bit_exact(msim, oracle, golden):
    msim.set_seed(12345)
    oracle.set_seed(12345)
    u = msim.mt_uniforms(300000)  # 480 regenerations
    assert np.array_equal(u, oracle.mt_uniforms(300000))
    assert np.array_equal(u[:2000], golden["mt_uniforms_12345_2000"])


def test_bad_seed_is_an_error_code(msim):
    with pytest.raises(msim.MsimError) as e:
        msim.rng_uniforms([0, 0, 0, 1, 2, 3], 0, 10, 1)
    assert e.value.code == -5
    with pytest.raises(msim.MsimError):
        msim.callPersonSimulation(10, seed=[4294967087] * 6)


# ------------------------------------------------------------------ person-r (config 3)

@pytest.mark.parametrize("n", [0, 1, 20, 255, 256, 257, 1000, 70001])
def test_person_rowlog_vs_oracle(pmsim, oracle, n):
    g, r = pmsim.callPersonSimulation(n), oracle.callPersonSimulation(n)
    assert list(g.keys()) == ["endTime", "event", "id", "startTime", "state"]  # std::map key order, person-r.cc:224
    assert len(g["id"]) == len(r["id"])
    assert_rowlog_close(g, r, TIME_RTOL)


def test_person_rowlog_golden(pmsim, golden):
    g = pmsim.callPersonSimulation(2000)
    want = {k: golden["person_n2000_" + k] for k in g}
    assert_rowlog_close(g, want, TIME_RTOL)


def test_person_other_seed_and_shard_offset(pmsim, oracle):
    import microsimulation_b200.api as api
    seed = [11, 22, 33, 44, 55, 66]
    api.person_run_device(seed, 123457, 30000, 1 | 4)
    g = api.person_fetch_rowlog()
    r = oracle.person_run(seed, 123457, 30000)
    assert_rowlog_close(g, r["rowlog"], TIME_RTOL)
    c = api.person_fetch_counts()
    assert np.array_equal(c[:8], r["counts"][:8])
    assert g["id"][0] >= 123457 and g["id"][-1] == 123457 + 30000 - 1


@pytest.mark.parametrize("rate", [0.0, 0.03])
def test_person_event_report_vs_oracle(pmsim, oracle, rate):
    n = 150000
    rep, counts = pmsim.person_report(12345, 0, n, rate)
    ref = oracle.person_run(12345, 0, n, rowlog=False, report=True, discount_rate=rate)
    assert np.array_equal(counts[:8], ref["counts"][:8])
    assert abs(counts[8] - ref["counts"][8]) <= 1e-12 * ref["counts"][8]
    assert_report_close(rep, ref["report"], rtol=1e-9)
    if rate == 0.0:
        assert np.array_equal(rep["ut"]["utility"], rep["pt"]["pt"])  # utility*(b-a) with utility = 1


def test_person_report_golden(pmsim, golden):
    rep, counts = pmsim.person_report(12345, 0, 2000, 0.03)
    assert np.array_equal(counts[:8], golden["person_n2000_counts"][:8])
    for tb in ("pt", "ut", "events", "prev"):
        for k, v in rep[tb].items():
            w = golden["person_n2000_report_%s_%s" % (tb, k)]
            if k in ("pt", "utility"):
                np.testing.assert_allclose(v, w, rtol=1e-9)
            else:
                assert np.array_equal(v, w), (tb, k)


def test_person_counts_1e6_exact(pmsim, golden):
    import microsimulation_b200.api as api
    api.person_run_device(12345, 0, 10 ** 6, 4)
    c = api.person_fetch_counts()
    assert np.array_equal(c[:8], golden["person_n1e6_counts"][:8])  # 991300, 8700, 105038, ...
    assert abs(c[8] - golden["person_n1e6_counts"][8]) <= 1e-12 * c[8]


def test_person_full_size_properties(pmsim, golden):
    """BASELINE config 3 (n = 1e7): properties that need no oracle run of that size."""
    import microsimulation_b200.api as api
    n = 10 ** 7
    res = api.person_run_device(12345, 0, n, 1 | 4)
    counts = api.person_fetch_counts()
    g = api.person_fetch_rowlog()
    assert res.n_rows == len(g["id"]) == int(counts[:8].sum())
    ids = g["id"]
    assert ids[0] == 0 and ids[-1] == n - 1
    d = np.diff(ids)
    assert ((d == 0) | (d == 1)).all()  # person order, nobody missing, rows of a person contiguous
    assert (g["startTime"] <= g["endTime"]).all()
    last = np.r_[d == 1, True]  # last row of each person
    assert np.isin(g["event"][last], (0, 1)).all()  # everybody ends with toDeath / toPCDeath
    assert not np.isin(g["event"][~last], (0, 1)).any()
    first = np.r_[True, d == 1]
    assert (g["startTime"][first] == 0).all() and (g["state"][first] == 0).all()
    cont = ~first
    assert np.array_equal(g["startTime"][cont], g["endTime"][np.flatnonzero(cont) - 1])  # intervals chain
    assert np.array_equal(np.bincount(g["event"].astype(int), minlength=8), counts[:8].astype(int))
    assert abs(g["endTime"].sum() - counts[8]) <= 1e-10 * counts[8]
    # prefix property: the first 1e6 persons are the n = 1e6 run
    head = ids < 10 ** 6
    assert np.array_equal(np.bincount(g["event"][head].astype(int), minlength=8),
                          golden["person_n1e6_counts"][:8].astype(int))
    # shard additivity of counts (the multi-GPU merge is this sum)
    tot = np.zeros(9)
    for k in range(4):
        api.person_run_device(12345, k * n // 4, n // 4, 4)
        tot += api.person_fetch_counts()
    assert np.array_equal(tot[:8], counts[:8])


def test_person_report_is_deterministic_in_integer_cells(pmsim):
    a, _ = pmsim.person_report(12345, 0, 300000, 0.0)
    b, _ = pmsim.person_report(12345, 0, 300000, 0.0)
    for tb in ("events", "prev"):
        assert np.array_equal(a[tb]["number"], b[tb]["number"])
    np.testing.assert_allclose(a["pt"]["pt"], b["pt"]["pt"], rtol=1e-13)


def test_person_schedules_agree_bit_for_bit(msim):
    """Staged vs generic schedule: same device arithmetic, so rows (times included), counts and every integer
    table are IDENTICAL; only the order of FP additions into pt/ut differs."""
    import microsimulation_b200.api as api
    n, out = 10 ** 6 + 77, {}
    for sched in ("generic", "staged"):
        api.set_person_schedule(sched)
        api.person_run_device(12345, 5, n, 1 | 4)
        rows, counts = api.person_fetch_rowlog(), api.person_fetch_counts()
        rep, _ = msim.person_report(12345, 5, n, 0.03)
        out[sched] = (rows, counts, rep)
    api.set_person_schedule("staged")
    _schedules_equal(out["generic"], out["staged"])


def test_person_wave_share_never_changes_results(msim):
    """Two blocks per SM: the favoured block's share of the persons (msim_set_wave_share; adapted from the waves'
    finishing times by default) moves persons between blocks, nothing else: rows, counts and integer tables are
    identical at the smallest, the largest and the adapted share, and equal to the generic schedule's."""
    import microsimulation_b200.api as api
    n, out = 2 * 10 ** 6 + 77, {}
    api.set_person_schedule("generic")
    api.person_run_device(12345, 5, n, 1 | 4)
    ref = (api.person_fetch_rowlog(), api.person_fetch_counts(), msim.person_report(12345, 5, n, 0.03)[0])
    api.set_person_schedule("staged")
    try:
        for share in (0.40, 0.65, 0.0, 0.0):  # adapted: the second run uses what the first one's clocks said
            api.set_wave_share(share) if share else None
            api.person_run_device(12345, 5, n, 1 | 4)
            used = api.last_wave_share()
            assert (abs(used - share) < 1e-3) if share else (0.40 <= used <= 0.65), (share, used)
            got = (api.person_fetch_rowlog(), api.person_fetch_counts(), msim.person_report(12345, 5, n, 0.03)[0])
            _schedules_equal(ref, got)
            if share == 0.65:
                api.set_wave_share(0.0)
        # a launch too small for two waves is not split
        api.person_run_device(12345, 0, 1000, 4)
        assert api.last_wave_share() == 0.0
        with pytest.raises(msim.MsimError):
            api.set_wave_share(0.9)
    finally:
        api.set_wave_share(0.0)


def _schedules_equal(x, y):
    (ra, ca, pa), (rb, cb, pb) = x, y
    for k in ra:
        assert np.array_equal(ra[k], rb[k]), k
    assert np.array_equal(ca[:8], cb[:8]) and abs(ca[8] - cb[8]) <= 1e-12 * ca[8]
    for tb in ("events", "prev"):
        for k in pa[tb]:
            assert np.array_equal(pa[tb][k], pb[tb][k]), (tb, k)
    for tb, k in (("pt", "pt"), ("ut", "utility")):
        assert np.array_equal(pa[tb]["age"], pb[tb]["age"]) and np.array_equal(pa[tb]["Key"], pb[tb]["Key"])
        np.testing.assert_allclose(pa[tb][k], pb[tb][k], rtol=1e-11)


@pytest.mark.parametrize("sched", ["staged", "generic"])
def test_person_rowlog_undersized_plan_is_recovered(msim, oracle, sched):
    """Row-log buffers are sized from a plan (rows per person); a run that needs more is repeated / re-expanded
    with the exact sizes and returns the same rows."""
    import microsimulation_b200.api as api
    api.set_person_schedule(sched)
    api.set_rowlog_plan(1.01, 0.01)
    try:
        g = msim.callPersonSimulation(300000)
    finally:
        api.set_rowlog_plan()
        api.set_person_schedule("staged")
    r = oracle.callPersonSimulation(300000)
    assert len(g["id"]) == len(r["id"])
    assert_rowlog_close(g, r, TIME_RTOL)


# ------------------------------------------------------------------ sequential-stream models (configs 1-2)

@pytest.mark.parametrize("n", [0, 1, 10, 2049, 100000])
def test_simple_person_vs_oracle(msim, oracle, n):
    g, r = msim.callSimplePerson(n), oracle.callSimplePerson(n)
    assert list(g.keys()) == ["endtime", "event", "id", "startTime", "state"]  # simple-example.cc:45 spells "endtime"
    assert len(g["id"]) == len(r["id"])
    assert_rowlog_close(g, r, TIME_RTOL, end_col="endtime")
    assert np.array_equal(msim.mt_uniforms(16), oracle.mt_uniforms(16))  # R's generator left in the same state


def test_simple_person_golden(msim, golden):
    g = msim.callSimplePerson(3000)
    assert_rowlog_close(g, {k: golden["simple_n3000_" + k] for k in g}, TIME_RTOL, end_col="endtime")


@pytest.mark.parametrize("n", [10, 20000, 100000])
def test_simple_person2_vs_oracle(msim, oracle, n):
    assert_report_close(msim.callSimplePerson2(n), oracle.callSimplePerson2(n), rtol=1e-9)


@pytest.mark.parametrize("n,cure,zsd", [(0, 0.1, 0.0), (10, 0.1, 0.0), (100000, 0.1, 0.0), (100000, 0.2, 0.0),
                                        (50000, 0.95, 0.0), (50000, 0.1, 0.5)])
def test_illness_death_vs_oracle(msim, oracle, n, cure, zsd):
    """BASELINE config 1 (n = 1e5, default rates) and the README's cure = 0.2 run; zsd > 0 adds norm_rand's two draws."""
    g, r = msim.callIllnessDeath(n, cure, zsd), oracle.callIllnessDeath(n, cure, zsd)
    assert_report_close(g, r, rtol=1e-9)
    assert np.array_equal(msim.mt_uniforms(8), oracle.mt_uniforms(8))


def test_illness_death_doc_kats(msim, kats):
    k = kats["illness_prev_1e5_cure02"]  # README.org:114-124
    num = msim.callIllnessDeath(100000, cure=0.2)["prev"]["number"]
    assert len(num) == k["rows"] and num.max() == k["max"] and round(num.mean()) == k["mean"]
    le = msim.callIllnessDeath(10)["pt"]["pt"].sum() / 10  # test/test_microsimulation.R:1610
    assert abs(le - kats["illness_death_n10_LE"]) < 1e-3
    ev = msim.callSimplePerson2(100000)["events"]  # README.org:95-107
    k2 = kats["simple2_events_1e5"]
    assert len(ev["number"]) == k2["rows"] and ev["number"].max() == k2["max"]
    assert [int((ev["event"] == e).sum()) for e in range(3)] == k2["by_event"]


def test_reports_golden(msim, golden):
    for name, rep in (("simple2_n20000", msim.callSimplePerson2(20000)),
                      ("illness_n20000_cure01", msim.callIllnessDeath(20000, 0.1, 0.0)),
                      ("illness_n20000_cure02_zsd05", msim.callIllnessDeath(20000, 0.2, 0.5))):
        for tb in ("pt", "ut", "events", "prev"):
            for k, v in rep[tb].items():
                w = golden["%s_%s_%s" % (name, tb, k)]
                if k in ("pt", "utility"):
                    np.testing.assert_allclose(v, w, rtol=1e-9)
                else:
                    assert np.array_equal(v, w), (name, tb, k)


def test_sequential_calls_continue_the_stream(msim, oracle):
    """Two .Calls without re-seeding: the second continues R's stream (Rcpp::RNGScope)."""
    import ctypes as C
    from oracle.pyoracle import Table
    msim.set_seed(777)
    oracle.set_seed(777)
    L = msim.load()
    from microsimulation_b200._lib import Table as MT
    for n in (333, 1000):
        a, b = MT(), Table()
        assert L.msim_callSimplePerson(C.c_int32(n), C.byref(a)) == 0
        assert oracle.lib.oracle_callSimplePerson(C.c_int32(n), C.byref(b)) == 0
        assert_rowlog_close(a.to_dict(), b.to_dict(), TIME_RTOL, end_col="endtime")
        L.msim_table_free(C.byref(a))


def test_simple_person_config2_properties(msim):
    """BASELINE config 2 (n = 1e6): structure checks at full size."""
    g = msim.callSimplePerson(10 ** 6)
    ids = g["id"]
    d = np.diff(ids)
    assert ids[0] == 0 and ids[-1] == 10 ** 6 - 1 and ((d == 0) | (d == 1)).all()
    per = np.bincount(ids.astype(int))
    assert per.min() == 1 and per.max() == 2
    last = np.r_[d == 1, True]
    assert np.isin(g["event"][last], (0, 2)).all()  # toOtherDeath / toCancerDeath end a life
    assert (g["event"][~last] == 1).all()           # a first of two rows is always toCancer


# ------------------------------------------------------------------ calibration (config 5)

def test_calibration_vs_oracle(msim, oracle, kats, golden):
    g = msim.callCalibrationPerson(10, 500)["raw"]
    r = oracle.callCalibrationPerson(10, 500)
    assert list(g["DiseaseFree"][:2]) == kats["calibration_seed10_n500_DiseaseFree_first2"]
    assert set(g) == set(r)
    for k in r:
        if k == "TimeAtRisk":
            np.testing.assert_allclose(g[k], r[k], rtol=1e-12)
        else:
            assert np.array_equal(g[k], r[k]), k
    g = msim.callCalibrationPerson(12345, 20000)["raw"]
    for k in g:
        w = golden["calib_seed12345_n20000_" + k]
        if k == "TimeAtRisk":
            np.testing.assert_allclose(g[k], w, rtol=1e-12)
        else:
            assert np.array_equal(g[k], w), k


def test_calibration_sweep_vs_oracle(msim, oracle):
    rs = np.random.RandomState(7)
    pars = np.array([4, 0.5, 0.05, 10, 3, 0.5]) * (1 + 0.4 * (rs.rand(6, 6) - 0.5))
    pars[1, 1] = 0.0  # sigm1 = 0: rnorm draws nothing
    pars[2, 5] = 0.0  # tau3 = 0
    pars[3, 2] = 0.0  # p2 = 0: nobody has disease potential -> PreClinical/Clinical never observed
    g = msim.calibration_sweep(12345, pars, 40000, first_person=1000)
    r = oracle.calibration_sweep(12345, 1000, 40000, pars)
    for a, b in zip(g, r):
        assert set(a) == set(b)
        for k in b:
            if k == "TimeAtRisk":
                np.testing.assert_allclose(a[k], b[k], rtol=1e-12)
            else:
                assert np.array_equal(a[k], b[k]), k


def test_calibration_sweep_shared_draws_vs_oracle(msim, oracle):
    """The default sweep kernel (calib_shared.cuh: a person's draws once for all parameter sets, no event queue) against
    the oracle's event loop: 6 sets (several persons per set and block), 40 (a ragged block), 300 (two set chunks)."""
    rs = np.random.RandomState(8)
    base = np.array([4, 0.5, 0.3, 10, 3, 0.5])
    for n_sets, n, first in ((6, 40000, 1000), (40, 6000, 0), (300, 1500, 77)):
        pars = base * (1 + 0.6 * (rs.rand(n_sets, 6) - 0.5))
        before = msim.launch_count()
        g = msim.calibration_sweep(12345, pars, n, first_person=first)
        assert msim.launch_count() - before == 2   # one sweep kernel + the TimeAtRisk fold: no repeat
        r = oracle.calibration_sweep(12345, first, n, pars)
        seen = set()
        for a, b in zip(g, r):
            assert set(a) == set(b)
            seen |= set(b)
            for k in b:
                if k == "TimeAtRisk":
                    np.testing.assert_allclose(a[k], b[k], rtol=1e-12)
                else:
                    assert np.array_equal(a[k], b[k]), (n_sets, k)
        assert "Clinical" in seen


def test_calibration_event_queues_agree_bit_for_bit(msim):
    """Shared draws without a queue (default), ordered array, binary heap, and the repeat path (default, then the heap
    again): identical integer tables; TimeAtRisk bit-identical between the event-loop kernels (each thread adds the
    same persons in the same order) and to 1e-12 against the shared-draws kernel, which adds in another order."""
    rs = np.random.RandomState(11)
    pars = np.array([4, 0.5, 0.05, 10, 3, 0.5]) * (1 + 0.4 * (rs.rand(16, 6) - 0.5))
    runs = {}
    try:
        for sched in ("staged", "queue", "generic", "repeat"):
            msim.set_person_schedule(sched)
            before = msim.launch_count()
            runs[sched] = msim.calibration_sweep(12345, pars, 100000, first_person=77)
            runs[sched + "_launches"] = msim.launch_count() - before
    finally:
        msim.set_person_schedule("staged")
    assert runs["repeat_launches"] == 4   # the first attempt and the heap repeat (+ folds)
    assert runs["generic_launches"] == runs["staged_launches"] == runs["queue_launches"] == 2
    for other in ("queue", "generic", "repeat"):
        for a, b in zip(runs["staged"], runs[other]):
            assert set(a) == set(b)
            for k in a:
                if k == "TimeAtRisk":
                    np.testing.assert_allclose(a[k], b[k], rtol=1e-12)
                else:
                    assert np.array_equal(a[k], b[k]), (other, k)
    for a, b in zip(runs["queue"], runs["generic"]):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    for a, b in zip(runs["generic"], runs["repeat"]):
        for k in a:
            assert np.array_equal(a[k], b[k]), k   # the repeat path ends with the heap kernel's result
    again = msim.calibration_sweep(12345, pars, 100000, first_person=77)
    for a, b in zip(runs["staged"], again):
        for k in a:
            assert np.array_equal(a[k], b[k]), k   # run to run: bit-identical, TimeAtRisk included


def test_calibration_config5_full_size_properties(msim):
    """BASELINE config 5 (256 parameter sets x 1e6 persons): properties that need no oracle run of that size."""
    import microsimulation_b200.api as api
    rs = np.random.RandomState(2024)
    base = np.array([4, 0.5, 0.05, 10, 3, 0.5])
    pars = base * (1 + 0.2 * (2 * rs.rand(256, 6) - 1))  # each component within +-20%
    n = 10 ** 6
    full = api.calibration_sweep(12345, pars, n)
    names = ["DiseaseFree", "Precursor", "PreClinical", "Clinical"]
    for r in full:
        occ = sum(r.get(nm, np.zeros(10)) for nm in names)
        assert (np.diff(occ) <= 0).all() and occ[0] <= n  # persons alive at 10, 20, ..., 100: only ever fewer
        assert len(r["TimeAtRisk"]) == 4 and (r["TimeAtRisk"] > 0).all() and (r["TimeAtRisk"] <= 80.0 * n).all()
    # the death time does not depend on the parameters (Gumbel from the person's 4th+ uniform): every set with
    # sigm1 != 0 sees the same persons alive at each decade (common random numbers, R/calibSim.R:12-14)
    alive = [sum(r.get(nm, np.zeros(10)) for nm in names) for r in full]
    assert all(np.array_equal(a, alive[0]) for a in alive)
    # shard additivity (the multi-GPU merge is this sum) and independence of the other sets in the launch
    a = api.calibration_sweep(12345, pars[:3], n // 2, first_person=0)
    b = api.calibration_sweep(12345, pars[:3], n - n // 2, first_person=n // 2)
    for k in range(3):
        for nm in names:
            if nm in full[k]:
                assert np.array_equal(a[k].get(nm, 0) + b[k].get(nm, 0), full[k][nm]), (k, nm)
        np.testing.assert_allclose(a[k]["TimeAtRisk"] + b[k]["TimeAtRisk"], full[k]["TimeAtRisk"], rtol=1e-12)


def test_person_config4_full_size_properties(msim, golden):
    """BASELINE config 4 as built (n = 1e8, EventReport with 3% discounting + counts): exact properties."""
    import microsimulation_b200.api as api
    n, rate = 10 ** 8, 0.03
    rep, counts = msim.person_report(12345, 0, n, rate)
    ev = rep["events"]
    by_kind = np.bincount(ev["event"].astype(int), weights=ev["number"], minlength=8)
    assert np.array_equal(by_kind, counts[:8])                 # the report's events table and the per-kind tally agree
    assert by_kind[0] + by_kind[1] == n                        # everybody ends with toDeath or toPCDeath
    prev = rep["prev"]
    at0 = prev["number"][(prev["Key"] == 0) & (prev["age"] == 0)]
    assert at0.size == 1 and at0[0] == n                       # everybody is Healthy at age 0
    pt, ut = rep["pt"], rep["ut"]
    assert np.array_equal(pt["Key"], ut["Key"]) and np.array_equal(pt["age"], ut["age"])
    assert (ut["utility"] <= pt["pt"] * (1 + 1e-12)).all() and (ut["utility"] > 0).all()  # discounting only shrinks
    healthy = (pt["Key"] == 0) & (pt["age"] < 100)
    assert (np.diff(pt["pt"][healthy]) < 0).all()              # Healthy person-time per year of age only falls
    # four shards of 2.5e7 persons: every integer cell adds up exactly (what the all-reduce relies on)
    tot_counts, dense, dims = np.zeros(9), None, None
    for k in range(4):
        res = api.person_run_device(12345, k * (n // 4), n // 4, api.MSIM_OUT_REPORT | api.MSIM_OUT_COUNTS, rate)
        d = api.dense_fetch(res.dims.n_doubles)
        dense, dims = (d if dense is None else dense + d), res.dims
        tot_counts += api.person_fetch_counts()
    merged = api.dense_to_report(dense, dims, rate)
    assert np.array_equal(tot_counts[:8], counts[:8])
    for tb in ("events", "prev"):
        for k in rep[tb]:
            assert np.array_equal(merged[tb][k], rep[tb][k]), (tb, k)
    np.testing.assert_allclose(merged["pt"]["pt"], rep["pt"]["pt"], rtol=1e-11)
    np.testing.assert_allclose(merged["ut"]["utility"], rep["ut"]["utility"], rtol=1e-11)
    # prefix property: the first 1e6 persons are the n = 1e6 run pinned against the reference's own engine
    api.person_run_device(12345, 0, 10 ** 6, api.MSIM_OUT_COUNTS)
    assert np.array_equal(api.person_fetch_counts()[:8], golden["person_n1e6_counts"][:8])


# ------------------------------------------------------------------ SummaryReport: Sick-Sicker (README.org:148-301)

def _summary_equal(g, r, rtol=1e-9):
    for tb, val in (("pt", "pt"), ("ut", "utility"), ("costs", "cost")):
        assert np.array_equal(g[tb]["Key"], r[tb]["Key"]) and np.array_equal(g[tb]["age"], r[tb]["age"]), tb
        np.testing.assert_allclose(g[tb][val], r[tb][val], rtol=rtol, atol=1e-9, err_msg=tb)
    for tb in ("events", "prev"):
        for k in r[tb]:
            assert np.array_equal(g[tb][k], r[tb][k]), (tb, k)
    assert g["indiv"]["utilities"].shape == r["indiv"]["utilities"].shape
    np.testing.assert_allclose(g["indiv"]["utilities"], r["indiv"]["utilities"], rtol=1e-12)
    np.testing.assert_allclose(g["indiv"]["costs"], r["indiv"]["costs"], rtol=1e-12)


def test_sick_sicker_readme_table(msim, kats):
    """README.org:333-337: mean discounted costs and QALYs, n = 1000, set.seed(12345), both arms."""
    import microsimulation_b200.api as api
    k = kats["sick_sicker_n1000"]
    for arm, trt in (("no_treatment", False), ("treatment", True)):
        r = api.callSim(1000, api.sick_sicker_param(trt))
        assert round(r["indiv"]["costs"].mean(), 1) == k[arm][0]
        assert round(r["indiv"]["utilities"].mean(), 3) == k[arm][1]


@pytest.mark.parametrize("n,trt,indivp,by", [(0, 0, 1, 1.0), (5, 0, 1, 1.0), (1000, 1, 1, 1.0), (20000, 0, 1, 1.0),
                                             (20000, 1, 0, 1.0), (5000, 0, 1, 0.5), (5000, 1, 1, 2.5)])
def test_sick_sicker_stream_regime_vs_oracle(msim, oracle, n, trt, indivp, by):
    import microsimulation_b200.api as api
    par = api.sick_sicker_param(bool(trt))
    par["partitionBy"] = by
    g = api.callSim(n, par, indivp=bool(indivp))
    r = oracle.sick_sicker(n, par, indivp=bool(indivp))
    _summary_equal(g, r)
    if n:  # R's generator is left where the reference leaves it: the next model continues the stream
        assert np.array_equal(api.mt_uniforms(4), _oracle_after_sick(oracle, n, par, indivp))


def _oracle_after_sick(oracle, n, par, indivp):
    import ctypes as C
    from oracle.pyoracle import SICK_SICKER_KEYS, Report
    oracle.set_seed(12345)
    p17 = (C.c_double * 17)(*[float(par[k]) for k in SICK_SICKER_KEYS])
    r = Report()
    assert oracle.lib.oracle_sick_sicker(C.c_int64(n), p17, C.c_int(int(indivp)), C.c_int(0), C.c_int64(0), C.byref(r), None,
                                         C.c_int64(0), C.byref(C.c_int64(0))) == 0
    oracle.lib.oracle_report_free(C.byref(r))
    return oracle.mt_uniforms(4)


@pytest.mark.parametrize("n,first", [(30000, 0), (10000, 123456)])
def test_sick_sicker_substream_regime_vs_oracle(msim, oracle, n, first):
    import microsimulation_b200.api as api
    par = api.sick_sicker_param(True)
    api.set_user_random_seed(12345)
    oracle.set_user_random_seed(12345)
    g = api.callSim(n, par, substreams=True, first_person=first)
    r = oracle.sick_sicker(n, par, substreams=True, first_person=first)
    _summary_equal(g, r)


@pytest.mark.parametrize("n,first,by,rate,indivp", [(0, 0, 50.0, 0.03, 1), (7, 0, 50.0, 0.03, 1), (20000, 0, 20.0, 0.03, 1),
                                                   (20000, 5000, 1.0, 0.0, 1), (30000, 0, 50.0, 0.05, 0)])
def test_summary_report_point_costs_vs_oracle(msim, oracle, n, first, by, rate, indivp):
    """SummaryReport::addPointCost (microsimulation.h:1105-1112) on the GPU: the reference's dev-script example with a
    one-off cost at cancer onset and a management cost every two years (test/test_microsimulation.R:1006-1093),
    against the oracle — cost rows (including cells that only ever received a point cost), per-individual totals."""
    import microsimulation_b200.api as api
    api.set_user_random_seed(12345)
    oracle.set_user_random_seed(12345)
    g = api.cancer_management(n, 8.0, rate, by, indivp=bool(indivp), first_person=first)
    r = oracle.cancer_management(n, 8.0, rate, by, indivp=bool(indivp), first_person=first)
    _summary_equal(g, r)
    if n >= 20000:
        assert g["costs"]["cost"].sum() > 0 and len(g["costs"]["cost"]) >= len(g["pt"]["pt"])
        # the Cancer state's costs are point costs only (the report's interval cost is never set: 0)
        np.testing.assert_allclose(g["indiv"]["costs"].sum(), g["costs"]["cost"].sum(), rtol=1e-12)


@pytest.mark.parametrize("rate,start_age,part", [(0.0, 0.0, (0.0, 100.0, 1.0)), (0.03, 0.0, (0.0, 100.0, 5.0)),
                                                  (0.03, 50.0, (0.0, 100.0, 10.0)), (0.05, 30.0, (20.0, 90.0, 2.5))])
def test_cost_report_add_vs_oracle(msim, oracle, rate, start_age, part):
    """CostReport::add / addPointCost (microsimulation.h:1265-1278) as a batch on the GPU against the oracle's
    CostReport: the wrap() table and the individual's running total (discounted from startReportAge on)."""
    import ctypes as C
    import microsimulation_b200.api as api
    rng = np.random.RandomState(7)
    n = 50000
    state = rng.randint(0, 5, n).astype(np.int32)
    # (times below the partition's first point have no bin in the reference: lower_bound() returns end(), :1270)
    time = np.round(rng.uniform(part[0], 110, n), 3)
    time[:100] = part[0] + part[2] * np.arange(100) % (part[1] - part[0])   # exactly on partition points
    cost = np.round(rng.uniform(0, 2e4, n), 2)
    cost[100:120] = 0.0
    got, cur = api.cost_report(state, time, cost, rate, start_age, part, n_state=8)
    rows = np.zeros(3 * 4096)
    ref_cur = C.c_double()
    dp = C.POINTER(C.c_double)
    k = oracle.lib.oracle_cost_report(C.c_double(rate), C.c_double(start_age), C.c_double(part[0]), C.c_double(part[1]), C.c_double(part[2]),
                                      C.c_int64(n), state.ctypes.data_as(C.POINTER(C.c_int)), time.ctypes.data_as(dp),
                                      cost.ctypes.data_as(dp), rows.ctypes.data_as(dp), C.c_int64(4096), C.byref(ref_cur))
    ref = rows[:3 * k].reshape(k, 3)
    assert np.array_equal(got["Key"], ref[:, 0]) and np.array_equal(got["age"], ref[:, 1])
    np.testing.assert_allclose(got["cost"], ref[:, 2], rtol=1e-11)
    np.testing.assert_allclose(cur, ref_cur.value, rtol=1e-11)
    with pytest.raises(msim.MsimError):
        api.cost_report(np.array([9], dtype=np.int32), [1.0], [1.0], n_state=8)


def test_sick_sicker_shards_add_up_to_the_single_run(msim):
    """first_person is the sharding knob of the substream regime (include/msim.h): the report's utility and cost
    carry over from one person to the next in the reference (README.org:252-281), so a shard starts with what the
    persons BEFORE it left — two shards must give exactly the single run's per-person values and, added, its tables."""
    import microsimulation_b200.api as api
    par = api.sick_sicker_param(True)
    n, cut = 40000, 17123
    api.set_user_random_seed(12345)
    whole = api.callSim(n, par, substreams=True, first_person=0)
    api.set_user_random_seed(12345)
    a = api.callSim(cut, par, substreams=True, first_person=0)
    api.set_user_random_seed(12345)
    b = api.callSim(n - cut, par, substreams=True, first_person=cut)
    for col in ("utilities", "costs"):
        assert np.array_equal(np.r_[a["indiv"][col], b["indiv"][col]], whole["indiv"][col]), col
    for tb, val in (("pt", "pt"), ("ut", "utility"), ("costs", "cost"), ("events", "number"), ("prev", "number")):
        keys = [k for k in whole[tb] if k != val]
        acc = {}
        for part in (a, b):
            for i in range(len(part[tb][val])):
                kt = tuple(part[tb][k][i] for k in keys)
                acc[kt] = acc.get(kt, 0.0) + part[tb][val][i]
        got = np.array([acc[tuple(whole[tb][k][i] for k in keys)] for i in range(len(whole[tb][val]))])
        assert len(acc) == len(whole[tb][val]), tb
        np.testing.assert_allclose(got, whole[tb][val], rtol=1e-12, err_msg=tb)


@pytest.mark.parametrize("substreams", [False, True])
def test_sick_sicker_event_queues_agree(msim, oracle, substreams):
    """Default (small ordered array, cancelled actions removed), heap.h's binary heap with lingering cancelled entries,
    and the repeat path (default, then the heap kernels again): the same report, and the same generator state after."""
    import microsimulation_b200.api as api
    par = api.sick_sicker_param(True)
    runs, after = {}, {}
    try:
        for sched in ("staged", "generic", "repeat"):
            msim.set_person_schedule(sched)
            if substreams:
                api.set_user_random_seed(12345)
            before = msim.launch_count()
            runs[sched] = api.callSim(8000, par, substreams=substreams, first_person=17 if substreams else 0)
            runs[sched + "_launches"] = msim.launch_count() - before
            after[sched] = api.user_random_seed() if substreams else api.mt_state().copy()
    finally:
        msim.set_person_schedule("staged")
    assert runs["repeat_launches"] == runs["staged_launches"] + runs["generic_launches"]
    for other in ("generic", "repeat"):
        _summary_equal(runs["staged"], runs[other], rtol=1e-12)
        assert np.array_equal(after["staged"], after[other])
        # per-person totals do not depend on any accumulation order: bit for bit
        assert np.array_equal(runs["staged"]["indiv"]["utilities"], runs[other]["indiv"]["utilities"])
        assert np.array_equal(runs["staged"]["indiv"]["costs"], runs[other]["indiv"]["costs"])
    if substreams:
        oracle.set_user_random_seed(12345)
    _summary_equal(runs["staged"], oracle.sick_sicker(8000, par, substreams=substreams, first_person=17 if substreams else 0))


def test_discounting_free_functions(msim, oracle):
    """discountedInterval / discountedPoint (microsimulation.h:811-836) as device functions: pow() differs from glibc's
    in the last bits only."""
    import microsimulation_b200.api as api
    rs = np.random.RandomState(3)
    a = rs.uniform(0, 90, 5000)
    b = a + rs.uniform(0, 30, 5000)
    y = rs.uniform(0.1, 5000, 5000)
    for rate in (0.0, 0.03, 0.05):
        want = np.array([oracle.discountedInterval(float(yy), float(aa), float(bb), rate) for yy, aa, bb in zip(y, a, b)])
        np.testing.assert_allclose(api.discountedInterval(y, a, b, rate), want, rtol=1e-12)
        wantp = y * (1.0 + rate) ** (-a) if rate > 0 else y
        np.testing.assert_allclose(api.discountedPoint(y, a, rate), wantp, rtol=1e-13)


def test_package_seed_advances_like_the_reference(msim, oracle):
    """Two calibration calls after ONE set.user.Random.seed: the second `new Rng()` is the next stream."""
    import ctypes as C
    msim.set_user_random_seed(99)
    oracle.set_user_random_seed(99)
    from microsimulation_b200._lib import Calibration as MC
    from oracle.pyoracle import Calibration as OC
    par = (C.c_double * 6)(4, 0.5, 0.05, 10, 3, 0.5)
    for _ in range(2):
        a, b = MC(), OC()
        assert msim.load().msim_callCalibrationSimulation(C.c_int32(3000), par, C.byref(a)) == 0
        assert oracle.lib.oracle_callCalibrationSimulation(C.c_int32(3000), par, C.byref(b)) == 0
        assert list(a.occupancy) == list(b.occupancy)


def test_calibration_person_mc_cores_like_mclapply(msim, oracle):
    """R/calibSim.R:16-39 with mc.cores = 3: chunk i runs its persons on the stream after chunk i-1's; the sum depends on
    mc.cores exactly as in the reference."""
    n, runpar = 4000, [4, 0.5, 0.05, 10, 3, 0.5]
    got = msim.callCalibrationPerson(777, n, runpar, mc_cores=3)
    states = ["DiseaseFree", "Precursor", "PreClinical", "Clinical"]
    mat, tar = np.zeros((10, 4)), np.zeros(4)
    seed = [777] * 6
    for i, n_chunk in enumerate(msim.calibration_chunks(n, 3)):
        if i:
            seed = msim.next_rng_stream(seed)   # checked against the reference's RngStream.cpp in test_host_logic.py
        d = oracle.callCalibrationPerson(seed, n_chunk, runpar)
        for j, nm in enumerate(states):
            if nm in d:
                mat[:, j] += d[nm]
        tar[:len(d["TimeAtRisk"])] += d["TimeAtRisk"]
    assert np.array_equal(got["StateOccupancy"], mat)
    np.testing.assert_allclose(got["TimeAtRisk"], tar, rtol=1e-12)
    one = msim.callCalibrationPerson(777, n, runpar, mc_cores=1)
    assert not np.array_equal(one["StateOccupancy"], mat)   # as in the reference, mc.cores changes the random numbers


def test_gpu_against_the_reference_linked_oracle(msim, oracle_ref):
    """The same comparisons with the oracle build that links the reference's OWN ssim.cc / RngStream.cpp / heap.h
    (oracle/_ref, prebuilt where /root/reference exists): event loop, heap order and generator are the reference's."""
    import microsimulation_b200.api as api
    assert oracle_ref.flavour == "reference"
    assert_rowlog_close(msim.callPersonSimulation(30000), oracle_ref.callPersonSimulation(30000))
    rep, counts = msim.person_report(12345, 500, 40000, 0.03)
    ref = oracle_ref.person_run(12345, 500, 40000, rowlog=False, report=True, discount_rate=0.03)
    assert np.array_equal(counts[:8], ref["counts"][:8])
    assert_report_close(rep, ref["report"])
    assert_report_close(msim.callIllnessDeath(50000, 0.1, 0.0), oracle_ref.callIllnessDeath(50000, 0.1, 0.0))
    assert_rowlog_close(msim.callSimplePerson(50000), oracle_ref.callSimplePerson(50000), end_col="endtime")
    g, r = msim.callCalibrationPerson(10, 20000)["raw"], oracle_ref.callCalibrationPerson(10, 20000)
    assert set(g) == set(r)
    for k in r:
        if k == "TimeAtRisk":
            np.testing.assert_allclose(g[k], r[k], rtol=1e-12)
        else:
            assert np.array_equal(g[k], r[k]), k
    par = api.sick_sicker_param(True)
    _summary_equal(api.callSim(3000, par), oracle_ref.sick_sicker(3000, par))


@pytest.mark.parametrize("start,finish,delta,max_time,sra,rate,out_ut", [
    (0.0, 100.0, 5.0, 1.0e100, 50.0, 0.03, True),      # 0..100 by 5, startReportAge 50
    (0.0, 100.0, 1.0, 1.0e100, 0.0, 0.03, True),       # the unit partition through the general entry (maxTime 1e100)
    (0.0, 90.0, 0.1 * 9, 1.0e6, 20.0, 0.05, True),     # points by repeated addition of an inexact delta
    (0.0, 100.0, 10.0, 1.0e100, 65.0, 0.0, False),     # no discounting, outputUtilities = false
])
def test_person_report_options_vs_oracle(msim, oracle, start, finish, delta, max_time, sra, rate, out_ut):
    """EventReport's surface beyond the built models' fixed partition (microsimulation.h:873-932): setPartition(start,
    finish, delta, maxTime), startReportAge, outputUtilities, and the per-individual discounted totals of addBrief /
    individualReset with their Means — the GPU against the oracle's EventReport, cell by cell and person by person."""
    import microsimulation_b200.api as api
    n = 60000
    api.set_report_options(start, finish, delta, max_time, sra, out_ut)
    try:
        rep, counts, indiv, means = api.person_report(12345, 77, n, rate, indiv=True)
    finally:
        api.set_report_options()
    ref, rcounts, rindiv, rmeans = oracle.person_run_options(12345, 77, n, rate, start, finish, delta, max_time, sra, out_ut)
    assert np.array_equal(counts[:8], rcounts[:8])
    for tb in ("pt", "events", "prev") + (("ut",) if out_ut else ()):
        assert_table_close(rep[tb], ref[tb], 1e-9)
    if not out_ut:
        assert len(rep["ut"]["utility"]) == 0
    assert np.array_equal(np.isnan(indiv), np.isnan(rindiv))     # NA exactly where the last event came before startReportAge
    assert 0 < np.isnan(indiv).sum() < n or sra == 0.0
    ok = ~np.isnan(indiv)
    np.testing.assert_allclose(indiv[ok], rindiv[ok], rtol=1e-11, atol=1e-12)
    assert means["n"] == rmeans["n"] == ok.sum()
    for k in ("mean", "var", "sd", "se", "sum", "sumsq"):
        np.testing.assert_allclose(means[k], rmeans[k], rtol=1e-10, err_msg=k)
    # the default is restored: the built models' partition again
    again, _ = api.person_report(12345, 77, 2000, 0.03)
    assert again["pt"]["age"].max() in (1.0e6, 100.0) and len(set(again["pt"]["age"])) > 50


def test_report_options_are_validated(msim):
    import microsimulation_b200.api as api
    for bad in ((0.0, 100.0, 0.0), (0.0, -1.0, 1.0), (0.0, 100.0, 0.5)):   # delta 0; finish < start; 201 points
        with pytest.raises(msim.MsimError):
            api.set_report_options(*bad)
    api.set_report_options()
    with pytest.raises(msim.MsimError):   # individual totals need the report
        api.person_run_device(12345, 0, 100, 4 | 8, 0.03)


def test_person_1e7_rows_and_report_against_the_reference_on_all_cores(msim):
    """BASELINE configs[2] at its full size, and the bench's report workload at a size where a 1-ulp libm difference
    could move an event across an age-bin edge: 1e7 persons on the GPU against the reference-linked oracle run on
    every host core (oracle/parallel.py) — every row (ids, states, events exactly, times to a few ulp) and every
    cell of the EventReport (events / prev exactly, pt / ut to 1e-9)."""
    import os
    import microsimulation_b200.api as api
    from oracle import parallel
    n = 10 ** 7
    api.person_run_device(12345, 0, n, 1 | 4)
    rows = api.person_fetch_rowlog()
    counts = api.person_fetch_counts()
    assert np.all(np.diff(rows["id"]) >= 0)
    ref = parallel.person_run_parallel(12345, 0, n, cores=os.cpu_count(), rows_to_check=rows, report=True,
                                       discount_rate=0.03, time_rtol=TIME_RTOL)
    assert ref["rows_ok"], ref["rows_msg"]
    assert ref["n_rows"] == len(rows["id"])
    assert np.array_equal(counts[:8], ref["counts"][:8])
    assert ref["max_time_rel_err"] <= TIME_RTOL
    del rows
    rep, counts2 = msim.person_report(12345, 0, n, 0.03)
    assert np.array_equal(counts2[:8], ref["counts"][:8])
    ok, worst, msg = parallel.report_equal(rep, ref["report"], rtol=1e-9)
    assert ok, msg
    assert_report_close(rep, ref["report"], rtol=1e-9)


def test_null_arguments_with_a_device(msim):
    """The same with the device initialised: argument checks that sit behind the device check are reached."""
    import os
    import fuzz_null_args
    os.environ["MSIM_FUZZ_INIT"] = "1"
    try:
        res = fuzz_null_args.run_all()
    finally:
        del os.environ["MSIM_FUZZ_INIT"]
    crashed = {k: v for k, v in res.items() if v[0] != 0}
    assert not crashed, crashed


def test_kernels_were_launched(msim):
    before = msim.launch_count()
    msim.callPersonSimulation(1000)
    assert msim.launch_count() > before


def test_peer_merge_entries_on_one_gpu(msim):
    """msim_peer_*: the merge across GPUs inside the launch (the multi-GPU behaviour itself is checked by
    scripts/multi_gpu_check.py under torchrun and by bench.py's merge_check at N > 1).  On one GPU: call order is
    checked, and a world of one rank changes nothing."""
    import microsimulation_b200.api as api
    from microsimulation_b200._lib import load
    api.peer_close()
    with pytest.raises(RuntimeError):
        api.peer_merge(1)                      # no mailbox yet
    with pytest.raises(RuntimeError):
        api.peer_export(9)                     # at most 8 ranks
    with pytest.raises(RuntimeError):
        api.peer_import(0, 1, [b"\0" * 64])    # export first
    rep0, counts0 = msim.person_report(12345, 100, 30000, 0.03)
    h = api.peer_export(1)
    assert len(h) == 64
    with pytest.raises(ValueError):
        api.peer_import(0, 1, [h, h])
    api.peer_import(0, 1, [h])
    with pytest.raises(RuntimeError):
        api.peer_merge(3)
    try:
        for mode in (1, 2):
            api.peer_merge(mode)
            rep, counts = msim.person_report(12345, 100, 30000, 0.03)
            api.peer_flush()
            assert np.array_equal(counts[:8], counts0[:8])
            for tb in ("events", "prev"):
                for k in rep0[tb]:
                    assert np.array_equal(rep[tb][k], rep0[tb][k])
    finally:
        api.peer_merge(0)
        api.peer_close()
    assert load().msim_peer_flush() == 0
