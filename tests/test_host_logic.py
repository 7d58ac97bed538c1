"""CPU-only checks of the product's host logic and of the device engine's
per-person code compiled for the host (tests/host_emulation): RNG jump tables
and strides, the event heap, report accumulation, the sequential-stream offset
resolution.  No kernel runs here; the GPU parity tests are tests/test_gpu_parity.py."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, assert_report_close, assert_rowlog_close
from oracle.pyoracle import Calibration, PersonShard, Report, Table


def test_library_loads_and_exports_every_declared_symbol():
    import microsimulation_b200 as m
    L = m.load()
    hdr = open(os.path.join(ROOT, "include", "msim.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(msim_[a-zA-Z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    for sym in declared:
        assert hasattr(L, sym), "libmsim_b200.so does not export " + sym
    assert sorted(m.SYMBOLS) == declared
    assert b"sm_100a" in L.msim_version()


def test_no_cpu_fallback_without_a_device():
    import torch
    import microsimulation_b200 as m
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(m.MsimError) as e:
        m.callPersonSimulation(5)
    assert e.value.code == -1  # MSIM_ERR_NO_DEVICE
    with pytest.raises(m.MsimError):
        m.callIllnessDeath(5)


@pytest.mark.parametrize("header", ["msim.h", "msim_plugin.h"])
def test_public_headers_are_plain_c(header):
    """The drop-in boundary is a C ABI: both headers compile as C99 (pedantic) and as C++11, with nothing but <stdint.h>."""
    import subprocess
    inc = os.path.join(ROOT, "include")
    for cc, std, lang in (("gcc", "-std=c99", "c"), ("g++", "-std=c++11", "c++")):
        r = subprocess.run([cc, std, "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + inc, "-fsyntax-only", "-x", lang, "-"],
                           input='#include "%s"\nint main(void) { return 0; }\n' % header, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    src = open(os.path.join(inc, header)).read()
    includes = re.findall(r"#include\s*[<\"]([^>\"]+)", src)
    assert set(includes) <= {"stdint.h", "stddef.h", "msim.h"}, includes   # no torch, CUDA, R or C++ headers


def test_null_arguments_are_error_codes_not_crashes():
    """Every entry point of include/msim.h with null pointers: an error code (or a no-op for the free functions)."""
    import fuzz_null_args
    res = fuzz_null_args.run_all()
    assert len(res) >= 30
    crashed = {k: v for k, v in res.items() if v[0] != 0}
    assert not crashed, crashed


def test_product_never_touches_the_oracle():
    """No import / include / dlopen of anything under oracle/ (or of the host emulation) in the product."""
    pkg = os.path.join(ROOT, "microsimulation_b200")
    bad = re.compile(r"(#include\s*[\"<][^\n]*oracle|import\s+oracle|from\s+oracle|pyoracle|libmsim_oracle|"
                     r"host_emulation/|libmsim_emu)")
    for top in (pkg, os.path.join(ROOT, "scripts"), os.path.join(ROOT, "glue"), os.path.join(ROOT, "examples"),
                os.path.join(ROOT, "include")):
        for dirpath, _, files in os.walk(top):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp", ".sh")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not bad.search(src), os.path.join(dirpath, f)


def test_host_rng_state_machine_matches_oracle(oracle):
    """.C table semantics: package seed, default stream, advance.substream (src/microsimulation.cc:33-75)."""
    import microsimulation_b200 as m
    m.set_user_random_seed(12345)
    oracle.set_user_random_seed(12345)
    assert m.user_random_seed()[1:] == [int(x) for x in oracle.get_user_random_seed()]
    for _ in range(3):
        m.next_user_random_substream()
        oracle.next_rng_substream()
    assert m.user_random_seed()[1:] == [int(x) for x in oracle.get_user_random_seed()]
    for seed, n in ((12345, 1000), ([1, 2, 3, 4, 5, 6], 99999999), (12345, -7), ([407, 9, 8, 7, 6, 5, 4], 3)):
        assert m.advance_substream(seed, n) == oracle.rng_advance_substream(seed, n)
    with pytest.raises(m.MsimError) as e:
        m.set_user_random_seed([0, 0, 0, 1, 1, 1])
    assert e.value.code == -5


def test_user_unif_rand_hook_matches_oracle(oracle):
    """`user_unif_rand` (src/microsimulation.cc:77-85): R-level runif() under RNGkind("user") draws from the default
    stream; next.user.Random.substream() moves it to the next substream."""
    import microsimulation_b200 as m
    m.set_user_random_seed(12345)
    want = oracle.rng_uniforms(12345, 0, 3, 7)
    assert np.array_equal(m.user_unif_rand(7), want[0])
    assert abs(want[0][0] - 0.1270111) < 1e-7  # the RngStream paper's first value for seed 12345 x 6
    m.next_user_random_substream()
    assert np.array_equal(m.user_unif_rand(5), want[1][:5])
    m.next_user_random_substream()
    assert np.array_equal(m.user_unif_rand(7), want[2])
    seed_after = m.user_random_seed()[1:]   # the state moved with the draws (r_get_user_random_seed returns Cg)
    m.set_user_random_seed(seed_after)
    oracle_after = oracle.rng_uniforms(seed_after, 0, 1, 2)[0]
    assert np.array_equal(m.user_unif_rand(2), oracle_after)


def test_next_rng_stream_matches_the_reference_jump(oracle, oracle_ref):
    """parallel::nextRNGStream (used by R/calibSim.R:21 and the RNGStream R class) is A^(2^127): 2^51 substreams of
    2^76 steps, which RngStream::AdvanceSubstream(51, 0) of the reference's own RngStream.cpp computes."""
    import microsimulation_b200 as m
    for seed in ([12345] * 6, [1, 2, 3, 4, 5, 6], [4294967086, 4294967086, 4294967086, 4294944442, 4294944442, 4294944442]):
        want = (C.c_double * 6)(*[float(x) for x in seed])
        assert oracle_ref.lib.oracle_rng_advance_e(want, 51, 0) == 0
        port = (C.c_double * 6)(*[float(x) for x in seed])
        assert oracle.lib.oracle_rng_advance_e(port, 51, 0) == 0
        assert m.next_rng_stream(seed) == [int(x) for x in want] == [int(x) for x in port]
    # the reference's own known answers (src/rngstream-example.cpp:9-13): first uniform of stream 1, of stream 2, and
    # of stream 1's second substream, through the hook R-level runif() uses
    s2 = m.next_rng_stream([12345] * 6)
    assert s2 == [3692455944, 1366884236, 2968912127, 335948734, 4161675175, 475798818]
    m.set_user_random_seed(12345)
    assert round(float(m.user_unif_rand(1)[0]), 6) == 0.127011
    m.set_user_random_seed(s2)
    assert round(float(m.user_unif_rand(1)[0]), 6) == 0.759582
    m.set_user_random_seed(12345)
    m.next_user_random_substream()
    assert round(float(m.user_unif_rand(1)[0]), 6) == 0.079399
    with pytest.raises(m.MsimError):
        m.next_rng_stream([0, 0, 0, 1, 1, 1])
    assert m.calibration_chunks(10, 3) == [4, 3, 3] and m.calibration_chunks(2, 4) == [1, 1, 0, 0]


def test_mrg_modular_arithmetic_against_128_bit_integers(emu):
    """mrg_fold_small / mrg_reduce / mrg_dot3 (csrc/mrg32k3a.cuh, the code the device runs) on random operands and
    on the extremes: residues m - 1, full high words, sums that carry out of 32 bits."""
    emu.emu_mrg_selftest.restype = C.c_int64
    assert emu.emu_mrg_selftest(C.c_uint64(12345), C.c_int64(3000000)) == 0


def test_mt_set_seed_matches_oracle(oracle, emu):
    import microsimulation_b200 as m
    m.set_seed(12345)
    st = (C.c_uint32 * 625)()
    emu.emu_mt_set_seed(C.c_uint32(12345), st)
    assert np.array_equal(m.mt_state(), np.array(st[:], dtype=np.uint32))
    u = np.empty(1500)
    emu.emu_mt_uniforms(st, C.c_int64(1500), u.ctypes.data_as(C.POINTER(C.c_double)))
    oracle.set_seed(12345)
    assert np.array_equal(u, oracle.mt_uniforms(1500))  # crosses two three-wave regenerations
    m.set_mt_state(m.mt_state())
    with pytest.raises(m.MsimError):
        m.set_mt_state([0] * 625)


@pytest.mark.parametrize("first,n,draws,grid", [(0, 3000, 4, 1), (5, 5000, 8, 7), (99999990, 700, 3, 2)])
def test_emulated_substream_uniforms(oracle, emu, first, n, draws, grid):
    seed = (C.c_double * 6)(*[12345.0] * 6)
    out = np.empty(n * draws)
    emu.emu_rng_uniforms(seed, C.c_int64(first), C.c_int64(n), C.c_int32(draws), C.c_int(grid),
                         out.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.array_equal(out.reshape(n, draws), oracle.rng_uniforms(12345, first, n, draws))


@pytest.mark.parametrize("first,n,grid", [(0, 20000, 3), (12345, 5000, 40), (0, 1, 1), (7, 255, 2), (0, 257, 1)])
def test_emulated_person_r(oracle, emu, first, n, grid):
    """Same templates and the same exp / log / pow (csrc/fastmath.cuh) as the kernel: ids, states and events must be
    BIT-exact, event times within a few ulp of the oracle's glibc values, report sums to 1e-12."""
    seed = (C.c_double * 6)(*[12345.0] * 6)
    sh = PersonShard(seed, first, n, 7, 0.03)
    t, r, cnt = Table(), Report(), (C.c_double * 9)()
    assert emu.emu_person_run(C.byref(sh), C.c_int(grid), C.byref(t), cnt, C.byref(r)) == 0
    e, er = t.to_dict(), r.to_dict()
    ref = oracle.person_run(12345, first, n, rowlog=True, report=True, discount_rate=0.03)
    assert_rowlog_close(e, ref["rowlog"], 4e-15)
    assert np.array_equal(np.array(cnt[:8]), ref["counts"][:8])
    assert_report_close(er, ref["report"], rtol=1e-12)
    emu.emu_table_free(C.byref(t))
    emu.emu_report_free(C.byref(r))


@pytest.mark.parametrize("first,n,grid", [(0, 20000, 3), (12345, 5000, 40), (0, 1, 1), (7, 255, 2), (0, 257, 1), (3, 0, 1),
                                          (0, 100000, 7)])
def test_emulated_person_r_staged_schedule(oracle, emu, first, n, grid):
    """The staged schedule (stage functions + per-warp queues of person_staged.cuh) gives the reference's
    rows (times to a few ulp: fastmath.cuh against glibc): regrouping persons between stages changes when arithmetic happens, not what is drawn."""
    seed = (C.c_double * 6)(*[12345.0] * 6)
    sh = PersonShard(seed, first, n, 7, 0.03)
    t, r, cnt = Table(), Report(), (C.c_double * 9)()
    assert emu.emu_person_run_staged(C.byref(sh), C.c_int(grid), C.byref(t), cnt, C.byref(r)) == 0
    e, er = t.to_dict(), r.to_dict()
    ref = oracle.person_run(12345, first, n, rowlog=True, report=True, discount_rate=0.03)
    assert_rowlog_close(e, ref["rowlog"], 4e-15)
    assert np.array_equal(np.array(cnt[:8]), ref["counts"][:8])
    if n:  # FP sums are taken in another order than the reference's person order
        assert_report_close(er, ref["report"], rtol=1e-10)
    emu.emu_table_free(C.byref(t))
    emu.emu_report_free(C.byref(r))


@pytest.mark.parametrize("first,n,grid,share", [(0, 20000, 4, 0.55), (12345, 5000, 40, 0.40), (7, 255, 2, 0.65),
                                                (0, 100000, 8, 0.5), (3, 40, 2, 0.6), (0, 33, 2, 0.5)])
def test_emulated_person_r_two_waves(oracle, emu, first, n, grid, share):
    """The staged kernel's two waves of blocks (the favoured blocks take the tiles [0, split), the others the rest, each
    wave from the substream state of its range's first person): every person still gets its own substream and is
    simulated once, whatever the share."""
    seed = (C.c_double * 6)(*[12345.0] * 6)
    sh = PersonShard(seed, first, n, 7, 0.03)
    t, r, cnt = Table(), Report(), (C.c_double * 9)()
    assert emu.emu_person_run_staged_waves(C.byref(sh), C.c_int(grid), C.c_double(share), C.byref(t), cnt, C.byref(r)) == 0
    e, er = t.to_dict(), r.to_dict()
    ref = oracle.person_run(12345, first, n, rowlog=True, report=True, discount_rate=0.03)
    assert_rowlog_close(e, ref["rowlog"], 4e-15)
    assert np.array_equal(np.array(cnt[:8]), ref["counts"][:8])
    assert_report_close(er, ref["report"], rtol=1e-10)
    emu.emu_table_free(C.byref(t))
    emu.emu_report_free(C.byref(r))
    assert emu.emu_person_run_staged_waves(C.byref(sh), C.c_int(3), C.c_double(0.5), C.byref(t), cnt, C.byref(r)) == -3  # odd grid


@pytest.mark.parametrize("n", [0, 1, 10, 700, 5000, 60000])
def test_emulated_sequential_stream_models(oracle, emu, n):
    """Per-position draw counts + chunked chain resolution == the reference's person-after-person consumption."""
    st = (C.c_uint32 * 625)()
    emu.emu_mt_set_seed(C.c_uint32(12345), st)
    t = Table()
    assert emu.emu_sequential(0, C.c_int64(n), C.c_double(0), C.c_double(0), st, C.byref(t), None) == 0
    e, ref = t.to_dict(), oracle.callSimplePerson(n)
    assert_rowlog_close(e, ref, 4e-15, end_col="endtime")  # times: fastmath.cuh against glibc
    # R's generator is left exactly where the reference leaves it
    u = np.empty(5)
    emu.emu_mt_uniforms(st, C.c_int64(5), u.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.array_equal(u, oracle.mt_uniforms(5))
    emu.emu_mt_set_seed(C.c_uint32(12345), st)
    r = Report()
    assert emu.emu_sequential(1, C.c_int64(n), C.c_double(0), C.c_double(0), st, None, C.byref(r)) == 0
    assert_report_close(r.to_dict(), oracle.callSimplePerson2(n), rtol=1e-13)
    for cure, zsd in ((0.1, 0.0), (0.2, 0.0), (0.9, 0.0), (0.1, 0.5)):
        emu.emu_mt_set_seed(C.c_uint32(12345), st)
        r = Report()
        assert emu.emu_sequential(2, C.c_int64(n), C.c_double(cure), C.c_double(zsd), st, None, C.byref(r)) == 0
        assert_report_close(r.to_dict(), oracle.callIllnessDeath(n, cure, zsd), rtol=1e-13)


def test_emulated_sequential_from_mid_batch_state(oracle, emu):
    """The MT position need not be at a regeneration boundary: a second call continues where the first
    stopped, on both sides (Rcpp::RNGScope writes R's state back, illness-death.cpp:75)."""
    st = (C.c_uint32 * 625)()
    emu.emu_mt_set_seed(C.c_uint32(42), st)
    oracle.set_seed(42)
    t1, t1_ref = Table(), Table()
    assert emu.emu_sequential(0, C.c_int64(400), C.c_double(0), C.c_double(0), st, C.byref(t1), None) == 0
    assert oracle.lib.oracle_callSimplePerson(C.c_int32(400), C.byref(t1_ref)) == 0
    assert 1 <= st[0] <= 624
    t2, t2_ref = Table(), Table()
    assert emu.emu_sequential(0, C.c_int64(300), C.c_double(0), C.c_double(0), st, C.byref(t2), None) == 0
    assert oracle.lib.oracle_callSimplePerson(C.c_int32(300), C.byref(t2_ref)) == 0
    a, b = t2.to_dict(), t2_ref.to_dict()
    assert_rowlog_close(a, b, 4e-15, end_col="endtime")


def test_emulated_calibration(oracle, emu):
    seed = (C.c_double * 6)(*[10.0] * 6)
    par = (C.c_double * 6)(4, 0.5, 0.05, 10, 3, 0.5)
    c = Calibration()
    assert emu.emu_calibration(seed, C.c_int64(0), C.c_int64(500), par, C.c_int(2), C.byref(c)) == 0
    got, ref = c.to_dict(), oracle.callCalibrationPerson(10, 500)
    assert set(got) == set(ref)
    for k in ref:
        if k == "TimeAtRisk":  # sums of event times (fastmath.cuh against glibc)
            np.testing.assert_allclose(got[k], ref[k], rtol=1e-13)
        else:
            assert np.array_equal(got[k], ref[k]), k
    # sigm1 = 0: rnorm draws nothing; tau3 = 0 likewise (calibperson-r.cc:102, 152)
    par0 = (C.c_double * 6)(4, 0.0, 0.3, 10, 3, 0.0)
    assert emu.emu_calibration(seed, C.c_int64(100), C.c_int64(2000), par0, C.c_int(3), C.byref(c)) == 0
    ref = oracle.calibration_sweep(10, 100, 2000, [[4, 0.0, 0.3, 10, 3, 0.0]])[0]
    got = c.to_dict()
    assert set(got) == set(ref)
    for k in ref:
        assert np.array_equal(got[k], ref[k]), k


def test_ordered_event_queue_pops_in_heap_order_or_says_so(emu):
    """engine.cuh OrderedEvents: the heap's pop order whenever no two pending actions compare equal; where two do, the
    `tie` flag (the device then repeats the run with the binary heap)."""
    tie = C.c_int(0)
    clean = 0
    for seed in range(1, 41):
        # many distinct times: equal pairs are rare; every pop must agree with the heap unless one was flagged
        d = emu.emu_ordered_vs_heap(C.c_uint32(seed), 4000, 1 << 20, 3, 0, C.byref(tie))
        assert d >= 0, d
        if not tie.value:
            assert d == 0
            clean += 1
    assert clean >= 30
    flagged = 0
    for seed in range(1, 21):
        # 6 times x 2 priorities: equal pairs all over; the queue must flag them (return -1 = one went unnoticed)
        d = emu.emu_ordered_vs_heap(C.c_uint32(seed), 500, 6, 2, 0, C.byref(tie))
        assert d >= 0, d
        flagged += tie.value
    assert flagged == 20
    # with cancellations: the heap keeps cancelled entries and skips them at pop time, the queue removes them; the
    # sequence of live pops is the same
    clean = 0
    for seed in range(1, 41):
        d = emu.emu_ordered_vs_heap(C.c_uint32(seed), 4000, 1 << 20, 3, 5, C.byref(tie))
        assert d >= 0, d
        if not tie.value:
            assert d == 0
            clean += 1
    assert clean >= 30


def test_emulated_calibration_ordered_queue(oracle, emu):
    """The calibration model over the ordered queue: same tables as over the heap, no equal pair in thousands of lives."""
    seed = (C.c_double * 6)(*[10.0] * 6)
    tie = C.c_int(0)
    for par, first, n in (((4, 0.5, 0.05, 10, 3, 0.5), 0, 3000), ((3, 0.8, 0.6, 4, 1, 0.9), 12345, 3000)):
        a, b = Calibration(), Calibration()
        p = (C.c_double * 6)(*par)
        assert emu.emu_calibration(seed, C.c_int64(first), C.c_int64(n), p, C.c_int(2), C.byref(a)) == 0
        assert emu.emu_calibration_ordered(seed, C.c_int64(first), C.c_int64(n), p, C.c_int(2), C.byref(b), C.byref(tie)) == 0
        assert tie.value == 0
        a, b = a.to_dict(), b.to_dict()
        assert set(a) == set(b)
        for k in a:
            assert np.array_equal(a[k], b[k]), k


def test_exp_rand_without_its_doubling_loop_is_bit_exact(emu):
    """rmath.cuh exp_rand: the number of doublings from u's exponent and the sum of ln 2's from a table of repeated
    additions, against R's loop — every exponent down to 2^-70, exact powers of two, values next to them, 2e5 random
    uniforms and the smallest values the generators produce; same result bits, same number of uniforms taken."""
    emu.emu_exp_rand_selftest.restype = C.c_int64
    rng = np.random.default_rng(11)
    first = []
    for e in range(1, 71):
        p = 2.0 ** -e
        first += [p, np.nextafter(p, 1.0), np.nextafter(p, 0.0), p * 1.5, p * 1.9999999]
    first += [2.328306549295727688e-10, 2.328306437080797e-10, 1.0 - 2.3283064365386963e-10, 0.5, 0.75, 0.9999999999]
    first += list(rng.random(200000)) + list(rng.random(20000) * 2.0 ** -rng.integers(1, 40, 20000))
    first = np.array(first)
    draws = 24   # the first uniform, then up to 17 more in the second loop (q[16] = 1 ends it)
    u = rng.random((first.size, draws))
    u[:, 0] = first
    assert ((u > 0) & (u < 1)).all()
    u = np.ascontiguousarray(u)
    bad = emu.emu_exp_rand_selftest(u.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(first.size), C.c_int(draws))
    assert bad == 0


def test_count_ladder_position_is_exact(emu):
    """calib_shared.cuh counts_before: how many Count events (10, 20, .., 100) lie strictly before t — by an estimate and
    two exact corrections; checked against the ten comparisons at and one ulp either side of every rung."""
    emu.emu_counts_before.restype = C.c_int
    tie = C.c_int(0)
    rng = np.random.default_rng(5)
    ts = [0.0, 5e-324, 1e-300, 9.999999999999998, 1e3, 1e300, np.inf]
    for j in range(1, 12):
        t = float(10 * j)
        ts += [t, np.nextafter(t, 0.0), np.nextafter(t, np.inf), np.nextafter(np.nextafter(t, 0.0), 0.0)]
    ts += list(rng.uniform(0.0, 130.0, 20000)) + list(10.0 * rng.integers(1, 11, 2000) * (1.0 + rng.integers(-3, 4, 2000) * 2.0 ** -52))
    for t in ts:
        got = emu.emu_counts_before(C.c_double(t), C.byref(tie))
        assert got == sum(10.0 * j < t for j in range(1, 11)), t
        # (an infinite time counts as coinciding: the sweep is then repeated through the event queue)
        assert tie.value == int(any(10.0 * j == t for j in range(1, 11)) or t == np.inf), t


def test_emulated_calibration_with_shared_draws(oracle, emu):
    """The sweep's straight-line form (calib_shared.cuh: a person's draws once for all parameter sets, the stages from
    comparisons of the event times) gives the event loop's tables — integer cells exactly — and flags coinciding times."""
    seed = (C.c_double * 6)(*[10.0] * 6)
    sets = [(4, 0.5, 0.05, 10, 3, 0.5), (3, 0.8, 0.6, 4, 1, 0.9), (2.5, 1.2, 0.9, 0.5, 0.5, 2.0), (5, 0.3, 0.2, 30, 4, 0.1),
            (3.5, 0.7, 1.0, 0.0, 2, 0.3), (4, 0.5, 0.5, 10, -1, 0.5)]
    flat = (C.c_double * (6 * len(sets)))(*[v for s in sets for v in s])
    tie = C.c_int(0)
    for first, n in ((0, 3000), (12345, 2500)):
        got = (Calibration * len(sets))()
        assert emu.emu_calibration_shared(seed, C.c_int64(first), C.c_int64(n), len(sets), flat, C.c_int(2), got, C.byref(tie),
                                          C.c_double(0.0)) == 0
        assert tie.value == 0
        progressed = 0
        for k, par in enumerate(sets):
            ref = Calibration()
            assert emu.emu_calibration(seed, C.c_int64(first), C.c_int64(n), (C.c_double * 6)(*par), C.c_int(2), C.byref(ref)) == 0
            a, b = got[k].to_dict(), ref.to_dict()
            assert set(a) == set(b), (k, set(a), set(b))
            for key in b:
                if key == "TimeAtRisk":
                    np.testing.assert_allclose(a[key], b[key], rtol=1e-14)
                else:
                    assert np.array_equal(a[key], b[key]), (k, key)
            progressed += int("Clinical" in b)
        assert progressed >= 3  # the chain was walked to its end under several sets
    # and against the oracle (the reference's engine restated) for one of them
    ref = oracle.calibration_sweep(10, 12345, 2500, [list(sets[1])])[0]
    a = got[1].to_dict()
    for key in ref:
        if key == "TimeAtRisk":
            np.testing.assert_allclose(a[key], ref[key], rtol=1e-13)
        else:
            assert np.array_equal(a[key], ref[key]), key
    # a death on a Count rung is a pair of equal times: flagged (the device then repeats the sweep with the heap)
    assert emu.emu_calibration_shared(seed, C.c_int64(0), C.c_int64(10), 1, flat, C.c_int(1), got, C.byref(tie), C.c_double(70.0)) == 0
    assert tie.value == 1
    # sets the straight-line form must not take: rnorm with sigma == 0 draws nothing, a negative rate schedules in the past
    for bad in ((4, 0.0, 0.05, 10, 3, 0.5), (4, 0.5, 0.05, 10, 3, 0.0), (4, 0.5, 0.05, 10, 0.0, 0.5), (4, 0.5, 0.05, -1, 3, 0.5),
                (np.nan, 0.5, 0.05, 10, 3, 0.5), (4, 0.5, 0.05, np.inf, 3, 0.5)):
        assert emu.emu_calibration_shared(seed, C.c_int64(0), C.c_int64(10), 1, (C.c_double * 6)(*bad), C.c_int(1), got,
                                          C.byref(tie), C.c_double(0.0)) == -7


@pytest.mark.parametrize("n,trt,indivp,substreams,by", [(5, 0, 1, 0, 1.0), (1000, 0, 1, 0, 1.0), (1000, 1, 1, 0, 1.0),
                                                        (700, 1, 0, 0, 1.0), (800, 0, 1, 1, 1.0), (0, 0, 1, 0, 1.0),
                                                        (600, 0, 1, 0, 0.5), (600, 1, 1, 0, 2.5)])
def test_emulated_sick_sicker_summary_report(oracle, emu, n, trt, indivp, substreams, by):
    """SummaryReport in O(1) cells per interval (weighted difference arrays, carried-over utility/cost) equals the
    reference's bin-by-bin walk: Sick-Sicker (README.org:148-301), both RNG regimes, several partitions."""
    from oracle.pyoracle import SICK_SICKER_KEYS, sick_sicker_param
    par = sick_sicker_param(bool(trt))
    par["partitionBy"] = by
    p17 = (C.c_double * 17)(*[float(par[k]) for k in SICK_SICKER_KEYS])
    seed = (C.c_double * 6)(*[12345.0] * 6)
    st = (C.c_uint32 * 625)()
    emu.emu_mt_set_seed(C.c_uint32(12345), st)
    r = Report()
    assert emu.emu_sick_sicker(C.c_int64(n), p17, C.c_int(indivp), C.c_int(substreams), C.c_int64(3 if substreams else 0),
                               seed, st, C.byref(r)) == 0
    e = r.to_dict()
    if substreams:
        oracle.set_user_random_seed(12345)
    ref = oracle.sick_sicker(n, par, indivp=bool(indivp), substreams=bool(substreams), first_person=3 if substreams else 0)
    for tb, val in (("pt", "pt"), ("ut", "utility"), ("costs", "cost")):
        assert np.array_equal(e[tb]["Key"], ref[tb]["Key"]) and np.array_equal(e[tb]["age"], ref[tb]["age"]), tb
        np.testing.assert_allclose(e[tb][val], ref[tb][val], rtol=1e-11, atol=1e-9, err_msg=tb)
    for tb in ("events", "prev"):
        for k in ref[tb]:
            assert np.array_equal(e[tb][k], ref[tb][k]), (tb, k)
    assert e["indiv"]["utilities"].shape == ref["indiv"]["utilities"].shape
    np.testing.assert_allclose(e["indiv"]["utilities"], ref["indiv"]["utilities"], rtol=1e-12)
    np.testing.assert_allclose(e["indiv"]["costs"], ref["indiv"]["costs"], rtol=1e-12)
    emu.emu_report_free(C.byref(r))


@pytest.mark.parametrize("rate,start_age", [(0.0, 0.0), (0.03, 0.0), (0.05, 50.0)])
def test_emulated_cost_report(oracle, emu, rate, start_age):
    """CostReport::add (microsimulation.h:1265-1278) on the SummaryReport cost cells vs the restated class."""
    rs = np.random.RandomState(9)
    n = 3000
    state = rs.randint(0, 4, n).astype(np.int32)
    time_ = np.r_[rs.uniform(0, 110, n - 3), [0.0, 50.0, 100.0]]
    cost = rs.uniform(10, 5000, n)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    outs = []
    for lib, fn in ((emu, "emu_cost_report"), (oracle.lib, "oracle_cost_report")):
        rows, cur = np.zeros(3 * 2000), C.c_double()
        k = getattr(lib, fn)(C.c_double(rate), C.c_double(start_age), C.c_double(0.0), C.c_double(100.0), C.c_double(1.0),
                             C.c_int64(n), state.ctypes.data_as(ip), time_.ctypes.data_as(dp), cost.ctypes.data_as(dp),
                             rows.ctypes.data_as(dp), C.c_int64(2000), C.byref(cur))
        assert k > 0
        outs.append((rows[:3 * k].reshape(k, 3), cur.value))
    (a, ca), (b, cb) = outs
    assert a.shape == b.shape and np.array_equal(a[:, :2], b[:, :2])
    np.testing.assert_allclose(a[:, 2], b[:, 2], rtol=1e-13)
    assert abs(ca - cb) <= 1e-12 * abs(cb)


def test_wave_share_is_validated_without_a_device():
    """msim_set_wave_share (scheduling of the staged person-r kernel): 0 = adapt, 0.40..0.65 = fixed, anything else is an
    argument error; no launch has happened, so there is no share to report."""
    import microsimulation_b200 as m
    import microsimulation_b200.api as api
    for ok in (0.0, -1.0, 0.40, 0.5, 0.65):
        api.set_wave_share(ok)
    for bad in (0.2, 0.39, 0.66, 1.0):
        with pytest.raises(m.MsimError) as e:
            api.set_wave_share(bad)
        assert e.value.code == -3, bad  # MSIM_ERR_ARG
    api.set_wave_share(0.0)
    assert api.last_wave_share() == 0.0
