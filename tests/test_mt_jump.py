"""Jump-ahead for R's Mersenne-Twister stream (microsimulation_b200/mt_jump.py, csrc/kernels.cuh mt_jump_kernel)."""
import os

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def phi():
    from microsimulation_b200 import mt_jump
    return mt_jump.characteristic_polynomial()


def test_characteristic_polynomial_annihilates_the_word_sequence(phi):
    from microsimulation_b200 import mt_jump as J
    assert phi.bit_length() - 1 == 19937 and phi & 1
    y = J.mt_words(J.seed_state(12345), 1 + J.DEGREE + 700)[1:]
    terms = [i for i in range(J.DEGREE + 1) if (phi >> i) & 1]
    for m in (0, 1, 17, 623, 650):
        acc = 0
        for i in terms:
            acc ^= y[m + i]
        assert acc == 0, m


def test_jump_polynomials_against_the_plain_recurrence(phi):
    """x[pS + k] = XOR over the set bits i of g_p of x[1 + k + i] for the polynomials the table is made of."""
    from microsimulation_b200 import mt_jump as J
    S = 4096   # small segments so that a pure-Python stream reaches them; the arithmetic is the table's
    polys = J.jump_polynomials(S, 4, phi)
    x = J.mt_words(J.seed_state(2024), 3 * S + 700)
    for p, g in enumerate(polys, start=1):
        terms = [i for i in range(J.DEGREE) if (g >> i) & 1]
        for k in (0, 1, 311, 623):
            acc = 0
            for i in terms:
                acc ^= x[1 + k + i]
            assert acc == x[p * S + k], (p, k)


def test_table_file_matches_the_generator(phi):
    from microsimulation_b200 import build, mt_jump as J
    path = build.build_mt_table()
    seg, polys = J.read_table(path)
    assert seg == J.SEGMENT and len(polys) == J.TABLE_SEGMENTS - 1
    want = J.jump_polynomials(J.SEGMENT, 3, phi)
    assert list(polys[0]) == J.poly_words(want[0]) and list(polys[1]) == J.poly_words(want[1])
    assert all(w >> 1 == 0 for w in [p[-1] for p in polys])   # degree < 19937: only bit 19936 may be set in the last word


@pytest.mark.gpu
@pytest.mark.parametrize("position,n", [(624, 1_000_000), (1, 250_000), (300, 3_000_000),
                                        (624, 20_000_000)])  # the last: beyond the table (255 segments), serial tail
def test_parallel_stream_equals_the_serial_one(msim, oracle, position, n):
    """The words behind a state, generated as jump-started segments, are R's: against the oracle's MT19937 and against
    the one-block generator."""
    import microsimulation_b200.api as api
    msim.set_seed(12345)
    st = msim.mt_state().copy()
    st[0] = position   # R's dummy[0]: words before it are used up
    assert api.set_mt_parallel(True), "mt_jump.bin was not loaded"
    msim.set_mt_state(st)
    par = api.mt_uniforms(n)
    after_par = msim.mt_state().copy()
    api.set_mt_parallel(False)
    try:
        msim.set_mt_state(st)
        ser = api.mt_uniforms(n)
        after_ser = msim.mt_state().copy()
    finally:
        api.set_mt_parallel(True)
    assert np.array_equal(par, ser) and np.array_equal(after_par, after_ser)
    if position == 624:
        oracle.set_seed(12345)
        assert np.array_equal(par, oracle.mt_uniforms(n))


@pytest.mark.gpu
def test_models_on_the_parallel_stream(msim, oracle):
    """Config 2 at full size and the illness-death model: same tables with the stream generated either way."""
    import microsimulation_b200.api as api
    runs = {}
    for on in (True, False):
        api.set_mt_parallel(on)
        try:
            runs[on] = (msim.callSimplePerson(1_000_000), msim.callIllnessDeath(300_000, 0.1, 0.0), msim.mt_state().copy())
        finally:
            api.set_mt_parallel(True)
    for k in runs[True][0]:
        assert np.array_equal(runs[True][0][k], runs[False][0][k]), k
    for tb in ("pt", "ut", "events", "prev"):
        for k in runs[True][1][tb]:
            assert np.array_equal(runs[True][1][tb][k], runs[False][1][tb][k]) or tb in ("pt", "ut"), (tb, k)
    assert np.array_equal(runs[True][2], runs[False][2])
    want = oracle.callSimplePerson(200_000)
    head = runs[True][0]["id"] < 200_000
    for k in ("id", "state", "event"):
        assert np.array_equal(runs[True][0][k][head], want[k]), k
