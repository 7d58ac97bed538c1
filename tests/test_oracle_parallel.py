"""oracle/parallel.py (test infrastructure): the oracle on several host cores over contiguous person ranges gives the
single-process result — what bench.py's CPU legs and the full-size GPU parity test rely on."""
import numpy as np

from conftest import assert_report_close
from oracle import parallel


def test_parallel_oracle_equals_single_process(oracle):
    n, first = 30000, 1234
    one = oracle.person_run(12345, first, n, rowlog=True, report=True, discount_rate=0.03)
    par = parallel.person_run_parallel(12345, first, n, cores=3, flavour="port", rows_to_check=one["rowlog"],
                                       report=True, discount_rate=0.03)
    assert par["rows_ok"], par["rows_msg"]
    assert par["n_rows"] == len(one["rowlog"]["id"])
    assert np.array_equal(par["counts"][:8], one["counts"][:8])
    assert_report_close(par["report"], one["report"], rtol=1e-12)
    ok, worst, msg = parallel.report_equal(par["report"], one["report"])
    assert ok and worst < 1e-12, msg


def test_parallel_oracle_row_check_notices_a_difference(oracle):
    n = 5000
    rows = oracle.person_run(12345, 0, n, rowlog=True)["rowlog"]
    bad = {k: v.copy() for k, v in rows.items()}
    bad["endTime"][len(bad["endTime"]) // 2] *= 1.0 + 1e-12
    par = parallel.person_run_parallel(12345, 0, n, cores=2, flavour="port", rows_to_check=bad, report=False)
    assert not par["rows_ok"] and "event times" in par["rows_msg"]
    bad = {k: v.copy() for k, v in rows.items()}
    bad["event"][7] = 1.0 - bad["event"][7]
    par = parallel.person_run_parallel(12345, 0, n, cores=2, flavour="port", rows_to_check=bad, report=False)
    assert not par["rows_ok"]


def test_report_equal_notices_count_and_sum_differences(oracle):
    rep = oracle.person_run(12345, 0, 4000, rowlog=False, report=True, discount_rate=0.03)["report"]
    other = {tb: {k: v.copy() for k, v in t.items()} for tb, t in rep.items()}
    assert parallel.report_equal(rep, other)[0]
    other["prev"]["number"][3] += 1
    assert not parallel.report_equal(rep, other)[0]
    other["prev"]["number"][3] -= 1
    other["ut"]["utility"][5] *= 1 + 1e-8
    assert not parallel.report_equal(rep, other)[0]
