"""ctypes binding of include/msim.h (the C ABI of libmsim_b200.so)."""
import ctypes as C
import os

import numpy as np

from . import build as _build

MSIM_OUT_ROWLOG, MSIM_OUT_REPORT, MSIM_OUT_COUNTS, MSIM_OUT_INDIV = 1, 2, 4, 8

ERRORS = {-1: "MSIM_ERR_NO_DEVICE", -2: "MSIM_ERR_CUDA", -3: "MSIM_ERR_ARG", -4: "MSIM_ERR_HEAP", -5: "MSIM_ERR_SEED",
          -6: "MSIM_ERR_STREAM", -7: "MSIM_ERR_STATE", -8: "MSIM_ERR_MODEL"}


class MsimError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "error"), code, msg))
        self.code = code


class Table(C.Structure):
    _fields_ = [("nrow", C.c_int64), ("ncol", C.c_int32), ("colnames", C.POINTER(C.c_char_p)),
                ("data", C.POINTER(C.c_double))]

    def to_dict(self, copy=True):
        # one view of the whole column-major table, then a row of it per column (a ctypes array per column costs more
        # than the copy itself for the report's small tables)
        n, k = int(self.nrow), int(self.ncol)
        names = [self.colnames[j].decode() for j in range(k)]
        if not n:
            return {name: np.zeros(0) for name in names}
        whole = np.ctypeslib.as_array((C.c_double * (n * k)).from_address(C.addressof(self.data.contents))).reshape(k, n)
        if copy:
            whole = whole.copy()
        return {name: whole[j] for j, name in enumerate(names)}


class Report(C.Structure):
    _fields_ = [(k, Table) for k in ("pt", "ut", "events", "prev", "costs", "indiv")]

    def to_dict(self):
        return {k: getattr(self, k).to_dict() for k in ("pt", "ut", "events", "prev", "costs", "indiv")}


class Calibration(C.Structure):
    _fields_ = [("occupancy", C.c_double * 40), ("present", C.c_int32 * 4), ("time_at_risk", C.c_double * 4),
                ("n_time_at_risk", C.c_int32)]

    def to_dict(self):
        names = ["DiseaseFree", "Precursor", "PreClinical", "Clinical"]
        out = {"TimeAtRisk": np.array(self.time_at_risk[:self.n_time_at_risk])}
        for s, nm in enumerate(names):
            if self.present[s]:
                out[nm] = np.array(self.occupancy[s * 10:(s + 1) * 10])
        return out


class DenseDims(C.Structure):
    _fields_ = [("n_state", C.c_int32), ("n_event", C.c_int32), ("n_bin", C.c_int32), ("n_doubles", C.c_int64)]


class ReportOptions(C.Structure):
    _fields_ = [("start", C.c_double), ("finish", C.c_double), ("delta", C.c_double), ("max_time", C.c_double),
                ("start_report_age", C.c_double), ("output_utilities", C.c_int32)]


class PersonShard(C.Structure):
    _fields_ = [("seed", C.c_double * 6), ("first_person", C.c_int64), ("n", C.c_int64), ("outputs", C.c_uint32),
                ("discount_rate", C.c_double)]


class PersonDeviceResult(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("capacity", C.c_int64), ("d_rowlog", C.c_void_p), ("d_dense", C.c_void_p),
                ("d_counts", C.c_void_p), ("dims", DenseDims)]


# every symbol include/msim.h declares (tests check that the library exports them all)
SYMBOLS = [
    "msim_table_free", "msim_report_free", "msim_last_error", "msim_version", "msim_init", "msim_shutdown",
    "msim_launch_count", "msim_set_stream", "msim_peer_export", "msim_peer_import", "msim_peer_merge", "msim_peer_flush", "msim_peer_result_ptr", "msim_peer_close", "msim_set_person_schedule", "msim_set_mt_parallel", "msim_set_rowlog_plan", "msim_set_wave_share", "msim_last_wave_share", "msim_set_report_options", "msim_person_fetch_indiv", "msim_r_create_current_stream", "msim_r_remove_current_stream",
    "msim_r_set_user_random_seed", "msim_r_get_user_random_seed", "msim_r_next_rng_substream",
    "msim_r_rng_advance_substream", "msim_next_rng_stream", "msim_user_unif_rand", "msim_mt_set_state", "msim_mt_get_state", "msim_mt_set_seed",
    "msim_callPersonSimulation", "msim_callSimplePerson", "msim_callSimplePerson2", "msim_callIllnessDeath",
    "msim_callCalibrationSimulation", "msim_callSim_SickSicker", "msim_callSim_CancerManagement", "msim_cost_report", "msim_discounted", "msim_rpexp", "msim_rpexp_gamma", "msim_interpolate", "msim_table_lookup", "msim_gsm_randU", "msim_gsm_randU0", "msim_gsm_eta", "msim_person_run_device", "msim_person_enqueue", "msim_sync",
    "msim_last_kernel_ms", "msim_timer_start", "msim_timer_stop", "msim_person_fetch_rowlog", "msim_person_fetch_counts", "msim_dense_to_report",
    "msim_dense_device_ptr", "msim_dense_fetch", "msim_calibration_sweep", "msim_calibration_sweep_device", "msim_calibration_fetch", "msim_calibration_unpack",
    "msim_rng_uniforms", "msim_mt_uniforms", "msim_microbench",
]

_lib = None


def lib_path():
    return _build.LIB


def load():
    """Load libmsim_b200.so (building it if sources changed).  Fails loudly if it cannot."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MSIM_B200_LIB") or _build.build()  # the override is for kernel-variant experiments
    L = C.CDLL(path)
    L.msim_last_error.restype = C.c_char_p
    L.msim_version.restype = C.c_char_p
    L.msim_launch_count.restype = C.c_int64
    L.msim_r_create_current_stream.restype = None
    L.msim_user_unif_rand.restype = C.POINTER(C.c_double)
    L.msim_r_remove_current_stream.restype = None
    L.msim_table_free.restype = None
    L.msim_report_free.restype = None
    L.msim_shutdown.restype = None
    L.msim_callIllnessDeath.argtypes = [C.c_int32, C.c_double, C.c_double, C.POINTER(Report)]
    L.msim_discounted.argtypes = [C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                  C.c_double, C.POINTER(C.c_double)]
    dp = C.POINTER(C.c_double)
    L.msim_rpexp.argtypes = [C.c_int, C.c_int32, dp, dp, C.c_int64, dp, dp, dp]
    L.msim_rpexp_gamma.argtypes = [C.c_int, C.c_int32, C.c_double, dp, dp, C.c_int64, dp, dp, dp, dp]
    L.msim_interpolate.argtypes = [C.c_int, C.c_int32, dp, dp, C.c_int64, dp, dp]
    L.msim_table_lookup.argtypes = [C.c_int32, C.c_int64, dp, C.c_int64, dp, dp]
    L.msim_callSim_CancerManagement.argtypes = [C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int64, C.POINTER(Report)]
    L.msim_cost_report.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int64,
                                   C.POINTER(C.c_int32), dp, dp, C.POINTER(Table), dp]
    L.msim_set_stream.argtypes = [C.c_void_p, C.c_int]
    L.msim_set_rowlog_plan.argtypes = [C.c_double, C.c_double]
    L.msim_set_wave_share.argtypes = [C.c_double]
    L.msim_last_wave_share.argtypes = [C.POINTER(C.c_double)]
    L.msim_dense_to_report.argtypes = [C.POINTER(C.c_double), C.POINTER(DenseDims), C.c_double, C.POINTER(Report)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise MsimError(rc, load().msim_last_error().decode())
