"""Every C-ABI entry point called with null pointers, one process per call (a crash must not take the test run down).
   python tests/fuzz_null_args.py            # table of exit codes
   python tests/fuzz_null_args.py <symbol>   # one call
Used by tests/test_host_logic.py (no device) and tests/test_gpu_parity.py (device present)."""
import ctypes as C, sys, subprocess, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
calls = {
 "msim_mt_set_state": "(None,)", "msim_mt_get_state": "(None,)", "msim_r_set_user_random_seed": "(None,)",
 "msim_r_get_user_random_seed": "(None,)", "msim_r_rng_advance_substream": "(None, None)", "msim_next_rng_stream": "(None,)",
 "msim_callPersonSimulation": "(None, C.c_int32(5), None)", "msim_callSimplePerson": "(C.c_int32(5), None)",
 "msim_callSimplePerson2": "(C.c_int32(5), None)", "msim_callIllnessDeath": "(C.c_int32(5), C.c_double(.1), C.c_double(0), None)",
 "msim_callCalibrationSimulation": "(C.c_int32(5), None, None)", "msim_callSim_SickSicker": "(C.c_int32(5), None, 1, 0, C.c_int64(0), None)",
 "msim_callSim_CancerManagement": "(C.c_int32(5), C.c_double(8), C.c_double(0.03), C.c_double(50), 1, C.c_int64(0), None)",
 "msim_cost_report": "(C.c_double(0.03), C.c_double(0), C.c_double(0), C.c_double(100), C.c_double(1), C.c_int32(8), C.c_int64(5), None, None, None, None, None)",
 "msim_calibration_sweep_device": "(None, C.c_int64(0), C.c_int64(5), C.c_int32(1), None, C.c_int32(0), C.c_int32(1), None, None)",
 "msim_calibration_fetch": "(None, C.c_int64(48))", "msim_calibration_unpack": "(None, C.c_int32(1), None)",
 "msim_calibration_sweep": "(None, C.c_int64(0), C.c_int64(5), C.c_int32(1), None, None)",
 "msim_discounted": "(0, C.c_int64(5), None, None, None, C.c_double(0.03), None)",
 "msim_rpexp": "(0, C.c_int32(3), None, None, C.c_int64(5), None, None, None)",
 "msim_rpexp_gamma": "(0, C.c_int32(3), C.c_double(1.0), None, None, C.c_int64(5), None, None, None, None)",
 "msim_interpolate": "(0, C.c_int32(3), None, None, C.c_int64(5), None, None)",
 "msim_table_lookup": "(C.c_int32(2), C.c_int64(3), None, C.c_int64(5), None, None)",
 "msim_gsm_randU": "(None, C.c_int64(5), None, None, None, C.c_double(10), None)",
 "msim_gsm_randU0": "(None, C.c_int64(5), None, None, C.c_double(10), None)", "msim_gsm_eta": "(None, C.c_int64(5), None, None, None)",
 "msim_person_enqueue": "(None, None)", "msim_person_run_device": "(None, None)", "msim_person_fetch_rowlog": "(None,)",
 "msim_person_fetch_counts": "(None,)", "msim_dense_to_report": "(None, None, C.c_double(0), None)",
 "msim_dense_device_ptr": "(None, None)", "msim_dense_fetch": "(None, C.c_int64(5))",
 "msim_rng_uniforms": "(None, C.c_int64(0), C.c_int64(1), C.c_int32(1), None)", "msim_mt_uniforms": "(C.c_int64(5), None)",
 # one null at a time: the other arguments are valid, so an earlier check cannot hide a missing one
 "msim_calibration_sweep#seed": "(None, C.c_int64(0), C.c_int64(5), C.c_int32(1), buf, buf)",
 "msim_rng_uniforms#seed": "(None, C.c_int64(0), C.c_int64(1), C.c_int32(1), buf)",
 "msim_callPersonSimulation#seed": "(None, C.c_int32(5), buf)",
 "msim_callPersonSimulation#out": "(ibuf, C.c_int32(5), None)",
 "msim_last_kernel_ms": "(None,)", "msim_timer_stop": "(None,)", "msim_dense_device_ptr#n": "(buf, None)",
 "msim_dense_device_ptr#p": "(None, buf)", "msim_dense_fetch#n": "(None, C.c_int64(9))",
 "msim_r_rng_advance_substream#n": "(buf, None)", "msim_r_rng_advance_substream#seed": "(None, ibuf)",
 "msim_set_report_options": "(None,)", "msim_person_fetch_indiv": "(None, C.c_int64(5), None)",
 "msim_microbench": "(2, C.c_int64(256), 1, None)", "msim_last_wave_share": "(None,)",
 "msim_table_free": "(None,)", "msim_report_free": "(None,)", "msim_set_stream": "(None, 0)",
}
if __name__ == "__main__" and len(sys.argv) > 1:
    import microsimulation_b200 as m
    L = m.load()
    L.msim_r_create_current_stream()
    if os.environ.get("MSIM_FUZZ_INIT"):
        m.init(0)
    f = getattr(L, sys.argv[1].split("#")[0]); f.argtypes = None
    buf, ibuf = (C.c_double * 512)(), (C.c_int32 * 16)(*([12345] * 16))
    rc = eval("f" + calls[sys.argv[1]])
    print("rc", rc)
    sys.exit(0)
def run_all():
    out = {}
    for name in calls:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True)
        out[name] = (r.returncode, r.stdout.strip()[-20:])
    return out


if __name__ == "__main__" and len(sys.argv) == 1:
    for name, (code, tail) in run_all().items():
        print("%-34s exit %4d  %s" % (name, code, tail))
