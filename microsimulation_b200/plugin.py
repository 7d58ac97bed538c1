"""Out-of-tree models (csrc/plugin.cuh): build a model's .cu into its own shared library and drive it.

The reference's counterpart is `Rcpp::sourceCpp` of a `cProcess` subclass plus a hand-written exported driver
(inst/doc/examples.org:254-312).  `build()` is the sourceCpp step (nvcc, sm_100a), `Plugin.run()` the driver.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from ._lib import MsimError, Report, Table

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUBSTREAMS, MT_STREAM = 0, 1
OUT_REPORT, OUT_VALUES = 2, 8


class PluginCall(C.Structure):
    _fields_ = [("device", C.c_int32), ("regime", C.c_int32), ("n", C.c_int64), ("first_person", C.c_int64),
                ("seed", C.c_double * 6), ("mt_state", C.POINTER(C.c_uint32)), ("max_draws", C.c_int32),
                ("mean_draws", C.c_double), ("consts", C.c_void_p), ("consts_bytes", C.c_uint64), ("outputs", C.c_uint32),
                ("n_state", C.c_int32), ("n_event", C.c_int32), ("discount_rate", C.c_double)]


class PluginResult(C.Structure):
    _fields_ = [("values", Table), ("report", Report), ("kernel_ms", C.c_float), ("launches", C.c_int64)]


def build(source, out=None, defines=(), quiet=True):
    """nvcc a model source against plugin.cuh; returns the library path.  Rebuilds when a source is newer."""
    source = os.path.abspath(source)
    out = out or os.path.join(os.path.dirname(source), "lib" + os.path.splitext(os.path.basename(source))[0] + ".so")
    csrc = os.path.join(ROOT, "microsimulation_b200", "csrc")
    deps = [source] + [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
           "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-I" + csrc, "-I" + os.path.join(ROOT, "include")]
    cmd += ["-D" + d for d in defines] + ["-o", out, source]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stderr)
    if not quiet:
        print(r.stderr)
    return out


class Plugin:
    def __init__(self, library, name, consts_type=None):
        # the plugin's own runtime looks for the Mersenne-Twister jump table next to ITS library; point it at ours
        table = os.path.join(ROOT, "microsimulation_b200", "mt_jump.bin")
        if os.path.exists(table):
            os.environ.setdefault("MSIM_MT_JUMP_TABLE", table)
        self.lib = C.CDLL(library)
        self.name = name
        self._run = getattr(self.lib, name + "_run")
        self._free = getattr(self.lib, name + "_free")
        self._free.restype = None
        self._err = getattr(self.lib, name + "_last_error")
        self._err.restype = C.c_char_p
        self.n_values = getattr(self.lib, name + "_n_values")()
        cb = getattr(self.lib, name + "_consts_bytes")
        cb.restype = C.c_ulong
        self.consts_bytes = cb()
        self.consts_type = consts_type
        self.last_kernel_ms = 0.0
        self.launches = 0

    def run(self, n, consts, regime=SUBSTREAMS, seed=(12345,) * 6, first_person=0, mt_state=None, max_draws=250,
            mean_draws=0.0, outputs=OUT_VALUES, n_state=1, n_event=1, discount_rate=0.0, device=0):
        """Returns (values dict or None, report dict or None).  In the MT regime `mt_state` (625 uint32: position and
        words, e.g. microsimulation_b200.mt_state() after set_seed) is advanced in place."""
        call, res = PluginCall(), PluginResult()
        call.device, call.regime, call.n, call.first_person = device, regime, int(n), int(first_person)
        s = list(seed) if hasattr(seed, "__len__") else [seed] * 6
        call.seed = (C.c_double * 6)(*[float(x) for x in (s * 6)[:6]])
        st = None
        if regime == MT_STREAM:
            st = np.ascontiguousarray(mt_state, dtype=np.uint32)
            assert st.size == 625
            call.mt_state = st.ctypes.data_as(C.POINTER(C.c_uint32))
        call.max_draws, call.mean_draws = int(max_draws), float(mean_draws)
        buf = consts if isinstance(consts, (bytes, bytearray)) else bytes(consts)
        cbuf = C.create_string_buffer(buf, len(buf))
        call.consts, call.consts_bytes = C.cast(cbuf, C.c_void_p), len(buf)
        call.outputs, call.n_state, call.n_event, call.discount_rate = outputs, n_state, n_event, discount_rate
        rc = self._run(C.byref(call), C.byref(res))
        if rc != 0:
            self._free(C.byref(res))
            raise MsimError(rc, self._err().decode())
        try:
            values = res.values.to_dict() if (outputs & OUT_VALUES) and self.n_values else None
            report = {k: getattr(res.report, k).to_dict() for k in ("pt", "ut", "events", "prev")} if outputs & OUT_REPORT else None
        finally:
            self._free(C.byref(res))
        self.last_kernel_ms, self.launches = float(res.kernel_ms), int(res.launches)
        if st is not None and mt_state is not None and isinstance(mt_state, np.ndarray) and mt_state is not st:
            mt_state[:] = st
        return values, report
