"""Build the CUDA extension in-tree: microsimulation_b200/libmsim_b200.so.

One translation unit (csrc/capi.cu, which includes the kernels) compiled by
nvcc for sm_100a only.  -fmad=false: no multiply-add contraction in any
expression that produces an event time (qnorm's rational polynomials, the
samplers), so device arithmetic is evaluated in the order the reference's CPU
code evaluates it.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmsim_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def sources():
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    out.append(os.path.join(HERE, "..", "include", "msim.h"))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    """Compile if sources are newer than the library.  Returns the .so path."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    if nvcc is None:
        if os.path.exists(LIB):
            return LIB  # GPU box without a toolkit on PATH: use the library that travelled with the snapshot
        raise RuntimeError("nvcc not found and %s is not built" % LIB)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "capi.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
