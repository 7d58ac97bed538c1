"""microsimulation_b200 — B200-native engine for the microsimulation hot path.

Independent per-individual discrete-event lifetimes (the reference's
ssim::Sim / cProcess / RngStream / EventReport path) simulated by hand-written
sm_100a kernels behind a C ABI (include/msim.h).  This package is the host-side
mirror of the reference's R wrappers over that ABI.
"""
from .api import *  # noqa: F401,F403
from ._lib import MsimError, load, lib_path, SYMBOLS  # noqa: F401
