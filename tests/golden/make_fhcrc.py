"""Writes tests/golden/fhcrc_tables.npz: the ten tables of the reference's data/fhcrcData.rda as numpy columns, read
with the package's own .rda reader (microsimulation_b200/rdata.py — no R here).  The GPU box has no /root/reference, so
the tests that look every row up on the device use this fixture.   python tests/golden/make_fhcrc.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from microsimulation_b200 import fhcrc_data  # noqa: E402

tables = fhcrc_data.load("/root/reference/data/fhcrcData.rda")
flat = {}
for name, tab in tables.items():
    for col, v in tab.items():
        flat["%s/%s" % (name, col)] = np.asarray(v, dtype=np.float64)
out = os.path.join(ROOT, "tests", "golden", "fhcrc_tables.npz")
np.savez_compressed(out, **flat)
print(out, os.path.getsize(out), "bytes,", len(flat), "columns")
