"""Generates tests/golden/*.json|npz.

Two kinds of fixtures:
  * doc_kats.json  — values PRINTED by the reference itself (README.org,
    inst/doc/examples.org, src/rngstream-example.cpp, test/test_microsimulation.R);
    typed in by hand below with their file:line.  These pin the oracle.
  * oracle_*.npz   — outputs of the oracle's `reference` flavour (oracle/_ref:
    the reference's own ssim.cc + RngStream.cpp + heap.h compiled where they lie
    under /root/reference, plus the restated Rcpp layer).  They travel to the GPU
    box, where /root/reference does not exist.
Run here (container with /root/reference):  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.pyoracle import Oracle, sick_sicker_param  # noqa: E402

DOC_KATS = {
    # src/rngstream-example.cpp:9-13
    "rngstream_first_uniforms": {"stream1": 0.127011, "stream2": 0.759582, "stream1_substream1": 0.079399},
    # inst/doc/examples.org:162-180
    "tick_tock": [[0.441808, 1.250275, 1.268724, 1.724134, 1.748072, 5.274728],
                  [1.257961, 3.075964, 3.302451, 3.749214, 5.001675, 5.398762, 5.486878, 7.209042, 9.621881]],
    # inst/doc/examples.org:341-360
    "doc_illness_death_times": [13.254234, 14.361154, 1.436288, 75.477631, 28.865643, 35.660268, 59.485505, 2.643472,
                                115.923108],
    "doc_illness_death_qaly_cost": [[11.117066, 11702.175], [1.335915, 1406.227], [28.686987, 30196.829],
                                    [26.600625, 28000.658], [31.094833, 32731.403]],
    # README.org:95-107  summary(callSimplePerson2(1e5)$events)
    "simple2_events_1e5": {"rows": 338, "by_event": [142, 100, 96], "max": 1722, "mean": 445.1, "median": 325.0,
                           "q1": 74.0, "q3": 721.8, "age_mean": 56.98},
    # README.org:114-124  summary(callIllnessDeath(1e5, cure=0.2)$prev)
    "illness_prev_1e5_cure02": {"rows": 200, "max": 100000, "mean": 34914, "median": 8096, "q1": 2030, "q3": 79029},
    # test/test_microsimulation.R:1602-1612
    "calibration_seed10_n500_DiseaseFree_first2": [422, 354],
    "illness_death_n10_LE": 64.96217,
    # README.org:398-434 (first run, Trt = FALSE): id, state, kind, previous, now
    "sick_sicker_trace": [
        [0, 0, 0, 0.000000, 1.554306], [0, 1, 1, 1.554306, 1.660357], [0, 2, 3, 1.660357, 2.137922],
        [1, 0, 0, 0.000000, 22.523279], [1, 1, 2, 22.523279, 23.538790], [1, 0, 0, 23.538790, 25.110528],
        [1, 1, 2, 25.110528, 25.529621], [1, 0, 4, 25.529621, 31.000000], [2, 0, 0, 0.000000, 1.279403],
        [2, 1, 2, 1.279403, 2.482209], [2, 0, 0, 2.482209, 4.467754], [2, 1, 2, 4.467754, 5.240185],
        [2, 0, 0, 5.240185, 6.123519], [2, 1, 2, 6.123519, 7.951863], [2, 0, 0, 7.951863, 10.967305],
        [2, 1, 2, 10.967305, 11.127479], [2, 0, 0, 11.127479, 15.295999], [2, 1, 1, 15.295999, 16.663526],
        [2, 2, 4, 16.663526, 31.000000], [3, 0, 0, 0.000000, 0.021088], [3, 1, 2, 0.021088, 1.648953],
        [3, 0, 0, 1.648953, 1.665451], [3, 1, 1, 1.665451, 2.825812], [3, 2, 3, 2.825812, 3.378253],
        [4, 0, 0, 0.000000, 1.548766], [4, 1, 2, 1.548766, 4.303079], [4, 0, 0, 4.303079, 5.128089],
        [4, 1, 2, 5.128089, 5.418233], [4, 0, 0, 5.418233, 9.734630], [4, 1, 2, 9.734630, 11.013894],
        [4, 0, 0, 11.013894, 26.388802], [4, 1, 2, 26.388802, 26.514306], [4, 0, 0, 26.514306, 27.803678],
        [4, 1, 2, 27.803678, 28.437418], [4, 0, 4, 28.437418, 31.000000]],
    # README.org:333-337 (n = 1e3): mean cost, mean QALY per arm
    "sick_sicker_n1000": {"no_treatment": [117835.7, 13.101], "treatment": [223871.9, 13.658]},
    # inst/doc/examples.org:535-546 via callSimplePerson(5): end times per row
    "simple_person_5": [45.90, 73.92, 65.27, 91.44, 61.42, 80.92, 98.92],
}


def main():
    with open(os.path.join(HERE, "doc_kats.json"), "w") as f:
        json.dump(DOC_KATS, f, indent=1)
    o = Oracle("ref")
    assert o.flavour == "reference"
    out = {}
    out["mrg_uniforms_seed12345_sub0_64x8"] = o.rng_uniforms(12345, 0, 64, 8)
    out["mrg_uniforms_seed12345_sub99999990_8x4"] = o.rng_uniforms(12345, 99999990, 8, 4)
    out["advance_substream_12345_1000"] = np.array(o.rng_advance_substream(12345, 1000))
    o.set_seed(12345)
    out["mt_uniforms_12345_2000"] = o.mt_uniforms(2000)
    r = o.person_run(12345, 0, 2000, rowlog=True, report=True, discount_rate=0.03)
    for k, v in r["rowlog"].items():
        out["person_n2000_" + k] = v
    out["person_n2000_counts"] = r["counts"]
    for tb in ("pt", "ut", "events", "prev"):
        for k, v in r["report"][tb].items():
            out["person_n2000_report_%s_%s" % (tb, k)] = v
    r = o.person_run(12345, 0, 10 ** 6, rowlog=False)
    out["person_n1e6_counts"] = r["counts"]
    r = o.callSimplePerson(3000)
    for k, v in r.items():
        out["simple_n3000_" + k] = v
    for name, rep in (("simple2_n20000", o.callSimplePerson2(20000)), ("illness_n20000_cure01", o.callIllnessDeath(20000, 0.1, 0.0)),
                      ("illness_n20000_cure02_zsd05", o.callIllnessDeath(20000, 0.2, 0.5))):
        for tb in ("pt", "ut", "events", "prev"):
            for k, v in rep[tb].items():
                out["%s_%s_%s" % (name, tb, k)] = v
    c = o.callCalibrationPerson(12345, 20000)
    for k, v in c.items():
        out["calib_seed12345_n20000_" + k] = v
    np.savez_compressed(os.path.join(HERE, "oracle_ref.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
