"""Generates tests/golden/*.npz by running the REFERENCE'S OWN kernels (compiled unmodified from /root/reference/src
into oracle/_ref/libbatchelor_ref.so, see oracle/Makefile) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; it only ever reads the committed .npz files.  Shapes follow the reference's own
tests (tests/testthat/test-mnn-correct.R:28-174: 400/1000 cells x 25 genes, sd 0.1, sigma in {0.1, 0.5, 1}).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import capi, host_oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    assert capi.have_ref(), "needs /root/reference to build oracle/_ref"
    rng = np.random.default_rng(10003)
    # ---- smooth_gaussian_kernel (via .compute_correction_vectors' inputs) ----
    data1 = rng.normal(scale=0.1, size=(400, 25))
    data2 = rng.normal(scale=0.1, size=(1000, 25))
    cases = {
        "vanilla": (np.arange(1, 11), np.arange(30, 20, -1), 0.1),
        "repeated": (np.r_[11, 12, 13, np.arange(1, 11)], np.r_[30, 30, 30, np.arange(30, 20, -1)], 0.1),
        "many": (np.arange(1, 201), np.arange(500, 300, -1), 0.1),
        "wide": (np.arange(1, 11), np.arange(30, 20, -1), 0.5),
    }
    out = {"data1": data1, "data2": data2}
    for name, (m1, m2, s2) in cases.items():
        averaged, second = host_oracle.average_correction(data1, m1, data2, m2)
        res = capi.ref_smooth_gaussian_kernel(averaged.T, second - 1, data2.T, s2)
        out[f"{name}_mnn1"] = m1.astype(np.int32)
        out[f"{name}_mnn2"] = m2.astype(np.int32)
        out[f"{name}_sigma"] = np.float64(s2)
        out[f"{name}_averaged"] = np.asfortranarray(averaged.T)
        out[f"{name}_index0"] = (second - 1).astype(np.int32)
        out[f"{name}_out"] = res
    np.savez_compressed(os.path.join(HERE, "smooth_gaussian_kernel.npz"), **out)

    # ---- adjust_shift_variance ----
    rng = np.random.default_rng(100032)
    d1 = rng.normal(scale=0.1, size=(25, 400))
    d2 = rng.normal(scale=0.1, size=(25, 1000))
    cv = rng.uniform(size=(1000, 25))
    out = {"data1": d1, "data2": d2, "vect": cv}
    for s in (1.0, 0.1):
        out[f"out_sigma_{s}"] = capi.ref_adjust_shift_variance(d1, d2, cv, s, np.arange(400), np.arange(1000))
    r1 = rng.choice(400, size=150, replace=False).astype(np.int32)
    r2 = rng.choice(1000, size=300, replace=False).astype(np.int32)
    out["r1"] = r1
    out["r2"] = r2
    out["out_restricted"] = capi.ref_adjust_shift_variance(d1, d2, cv, 1.0, r1, r2)
    np.savez_compressed(os.path.join(HERE, "adjust_shift_variance.npz"), **out)

    # ---- find_mutual_nns: neighbour matrices from the fp64 oracle search, pairs from the reference kernel ----
    rng = np.random.default_rng(1200004)
    X1 = rng.normal(size=(700, 10)).astype(np.float32).astype(np.float64)
    X2 = (rng.normal(size=(900, 10)) + 0.5).astype(np.float32).astype(np.float64)
    w21, dist21 = capi.query_knn(X2, X1, 15)
    w12, _ = capi.query_knn(X1, X2, 10)
    first, second = capi.ref_find_mutual_nns(w21, w12)
    np.savez_compressed(os.path.join(HERE, "find_mutual_nns.npz"), X1=X1, X2=X2, w21=w21, w12=w12, dist21=dist21, first=first, second=second)
    print("golden fixtures written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
