"""BASELINE.json configs[2]: mnnCorrect on 2 batches x N cells x 2000 HVGs (cosine normalisation, gene-space MNN search,
Gaussian smoothing, shift-variance adjustment) through the public API, with per-stage times and SAMPLED parity checks
against the CPU oracle (a full CPU pass at N = 200k would take days).

    python tools/run_config3.py --cells 200000 [--genes 2000] [--json out.json]
"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import batchelor_b200 as bb
from batchelor_b200 import _lib, synth


def run(cells, genes=2000, k=20, sigma=0.1, check=True, nsample=8, quiet=False):
    t0 = time.perf_counter()
    A, B = synth.gene_batches(2, [cells, cells], G=genes)
    A = np.asfortranarray(A); B = np.asfortranarray(B)        # R layout: [genes x cells] column-major
    t_gen = time.perf_counter() - t0
    timings = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = bb.mnnCorrect(A, B, k=k, sigma=sigma, _timings=timings)
    torch.cuda.synchronize(); total = time.perf_counter() - t0
    path, e0, e1 = C.c_int(0), C.c_double(0), C.c_double(0)
    _lib.call("b200mnn_smooth_last_check", C.byref(path), C.byref(e0), C.byref(e1))
    out = {"workload": f"mnnCorrect 2 x {cells} cells x {genes} genes, k={k}, sigma={sigma}, cos.norm in/out, var.adj (BASELINE.json configs[2])",
           "seconds": total, "cells_per_sec": 2 * cells / total, "mnn_pairs": int(res.merge_info["pairs"][0]["left"].shape[0]),
           "mnn_cells": timings.get("mnn_cells"), "stages_s": {k_: round(v, 4) for k_, v in timings.items() if not k_.startswith("_") and k_ != "mnn_cells"},
           "smoothing_path": {1: "tensor (accepted by the in-call fp64 sample check)", 2: "fp64", 3: "tensor rejected -> fp64"}.get(path.value),
           "smoothing_sample_check": {"rows_rel": e0.value, "log_density_abs": e1.value}, "synth_seconds": round(t_gen, 2)}
    if check:
        from oracle import capi
        rng = np.random.default_rng(3)
        d1, d2, vect, r1, r2, scaling = timings["_shiftvar_io"][0]
        h1 = d1.cpu().numpy(); h2 = d2.cpu().numpy()          # the cosine-normalised inputs exactly as the device saw them
        # (1) MNN pairs of sampled batch-1 cells: exact brute-force kNN (fp64, reference summation order) of those cells in
        # batch 2 and of every neighbour they name back in batch 1
        s1 = np.sort(rng.choice(cells, size=min(nsample, cells), replace=False))
        t0 = time.perf_counter()
        w21, _ = capi.query_knn(h2, h1[s1], k)                # 1-based ids into batch 2
        back = np.unique(w21) - 1
        w12, _ = capi.query_knn(h1, h2[back], k)
        lookup = {int(b): set(w12[i].tolist()) for i, b in enumerate(back)}
        want = [(int(f) + 1, int(v)) for f, row in zip(s1, w21) for v in row if (int(f) + 1) in lookup[int(v) - 1]]
        pl, pr = res.merge_info["pairs"][0]["left"], res.merge_info["pairs"][0]["right"] - cells
        sel = np.isin(pl, s1 + 1)
        got = list(zip(pl[sel].tolist(), pr[sel].tolist()))
        out["parity_pairs"] = {"sampled_cells": int(s1.size), "pairs_checked": len(want), "equal_and_in_order": got == want,
                               "oracle_seconds": round(time.perf_counter() - t0, 1)}
        # (2) shift variance of sampled cells: the reference's per-cell loop (C restatement pinned to the reference object code)
        s2 = np.sort(rng.choice(cells, size=min(nsample, cells), replace=False))
        t0 = time.perf_counter()
        ref = capi.adjust_shift_variance_cells(h1.T, h2.T, vect[torch.from_numpy(s2).cuda()].cpu().numpy(), s2, sigma, r1.cpu().numpy(), r2.cpu().numpy())
        gotv = scaling[torch.from_numpy(s2).cuda()].cpu().numpy()
        out["parity_shift_variance"] = {"sampled_cells": int(s2.size), "bit_identical": bool(np.array_equal(ref, gotv)),
                                        "max_rel_err": float(np.max(np.abs(ref - gotv) / np.maximum(np.abs(ref), 1e-300))),
                                        "oracle_seconds": round(time.perf_counter() - t0, 1)}
        # (3) smoothing: sampled output rows through the library's fp64 difference-form path (<= 1e-10 of the reference object
        # code on the goldens) on the same MNN set: mat' = [MNN cells; sampled cells], so only the density pass is large
        averaged, uniq, rd, cor = timings["_smooth_io"][0]
        s3 = np.sort(rng.choice(cells, size=min(8, cells), replace=False))
        t0 = time.perf_counter()
        from batchelor_b200 import device as dev
        rows = torch.cat([uniq.long(), torch.from_numpy(s3).cuda()])
        os.environ["B200MNN_SMOOTH"] = "fp64"
        try:
            ref64 = dev.smooth_gaussian_kernel(averaged, torch.arange(uniq.shape[0], device=uniq.device, dtype=torch.int32), rd.index_select(0, rows).contiguous(), sigma)
        finally:
            del os.environ["B200MNN_SMOOTH"]
        refrows = ref64[uniq.shape[0]:]
        gotrows = cor.index_select(0, torch.from_numpy(s3).cuda())
        worst = float(((refrows - gotrows).abs().amax(dim=1) / refrows.abs().amax(dim=1)).max().item())
        torch.cuda.synchronize()
        out["parity_smoothing"] = {"sampled_rows": int(s3.size), "max_rel_err": worst, "checker": "fp64 difference-form path of the library on the same MNN set",
                                   "seconds": round(time.perf_counter() - t0, 1)}
    if not quiet:
        print(json.dumps(out))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=20000)
    ap.add_argument("--genes", type=int, default=2000)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    r = run(a.cells, a.genes, check=not a.no_check)
    if a.json:
        json.dump(r, open(a.json, "w"), indent=1)
