"""BASELINE.json configs[3] and [4] through the public API (one C-ABI call, all visible GPUs driven from this process):
  --config 4: fastMNN post-PCA path (reducedMNN) on 8 batches x 500k cells x 50 PCs, hierarchical merge.order
  --config 5: reducedMNN on 16 batches x 625k cells x 50 PCs, k = 30, progressive order
    python tools/run_config45.py --config 4 --devices 8 [--scale 1.0] [--json out.json]
Prints one JSON line: seconds (warmed, mean of --reps), cells/s, pair counts per merge, SHA-1 of the corrected matrix and of
the pair lists (equal across --devices values = the multi-GPU result is the single-GPU result), and a sampled parity check
of the first merge's MNN pairs against the CPU KMKNN port."""
import argparse, hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4, choices=[4, 5])
ap.add_argument("--devices", type=int, default=0, help="0 = all visible")
ap.add_argument("--scale", type=float, default=1.0, help="fraction of the configured cells per batch")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--no-check", action="store_true")
ap.add_argument("--json", default=None)
a = ap.parse_args()
if a.devices > 0:
    os.environ["B200MNN_DEVICES"] = str(a.devices)
import torch
import batchelor_b200 as bb
from batchelor_b200 import synth

if a.config == 4:
    nb, per, k = 8, int(500_000 * a.scale), 20
    order = [[[1, 2], [3, 4]], [[5, 6], [7, 8]]]
    name = f"reducedMNN (fastMNN post-PCA path) 8 batches x {per} cells x 50 PCs, k=20, hierarchical merge.order (BASELINE.json configs[3])"
else:
    nb, per, k = 16, int(625_000 * a.scale), 30
    order = None
    name = f"reducedMNN 16 batches x {per} cells x 50 PCs, k=30, progressive merge order (BASELINE.json configs[4])"
t0 = time.perf_counter()
batches = synth.pc_batches(nb, per, d=50)
t_gen = time.perf_counter() - t0
ndev = torch.cuda.device_count() if a.devices <= 0 else min(a.devices, torch.cuda.device_count())
bb.reducedMNN(*[b[: max(2000, per // 50)] for b in batches], k=k, merge_order=order)      # warm-up: modules, pools on every device
times = []
for _ in range(a.reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = bb.reducedMNN(*batches, k=k, merge_order=order)
    torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
dt = float(np.mean(times))
h = hashlib.sha1(np.ascontiguousarray(res.corrected).tobytes()).hexdigest()
hp = hashlib.sha1(b"".join(np.ascontiguousarray(p["left"]).tobytes() + np.ascontiguousarray(p["right"]).tobytes() for p in res.merge_info["pairs"])).hexdigest()
out = {"workload": name, "devices": ndev, "seconds": dt, "seconds_each": [round(t, 3) for t in times], "cells_per_sec": nb * per / dt,
       "merges": nb - 1, "mnn_pairs_per_merge": [int(p["left"].shape[0]) for p in res.merge_info["pairs"]],
       "batch_size": [round(float(x), 4) for x in res.merge_info["batch_size"]], "corrected_sha1": h, "pairs_sha1": hp,
       "api": "batchelor_b200.reducedMNN -> b200mnn_reduced_mnn (host matrices in, corrected host matrix + pairs out)", "synth_seconds": round(t_gen, 1)}
if not a.no_check:
    from oracle import capi
    # first merge = two leaf batches, no orthogonalisation yet: its pairs are findMutualNN(left, right)
    li, ri = res.merge_info["left"][0][0], res.merge_info["right"][0][0]
    L, R = batches[li - 1], batches[ri - 1]
    rng = np.random.default_rng(1)
    s = np.sort(rng.choice(L.shape[0], size=min(2000, L.shape[0]), replace=False))
    t0 = time.perf_counter()
    iR, iL = capi.Kmknn(R), capi.Kmknn(L)
    w21, _ = iR.query(np.ascontiguousarray(L[s]), k)
    back = np.unique(w21) - 1
    w12, _ = iL.query(np.ascontiguousarray(R[back]), k)
    look = {int(b): set(w12[i].tolist()) for i, b in enumerate(back)}
    want = [(int(f) + 1, int(v)) for f, row in zip(s, w21) for v in row if (int(f) + 1) in look[int(v) - 1]]
    origin = res.batch
    off_l = int(np.nonzero(origin == li)[0][0]); off_r = int(np.nonzero(origin == ri)[0][0])
    pl, pr = res.merge_info["pairs"][0]["left"] - off_l, res.merge_info["pairs"][0]["right"] - off_r
    sel = np.isin(pl, s + 1)
    got = list(zip(pl[sel].tolist(), pr[sel].tolist()))
    out["parity_first_merge_pairs"] = {"sampled_cells": int(s.size), "pairs_checked": len(want), "equal_and_in_order": got == want,
                                       "oracle_seconds": round(time.perf_counter() - t0, 1)}
print(json.dumps(out))
if a.json:
    json.dump(out, open(a.json, "w"), indent=1)
