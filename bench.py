#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native MNN hot path.

Workload (BASELINE.json configs[1]): findMutualNN of 1M vs 1M cells x 50 PCs, k1 = k2 = 20, exact search, synthetic
Gaussian-mixture batches (batchelor_b200/synth.py).  Metric: MNN kNN queries/s = (n1 + n2) / time of one findMutualNN
(both exact searches + mutual-pair extraction).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (KMKNN port, all host threads)

One JSON line on stdout (rank 0).  `value` = device-resident throughput (inputs in HBM when the timed region starts, CUDA
events, max over ranks); `e2e` = the same metric through the host-buffer C ABI with pinned host inputs, H2D/D2H inside the
timed region; `roofline` = algorithmic flops of the dominant kernel / its CUDA-event time vs the measured bf16 peak;
`cpu_baseline` = the CPU port timed on a bounded sample on this box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mnn_knn_queries_per_sec"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=1_000_000, help="cells per batch (BASELINE config: 1M)")
    ap.add_argument("--dims", type=int, default=50)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--workload", default="mixture", choices=["mixture", "single_blob"],
                    help="mixture: 32-component Gaussian mixture (SURVEY section 8d); single_blob: one component (no cluster structure)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline query sample")
    ap.add_argument("--parity-queries", type=int, default=50_000, help="queries per direction checked against the CPU KMKNN port")
    ap.add_argument("--config3-cells", type=int, default=20_000, help="cells per batch of the mnnCorrect (config 3) leg; 200000 = full size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the secondary legs (dense scan, single blob, fastMNN, config 3, kernels)")
    ap.add_argument("--no-fastmnn", action="store_true", help="skip the secondary fastMNN cells/s measurement")
    return ap.parse_args()


def workload_config(args):
    """Identical for both arms (`--impl b200` and `--impl reference`)."""
    return {
        "workload": f"findMutualNN {args.cells} vs {args.cells} cells x {args.dims} PCs, k1=k2={args.k}, exact search "
                    f"(BASELINE.json configs[1])",
        "cells_per_batch": args.cells, "dims": args.dims, "k": args.k, "data_regime": args.workload,
        "l2_policy": "inputs larger than L2 (2 x %.0f MB fp64 + %.0f MB fp16 operands vs 126 MB L2)" % (
            args.cells * args.dims * 8 / 1e6, 2 * args.cells * 256 / 1e6),
    }


def make_batches(args, workload=None):
    from batchelor_b200 import synth
    w = workload or args.workload
    return synth.pc_batches(2, args.cells, d=args.dims, ncomp=32 if w == "mixture" else 1)


# ------------------------------------------------------------------------------------------------------------------
# clocks during the timed region (profiling recipe's nvidia-smi line)
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # the sampler only runs inside the timed region, so every sample is "under load"
        return {"sm_mhz": float(np.median(sm)), "sm_mhz_min": float(min(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(power)), "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p.get("bf16_tflops_sustained", p.get("bf16_tflops")), "hbm_gbs": p.get("hbm_gbs"),
                "source": "MEASURED_PEAKS.json (bf16_tflops_sustained: the kernel runs for >100 ms inside the step)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback of /opt/skills/guides/B200_PROFILING.md (sustained)"}


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline: KMKNN port on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------------------------
def cpu_knn_sample(b1, b2, k, target_seconds, steps=1, warmup=0):
    """Builds the KMKNN index on batch 2, then times `steps` query samples from batch 1.  Returns a dict with the
    per-step sample times and the extrapolated findMutualNN throughput (both searches incl. index builds)."""
    from oracle import capi

    threads = os.cpu_count() or 1
    n1, n2 = b1.shape[0], b2.shape[0]
    t0 = time.perf_counter()
    index = capi.Kmknn(b2, nthreads=threads)
    t_build = time.perf_counter() - t0
    pilot = min(256, n1)
    t0 = time.perf_counter()
    index.query(b1[:pilot], k, nthreads=threads, want_dist=False)
    per_q = (time.perf_counter() - t0) / pilot
    sample = int(max(256, min(n1, target_seconds / max(per_q, 1e-9))))
    times = []
    rng = np.random.default_rng(99)
    for s in range(warmup + steps):
        rows = rng.choice(n1, size=sample, replace=False)
        q = np.ascontiguousarray(b1[rows])
        t0 = time.perf_counter()
        index.query(q, k, nthreads=threads, want_dist=False)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    t_q = float(np.mean(times)) / sample
    full = 2 * t_build + (n1 + n2) * t_q          # two index builds + every cell queried once
    return {"value": (n1 + n2) / full, "threads": threads, "sample": sample, "t_build": t_build, "t_per_query": t_q,
            "step_times": times, "full_estimate_s": full}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    b1, b2 = make_batches(args)
    per_step = max(3.0, min(30.0, 150.0 / max(1, args.steps + args.warmup)))
    r = cpu_knn_sample(b1, b2, args.k, per_step, steps=args.steps, warmup=args.warmup)
    sample_desc = (f"KMKNN port (oracle/kmknn_port.cpp; BiocNeighbors itself is not in this image): index built once on the full "
                   f"{args.cells}-cell batch ({r['t_build']:.1f} s), each step = {r['sample']} random queries of the other batch on "
                   f"{r['threads']} threads; value = (n1+n2) / (2*build + (n1+n2)*time_per_query)")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(r["step_times"])), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": sample_desc},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
# this repo's CUDA path
# ------------------------------------------------------------------------------------------------------------------
NCU_SUMMARY = "profiles/r2_ncu_cand_ts_summary.csv"
NCU_SUMMARY_FALLBACK = "profiles/r1_ncu_cand_ts_final_summary.csv"


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full summary
    (a number taken under a profiler cannot be measured inside the timed run); returns (bytes or None, file)."""
    for rel in (NCU_SUMMARY, NCU_SUMMARY_FALLBACK):
        path = os.path.join(ROOT, rel)
        try:
            tot = 0.0
            for line in open(path):
                f = line.strip().split(",")
                if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(f[1]) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[f[2]]
            if tot:
                return tot, rel
        except OSError:
            continue
    return None, None


def parity_gate(args, b1, b2, d1, d2, w21, w12, first, second, dev):
    """Outside the timed region: the result the timed steps produced, checked at the benchmarked size.
    (i) neighbour ids AND distances of `--parity-queries` sampled queries per direction against the CPU KMKNN port (exact
    search, fp64, ties by index); (ii) the GPU's pair list against the reference's pair-extraction algorithm run on the CPU
    over the GPU's full index matrices (order included)."""
    import torch
    from oracle import capi

    t0 = time.perf_counter()
    threads = os.cpu_count() or 1
    rng = np.random.default_rng(2024)
    out = {"sampled_queries": 0, "knn_mismatch_slots": 0, "distance_mismatch_slots": 0}
    for X, Q, dX, dQ, w in ((b2, b1, d2, d1, w21), (b1, b2, d1, d2, w12)):
        ns = int(min(args.parity_queries, Q.shape[0]))
        rows = np.sort(rng.choice(Q.shape[0], size=ns, replace=False))
        index = capi.Kmknn(X, nthreads=threads)
        want_idx, want_dist = index.query(np.ascontiguousarray(Q[rows]), args.k, nthreads=threads, want_dist=True)
        del index
        rt = torch.from_numpy(rows).to(dQ.device)
        got_timed = w.index_select(0, rt).cpu().numpy() + 1                     # what the timed steps computed (ids only)
        gi, gd = dev.query_knn(dX, dQ.index_select(0, rt), args.k, want_dist=True)   # same kernels, distances requested
        out["sampled_queries"] += ns
        out["knn_mismatch_slots"] += int((got_timed != want_idx).sum()) + int((gi.cpu().numpy() + 1 != want_idx).sum())
        out["distance_mismatch_slots"] += int((gd.cpu().numpy() != want_dist).sum())
    f, s2 = capi.find_mutual_nns(np.asfortranarray(w21.cpu().numpy() + 1), np.asfortranarray(w12.cpu().numpy() + 1))
    out["pairs"] = int(f.size)
    out["pairs_equal"] = bool(np.array_equal(f, first.cpu().numpy() + 1) and np.array_equal(s2, second.cpu().numpy() + 1))
    out["checker"] = "CPU KMKNN port (oracle/kmknn_port.cpp) on %d threads; pair extraction: oracle/mnn_oracle.c" % threads
    out["seconds"] = round(time.perf_counter() - t0, 1)
    out["green"] = out["knn_mismatch_slots"] == 0 and out["distance_mismatch_slots"] == 0 and out["pairs_equal"]
    return out


def timed_steps(step, steps, barrier):
    import torch
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    barrier()
    return ev0.elapsed_time(ev1)


def secondary_legs(args, b1, b2, peaks):
    """Single-GPU legs next to the headline: the data-independent regimes of the search, the fastMNN half of the BASELINE
    metric, config 3 (mnnCorrect) and its two gene-space kernels with their own rooflines."""
    import torch

    import batchelor_b200 as bb
    from batchelor_b200 import _lib, api, device as dev

    legs = {}
    cuda = torch.device("cuda", torch.cuda.current_device())
    n1, n2 = b1.shape[0], b2.shape[0]
    sync = torch.cuda.synchronize

    def knn_leg(x1, x2, steps=3):
        step = lambda: dev.find_mutual_nn(x1, x2, args.k, args.k, sharded=False)
        step(); step()
        _lib.call("b200mnn_profile_enable", 1)
        ms = timed_steps(step, steps, sync)
        kms, kl, kf, kex = C.c_double(0), C.c_int64(0), C.c_double(0), C.c_double(0)
        _lib.call("b200mnn_profile_collect", C.byref(kms), C.byref(kl), C.byref(kf))
        _lib.call("b200mnn_profile_collect_executed", C.byref(kex))
        _lib.call("b200mnn_profile_enable", 0)
        return {"value": steps * (n1 + n2) / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
                "algorithmic_tflops": kf.value / max(kms.value * 1e-3, 1e-12) / 1e12, "executed_tflops": kex.value / max(kms.value * 1e-3, 1e-12) / 1e12,
                "frac": kf.value / max(kms.value * 1e-3, 1e-12) / 1e12 / peaks["bf16_tflops"],
                "frac_executed": kex.value / max(kms.value * 1e-3, 1e-12) / 1e12 / peaks["bf16_tflops"],
                "executed_over_algorithmic": kex.value / max(kf.value, 1.0), "kernel_share_of_step": kms.value / ms}

    # (1) dense scan of the same data: cluster pruning off -> every 128 x 128 score tile is computed
    d1, d2 = torch.from_numpy(b1).to(cuda), torch.from_numpy(b2).to(cuda)
    os.environ["B200MNN_PRUNE"] = "0"
    try:
        legs["roofline_dense"] = dict(knn_leg(d1, d2), note="B200MNN_PRUNE=0: same workload, no cluster pruning; executed == algorithmic x padding (d=50 -> K=64)")
    finally:
        del os.environ["B200MNN_PRUNE"]
    # (2) data without cluster structure: one Gaussian blob (the pruning bounds exclude little)
    if args.workload == "mixture":
        s1, s2 = make_batches(args, "single_blob")
        e1, e2 = torch.from_numpy(s1).to(cuda), torch.from_numpy(s2).to(cuda)
        legs["single_blob"] = dict(knn_leg(e1, e2), note="same shape, one mixture component (batch 2 shifted): default settings")
        del e1, e2, s1, s2

    # (3) fastMNN cells/s (post-PCA path = reducedMNN semantics), warmed, host matrices in and out
    if not args.no_fastmnn:
        api.reducedMNN(b1, b2, k=args.k)
        times = []
        for _ in range(3):
            sync(); t0 = time.perf_counter()
            res = api.reducedMNN(b1, b2, k=args.k)
            sync(); times.append(time.perf_counter() - t0)
        dt = float(np.mean(times))
        legs["fastmnn_cells_per_sec"] = {"value": (n1 + n2) / dt, "unit": "cells/s", "seconds": dt, "seconds_each": [round(t, 4) for t in times],
                                         "api": "batchelor_b200.reducedMNN (host in/out, one merge: MNN search, averaging, centring, tricube search + smoothing)",
                                         "mnn_pairs": int(res.merge_info["pairs"][0]["left"].shape[0]), "batch_size": float(res.merge_info["batch_size"][0])}
        if not args.no_cpu_baseline:
            from oracle import capi, host_oracle as ho
            ns = 20000
            c1, c2 = np.ascontiguousarray(b1[:ns]), np.ascontiguousarray(b2[:ns])
            def knn(X, Q, k):
                return capi.Kmknn(X).query(Q, k)
            t0 = time.perf_counter()
            ho.reduced_mnn([c1, c2], k=args.k, knn=knn, mutual=lambda a, b, k1, k2: capi.find_mutual_nns(knn(b, a, k2)[0], knn(a, b, k1)[0]))
            legs["fastmnn_cells_per_sec"]["cpu_port"] = {"value": 2 * ns / (time.perf_counter() - t0), "unit": "cells/s", "cores": os.cpu_count(),
                                                         "sample": f"numpy restatement of the merge (oracle/host_oracle.py) + KMKNN port on 2 x {ns} cells"}
    del d1, d2
    torch.cuda.empty_cache()

    # (4) config 3: mnnCorrect through the public API with sampled parity
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import run_config3
    if args.config3_cells > 0:
        run_config3.run(min(args.config3_cells, 4000), check=False, quiet=True)     # warm-up (memory pools, module load)
        legs["config3_mnn_correct"] = run_config3.run(args.config3_cells, quiet=True)

    # (5) the two gene-space kernels alone, device resident, with their own rooflines
    from batchelor_b200 import synth
    G, nc = 2000, 20000
    A, B = synth.gene_batches(2, [nc, nc], G=G)
    g1 = dev.cosine_norm(torch.from_numpy(np.ascontiguousarray(A.T)).to(cuda))[0]
    g2 = dev.cosine_norm(torch.from_numpy(np.ascontiguousarray(B.T)).to(cuda))[0]
    gen = torch.Generator(device="cuda").manual_seed(1)
    vect = torch.randn((nc, G), dtype=torch.float64, device=cuda, generator=gen) * 0.01
    r = torch.arange(nc, device=cuda, dtype=torch.int32)
    dev.adjust_shift_variance(g1, g2, vect, 0.1, r, r)
    sync(); t0 = time.perf_counter()
    sv = dev.adjust_shift_variance(g1, g2, vect, 0.1, r, r)
    sync(); dt = time.perf_counter() - t0
    pairs = float(nc) * (2 * nc)
    leg = {"workload": f"adjust_shift_variance {nc} x {nc} cells x {G} genes, sigma=0.1 (device resident)", "seconds": dt,
           "cells_per_sec": nc / dt, "pairs_per_sec": pairs / dt,
           "roofline": {"bound": "fp64 pipe", "achieved": 2 * 2 * pairs * G / dt / 1e12, "peak": 40.0, "unit": "TFLOP/s (fp64)",
                        "frac": 2 * 2 * pairs * G / dt / 1e12 / 40.0,
                        "note": "algorithmic work = two fp64 FMAs per (cell, comparison cell, gene): projection and Gram entry (SURVEY A7); "
                                "peak = B200 nominal fp64 rate (no measured fp64 figure in MEASURED_PEAKS.json)"}}
    if not args.no_cpu_baseline:
        from oracle import capi
        ncpu = 24
        cells = np.arange(0, nc, nc // ncpu)[:ncpu].astype(np.int64)
        h1, h2 = g1.cpu().numpy(), g2.cpu().numpy()
        t0 = time.perf_counter()
        ref = capi.adjust_shift_variance_cells(h1.T, h2.T, vect[torch.from_numpy(cells).cuda()].cpu().numpy(), cells, 0.1, np.arange(nc), np.arange(nc))
        tc = time.perf_counter() - t0
        leg["cpu_baseline"] = {"value": ncpu / tc, "unit": "cells/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"{ncpu} cells of the same problem through the reference's per-cell loop (oracle/mnn_oracle.c, OpenMP over cells)"}
        leg["parity"] = {"sampled_cells": ncpu, "bit_identical": bool(np.array_equal(ref, sv[torch.from_numpy(cells).cuda()].cpu().numpy()))}
    legs["shift_variance_kernel"] = leg

    nm = 8000
    avg = torch.randn((nm, G), dtype=torch.float64, device=cuda, generator=gen) * 0.01
    idx = torch.sort(torch.randperm(nc, device=cuda, generator=gen)[:nm])[0].to(torch.int32)
    from batchelor_b200.csrc_hooks import gemm_profile
    dev.smooth_gaussian_kernel(avg, idx, g2, 0.1)
    gemm_profile(True)
    sync(); t0 = time.perf_counter()
    dev.smooth_gaussian_kernel(avg, idx, g2, 0.1)
    sync(); dt = time.perf_counter() - t0
    gms, gl, gf = gemm_profile(False)
    path, e0, e1 = C.c_int(0), C.c_double(0), C.c_double(0)
    _lib.call("b200mnn_smooth_last_check", C.byref(path), C.byref(e0), C.byref(e1))
    alg = 2.0 * nm * nc * (2 * G) + 2.0 * nm * nm * G
    legs["smoothing_kernel"] = {
        "workload": f"smooth_gaussian_kernel {nc} cells x {nm} MNN cells x {G} genes, sigma=0.1 (device resident)", "seconds": dt,
        "cells_per_sec": nc / dt, "path": path.value, "sample_check": {"rows_rel": e0.value, "log_density_abs": e1.value},
        "roofline": {"bound": "tensor", "achieved": gf / max(gms * 1e-3, 1e-12) / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": gf / max(gms * 1e-3, 1e-12) / 1e12 / peaks["bf16_tflops"], "gemm_ms": gms, "gemm_launches": gl,
                     "algorithmic_tflops_whole_call": alg / dt / 1e12,
                     "note": "achieved = executed fp16 tensor flops of the split-fp16 GEMM launches (3 terms, padded tiles) / their CUDA-event time; "
                             "the whole call also holds the operand split, soft-max and the fp64 sample check"}}
    return legs


def run_b200(args):
    import torch
    import torch.distributed as dist

    import batchelor_b200 as bb
    from batchelor_b200 import _lib, device as dev

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: batchelor_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    bb.load()

    b1, b2 = make_batches(args)                                     # same seeds on every rank
    n1, n2 = b1.shape[0], b2.shape[0]
    d1, d2 = torch.from_numpy(b1).to(device), torch.from_numpy(b2).to(device)
    torch.cuda.synchronize()

    def step():
        return dev.find_mutual_nn(d1, d2, args.k, args.k, sharded=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        first, second, w21, w12 = step()
    npairs = int(first.shape[0])

    sampler = ClockSampler(local_rank)
    launches0 = dev.launches()
    _lib.call("b200mnn_profile_enable", 1)
    if rank == 0:
        sampler.start()
    total_ms_local = timed_steps(step, args.steps, barrier)
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([total_ms_local], dtype=torch.float64, device=device)
    kms, kl, kf = C.c_double(0), C.c_int64(0), C.c_double(0)
    _lib.call("b200mnn_profile_collect", C.byref(kms), C.byref(kl), C.byref(kf))
    kex = C.c_double(0)
    _lib.call("b200mnn_profile_collect_executed", C.byref(kex))
    _lib.call("b200mnn_profile_enable", 0)
    launches = dev.launches() - launches0
    kstat = torch.tensor([kms.value, float(kl.value), kf.value, kex.value], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        kmax = kstat.clone(); dist.all_reduce(kmax, op=dist.ReduceOp.MAX)
        ksum = kstat.clone(); dist.all_reduce(ksum, op=dist.ReduceOp.SUM)
        kernel_ms_max, kernel_launches, kernel_flops, kernel_executed = float(kmax[0]), float(ksum[1]), float(ksum[2]), float(ksum[3])
    else:
        kernel_ms_max, kernel_launches, kernel_flops, kernel_executed = kms.value, float(kl.value), kf.value, kex.value
    total_ms = float(ms.item())
    value = args.steps * (n1 + n2) / (total_ms / 1e3)

    # ---- parity gate at the benchmarked size (outside the timed region) ----
    parity = None
    if not args.no_parity:
        if world > 1:
            # every rank must hold the same full result, and it must be the single-GPU result
            f1, s1_, a1, c1_ = dev.find_mutual_nn(d1, d2, args.k, args.k, sharded=False)
            same = torch.tensor([int(torch.equal(f1, first) and torch.equal(s1_, second) and torch.equal(a1, w21) and torch.equal(c1_, w12))],
                                device=device)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            sharded_equal = bool(same.item())
        if rank == 0:
            parity = parity_gate(args, b1, b2, d1, d2, w21, w12, first, second, dev)
            if world > 1:
                parity["sharded_equals_single_gpu_on_every_rank"] = sharded_equal
                parity["green"] = parity["green"] and sharded_equal
        barrier()

    # ---- end to end: pinned host buffers, H2D + compute + D2H inside the timed region ----
    h1 = torch.from_numpy(b1).pin_memory(); h2 = torch.from_numpy(b2).pin_memory()
    e2e_times = []
    d2h = 0
    E2E_WARMUP = 2   # the first calls of this path populate the stream-ordered memory pools of its own streams
    for it in range(E2E_WARMUP + args.e2e_steps):
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            res = bb.findMutualNN(h1.numpy(), h2.numpy(), k1=args.k, k2=args.k)   # the public, reference-facing call
            d2h = int(res["first"].nbytes + res["second"].nbytes)
        else:
            # every rank uploads only its block of rows of both batches; NCCL all-gathers replicate them over NVLink
            x1 = dev.upload_sharded(h1, device); x2 = dev.upload_sharded(h2, device)
            f, s, _, _ = dev.find_mutual_nn(x1, x2, args.k, args.k, sharded=True)
            if rank == 0:                       # the caller's process receives the pair lists
                fh, sh = f.cpu(), s.cpu()
                d2h = int(fh.numel() * 4 + sh.numel() * 4)
        barrier()
        if it >= E2E_WARMUP:
            e2e_times.append(time.perf_counter() - t0)
    e2e_t = torch.tensor([float(np.mean(e2e_times))], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = (n1 + n2) / float(e2e_t.item())
    h2d = int(b1.nbytes + b2.nbytes)          # whole job: each row crosses PCIe once (sharded upload at N > 1)
    del h1, h2

    legs = {}
    if world == 1 and not args.no_legs:
        del d1, d2
        torch.cuda.empty_cache()
        legs = secondary_legs(args, b1, b2, measured_peaks())

    if rank == 0:
        peaks = measured_peaks()
        # per-GPU figure: algorithmic flops of one rank's launches over the slowest rank's summed kernel time (TFLOP/s)
        achieved = (kernel_flops / world) / max(kernel_ms_max * 1e-3, 1e-12) / 1e12
        executed = (kernel_executed / world) / max(kernel_ms_max * 1e-3, 1e-12) / 1e12
        traffic, traffic_file = ncu_traffic_bytes()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "fp16 tensor-core scoring (one-term tier, fp16x3 re-score of uncertified queries; fp32 accumulate) + fp64 exact re-rank", "data": "synthetic",
            "config": workload_config(args),
            "parallelism": (f"{world} GPUs: the two searches split over two groups of ranks, query rows sharded inside a group, reference "
                            "batch replicated, one NCCL all-gather of the per-shard neighbour indices (device.direction_split)"
                            if world > 1 else "single GPU"),
            "mnn_pairs": npairs,
            "parity": parity,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * float(e2e_t.item()), "ms_each": [round(1e3 * t, 2) for t in e2e_times],
                    "api": "batchelor_b200.findMutualNN -> b200mnn_find_mutual_nn (host buffers)" if world == 1 else
                           "device.upload_sharded (each rank uploads its rows, NCCL all-gather) + device.find_mutual_nn (sharded) + pair lists to rank 0's host"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_tflops"], "traffic": traffic,
                         "traffic_note": f"DRAM bytes (read + write) of one first-tier launch at 1M x 1M from the committed ncu --set full capture ({traffic_file}); "
                                         "a profiler figure, so stamped from the file, not measured in this run; captured from the working tree "
                                         "committed as 916a58c (the epilogue of this kernel instance has not changed since; later commits only added the 64-candidate variant)",
                         "kernel": "knn_candidates_ts_kernel<1,1,1> (tcgen05 TS mode: query operand in TMEM, one-term fp16 tier, per-thread register-list epilogue; its three-term launches on the uncertified queries are included)",
                         "kernel_ms_per_launch": kernel_ms_max / max(kernel_launches / world, 1),
                         "kernel_share_of_step": kernel_ms_max / total_ms,
                         "algorithmic_flops_per_launch": 2.0 * (n1 * (2 if world > 1 else 1) / world) * n2 * args.dims,
                         "executed_tflops": executed, "frac_executed": executed / peaks["bf16_tflops"],
                         "executed_over_algorithmic": kernel_executed / max(kernel_flops, 1.0),
                         "note": "frac = ALGORITHMIC (brute-force) 2*nq*n*d flops of the search / summed time of the candidate-scoring "
                                 "launches / peak, per GPU.  The search is exact but cluster-pruned (KMKNN-style, csrc/knn_cluster.cu): "
                                 "only the 128x128 score tiles whose lower bound cannot exclude them are computed, so the algorithmic "
                                 "rate can exceed the tensor peak; frac_executed is the HARDWARE fraction (tcgen05.mma work actually "
                                 "issued); roofline_dense / single_blob below are the data-independent regimes.",
                         "peak_source": peaks["source"]},
        }
        line.update(legs)
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_knn_sample(b1, b2, args.k, args.cpu_seconds)
            line["cpu_baseline"] = {
                "value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                "sample": (f"KMKNN port (oracle/kmknn_port.cpp): index built on the full {n2}-cell batch in {r['t_build']:.1f} s, "
                           f"{r['sample']} random queries timed on {r['threads']} threads ({r['t_per_query'] * 1e3:.3f} ms/query); "
                           f"value = (n1+n2) / (2*build + (n1+n2)*time_per_query) = {r['full_estimate_s']:.0f} s per findMutualNN")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
