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
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline query sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fastmnn", action="store_true", help="skip the secondary fastMNN cells/s measurement")
    return ap.parse_args()


def workload_config(args, extra=None):
    cfg = {
        "workload": f"findMutualNN {args.cells} vs {args.cells} cells x {args.dims} PCs, k1=k2={args.k}, exact search "
                    f"(BASELINE.json configs[1])",
        "cells_per_batch": args.cells, "dims": args.dims, "k": args.k,
        "l2_policy": "inputs larger than L2 (2 x %.0f MB fp64 + %.0f MB fp16 operands vs 126 MB L2)" % (
            args.cells * args.dims * 8 / 1e6, 2 * args.cells * 256 / 1e6),
    }
    if extra:
        cfg.update(extra)
    return cfg


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
    from batchelor_b200 import synth

    b1, b2 = synth.pc_batches(2, args.cells, d=args.dims)
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
def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu summary (None if absent)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_ncu_cand_ts_final_summary.csv")
    try:
        tot = 0.0
        for line in open(path):
            f = line.strip().split(",")
            if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[1]) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[f[2]]
        return tot or None
    except OSError:
        return None


def run_b200(args):
    import torch
    import torch.distributed as dist

    import batchelor_b200 as bb
    from batchelor_b200 import _lib, device as dev, synth

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

    b1, b2 = synth.pc_batches(2, args.cells, d=args.dims)          # replicated on every rank (same seeds)
    n1, n2 = b1.shape[0], b2.shape[0]
    d1, d2 = torch.from_numpy(b1).to(device), torch.from_numpy(b2).to(device)
    torch.cuda.synchronize()

    def step():
        first, second, _, _ = dev.find_mutual_nn(d1, d2, args.k, args.k, sharded=True)
        return first, second

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        first, second = step()
    npairs = int(first.shape[0])

    sampler = ClockSampler(local_rank)
    launches0 = dev.launches()
    _lib.call("b200mnn_profile_enable", 1)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
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

    # ---- end to end: host buffers (pinned), H2D + compute + D2H inside the timed region ----
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
            x1 = h1.to(device, non_blocking=True); x2 = h2.to(device, non_blocking=True)
            f, s, _, _ = dev.find_mutual_nn(x1, x2, args.k, args.k, sharded=True)
            fh, sh = f.cpu(), s.cpu()
            d2h = int(fh.numel() * 4 + sh.numel() * 4) * world
        barrier()
        if it >= E2E_WARMUP:
            e2e_times.append(time.perf_counter() - t0)
    e2e_t = torch.tensor([float(np.mean(e2e_times))], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = (n1 + n2) / float(e2e_t.item())
    h2d = int((b1.nbytes + b2.nbytes) * world)

    # ---- secondary figure of the BASELINE metric: fastMNN cells/s = post-PCA path (reducedMNN semantics: MNN search,
    # correction averaging, centring along the batch vector, tricube search + smoothing) on the same two batches ----
    fast = None
    if world == 1 and not args.no_fastmnn:
        from batchelor_b200 import api
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = api.reducedMNN(b1, b2, k=args.k)      # host matrices in, corrected host matrix out
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        fast = {"value": (n1 + n2) / dt, "unit": "cells/s", "seconds": dt, "api": "batchelor_b200.reducedMNN (host in/out, one merge)",
                "mnn_pairs": int(res.merge_info["pairs"][0]["left"].shape[0]), "batch_size": float(res.merge_info["batch_size"][0])}

    if rank == 0:
        peaks = measured_peaks()
        # per-GPU figure: algorithmic flops of one rank's launches over the slowest rank's summed kernel time (TFLOP/s)
        achieved = (kernel_flops / world) / max(kernel_ms_max * 1e-3, 1e-12) / 1e12
        executed = (kernel_executed / world) / max(kernel_ms_max * 1e-3, 1e-12) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "fp16 tensor-core scoring (one-term tier, fp16x3 re-score of uncertified queries; fp32 accumulate) + fp64 exact re-rank", "data": "synthetic",
            "config": workload_config(args, {"parallelism": f"query rows sharded over {world} GPU(s), reference batch replicated, "
                                                            f"NCCL all-gather of per-shard top-k" if world > 1 else "single GPU",
                                             "mnn_pairs": npairs}),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * float(e2e_t.item()), "ms_each": [round(1e3 * t, 2) for t in e2e_times],
                    "api": "batchelor_b200.findMutualNN -> b200mnn_find_mutual_nn (host buffers)" if world == 1 else
                           "pinned host -> device copies + device.find_mutual_nn (sharded) + pair lists back to host"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_tflops"], "traffic": ncu_traffic_bytes(),
                         "traffic_note": "DRAM bytes (read + write) of one first-tier launch at 1M x 1M from the committed ncu --set full "
                                         "capture (profiles/r1_ncu_cand_ts_final_summary.csv); the 28.9 GB of operand tiles it reads come from L2",
                         "kernel": "knn_candidates_kernel (tcgen05 fp16x3 scoring + top-k filter)",
                         "kernel_ms_per_launch": kernel_ms_max / max(kernel_launches / world, 1),
                         "kernel_share_of_step": kernel_ms_max / total_ms,
                         "algorithmic_flops_per_launch": 2.0 * (n1 / world) * n2 * args.dims,
                         "executed_tflops": executed, "frac_executed": executed / peaks["bf16_tflops"],
                         "executed_over_algorithmic": kernel_executed / max(kernel_flops, 1.0),
                         "note": "frac = ALGORITHMIC (brute-force) 2*nq*n*d flops of the search / summed time of the candidate-scoring "
                                 "launches / peak, per GPU.  The search is exact but cluster-pruned (KMKNN-style, csrc/knn_cluster.cu): "
                                 "only the 128x128 score tiles whose lower bound cannot exclude them are computed, so the algorithmic "
                                 "rate can exceed the tensor peak; executed_tflops / frac_executed count the tcgen05.mma work actually "
                                 "issued (one-term fp16 tier + three-term re-score of the uncertified queries).  The scored tiles all "
                                 "belong to the queries' own mixture component, where the kernel is bound by the top-k insertion "
                                 "path of its epilogue, not by the tensor pipe (profiles/).",
                         "peak_source": peaks["source"]},
        }
        if fast is not None:
            line["fastmnn_cells_per_sec"] = fast
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
