"""Turns ncu outputs pulled from the GPU box into the small CSV summaries kept under profiles/.
  summarise_ncu.py launches <launches.csv> <out.csv>      per-kernel totals of a --metrics gpu__time_duration.sum launch list
  summarise_ncu.py full <capture.ncu-rep> <out.csv>       key counters of every launch of an `ncu --set full` capture"""
import collections, csv, subprocess, sys

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.sum")


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] == "us" else v if r[ui] == "ms" else v * 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("kernel,launches,total_ms,share\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{t:.4f},{t / tot:.4f}\n")
        f.write(f"\"TOTAL (cold-cache, serialised: compare shares)\",{sum(a[0] for a in agg.values())},{tot:.4f},1.0\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [i for i, h in enumerate(hdr) if h in KEYS or any(h.startswith(k) for k in KEYS)]
    ni = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        for r in data:
            f.write(f"# {r[ni][:150]}\n")
            for i in cols:
                f.write(f"{hdr[i]},{r[i]},{units[i]}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
