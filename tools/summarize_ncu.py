"""Summaries of ncu outputs for profiles/ (tracked): launch list -> per-kernel totals; .ncu-rep -> key raw metrics."""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", "")) / 1e6
        except ValueError:
            continue
        a = agg.setdefault(r[ki][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# source: {path} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare shares)")
    print(f"{'kernel':92s} {'launches':>8s} {'total ms':>10s} {'avg ms':>8s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:92s} {v[0]:8d} {v[1]:10.2f} {v[1] / v[0]:8.3f} {v[1] / tot:6.3f}")


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# source: {path} (ncu --set full --clock-control none --import-source on)")
    ni = hdr.index("Kernel Name")
    print("kernels:", [r[ni][:60] for r in rows[2:]])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:78s} {units[i]:16s} {[r[i] for r in rows[2:]]}")
    for i, h in enumerate(hdr):  # warps stalled per issue-active cycle, by reason (the full set reports ratios, not percentages)
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                v = [float((r[i] or "0").replace(",", "")) for r in rows[2:]]
            except ValueError:
                continue
            if max(v) >= 0.1 and "selected_per" not in h.replace("not_selected", ""):
                print(f"stalled warps per issue: {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {v}")
    for i, h in enumerate(hdr):
        if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct"):
            v = [float(r[i] or 0) for r in rows[2:]]
            if max(v) > 3:
                print(f"stall {h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''):40s} {v}")


if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else rep)(sys.argv[2])
