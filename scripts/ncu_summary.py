"""Summarise ncu output for profiles/: (1) a launch-list CSV (gpu__time_duration per launch) -> per-kernel totals and
shares; (2) a --set full .ncu-rep -> the metrics the roofline discussion uses.  Runs where ncu is installed (no GPU
needed to READ a report).  Usage:
  python scripts/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
  python scripts/ncu_summary.py report  gpurun_out/x.ncu-rep   > profiles/rNN_x_ncu.txt
  python scripts/ncu_summary.py traffic gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/rNN_traffic.json
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_tensor.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
    "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0][:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print("# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised:"
          " compare SHARES, not absolutes)")
    print("%6s %12s %7s  %s" % ("count", "total_us", "share", "kernel"))
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%6d %12.1f %6.1f%%  %s" % (c, t / 1e3, 100 * t / tot, k))
    print("total_us %.1f  launches %d" % (tot / 1e3, sum(a[0] for a in agg.values())))


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("=== " + r[idx["Kernel Name"]].split("(")[0])
        for m in METRICS:
            if m in idx:
                print("  %-75s %s %s" % (m, r[idx[m]], rows[1][idx[m]]))
        print()


def traffic(paths):
    """Per kernel (first launch of each name): DRAM bytes, warp instructions, issue-active share, top stall reasons.
    bench.py reads this file for roofline.traffic and roofline.issue_slot."""
    import json
    import re

    kernels = {}
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        idx = {h: i for i, h in enumerate(rows[0])}
        units = rows[1]

        def val(r, m, to=None):
            v = float(r[idx[m]].replace(",", ""))
            u = units[idx[m]]
            if to == "byte":
                v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            if to == "us":
                v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
            return v

        for r in rows[2:]:
            name = re.sub(r"^(void\s+)?(\(anonymous namespace\)::)?", "", r[idx["Kernel Name"]].split("(")[0]).strip()
            name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
            if name in kernels:
                continue
            rd, wr = val(r, "dram__bytes_read.sum", "byte"), val(r, "dram__bytes_write.sum", "byte")
            stalls = sorted(((val(r, m), m.split("stalled_")[1].split("_per_issue")[0]) for m in idx
                             if m.startswith("smsp__average_warps_issue_stalled_") and m.endswith("_per_issue_active.ratio")
                             and "selected" not in m), reverse=True)
            kernels[name] = {
                "dram_bytes": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                "time_us": round(val(r, "gpu__time_duration.sum", "us"), 1),
                "issue_active_pct": round(val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
                "sm_cycles_active_over_elapsed": round(val(r, "sm__cycles_active.avg") / val(r, "sm__cycles_elapsed.max"), 2),
                "warp_inst": int(val(r, "smsp__inst_executed.sum")),
                "registers": int(val(r, "launch__registers_per_thread")),
                "top_stalls": [s for _, s in stalls[:3]],
            }
    print(json.dumps({"source": "ncu --set full --clock-control none captures: " + ", ".join(paths) +
                      " (bench scene: 300k Gaussians, 1024x667, one view); written by scripts/ncu_summary.py traffic",
                      "kernels": kernels}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2:])
    else:
        {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
