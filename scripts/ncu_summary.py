#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches.csv        > profiles/<name>_launches.txt
  python scripts/ncu_summary.py full     gpurun_out/prof.ncu-rep        > profiles/<name>_full.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    seq = [(r[ki].split("(")[0].replace("void ", ""), float(r[vi].replace(",", ""))) for r in rows[1:] if r[mi] == "gpu__time_duration.sum"]
    print(f"# {path}: {len(seq)} launches (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)")
    gi = hdr.index("Grid Size")
    grids = [int(r[gi].strip("()").split(",")[0]) for r in rows[1:] if r[mi] == "gpu__time_duration.sum"]
    starts = [i for i, (k, _) in enumerate(seq) if k.startswith("k_raygen") or k.startswith("k_vx_clear")]
    steps = [(starts[j], starts[j + 1]) for j in range(len(starts) - 1)]

    def trav_grid(a, b):
        g = [grids[i] for i in range(a, b) if seq[i][0].startswith("k_traverse2")]
        return max(g) if g else 0

    def show(a, b, title):
        print(f"# {title} (launches {a}..{b - 1}, k_traverse2 grid {trav_grid(a, b)} blocks):")
        tot = defaultdict(float)
        for k, v in seq[a:b]:
            print(f"{k:44s} {v / 1000:10.1f} us")
            tot[k] += v
        T = sum(tot.values())
        print("# shares of the step:")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            print(f"{k:44s} {v / 1000:10.1f} us {v / T * 100:6.1f} %")

    if steps:
        # bench.py runs the same step in two modes: serial (full persistent grids; the pass the per-kernel numbers and the
        # roofline come from) and pipelined (each lane's traversal grid = its share of the resident blocks; ncu serialises the
        # lanes, so those launches look slow here although they overlap in the real run)
        plain = [st for st in steps if not any("<1" in seq[i][0] for i in range(*st))] or steps
        full_grid = max(trav_grid(*st) for st in plain)
        serial = [st for st in plain if trav_grid(*st) == full_grid]
        show(*serial[-1], "one full step of the serial pass")
        lane = [st for st in plain if 0 < trav_grid(*st) < full_grid]
        if lane:
            show(*lane[-1], "one step of the pipelined pass (lane-sized grids, serialised by ncu)")
    tot = defaultdict(lambda: [0, 0.0])
    for k, v in seq:
        tot[k][0] += 1
        tot[k][1] += v
    print("# all launches:")
    for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} n={n:4d} total {v / 1000:10.1f} us  avg {v / n / 1000:9.1f} us")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: ncu --set full --clock-control none --import-source on; {len(rows) - 2} launch(es)")
    for name in ["Kernel Name"] + METRICS:
        if name in hdr:
            i = hdr.index(name)
            print(f"{name:84s} [{units[i]:12s}] " + "  ".join(r[i] for r in rows[2:]))


def raw_rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]


def traffic(trav_rep, shade_rep, out_json):
    """profiles/traffic.json, read by bench.py: DRAM bytes per launch and DRAM utilisation of the captured traversal launches
    (mean over the capture) and of the shade kernel -- regenerated from THIS round's captures, never carried over."""
    import json

    def num(r, k):
        return float(r[k].replace(",", ""))
    t = raw_rows(trav_rep)
    s_ = raw_rows(shade_rep)
    dram = [num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum") for r in t]
    unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    rows = list(csv.reader(subprocess.run(["ncu", "-i", trav_rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    u = unit[rows[1][rows[0].index("dram__bytes_read.sum")]]
    doc = {
        "kernel": t[0]["Kernel Name"],
        "dram_bytes_per_launch": sum(dram) / len(dram) * u,
        "dram_frac": sum(num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") for r in t) / len(t) / 100.0,
        "l1tex_throughput_frac": sum(num(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed") for r in t) / len(t) / 100.0,
        "issue_active_frac": sum(num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") for r in t) / len(t) / 100.0,
        "active_lanes_per_instruction": sum(num(r, "smsp__thread_inst_executed_per_inst_executed.ratio") for r in t) / len(t),
        "shade_dram_frac": sum(num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") for r in s_) / len(s_) / 100.0,
        "launches_captured": len(t),
        "source": f"{trav_rep} / {shade_rep}: ncu --set full --clock-control none (scripts/gpu_session.sh ncu), bench workload, sample 2, bounces 1-3",
    }
    json.dump(doc, open(out_json, "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
