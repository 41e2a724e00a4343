#!/usr/bin/env python
"""Summarise ncu output for profiles/: (1) a launch list CSV (--metrics gpu__time_duration.sum) -> per-kernel share of the
step; (2) a `--set full` .ncu-rep -> the handful of metrics DESIGN.md quotes (duration, DRAM traffic, issue/occupancy,
stall reasons). Usage: tools/ncu_summary.py launches <csv> | full <file.ncu-rep> [...]"""
import collections, csv, json, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__cycles_active.avg", "sm__cycles_elapsed.max"]
STALL = "smsp__average_warps_issue_stalled_"


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H = rows[h]
    d = collections.defaultdict(list)
    for r in rows[h + 1:]:
        rec = dict(zip(H, r))
        d[rec["Kernel Name"].split("(")[0].replace("void ", "")].append(float(rec["Metric Value"]) / 1e3)
    tot = sum(sum(v) for v in d.values())
    out = {k: {"launches": len(v), "mean_us": round(sum(v) / len(v), 2), "share": round(sum(v) / tot, 4)}
           for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1]))}
    print(json.dumps(out, indent=1))


def _num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return s


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(H, r))
        rec = {"kernel": d["Kernel Name"].split("(")[0]}
        for w in WANT:
            if w in d:
                rec[w] = [_num(d[w]), U[H.index(w)]]
        rec["stalls_per_issue"] = {k[len(STALL):].replace("_per_issue_active.ratio", ""): round(_num(d[k]), 3)
                                   for k in H if k.startswith(STALL) and k.endswith("_per_issue_active.ratio")
                                   and "not_issued" not in k and _num(d[k]) >= 0.05}
        out.append(rec)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        for p in sys.argv[2:]:
            full(p)
