#!/usr/bin/env python
"""Hot source lines of a kernel from an `ncu --set full --import-source on` report: warp-stall samples per CUDA source
line (barrier-wait samples of idle warps listed separately). Usage: tools/ncu_hot_lines.py rep.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = None; fname = None; seen_kernel = 0
agg = collections.OrderedDict()
cur = None
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        seen_kernel += 1
        if seen_kernel > 1: break          # first launch only
        continue
    if r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": H = r; continue
    if H is None or len(r) < len(H) - 2: continue
    if r[0] != "":                           # a source line (its SASS rows follow)
        cur = (fname, int(r[0])); agg.setdefault(cur, [r[1], 0, 0, 0]); continue
    if cur is None or r[2] in ("...", ""): continue
    d = dict(zip(H[2:], r[2:]))
    try:
        n = int(d["Warp Stall Sampling (All Samples)"]); b = int(d.get("stall_barrier", 0) or 0); ex = int(d["Instructions Executed"])
    except (ValueError, KeyError):
        continue
    agg[cur][1] += n; agg[cur][2] += b; agg[cur][3] += ex
tot = sum(v[1] for v in agg.values()); totb = sum(v[2] for v in agg.values())
print(f"samples {tot}, of which barrier-wait {totb}; below: non-barrier samples per source line")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -(kv[1][1] - kv[1][2]))[:top]:
    print(f"{v[1]-v[2]:6d} {100*(v[1]-v[2])/max(tot-totb,1):5.1f}% inst {v[3]:9d} {f}:{ln}: {v[0].strip()[:100]}")
