#!/usr/bin/env python3
"""Summarise an ncu report per CUDA source line: instructions, avg active threads, stall samples.
usage: tools/ncu_lines.py report.ncu-rep [kernel-regex] [top N]"""
import csv, subprocess, sys, io, collections
import re
rep = sys.argv[1]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
kre = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
cur_fn = ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None; agg = collections.OrderedDict(); tot_i = tot_t = tot_s = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name": cur_fn = r[1]; continue
    if r and r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; idx_src2 = [i for i, h in enumerate(r) if h == "Source"]; continue
    if hdr is None or not r or r[0] in ("", "0") or len(r) < len(hdr): continue
    if kre and not kre.search(cur_fn): continue
    try:
        ln = int(r[0]); inst = float(r[hdr["Instructions Executed"]]); thr = float(r[hdr["Thread Instructions Executed"]]); smp = float(r[hdr["# Samples"]])
    except ValueError:
        continue
    key = (cur_file, ln, r[1].strip()[:90])
    a = agg.setdefault(key, [0, 0, 0]); a[0] += inst; a[1] += thr; a[2] += smp
    tot_i += inst; tot_t += thr; tot_s += smp
print(f"total warp-inst {tot_i:.3g}  thread-inst {tot_t:.3g}  avg active {tot_t/max(tot_i,1):.1f}  samples {tot_s:.0f}")
for (f, ln, src), (i, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{100*s/max(tot_s,1):5.1f}% smp {100*i/max(tot_i,1):5.1f}% inst  act {t/max(i,1):5.1f}  {f}:{ln:<4d} {src}")

# coarse regions of the trace kernel (file, first line, last line, label)
REGIONS = [("vpt_math.cuh", 60, 110, "philox + rng.next"), ("vpt_math.cuh", 1, 59, "float3 / pinned math helpers"),
           ("vpt_walk.cuh", 28, 49, "aabb intersect / contains"), ("vpt_walk.cuh", 50, 111, "sphere / closest_object"),
           ("vpt_walk.cuh", 112, 165, "octree locate / skip"), ("vpt_walk.cuh", 166, 230, "volume coord + tex lookups"),
           ("vpt_walk.cuh", 231, 300, "hg / sun dir"), ("vpt_trace.cuh", 75, 160, "ray record load/store"),
           ("vpt_trace.cuh", 161, 210, "walk_step body"), ("vpt_trace.cuh", 211, 232, "begin_ratio_walk"),
           ("vpt_trace.cuh", 233, 395, "advance (integrator glue)"), ("vpt_trace.cuh", 396, 430, "write_sample / kernel prologue"),
           ("vpt_trace.cuh", 431, 475, "refill from queue"), ("vpt_trace.cuh", 476, 510, "vote + service round"),
           ("vpt_trace.cuh", 511, 545, "stepping loop control"), ("vpt_trace.cuh", 546, 600, "epilogue / counters")]
reg = collections.OrderedDict()
for (f, ln, src), (i, t, s) in agg.items():
    lab = next((r[3] for r in REGIONS if r[0] == f and r[1] <= ln <= r[2]), f)
    a = reg.setdefault(lab, [0, 0, 0]); a[0] += i; a[1] += t; a[2] += s
print("\nby region:")
for lab, (i, t, s) in sorted(reg.items(), key=lambda kv: -kv[1][2]):
    print(f"{100*s/max(tot_s,1):5.1f}% smp {100*i/max(tot_i,1):5.1f}% inst  act {t/max(i,1):5.1f}  {lab}")
