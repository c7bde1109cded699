#!/usr/bin/env python3
"""Turn an .ncu-rep into the markdown summary committed under profiles/ (key metrics, stall mix, hot source lines).
usage: tools/ncu_summary.py report.ncu-rep [title] > profiles/xxx.md"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; title = sys.argv[2] if len(sys.argv) > 2 else rep
def run(*a): return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout
raw = list(csv.reader(io.StringIO(run("--page", "raw", "--csv"))))
hdr, units, data = raw[0], raw[1], raw[2:]
ki = hdr.index("Kernel Name")
want = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
        ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs/SM"), ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs/SM"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "avg active threads / instruction"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"), ("l1tex__t_sector_hit_rate.pct", "L1/TEX hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (expected 0: no dense contraction on this path)")]
print(f"# {title}\n\nSource: `{rep.split('/')[-1]}` (`ncu --set full --clock-control none --import-source on`, B200).  Per-launch values; times under ncu are serialised / cold-cache.\n")
print("| metric | " + " | ".join(d[ki].split("(")[0] for d in data) + " |"); print("|---|" + "---|" * len(data))
for m, name in want:
    if m in hdr:
        i = hdr.index(m); print(f"| {name} [{units[i]}] | " + " | ".join(d[i] for d in data) + " |")
src = list(csv.reader(io.StringIO(run("--page", "source", "--csv"))))
# stall mix per kernel
kern = None; h = None; st = collections.OrderedDict()
for r in src:
    if len(r) >= 2 and r[0] == "Kernel Name": kern = r[1].split("(")[0]; st[kern] = collections.Counter(); continue
    if r and r[0] == "Address": h = {x: i for i, x in enumerate(r)}; continue
    if h is None or kern is None or len(r) < len(h): continue
    for k, i in h.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            try: st[kern][k] += float(r[i])
            except ValueError: pass
print("\n## Warp stall sampling (share of samples)\n")
for k, c in st.items():
    S = sum(c.values()) or 1
    print(f"- `{k}`: " + ", ".join(f"{n[6:]} {100*v/S:.1f}%" for n, v in c.most_common(7)))
# hot source lines
out = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "cuda,sass"))))
cur = None; hdr2 = None; agg = collections.OrderedDict(); kern = None; tot = collections.Counter()
for r in out:
    if len(r) >= 2 and r[0] == "Function Name": kern = r[1].split("(")[0]; continue
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr2 = {x: i for i, x in enumerate(r)}; continue
    if hdr2 is None or not r or r[0] in ("", "0") or len(r) < len(hdr2): continue
    try: ln = int(r[0]); inst = float(r[hdr2["Instructions Executed"]]); thr = float(r[hdr2["Thread Instructions Executed"]]); smp = float(r[hdr2["# Samples"]])
    except ValueError: continue
    a = agg.setdefault((kern, cur, ln, r[1].strip()[:100]), [0, 0, 0]); a[0] += inst; a[1] += thr; a[2] += smp; tot[kern] += smp
print("\n## Hottest source lines (by stall samples)\n")
for k in tot:
    print(f"### `{k}`\n\n| samples | instr share | avg active lanes | line |\n|---|---|---|---|")
    ti = sum(v[0] for kk, v in agg.items() if kk[0] == k) or 1
    rows = sorted(((kk, v) for kk, v in agg.items() if kk[0] == k), key=lambda kv: -kv[1][2])[:18]
    for (kk, f, ln, s), (i, t, sm) in rows:
        print(f"| {100*sm/max(tot[k],1):.1f}% | {100*i/ti:.1f}% | {t/max(i,1):.1f} | `{f}:{ln}` {s.replace('|', '/')} |")
    print()
