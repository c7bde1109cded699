#!/bin/bash
# cell table (one sector per look-up): tests, then cfg 4 with it against the texture path, plus one ncu capture of the cell kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== brick / cell tests"; timeout 600 python -m pytest tests/test_bricks_gpu.py -q -s 2>&1 | grep -E "cell mode|software filter, tex|fast vs|passed|failed|Error|assert" | head
show='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), round(d["ms_per_step"],3), {k: round(v,3) for k,v in r["kernel_ms_per_step"].items()}, "parity", (d.get("parity") or {}).get("flipped_frac"), (d.get("parity") or {}).get("max_abs"))'
echo "== cfg4 cell table"; timeout 600 python bench.py --config 4 --cells --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/q.err | tail -1 | tee gpurun_out/r02r_bench_cfg4_cells.json | python -c "$show"; tail -1 gpurun_out/q.err | cut -c1-300
echo "== cfg4 512^3 cell table"; VPT_BENCH_GRID=512 timeout 600 python bench.py --config 4 --cells --steps 3 --warmup 3 --no-cpu-baseline --no-parity 2>gpurun_out/q.err | tail -1 | python -c "$show"
echo "== ncu cell kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace" -s 3 -c 1 -f -o gpurun_out/prof_r02_cfg4_cells python bench.py --config 4 --cells --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu4c.log 2>&1; tail -1 gpurun_out/ncu4c.log | cut -c1-150
