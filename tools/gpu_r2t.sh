#!/bin/bash
# A/B: paired transmittance steps in the 2-rays-per-lane lean kernel (cfg 4: texture path and cell table), parity, tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), round(d["ms_per_step"],3), {k: round(v,3) for k,v in r["kernel_ms_per_step"].items()}, "parity", (d.get("parity") or {}).get("flipped_frac"), (d.get("parity") or {}).get("max_abs"), "lookups/sample", round(r["density_lookups_per_sample"],3))'
echo "== cfg4 texture path, paired"; timeout 600 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/q.err | tail -1 | tee gpurun_out/r02t_bench_cfg4.json | python -c "$show"; tail -1 gpurun_out/q.err | cut -c1-200
echo "== cfg4 texture path, unpaired"; VPT_LIB_NAME=libvpt_nopair.so timeout 600 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline --no-parity 2>gpurun_out/q.err | tail -1 | python -c "$show"
echo "== cfg4 cell table, paired"; timeout 600 python bench.py --config 4 --cells --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/q.err | tail -1 | tee gpurun_out/r02t_bench_cfg4_cells.json | python -c "$show"; tail -1 gpurun_out/q.err | cut -c1-200
echo "== tests"; timeout 900 python -m pytest tests/test_bricks_gpu.py tests/test_parity_gpu.py -q 2>&1 | tail -2
