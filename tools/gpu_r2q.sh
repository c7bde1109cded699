#!/bin/bash
# A/B: look-ahead texture request in the 2-rays-per-lane lean kernel (cfg 4, 1024^3 and 512^3)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), round(d["ms_per_step"],3), {k: round(v,3) for k,v in r["kernel_ms_per_step"].items()}, "parity", (d.get("parity") or {}).get("flipped_frac"))'
echo "== cfg4 look-ahead"; timeout 600 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/q.err | tail -1 | tee gpurun_out/r02q_bench_cfg4.json | python -c "$show"; tail -1 gpurun_out/q.err | cut -c1-200
echo "== cfg4 without"; VPT_LIB_NAME=libvpt_noahead.so timeout 600 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline --no-parity 2>gpurun_out/q.err | tail -1 | python -c "$show"; tail -1 gpurun_out/q.err | cut -c1-200
echo "== cfg4 512^3 look-ahead / without"; VPT_BENCH_GRID=512 timeout 600 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline --no-parity 2>gpurun_out/q.err | tail -1 | python -c "$show"
VPT_BENCH_GRID=512 VPT_LIB_NAME=libvpt_noahead.so timeout 600 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline --no-parity 2>gpurun_out/q.err | tail -1 | python -c "$show"
echo "== procedural parity test"; timeout 600 python -m pytest tests/test_bricks_gpu.py -q 2>&1 | tail -2
