#!/bin/bash
# one point location per step call + out-of-line sphere test: all configurations, scheduler threshold sweep, parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in r["kernel_ms_per_step"].items()}, "parity", (d.get("parity") or {}).get("flipped_frac"), "simt", round(r["step_loop_simt_efficiency"],3))'
echo "== cfg2"; timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>gpurun_out/q.err | tail -1 | tee gpurun_out/r02x_bench_cfg2.json | python -c "$show"
for m in 12 16 26; do echo "== cfg2 sched_min_lanes $m"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity --sched-min-lanes $m 2>gpurun_out/q.err | tail -1 | python -c "$show"; done
echo "== cfg3"; timeout 600 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/q.err | tail -1 | tee gpurun_out/r02x_bench_cfg3.json | python -c "$show"
echo "== cfg4"; timeout 600 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/q.err | tail -1 | tee gpurun_out/r02x_bench_cfg4.json | python -c "$show"
echo "== cfg5"; timeout 600 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/q.err | tail -1 | tee gpurun_out/r02x_bench_cfg5.json | python -c "$show"
echo "== cfg1"; timeout 600 python bench.py --config 1 --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/q.err | tail -1 | tee gpurun_out/r02x_bench_cfg1.json | python -c "$show"
