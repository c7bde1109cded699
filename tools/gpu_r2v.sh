#!/bin/bash
# tile loop in k_generate: headline unchanged?  one pass per call faster?  bitwise tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in r["kernel_ms_per_step"].items()}, "parity", (d.get("parity") or {}).get("flipped_frac"))'
echo "== cfg2"; timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>gpurun_out/q.err | tail -1 | python -c "$show"
echo "== cfg2, one pass per call"; timeout 600 python bench.py --chunk 1 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/q.err | tail -1 | tee gpurun_out/r02v_bench_cfg2_progressive.json | python -c "$show"
echo "== cfg2, 4 passes per call"; timeout 600 python bench.py --chunk 4 --steps 10 --warmup 3 --no-cpu-baseline --no-parity 2>gpurun_out/q.err | tail -1 | python -c "$show"
echo "== cfg1"; timeout 600 python bench.py --config 1 --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/q.err | tail -1 | python -c "$show"
echo "== tests"; timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden_gpu.py -q 2>&1 | tail -2
