#!/bin/bash
# round 2, GPU session B: regenerate goldens (unmodified reference kernel), full GPU suite, brick-test details, bench cfg 2 / cfg 4 (parity, fast,
# reference), ncu traffic of the cfg-4 trace kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== goldens"; timeout 300 python tests/golden/make_golden.py 2>&1 | tail -3; cp gpurun_out/golden/*.npz tests/golden/
echo "== pytest gpu"; timeout 1500 python -X faulthandler -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== brick details"; timeout 600 python -m pytest tests/test_bricks_gpu.py -q -s -k "fill or statistics or converged" 2>&1 | grep -E "fill:|fast vs|converged|assert|Error|passed|failed" | head -30
echo "== pytest cpu oracle vs new goldens"; timeout 600 python -m pytest tests/test_oracle_cpu.py -q 2>&1 | tail -2
echo "== bench cfg2 ours"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/b.err | tail -1 | tee gpurun_out/r02b_bench_cfg2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_step'], d['parity'])"; tail -2 gpurun_out/b.err
echo "== bench cfg4 parity-mode"; timeout 1200 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b4.err | tail -1 | tee gpurun_out/r02b_bench_cfg4_parity.json | cut -c1-3000; tail -3 gpurun_out/b4.err
echo "== bench cfg4 fast-mode"; timeout 1200 python bench.py --config 4 --fast --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b4f.err | tail -1 | tee gpurun_out/r02b_bench_cfg4_fast.json | cut -c1-3000; tail -3 gpurun_out/b4f.err
echo "== bench cfg4 reference"; timeout 1200 python bench.py --config 4 --impl reference --steps 2 --warmup 3 2>gpurun_out/b4r.err | tail -1 | tee gpurun_out/r02b_bench_cfg4_ref.json | cut -c1-600; tail -3 gpurun_out/b4r.err
echo "== ncu cfg4 trace kernels (parity + fast)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_trace" -s 3 -c 1 -f -o gpurun_out/prof_r02b_cfg4_tex python bench.py --config 4 --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu4.log 2>&1; tail -1 gpurun_out/ncu4.log | cut -c1-200
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_trace_brick" -s 3 -c 1 -f -o gpurun_out/prof_r02b_cfg4_brick python bench.py --config 4 --fast --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu4f.log 2>&1; tail -1 gpurun_out/ncu4f.log | cut -c1-200
ls -la gpurun_out | tail -14
