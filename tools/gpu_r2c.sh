#!/bin/bash
# round 2, GPU session C: brick diagnostics + cfg2 bench after reverting the Philox carry + other configs (1, 3, 5) and level A timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== brick tests"; timeout 600 python -m pytest tests/test_bricks_gpu.py -q -s 2>&1 | grep -E "reference fill|fill:|software filter|fast vs|converged|assert |Error|passed|failed" | head -30
echo "== bench cfg2 ours"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/b.err | tail -1 | tee gpurun_out/r02c_bench_cfg2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_step'], d['parity'])"; tail -2 gpurun_out/b.err
for c in 1 3 5; do
  echo "== bench cfg$c ours"; timeout 900 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/b$c.err | tail -1 | tee gpurun_out/r02c_bench_cfg$c.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_step'], d['parity'])"; tail -2 gpurun_out/b$c.err
  echo "== bench cfg$c reference"; timeout 900 python bench.py --config $c --impl reference --steps 2 --warmup 3 2>gpurun_out/br$c.err | tail -1 | tee gpurun_out/r02c_bench_cfg${c}_ref.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"; tail -2 gpurun_out/br$c.err
done
echo "== level A cfg1"; timeout 600 python bench.py --config 1 --level-a --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/la1.err | tail -1 | tee gpurun_out/r02c_bench_cfg1_levelA.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['parity'])"; tail -2 gpurun_out/la1.err
echo "== level A cfg2"; timeout 600 python bench.py --config 2 --level-a --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/la2.err | tail -1 | tee gpurun_out/r02c_bench_cfg2_levelA.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['parity'])"; tail -2 gpurun_out/la2.err
ls gpurun_out | tail -20
