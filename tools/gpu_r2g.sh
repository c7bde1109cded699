#!/bin/bash
# round 2, GPU session G: exact software filter in the brick path, progressive mode, compute_61 JIT reference row, 100-step headline, final ncu captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== brick tests"; timeout 1200 python -m pytest tests/test_bricks_gpu.py -q -s 2>&1 | grep -E "fill:|software filter|fast vs|converged|assert |Error|passed|failed" | head -20
show='import sys,json; d=json.loads(sys.stdin.read()); r=d.get("roofline") or {}; print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in (r.get("kernel_ms_per_step") or {}).items()}, "parity", (d.get("parity") or {}), "bricks/lookup", r.get("bricks_staged_per_lookup"))'
echo "== bench cfg2, 100 steps"; timeout 900 python bench.py --steps 100 --warmup 5 2>gpurun_out/b.err | tail -1 | tee gpurun_out/r02g_bench_cfg2.json | python -c "$show"; tail -2 gpurun_out/b.err
echo "== bench cfg2 progressive (1 pass per call)"; timeout 900 python bench.py --chunk 1 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/b.err | tail -1 | tee gpurun_out/r02g_bench_cfg2_progressive.json | python -c "$show"; tail -2 gpurun_out/b.err
echo "== reference cfg2, compute_61 PTX JIT"; timeout 900 python bench.py --impl reference --ref-jit61 --steps 3 --warmup 3 2>gpurun_out/b.err | tail -1 | tee gpurun_out/r02g_bench_cfg2_ref_jit61.json | python -c "$show"; tail -2 gpurun_out/b.err
echo "== reference cfg2, sm_100a"; timeout 900 python bench.py --impl reference --steps 3 --warmup 3 2>gpurun_out/b.err | tail -1 | tee gpurun_out/r02g_bench_cfg2_ref.json | python -c "$show"; tail -2 gpurun_out/b.err
echo "== bench cfg4 brick mode"; timeout 900 python bench.py --config 4 --fast --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b4f.err | tail -1 | tee gpurun_out/r02g_bench_cfg4_brick.json | python -c "$show"; tail -2 gpurun_out/b4f.err
echo "== bench cfg5 (auto: 2 rays per lane)"; timeout 900 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b5.err | tail -1 | tee gpurun_out/r02g_bench_cfg5.json | python -c "$show"
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
echo "== ncu full cfg2"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_generate|k_resolve" -s 9 -c 3 -f -o gpurun_out/prof_r02_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-200
echo "== ncu full cfg4 (2 rays per lane)"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_trace" -s 3 -c 1 -f -o gpurun_out/prof_r02_cfg4 python bench.py --config 4 --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu4.log 2>&1; tail -1 gpurun_out/ncu4.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
