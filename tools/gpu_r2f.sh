#!/bin/bash
# round 2, GPU session F: texture-unit blend probe, hybrid scatter placement, rays-per-lane policy (2 for grids beyond L2), mu_s_min
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== blend probe"; timeout 600 python tools/tex_blend_probe.py > gpurun_out/r02f_tex_blend_probe.txt 2>&1; cat gpurun_out/r02f_tex_blend_probe.txt | head -70
echo "== brick + atmosphere tests"; timeout 1200 python -m pytest tests/test_bricks_gpu.py tests/test_atmosphere_gpu.py -q -s 2>&1 | grep -E "reference fill|fill:|software filter|fast vs|converged|median rel|own-table|own precompute|differs|assert |Error|passed|failed" | head -40
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
show='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in r["kernel_ms_per_step"].items()}, "parity", (d.get("parity") or {}).get("flipped_frac"), "bricks/lookup", r.get("bricks_staged_per_lookup"))'
echo "== bench cfg2"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/b.err | tail -1 | tee gpurun_out/r02f_bench_cfg2.json | python -c "$show"; tail -2 gpurun_out/b.err
echo "== bench cfg2 slots=2"; VPT_BENCH_OPTIONS=trace_slots=2 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/b.err | tail -1 | python -c "$show"; tail -2 gpurun_out/b.err
echo "== bench cfg4 (auto: 2 rays per lane)"; timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b4.err | tail -1 | tee gpurun_out/r02f_bench_cfg4.json | python -c "$show"; tail -2 gpurun_out/b4.err
echo "== bench cfg4 slots=3"; VPT_BENCH_OPTIONS=trace_slots=3 timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b4.err | tail -1 | python -c "$show"; tail -2 gpurun_out/b4.err
echo "== bench cfg4 512^3 slots 3 / 2"; for s in 3 2; do VPT_BENCH_GRID=512 VPT_BENCH_OPTIONS=trace_slots=$s timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b4.err | tail -1 | python -c "$show"; done
echo "== bench cfg3"; timeout 900 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b3.err | tail -1 | tee gpurun_out/r02f_bench_cfg3.json | python -c "$show"
echo "== bench cfg3 slots=2"; VPT_BENCH_OPTIONS=trace_slots=2 timeout 900 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b3.err | tail -1 | python -c "$show"
echo "== bench cfg5"; timeout 900 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b5.err | tail -1 | tee gpurun_out/r02f_bench_cfg5.json | python -c "$show"
echo "== bench cfg5 slots=2"; VPT_BENCH_OPTIONS=trace_slots=2 timeout 900 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b5.err | tail -1 | python -c "$show"
echo "== bench cfg1"; timeout 900 python bench.py --config 1 --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/b1.err | tail -1 | tee gpurun_out/r02f_bench_cfg1.json | python -c "$show"
