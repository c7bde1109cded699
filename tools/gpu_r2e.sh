#!/bin/bash
# round 2, GPU session E: exact-coordinate brick filter, atmosphere params, fill vs -G reference, scatter relocation effect (cfg 2 / cfg 4), k_trace occupancy A/B on cfg 4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== brick + atmosphere tests"; timeout 1200 python -m pytest tests/test_bricks_gpu.py tests/test_atmosphere_gpu.py -q -s 2>&1 | grep -E "reference fill|fill:|software filter|fast vs|converged|median rel|own-table|own precompute|differs|assert |Error|passed|failed" | head -40
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
show='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in r["kernel_ms_per_step"].items()}, "parity", (d.get("parity") or {}).get("flipped_frac"), "bricks/lookup", r.get("bricks_staged_per_lookup"))'
echo "== bench cfg2"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/b.err | tail -1 | tee gpurun_out/r02e_bench_cfg2.json | python -c "$show"; tail -2 gpurun_out/b.err
echo "== bench cfg4 parity"; timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b4.err | tail -1 | tee gpurun_out/r02e_bench_cfg4_parity.json | python -c "$show"; tail -2 gpurun_out/b4.err
echo "== bench cfg4 fast"; timeout 900 python bench.py --config 4 --fast --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b4f.err | tail -1 | tee gpurun_out/r02e_bench_cfg4_fast.json | python -c "$show"; tail -2 gpurun_out/b4f.err
echo "== A/B s2c5: cfg2, cfg4, cfg3"
VPT_LIB_NAME=libvpt_s2c5.so timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab.err | tail -1 | python -c "$show"; tail -2 gpurun_out/ab.err
VPT_LIB_NAME=libvpt_s2c5.so timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab.err | tail -1 | python -c "$show"; tail -2 gpurun_out/ab.err
VPT_LIB_NAME=libvpt_s2c5.so timeout 900 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab.err | tail -1 | python -c "$show"; tail -2 gpurun_out/ab.err
echo "== cfg3 default"; timeout 900 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b3.err | tail -1 | python -c "$show"
