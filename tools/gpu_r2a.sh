#!/bin/bash
# round 2, GPU session A: tests -> smoke -> bench (ours, reference) -> ncu launch list + full captures of the three frame kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== pytest gpu"; timeout 1500 python -X faulthandler -m pytest tests -m gpu -q 2>&1 | tail -40
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_ours.err | tail -1 | tee gpurun_out/r02a_bench_ours.json | cut -c1-2500; tail -3 gpurun_out/bench_ours.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r02a_bench_ref.json | cut -c1-300
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_generate|k_resolve" -s 9 -c 3 -f -o gpurun_out/prof_r02a python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out | tail -12
