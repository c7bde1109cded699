#!/bin/bash
# final ncu evidence for the round: launch list of the default bench command + full capture of the three frame kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-120
echo "== ncu full cfg2"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_generate|k_resolve" -s 9 -c 3 -f -o gpurun_out/prof_r02_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-120
