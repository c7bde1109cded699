#!/bin/bash
# GPU session: goldens -> gpu tests -> bench (ours, reference) -> ncu launch list + full capture of k_trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== goldens"; timeout 600 python tests/golden/make_golden.py 2>&1 | tail -5
if [ "$1" == "golden" ]; then cp gpurun_out/golden/*.npz tests/golden/; fi
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
echo "== bench ours"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_ours.json
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_ref.json
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
echo "== ncu full k_trace"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 8 -c 2 -f -o gpurun_out/prof_trace python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
