#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 300 python tools/gpu_compare.py "$@" 2>&1 | grep -E "accum|depth|raw |bluenoise|passes:|Error|error|Traceback" | head -12; }
run 256 256 2 ray_depth=3 volume_depth=3
run 640 360 3 ray_depth=100
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for C in 1 16; do
echo "== bench ours chunk=$C"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sched-min-lanes 16 --chunk $C 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
print('kernel ms/step', {k: round(v,2) for k,v in r['kernel_ms_per_step'].items()}, 'simt', round(r['step_loop_simt_efficiency'],3), 'cnt', r['trace_counters'])
"
done
echo "== ncu full k_trace"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_generate" -s 9 -c 2 -f -o gpurun_out/prof_trace2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log | cut -c1-200
