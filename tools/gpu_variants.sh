#!/bin/bash
cd "$(dirname "$0")/.."
for lib in libvpt_s3c4.so libvpt_s2c5.so libvpt_s2c6.so libvpt_s2c8.so; do
for L in 8 16; do
echo "== $lib sched_min_lanes=$L"; VPT_LIB_NAME=$lib timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sched-min-lanes $L 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'kernel ms/step', {k: round(v,2) for k,v in r['kernel_ms_per_step'].items()}, 'simt', round(r['step_loop_simt_efficiency'],3))
"
done; done
