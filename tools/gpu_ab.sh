#!/bin/bash
# A/B of library variants: tools/gpu_ab.sh libA.so libB.so ...   (first: parity run of the default library)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== parity (${PARITY_LIB:-default lib})"; VPT_LIB_NAME=${PARITY_LIB:-libvpt_b200.so} timeout 900 python -m pytest tests -m gpu -x -q -k "golden or live or fused or partition or point_lights" 2>&1 | tail -3
for lib in "$@"; do
echo "== $lib"; VPT_LIB_NAME=$lib timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'kernel ms/step', {k: round(v,2) for k,v in r['kernel_ms_per_step'].items()}, 'simt', round(r['step_loop_simt_efficiency'],3))
"
done
