#!/bin/bash
# A/B of bench options: tools/gpu_ab2.sh "<opts A>" "<opts B>" ...
cd "$(dirname "$0")/.."
echo "== parity"; timeout 900 python -m pytest tests -m gpu -x -q -k "golden or live or fused or partition or point_lights or lean or fireball or smoke" 2>&1 | tail -3
for o in "$@" "$@"; do
echo "== bench $o"; timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline $o 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'kernel ms/step', {k: round(v,2) for k,v in r['kernel_ms_per_step'].items()})
"
done
