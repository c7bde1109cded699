#!/bin/bash
# scheduler threshold 26 (new default) against 20 on the configurations that were measured with 20
cd "$(dirname "$0")/.."
show='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), round(d["ms_per_step"],3), {k: round(v,3) for k,v in r["kernel_ms_per_step"].items()})'
for cfg in 3 5 4; do for m in 26 20; do
echo "== cfg$cfg sched_min_lanes $m"; timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-parity --sched-min-lanes $m 2>/dev/null | tail -1 | python -c "$show"
done; done
