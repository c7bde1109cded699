#!/bin/bash
cd "$(dirname "$0")/.."
show='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), round(d["ms_per_step"],3), {k: round(v,3) for k,v in r["kernel_ms_per_step"].items()}, "simt", round(r["step_loop_simt_efficiency"],3))'
for lib in libvpt_b200.so libvpt_o1.so; do for m in 20 26 30; do
echo "== $lib sched_min_lanes $m"; VPT_LIB_NAME=$lib timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-parity --sched-min-lanes $m 2>/dev/null | tail -1 | python -c "$show"
done; done
