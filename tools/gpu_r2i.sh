#!/bin/bash
# round 2, GPU session I (N GPUs): bench at N (and 1) exactly as the driver launches it; every run under its own short timeout
cd "$(dirname "$0")/.."
N=${1:-2}; shift
CFGS=${*:-2}
mkdir -p gpurun_out
show='import sys,json; d=json.loads(sys.stdin.read()); r=d.get("roofline") or {}; print("N", d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in (r.get("kernel_ms_per_step") or {}).items()}, "parity", {k: v for k, v in (d.get("parity") or {}).items() if k in ("flipped_frac", "gathered_equals_single_gpu_bitwise")}); print("   by rank:", r.get("kernel_ms_per_step_by_rank"))'
for cfg in $CFGS; do
  steps=20; [ $cfg != 2 ] && steps=3
  echo "== bench cfg$cfg N=$N"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + cfg)) bench.py --config $cfg --gpus $N --steps $steps --warmup 3 2>gpurun_out/s.err | grep '^{' | tail -1 | tee gpurun_out/r02i_scale_cfg${cfg}_n$N.json | python -c "$show"
  grep -v "^W0\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/s.err | tail -3 | cut -c1-300
done
