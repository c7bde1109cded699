#!/bin/bash
# round 2, GPU session H (N GPUs): brick tests with the fitted coordinate rule, NCCL bitwise test, mgpu_check, bench at 1..N GPUs exactly as the driver launches it
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== brick tests"; timeout 1200 python -m pytest tests/test_bricks_gpu.py -q -s 2>&1 | grep -E "software filter|fast vs|converged|assert |Error|passed|failed" | head -20
echo "== NCCL bitwise test"; timeout 900 python -m pytest tests/test_scale_gpu.py -q -k two_rank 2>&1 | tail -3
echo "== mgpu_check N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -4
show='import sys,json; d=json.loads(sys.stdin.read()); r=d.get("roofline") or {}; print("N", d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in (r.get("kernel_ms_per_step") or {}).items()}, "parity", {k: v for k, v in (d.get("parity") or {}).items() if k in ("flipped_frac", "gathered_equals_single_gpu_bitwise")}, d["config"].get("collective", "")[:80])'
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  echo "== bench cfg2 N=$n"
  if [ $n == 1 ]; then timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 2>gpurun_out/s.err | tail -1 | tee gpurun_out/r02h_scale_n1.json | python -c "$show"
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 20 --warmup 3 2>gpurun_out/s.err | grep '^{' | tail -1 | tee gpurun_out/r02h_scale_n$n.json | python -c "$show"; fi
  tail -1 gpurun_out/s.err | cut -c1-200
done
if [ $N -ge 8 ]; then
  for cfg in 4 5 3; do
    echo "== bench cfg$cfg N=8"
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29540 + cfg)) bench.py --config $cfg --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/s.err | grep '^{' | tail -1 | tee gpurun_out/r02h_scale_cfg${cfg}_n8.json | python -c "$show"
    tail -1 gpurun_out/s.err | cut -c1-200
  done
fi
