#!/bin/bash
# round 2, GPU session L (2 GPUs): peer-memory exchange -- bitwise check, then bench N=2 with it and with NCCL
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== mgpu_check"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM\|^$" | tail -8 | cut -c1-300
show='import sys,json; d=json.loads(sys.stdin.read()); r=d.get("roofline") or {}; print("N", d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in (r.get("kernel_ms_per_step") or {}).items()}, "parity", {k: v for k, v in (d.get("parity") or {}).items() if k in ("flipped_frac", "gathered_equals_single_gpu_bitwise")}, d["config"].get("collective", "")[:60])'
for ex in p2p nccl; do
  echo "== bench cfg2 N=2 exchange=$ex"
  VPT_EXCHANGE=$ex timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/s.err | grep '^{' | tail -1 | tee gpurun_out/r02l_n2_$ex.json | python -c "$show"
  grep -v "^W0\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/s.err | tail -3 | cut -c1-300
done
echo "== NCCL pytest"; timeout 600 python -m pytest tests/test_scale_gpu.py -q -k two_rank 2>&1 | tail -2
