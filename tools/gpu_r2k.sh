#!/bin/bash
# round 2, GPU session K (8 GPUs): exchange check at N = 8, cfg 2 at N = 8 and 4, cfg 4 and cfg 5 at N = 8 -- launched exactly as the driver does
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== mgpu_check N=8"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM\|^$" | tail -6 | cut -c1-300
bash tools/gpu_r2i.sh 8 2 4 5
bash tools/gpu_r2i.sh 4 2
echo "== cfg2 N=8 over NCCL"; VPT_EXCHANGE=nccl bash tools/gpu_r2i.sh 8 2 | sed 's/^/   /'
