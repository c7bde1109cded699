#!/bin/bash
# round 2, GPU session K (8 GPUs): cfg 2 at N = 8 and 4, cfg 4 and cfg 5 at N = 8 -- launched exactly as the driver does
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/gpu_r2i.sh 8 2 4 5
bash tools/gpu_r2i.sh 4 2
