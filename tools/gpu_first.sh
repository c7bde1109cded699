#!/bin/bash
# first GPU contact: sanity + parity sweeps (dev tool)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
run() { echo "=== $*"; timeout 300 python tools/gpu_compare.py "$@" 2>&1 | tail -25; echo "rc=$?"; }
run 256 256 1
run 256 256 4
run 256 256 4 fused=0
run 256 256 2 ray_depth=3 volume_depth=3
run 256 256 2 g=0.6 density_mult=2.0
run 512 512 1
run 1920 1080 2
