#!/bin/bash
# round 2, GPU session D: Bruneton precompute tests, texture-filter probe, fill test vs the -G reference kernel, 2-GPU NCCL check if available
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== filter probe"; timeout 300 python tools/tex_filter_probe.py 2>&1 | tail -40
echo "== atmosphere tests"; timeout 1200 python -m pytest tests/test_atmosphere_gpu.py -q -s 2>&1 | grep -vE "^\s*$" | tail -40
echo "== fill test"; timeout 300 python -m pytest tests/test_bricks_gpu.py -q -s -k fill 2>&1 | grep -E "reference fill|fill:|passed|failed|Error" | head
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
