#!/bin/bash
cd "$(dirname "$0")/.."
for F in 0 1 2 4 7; do echo "=== debug_flags=$F"; timeout 300 python tools/gpu_compare.py 640 360 3 ray_depth=100 debug_flags=$F 2>&1 | grep -E "accum|raw |   px" | head -8; done
