#!/bin/bash
cd "$(dirname "$0")/.."
show='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), round(d["ms_per_step"],3), {k: round(v,3) for k,v in r["kernel_ms_per_step"].items()})'
echo "== cfg4 cells, L2 fetch granularity 32"; VPT_CELL_L2_32=1 timeout 600 python bench.py --config 4 --cells --steps 3 --warmup 3 --no-cpu-baseline --no-parity 2>gpurun_out/q.err | tail -1 | python -c "$show"; tail -1 gpurun_out/q.err | cut -c1-200
echo "== ncu dram bytes with granularity 32"; VPT_CELL_L2_32=1 timeout 600 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:"k_trace" -s 3 -c 1 python bench.py --config 4 --cells --steps 1 --warmup 3 --no-cpu-baseline --no-parity 2>&1 | grep -E "dram__bytes_read|gpu__time_duration" | head -3
echo "== brick tests"; timeout 600 python -m pytest tests/test_bricks_gpu.py -q 2>&1 | tail -2
