#!/bin/bash
# A/B: chunk length at N=2 (auto = 64) against 32, and 64 at N=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show='import sys,json; d=json.loads(sys.stdin.read()); r=d.get("roofline") or {}; print("N", d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in (r.get("kernel_ms_per_step") or {}).items()}, "launches/step", r.get("launches_per_step"))'
for ch in 32 16; do
echo "== N=2 chunk $ch"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + ch)) bench.py --gpus 2 --steps 20 --warmup 3 --chunk $ch --no-parity 2>gpurun_out/s.err | grep '^{' | tail -1 | python -c "$show"
done
echo "== N=1 chunk 64"; timeout 300 python bench.py --steps 20 --warmup 3 --chunk 64 --no-parity --no-cpu-baseline 2>gpurun_out/s.err | tail -1 | python -c "$show"
echo "== N=1 chunk 16"; timeout 300 python bench.py --steps 20 --warmup 3 --chunk 16 --no-parity --no-cpu-baseline 2>gpurun_out/s.err | tail -1 | python -c "$show"
