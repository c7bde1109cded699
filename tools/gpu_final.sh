#!/bin/bash
# final validation of the round: GPU suite, smoke(), both bench arms with default flags
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$1" != "bench" ]; then
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
fi
echo "== bench (default flags)"; t0=$SECONDS; timeout 900 python bench.py 2>gpurun_out/f.err | tail -1 > gpurun_out/r02_final_bench.json; echo "$((SECONDS - t0)) s wall"
python -c "
import json; d=json.load(open('gpurun_out/r02_final_bench.json')); print({k: d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','higher_is_better','scaling','vs_baseline','dtype','data','gpu_launches')}); print(d['e2e']); print(d['clocks']); print({k: d['roofline'][k] for k in ('bound','achieved','peak','unit','frac','traffic')}); print(d['cpu_baseline']); print(d['parity'])"
echo "== bench --impl reference (default flags)"; t0=$SECONDS; timeout 900 python bench.py --impl reference 2>gpurun_out/f.err | tail -1 > gpurun_out/r02_final_bench_ref.json; echo "$((SECONDS - t0)) s wall"
python -c "
import json; d=json.load(open('gpurun_out/r02_final_bench_ref.json')); print({k: d[k] for k in ('impl','value','unit','ms_per_step','steps')}); print(d['cpu_baseline']); print(d['e2e'])"
