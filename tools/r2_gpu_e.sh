#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 4 > gpurun_out/r2e_ref.json 2>/dev/null; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench.json')); r=json.load(open('gpurun_out/r2e_ref.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
print('ref',r['value'],r['ms_per_step'],r['config']['workload']==d['config']['workload'])
print({k:round(v,3) for k,v in d['kernels_ms_per_step'].items()})
PY
