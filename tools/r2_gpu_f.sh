# GPU parity tests, the other BASELINE configs, a short default-shape bench with per-kernel times
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/configs_bench.py 2>&1 | tail -8
timeout 900 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/quick.json 2> gpurun_out/quick.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/quick.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), {k: round(x,2) for k,x in d["kernels_ms_per_step"].items() if x>0.03})
PY
tail -2 gpurun_out/quick.err
