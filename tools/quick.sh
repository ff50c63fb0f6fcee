# usage: quick.sh [env assignments...] : GPU parity tests, then the short default-shape bench (no e2e / CPU arm)
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
env "$@" timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/quick.json 2> gpurun_out/quick.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/quick.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), {k: round(x,2) for k,x in d["kernels_ms_per_step"].items() if x>0.1})
PY
tail -2 gpurun_out/quick.err
