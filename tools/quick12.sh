# short default-shape bench over frames 5-16 (the window the round-2 A/B numbers in DESIGN.md use); env assignments as arguments
mkdir -p gpurun_out
env "$@" timeout 150 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/quick12.json 2> gpurun_out/quick12.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/quick12.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), {k: round(x,2) for k,x in d["kernels_ms_per_step"].items() if x>0.03})
PY
tail -2 gpurun_out/quick12.err
