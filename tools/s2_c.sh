# quick bench only (deferred step), per-kernel times; env assignments as arguments
mkdir -p gpurun_out
env "$@" timeout 900 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/quick_c.json 2> gpurun_out/quick_c.err
python - <<PY
import json
d=json.load(open("gpurun_out/quick_c.json"))
print("$@ value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), {k: round(x,2) for k,x in d["kernels_ms_per_step"].items() if x>0.05})
PY
tail -2 gpurun_out/quick_c.err
