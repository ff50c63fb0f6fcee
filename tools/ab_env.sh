# usage: ab_env.sh "ENV1=.." "ENV2=.." ... : runs the quick bench once per environment assignment
mkdir -p gpurun_out
i=0
for e in "$@"; do
i=$((i+1))
env $e timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
python - <<PY
import json
d=json.load(open("gpurun_out/ab_$i.json"))
print("$e value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), {k: round(x,2) for k,x in d["kernels_ms_per_step"].items() if x>0.3})
PY
tail -2 gpurun_out/ab_$i.err
done
