mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for mode in ${MODES:-ws128 ws}; do
EKFSLAM_DOWNDATE=$mode timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$mode.json"))
print("$mode value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), {k: round(x,2) for k,x in d["kernels_ms_per_step"].items() if x>0.4})
PY
tail -3 gpurun_out/bench_$mode.err
done
