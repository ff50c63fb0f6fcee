# session-2 check: GPU tests, then the quick bench with the hi downdate deferred (default) and the two-pass step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for d in 1 0; do
  EKFSLAM_DEFER_HI=$d timeout 900 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/quick_d$d.json 2> gpurun_out/quick_d$d.err
  python - <<PY
import json
d=json.load(open("gpurun_out/quick_d$d.json"))
print("DEFER=$d value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), {k: round(x,2) for k,x in d["kernels_ms_per_step"].items() if x>0.05})
PY
  tail -2 gpurun_out/quick_d$d.err
done
