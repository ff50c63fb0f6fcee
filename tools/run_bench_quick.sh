mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for fuse in ${FUSES:-1 0}; do
EKFSLAM_FUSE=$fuse timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_f$fuse.json 2> gpurun_out/bench_f$fuse.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_f$fuse.json"))
print("fuse=$fuse value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), {k: round(x,2) for k,x in d["kernels_ms_per_step"].items() if x>0.3})
PY
tail -3 gpurun_out/bench_f$fuse.err
done
