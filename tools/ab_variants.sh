# usage: ab_variants.sh name1 name2 ...  (ekf-slam_b200/variants/<name>.so); env EKFSLAM_DOWNDATE passes through
mkdir -p gpurun_out
for v in "$@"; do
cp ekf-slam_b200/variants/$v.so ekf-slam_b200/libekfslam.so
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
python - <<PY
import json
d=json.load(open("gpurun_out/ab_$v.json"))
print("$v value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), {k: round(x,2) for k,x in d["kernels_ms_per_step"].items() if x>0.4})
PY
tail -2 gpurun_out/ab_$v.err
done
