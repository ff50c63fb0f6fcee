# usage: ab_build.sh "<nvcc flags A>" "<nvcc flags B>" ...: rebuild libekfslam.so with each flag set ON THE GPU BOX and run the quick bench
mkdir -p gpurun_out
for flags in "$@"; do
  bash ekf-slam_b200/csrc/build.sh $flags > /dev/null 2>&1
  timeout 150 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab.json"))
print("[$flags] ms/step", round(d["ms_per_step"],3), {k: round(x,3) for k,x in d["kernels_ms_per_step"].items() if k in ("k_chol","k_chol_hi","k_w","k_w_hi","k_upd_S","k_hp")})
PY
done
bash ekf-slam_b200/csrc/build.sh > /dev/null 2>&1
