# Round-end style check on one GPU: GPU tests, smoke, default bench (both arms).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
r=json.load(open("gpurun_out/bench_ref.json")); d=json.load(open("gpurun_out/bench_default.json"))
print("reference arm:", round(r["value"],1), r["cpu_baseline"]["cores"], "threads")
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],2), "launches", d["gpu_launches"])
print("roofline", {k: d["roofline"][k] for k in ("bound","achieved","peak","frac","traffic","kernel_share_of_step")})
print("cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
print("clocks", d["clocks"])
print({k: round(v,2) for k,v in d["kernels_ms_per_step"].items()})
PY
tail -2 gpurun_out/bench_default.err
