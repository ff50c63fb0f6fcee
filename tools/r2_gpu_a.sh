#!/bin/bash
# round 2, call A: bench with extras; cfg4 ncu evidence (downdate at N=500: --set full; whole-step launch list)
mkdir -p gpurun_out
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.json
python tools/cfg4_once.py 8 > gpurun_out/r2a_cfg4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_downdate -s 4 -c 4 -o gpurun_out/r2_cfg4_downdate python tools/cfg4_once.py 8 > gpurun_out/r2a_ncu1.log 2>&1
tail -n 2 gpurun_out/r2a_ncu1.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_cfg4_launches.csv python tools/cfg4_once.py 8 > gpurun_out/r2a_ncu2.log 2>&1
tail -n 2 gpurun_out/r2a_ncu2.log; wc -l gpurun_out/r2_cfg4_launches.csv
