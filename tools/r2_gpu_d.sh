#!/bin/bash
mkdir -p gpurun_out
python tools/cfg4_once.py 8 > gpurun_out/r2d_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_cfg4_launches.csv python tools/cfg4_once.py 8 > gpurun_out/r2d_ncu.log 2>&1
tail -n 2 gpurun_out/r2d_ncu.log; wc -l gpurun_out/r2d_cfg4_launches.csv
