#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_latency.py tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q 2>&1 | tail -25
python tools/configs_bench.py > gpurun_out/r2b_configs.log 2>&1; tail -8 gpurun_out/r2b_configs.log
cp gpurun_out/configs.json gpurun_out/r2b_configs.json
