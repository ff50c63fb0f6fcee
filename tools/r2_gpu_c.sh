#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "large_map or blocked64 or n100_many" 2>&1 | tail -15
python tools/cfg4_once.py 8 2>&1 | tail -2
EKFSLAM_CHOL_BIG=0 python tools/cfg4_once.py 8 2>&1 | tail -2
python tools/cfg4_once.py 32 2>&1 | tail -2
