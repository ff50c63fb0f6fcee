#!/bin/bash
# Links the MEX gateway with the minimal mxArray runtime (mex/stub/mexrt.c) and libekfslam.so into
# mex/_build/libekfslam_mextest.so, which tests/test_mex_gateway.py drives through ctypes.
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
mkdir -p "$HERE/_build"
gcc -O1 -g -shared -fPIC -Wall -I"$HERE/stub" -I"$HERE/../include" "$HERE/ekfslam_mex.c" "$HERE/stub/mexrt.c" \
    -L"$HERE/../ekf-slam_b200" -lekfslam -Wl,-rpath,'$ORIGIN/../../ekf-slam_b200' -lm -o "$HERE/_build/libekfslam_mextest.so"
echo "built $HERE/_build/libekfslam_mextest.so"
