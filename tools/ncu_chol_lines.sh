# ncu --set full (with source) of the 112-row resident Cholesky of the li update at the FULL bench shape; per-line stall digest on the box
mkdir -p gpurun_out
BENCH="python bench.py --batch 4096 --steps 2 --warmup 6 --no-e2e --no-cpu-baseline --no-extra"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_chol_sm<\(int\)112, \(int\)256, \(int\)0' -s ${1:-6} -c 1 -o gpurun_out/chl $BENCH > gpurun_out/ncu_chl.log 2>&1
tail -n 2 gpurun_out/ncu_chl.log
python tools/ncu_digest.py gpurun_out/chl.ncu-rep 40 > gpurun_out/chl_digest.txt 2>&1
python tools/ncu_lines.py gpurun_out/chl.ncu-rep k_chol_sm 60 > gpurun_out/chl_lines.txt 2>&1
head -70 gpurun_out/chl_lines.txt
