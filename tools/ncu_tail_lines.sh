# ncu --set full (with source) of the latency-bound kernels of one step at the FULL bench shape (frame 7): per-line stall digest on the box
mkdir -p gpurun_out
BENCH="python bench.py --batch 4096 --steps 2 --warmup 6 --no-e2e --no-cpu-baseline --no-extra"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_ransac|k_upd_S|k_innov_gather|k_predict' -s ${1:-36} -c 6 -o gpurun_out/tail $BENCH > gpurun_out/ncu_tail.log 2>&1
tail -n 2 gpurun_out/ncu_tail.log
python tools/ncu_digest.py gpurun_out/tail.ncu-rep 12 > gpurun_out/tail_digest.txt 2>&1
python tools/ncu_lines.py gpurun_out/tail.ncu-rep 22 > gpurun_out/tail_lines.txt 2>&1
