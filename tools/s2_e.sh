# ncu --set full of the deferral's correction kernels at the full batch (frames 7-8)
mkdir -p gpurun_out
BENCH="python bench.py --batch 4096 --steps 2 --warmup 6 --no-e2e --no-cpu-baseline --no-extra"
ncu --set full --clock-control none --import-source on -k regex:"k_gcorr|k_vpend|k_ransac" -s 18 -c 3 -o gpurun_out/corr $BENCH > gpurun_out/ncu_corr.log 2>&1
tail -n 2 gpurun_out/ncu_corr.log
python tools/ncu_digest.py gpurun_out/corr.ncu-rep 22 > gpurun_out/corr_digest.txt 2>&1
