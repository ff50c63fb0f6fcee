mkdir -p gpurun_out
BENCH="python bench.py --batch 512 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$BENCH > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $BENCH > gpurun_out/ncu_list.log 2>&1
$BENCH > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_downdate|k_hp|k_chol|k_w|k_ransac" -s 10 -c 10 -o gpurun_out/prof_r1a $BENCH > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_list.log gpurun_out/ncu_full.log
ls -la gpurun_out
