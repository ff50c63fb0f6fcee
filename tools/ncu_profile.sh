# Round evidence: (1) launch list of a short default-shape bench run, (2) --set full capture of every kernel of one
# step at the FULL bench shape (B=4096, N=100).  Each ncu run is preceded by the same command without ncu (exit
# code checked).  Matching launches per step: predict 1, features 2, hp 2, innov 2, rescue_gate 1, ransac 1, upd_S 2,
# chol_sm 5, chol 2, w_small 2, gemm 2, wfix 2, downdate 2 = 26; two warm-up steps are skipped.
mkdir -p gpurun_out
R=${ROUND:-r1}
BENCH="python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline"
$BENCH > gpurun_out/plain_$R.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv $BENCH > gpurun_out/ncu_list_$R.log 2>&1
$BENCH > gpurun_out/plain2_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_downdate|k_hp|k_chol|k_gemm|k_w_small|k_wfix|k_ransac|k_upd_S|k_predict|k_features|k_innov" -s ${SKIP:-54} -c ${CNT:-27} -o gpurun_out/prof_$R $BENCH > gpurun_out/ncu_full_$R.log 2>&1
tail -n 2 gpurun_out/ncu_list_$R.log gpurun_out/ncu_full_$R.log
