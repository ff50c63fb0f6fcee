# As ncu_profile.sh (--set full capture of one step at the full bench shape), but the report is summarised ON the GPU
# box (gpurun copies back at most 64 MiB): summary table, per-kernel digest, downdate role split and DRAM traffic JSON
# are written as text; the .ncu-rep itself is kept only if it is small enough.
mkdir -p gpurun_out
R=${ROUND:-r1}
BENCH="python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-extra"
$BENCH > gpurun_out/plain2_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_downdate|k_hp|k_chol|k_gemm|k_w_small|k_wfix|k_ransac|k_upd_S|k_predict|k_features|k_innov" -s ${SKIP:-54} -c ${CNT:-27} -o /tmp/prof_$R $BENCH > gpurun_out/ncu_full_$R.log 2>&1
tail -n 2 gpurun_out/ncu_full_$R.log
python tools/ncu_summary.py /tmp/prof_$R.ncu-rep gpurun_out/ncu_${R}_summary.md > gpurun_out/ncu_${R}_summary.log 2>&1
python tools/ncu_digest.py /tmp/prof_$R.ncu-rep 8 > gpurun_out/ncu_${R}_digest.txt 2>&1
python tools/ncu_roles.py /tmp/prof_$R.ncu-rep > gpurun_out/ncu_${R}_downdate_roles.txt 2>&1
python tools/ncu_traffic.py /tmp/prof_$R.ncu-rep gpurun_out/ncu_traffic_$R.json "B=4096, N=100, one step (frame 3), kernels of round ${R}" > gpurun_out/ncu_traffic_$R.log 2>&1
sz=$(stat -c %s /tmp/prof_$R.ncu-rep)
if [ "$sz" -lt 50000000 ]; then cp /tmp/prof_$R.ncu-rep gpurun_out/; fi
ls -la gpurun_out | tail -8
