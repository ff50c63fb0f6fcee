# usage: ncu_kernel_full.sh <kernel regex> <tag> [skip] [count]: ncu --set full of one kernel at the FULL bench shape (frames 7-8), digest on the box
mkdir -p gpurun_out
BENCH="python bench.py --batch 4096 --steps 2 --warmup 6 --no-e2e --no-cpu-baseline --no-extra"
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-6} -c ${4:-1} -o gpurun_out/$2 $BENCH > gpurun_out/ncu_$2.log 2>&1
tail -n 2 gpurun_out/ncu_$2.log
python tools/ncu_digest.py gpurun_out/$2.ncu-rep 40 > gpurun_out/$2_digest.txt 2>&1
