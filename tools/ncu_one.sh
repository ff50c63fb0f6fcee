# usage: ncu_one.sh <kernel regex> <out name> [env assignments]
mkdir -p gpurun_out
BENCH="python bench.py --batch 512 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
env $3 $BENCH > gpurun_out/plain_$2.log 2>&1 &&
env $3 ncu --set full --clock-control none --import-source on -k regex:"$1" -s 4 -c 2 -o gpurun_out/$2 $BENCH > gpurun_out/ncu_$2.log 2>&1
tail -n 3 gpurun_out/ncu_$2.log
