# usage: launch_list.sh <tag> [env assignments]: ncu launch list (device time per launch) of a short default-shape run
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup ${WARM:-2} --no-e2e --no-cpu-baseline"
env "${@:2}" $BENCH > gpurun_out/plain_$1.log 2>&1 &&
env "${@:2}" ncu --metrics gpu__time_duration.sum --clock-control none -c ${CNT:-400} --csv --log-file gpurun_out/launches_$1.csv $BENCH > gpurun_out/ncu_list_$1.log 2>&1
tail -n 1 gpurun_out/ncu_list_$1.log
