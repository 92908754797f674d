#!/bin/bash
# Profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then one full capture per hot kernel.
# Usage (under gpurun):  bash profiles/run_ncu.sh <tag>
set -u
TAG=${1:-r1}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_EXTRA:-}"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
if [ -z "${SKIP_LIST:-}" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
fi
ncu --set full --clock-control none --import-source on -k regex:abr_rollout_kernel -s 4 -c 1 -f -o gpurun_out/rollout_$TAG $CMD > gpurun_out/ncu_rollout_$TAG.log 2>&1
echo "rollout capture rc=$?"
if [ -z "${SKIP_MPC:-}" ]; then
ncu --set full --clock-control none --import-source on -k regex:abr_mpc_kernel -s 3 -c 1 -f -o gpurun_out/mpc_$TAG $CMD > gpurun_out/ncu_mpc_$TAG.log 2>&1
echo "mpc capture rc=$?"
fi
if [ -n "${WITH_STEP:-}" ]; then
ncu --set full --clock-control none --import-source on -k regex:abr_step_kernel -s 10 -c 1 -f -o gpurun_out/step_$TAG $CMD --no-mpc > gpurun_out/ncu_step_$TAG.log 2>&1
echo "step capture rc=$?"
fi
ls -la gpurun_out/
