#!/bin/bash
# quick A/B of the fused episode kernel: parity tests of the chunk-step path, a short bench, and the executed-instruction
# count of one launch (ncu, three metrics only).  Usage (under gpurun): bash profiles/quick_rollout.sh <tag>
TAG=${1:-x}
python -m pytest tests/test_gpu_env.py -m gpu -x -q > gpurun_out/q_${TAG}_pytest.log 2>&1; tail -2 gpurun_out/q_${TAG}_pytest.log
python bench.py --steps 20 --warmup 5 --no-mpc --no-step-form --no-cpu-baseline > gpurun_out/q_${TAG}_bench.log 2>gpurun_out/q_${TAG}_bench.err || tail -5 gpurun_out/q_${TAG}_bench.err
python profiles/show_bench.py gpurun_out/q_${TAG}_bench.log | head -1
python profiles/time_steps.py 4736 65536 > gpurun_out/q_${TAG}_steps.log 2>&1; cat gpurun_out/q_${TAG}_steps.log
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:abr_rollout_kernel -s 4 -c 1 python bench.py --steps 3 --warmup 3 --no-mpc --no-step-form --no-cpu-baseline 2>&1 | grep -E "inst_executed|time_duration|issue_active" | tee gpurun_out/q_${TAG}_ncu.log
