#!/bin/bash
# A/B of library variants built with ABR_LIB_SUFFIX=_<name> ABR_EXTRA_NVCC_FLAGS=...: kernel time from bench.py and the
# executed-instruction count / occupancy from ncu.  Usage (under gpurun): VARIANTS="v0:flags v1:flags" bash profiles/ab_variants.sh
IFS=';' read -ra SPECS <<< "$VARIANTS"
for spec in "${SPECS[@]}"; do
  name=${spec%%:*}; flags=${spec#*:}
  export ABR_LIB_SUFFIX=_$name ABR_EXTRA_NVCC_FLAGS="$flags"
  python bench.py --steps 20 --warmup 5 --no-mpc --no-step-form --no-cpu-baseline > gpurun_out/ab_${name}.log 2>gpurun_out/ab_${name}.err || tail -3 gpurun_out/ab_${name}.err
  echo "$name: $(python profiles/show_bench.py gpurun_out/ab_${name}.log | head -1)"
  if [ -z "$NO_NCU" ]; then
  ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:abr_rollout_kernel -s 4 -c 1 python bench.py --steps 3 --warmup 3 --no-mpc --no-step-form --no-cpu-baseline 2>&1 | grep -E "inst_executed|time_duration|issue_active|occupancy_limit|registers_per|warps_active" | awk '{printf "   %s %s\n", $1, $NF}'
  fi
done
