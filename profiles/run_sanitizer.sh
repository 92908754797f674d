#!/bin/bash
# compute-sanitizer passes over the kernels that stage rows in shared memory with TMA + mbarrier (abr_step.cu) and
# that share a parent-state cache between warps (abr_mpc.cu).  Usage (under gpurun): bash profiles/run_sanitizer.sh <tag>
# Each tool runs smoke() (fused episode on the shared-memory path, robust MPC, whole-run entry points) and a slice of
# the parity tests that exercises: per-step kernel with restaging, fused episode sorted + mixed blocks, live mode,
# long traces (opt-in shared memory), horizon-7 block-per-session MPC.
set -u
TAG=${1:-r2}
OUT=gpurun_out/sanitizer_$TAG
mkdir -p $OUT
TESTS='tests/test_gpu_env.py::test_step_kernel_shared_memory_trace_path tests/test_gpu_env.py::test_fused_rollout_shared_memory_trace_path tests/test_gpu_env.py::test_fused_live_episode_matches_oracle tests/test_gpu_env.py::test_long_traces_take_the_opt_in_shared_memory_path_or_fall_back tests/test_gpu_env.py::test_trace_sorted_order_is_bit_identical_to_the_callers_order tests/test_gpu_mpc.py::test_horizon7_block_per_session tests/test_gpu_mpc.py::test_mode1_robust_batch_matches_oracle'
for TOOL in memcheck racecheck synccheck initcheck; do
  EXTRA=""
  [ $TOOL = memcheck ] && EXTRA="--leak-check no"
  [ $TOOL = initcheck ] && EXTRA="--track-unused-memory no"
  timeout 900 compute-sanitizer --tool $TOOL $EXTRA --error-exitcode 86 --log-file $OUT/${TOOL}_smoke.log \
      python __graft_entry__.py smoke > $OUT/${TOOL}_smoke.out 2>&1
  echo "$TOOL smoke rc=$?" | tee -a $OUT/summary.txt
  timeout 1500 compute-sanitizer --tool $TOOL $EXTRA --error-exitcode 86 --log-file $OUT/${TOOL}_tests.log \
      python -m pytest $TESTS -x -q -m gpu > $OUT/${TOOL}_tests.out 2>&1
  echo "$TOOL tests rc=$?" | tee -a $OUT/summary.txt
  tail -2 $OUT/${TOOL}_tests.out | tee -a $OUT/summary.txt
  grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" $OUT/${TOOL}_smoke.log $OUT/${TOOL}_tests.log | tee -a $OUT/summary.txt
done
