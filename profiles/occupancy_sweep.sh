#!/bin/bash
# The fused episode against the number of resident blocks per SM (64-thread blocks, 148 SMs): 37 888 sessions = 4 blocks
# per SM (2 warps per scheduler), 47 360 = 5 (3,3,2,2), 56 832 = 6 (3 each), 65 536 = 6.9 (4,4,3,3 on 136 SMs).
# steps=1 is the launch's fixed cost, steps=48 the bench episode.  Usage (under gpurun): bash profiles/occupancy_sweep.sh
for n in 37888 47360 56832 65536; do
  for s in 1 48; do
    python profiles/time_rollout.py 200 sessions=$n steps=$s
  done
done
