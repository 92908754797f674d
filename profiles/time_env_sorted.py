"""Fused episode (reset + rollout) for the layouts of a 65 536-session batch: sorted by the caller, sorted with an identity
order installed, interleaved, interleaved kept sorted by the environment.  usage: python profiles/time_env_sorted.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from abrsimulator_b200 import synth
from abrsimulator_b200.env import BatchedABREnv

V, N = 48, 65536
bitrates, sizes = synth.make_video(V)
bw, tl, ti = synth.make_traces(1024, 2048)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
env = BatchedABREnv(bw, sizes, bitrates, int(sys.argv[1]) if len(sys.argv) > 1 else N, trace_len=tl, trace_interval=ti)
out = {k: torch.empty(V, N, dtype=torch.float64, device=dev) for k in ("delay", "sleep", "buffer", "rebuffer", "reward")}
out["end_of_video"] = torch.empty(V, N, dtype=torch.uint8, device=dev)
tid_s, off_s = synth.make_sessions(N, 1024, 2048, group=64)
tid_i, off_i = synth.make_sessions(N, 1024, 2048, group=1)
ident = torch.arange(N, dtype=torch.int32, device=dev)
cases = [("sorted, no order", tid_s, off_s, None, False), ("sorted, identity order installed", tid_s, off_s, ident, False),
         ("interleaved", tid_i, off_i, None, False), ("interleaved, env sorts", tid_i, off_i, None, True)]
for name, tid, off, perm, sort in cases:
    tid_d, off_d = torch.from_numpy(tid).to(dev), torch.from_numpy(off).to(dev)
    ms = []
    for it in range(12):
        flush.fill_(1)
        env.set_order(perm)
        env.reset(tid_d, off_d, sort_by_trace=sort)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        env.rollout("random", V, seed=7, out=out)
        k1.record()
        k1.synchronize()
        if it >= 4:
            ms.append(k0.elapsed_time(k1))
    tr = env.state("trace_id").cpu().numpy()
    runs = int((np.diff(tr) != 0).sum()) + 1
    print(f"{name:36s} {1e3 * sum(ms) / len(ms):7.1f} us   trace runs in environment order: {runs}  reward sum {float(out['reward'].sum()):.6f}")
