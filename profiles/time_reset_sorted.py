"""Device time of reset(sort_by_trace=True) (abr_env_reset_sorted: counting sort + gather + reset) against a plain
reset, 65 536 interleaved sessions over 1 024 traces, L2 flushed.  usage: python profiles/time_reset_sorted.py"""
import sys

import torch

sys.path.insert(0, ".")
from abrsimulator_b200 import synth
from abrsimulator_b200.env import BatchedABREnv

V, N = 48, 65536
bitrates, sizes = synth.make_video(V)
bw, tl, ti = synth.make_traces(1024, 2048)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
env = BatchedABREnv(bw, sizes, bitrates, N, trace_len=tl, trace_interval=ti)
tid, off = synth.make_sessions(N, 1024, 2048, group=1)
tid_d, off_d = torch.from_numpy(tid).to(dev), torch.from_numpy(off).to(dev)
for name, sort in (("plain reset", False), ("reset, environment sorts by trace", True)):
    ms = []
    for it in range(12):
        flush.fill_(1)
        flush.fill_(2)
        env.set_order(None)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        env.reset(tid_d, off_d, sort_by_trace=sort)
        k1.record()
        k1.synchronize()
        if it >= 4:
            ms.append(k0.elapsed_time(k1))
    print(f"{name:36s} {1e3 * sorted(ms)[len(ms) // 2]:7.1f} us")
