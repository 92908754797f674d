import cProfile, pstats, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from abrsimulator_b200 import synth
from abrsimulator_b200.env import BatchedABREnv
N, V = 65536, 48
bitrates, sizes = synth.make_video(V)
bw, tl, ti = synth.make_traces(1024, 2048)
env = BatchedABREnv(bw, sizes, bitrates, N, trace_len=tl, trace_interval=ti)
tid, off = synth.make_sessions(N, 1024, 2048, group=64)
tid_p, off_p = torch.from_numpy(tid).pin_memory(), torch.from_numpy(off).pin_memory()
out = dict(qoe_cost=torch.zeros(N, dtype=torch.float64).pin_memory(), stats=torch.zeros(10, dtype=torch.float64).pin_memory())
f = lambda: env.run_host("random", V, tid_p, off_p, seed=7, want_acc=False, out=out)
for _ in range(20): f()
import time
t0 = time.perf_counter()
for _ in range(500): f()
print("torch pinned: %.1f us" % ((time.perf_counter() - t0) / 500 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(500): f()
pr.disable(); pstats.Stats(pr).sort_stats('tottime').print_stats(10)
