"""Mean time of one robust-MPC launch (horizon 5, 131 072 sessions, the bench's state) — the A/B number for variants of
abr_mpc_kernel.  usage: python profiles/time_mpc.py [exhaustive]"""
import sys

import torch

sys.path.insert(0, ".")
from abrsimulator_b200 import synth
from abrsimulator_b200.env import BatchedABREnv

M, V = 131072, 48
bitrates, sizes = synth.make_video(V)
bw, tl, ti = synth.make_traces(1024, 2048)
env = BatchedABREnv(bw, sizes, bitrates, M, trace_len=tl, trace_interval=ti, track_history=1, track_acc=1)
tid, off = synth.make_sessions(M, 1024, 2048, group=64)
env.reset(tid, off)
env.rollout("bba", 8, want=())
act = torch.empty(M, dtype=torch.int32, device="cuda")
ex = len(sys.argv) > 1 and sys.argv[1] == "exhaustive"
for _ in range(3):
    env.mpc_decide(5, "robust", out=act, exhaustive=ex)
ms = []
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.mpc_decide(5, "robust", out=act, exhaustive=ex)
    e1.record()
    e1.synchronize()
    ms.append(e0.elapsed_time(e1))
ms.sort()
print(f"{'exhaustive' if ex else 'branch and bound'}: median {ms[10]:.4f} ms per launch = {M / ms[10] / 1e3:.1f} M decisions/s")
