"""Mean CUDA-event time of the bench's fused episode (abr_env_run: reset + 48 chunks + statistics, 65 536 sessions, L2
flushed in front of every launch) over many launches — the A/B number for kernel variants (the event timer's granularity
is ~1 us, a single launch says little).  usage: python profiles/time_rollout.py [launches] [param=value ...]"""
import sys

import torch

sys.path.insert(0, ".")
from abrsimulator_b200 import synth, _lib
from abrsimulator_b200.env import BatchedABREnv

V = 48
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 and "=" not in sys.argv[1] else 200
params = {a.split("=")[0]: float(a.split("=")[1]) for a in sys.argv[1:] if "=" in a}
steps = int(params.pop("steps", 48))          # episode length of the timed launch (steps=1: the launch's fixed cost)
want_stats = not params.pop("nostats", 0)     # nostats=1: no statistics reduction inside the kernel
n_sessions = int(params.pop("sessions", 65536))
bitrates, sizes = synth.make_video(V)
bw, tl, ti = synth.make_traces(1024, 2048)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
N = n_sessions
env = BatchedABREnv(bw, sizes, bitrates, N, trace_len=tl, trace_interval=ti, **params)
tid, off = synth.make_sessions(N, 1024, 2048, group=64)
tid_d, off_d = torch.from_numpy(tid).to(dev), torch.from_numpy(off).to(dev)
out = {k: torch.empty(steps, N, dtype=torch.float64, device=dev) for k in ("delay", "sleep", "buffer", "rebuffer", "reward")}
out["end_of_video"] = torch.empty(steps, N, dtype=torch.uint8, device=dev)
stats = torch.empty(_lib.NUM_STATS, dtype=torch.float64, device=dev)
ev = []
for it in range(n_iter + 10):
    flush.fill_(1)
    flush.fill_(2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.run("random", steps, tid_d, off_d, seed=7, out=out, qoe_cost=False, stats=stats if want_stats else False)
    e1.record()
    ev.append((e0, e1))
torch.cuda.synchronize()
ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev[10:])
print(f"steps={steps} stats={int(want_stats)} sessions={N}: {n_iter} launches: mean {1e3 * sum(ms) / len(ms):.2f} us, median {1e3 * ms[len(ms) // 2]:.2f} us, "
      f"p10 {1e3 * ms[len(ms) // 10]:.2f} us, p90 {1e3 * ms[9 * len(ms) // 10]:.2f} us   reward sum {float(out['reward'].sum()):.6f}")
