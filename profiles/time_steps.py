"""Fixed cost vs per-step cost of the fused episode kernel: CUDA-event time of abr_env_run for several episode lengths.
usage: python profiles/time_steps.py [sessions ...] [param=value ...]   (AbrParams fields, e.g. max_buffer=1e9: no session
ever sleeps; max_buffer=0: every step sleeps)"""
import sys

import torch

sys.path.insert(0, ".")
from abrsimulator_b200 import synth
from abrsimulator_b200.env import BatchedABREnv

V = 48
bitrates, sizes = synth.make_video(V)
bw, tl, ti = synth.make_traces(1024, 2048)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
params = {a.split("=")[0]: float(a.split("=")[1]) for a in sys.argv[1:] if "=" in a}
for N in [int(x) for x in sys.argv[1:] if "=" not in x] or [4736, 65536]:
    env = BatchedABREnv(bw, sizes, bitrates, N, trace_len=tl, trace_interval=ti, **params)
    tid, off = synth.make_sessions(N, 1024, 2048, group=64)
    tid_d, off_d = torch.from_numpy(tid).to(dev), torch.from_numpy(off).to(dev)
    row = []
    for steps in (1, 2, 8, 24, 48, 96):
        out = {k: torch.empty(steps, N, dtype=torch.float64, device=dev) for k in ("delay", "sleep", "buffer", "rebuffer", "reward")}
        out["end_of_video"] = torch.empty(steps, N, dtype=torch.uint8, device=dev)
        ms = []
        for it in range(8):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            env.run("random", steps, tid_d, off_d, seed=7, out=out, qoe_cost=False, stats=False)
            e1.record()
            e1.synchronize()
            if it >= 3:
                ms.append(e0.elapsed_time(e1))
        row.append((steps, 1e3 * sorted(ms)[len(ms) // 2]))
    print(N, " ".join(f"steps={s}: {t:.1f}us" for s, t in row))
    (s0, t0), (s1, t1) = row[2], row[4]
    print(f"   per step {(t1 - t0) / (s1 - s0) * 1e3:.0f} ns = {(t1 - t0) / (s1 - s0) * 1965:.0f} cycles; fixed {t0 - s0 * (t1 - t0) / (s1 - s0):.1f} us")
