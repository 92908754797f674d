"""Mean time of the per-step kernel (abr_env_step, 4 Mi trace-sorted sessions, fp64 outputs) over back-to-back launches —
the A/B number for variants of abr_step_kernel.  usage: python profiles/time_step.py [launches] [sessions] [sessions per trace run; 1 = interleaved]"""
import sys

import torch

sys.path.insert(0, ".")
from abrsimulator_b200 import synth
from abrsimulator_b200.env import BatchedABREnv, StepResult

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 96
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
V, A = 48, 6
bitrates, sizes = synth.make_video(V)
bw, tl, ti = synth.make_traces(1024, 2048)
dev = torch.device("cuda", 0)
env = BatchedABREnv(bw, sizes, bitrates, M, trace_len=tl, trace_interval=ti)
g = torch.Generator(device=dev)
g.manual_seed(1)
acts = torch.randint(0, A, (8, M), dtype=torch.int32, device=dev, generator=g)
out = StepResult(*[torch.empty(M, dtype=torch.float64, device=dev) for _ in range(5)], None,
                 torch.empty(M, dtype=torch.uint8, device=dev), None)
group = int(sys.argv[3]) if len(sys.argv) > 3 else max(256, M // 1024)
tid, off = synth.make_sessions(M, 1024, 2048, group=group)
env.reset(tid, off)
for t in range(8):
    env.step(acts[t % 8], out=out)
res = []
for block in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(reps):
        env.step(acts[t % 8], out=out)
    e1.record()
    e1.synchronize()
    res.append(e0.elapsed_time(e1) / reps)
bytes_per = 40 + 4 + 36 + 41
print(f"{M} sessions (runs of {group}): " + " ".join(f"{1e3 * r:.2f}" for r in res) + f" us per launch; best {M * bytes_per / (min(res) * 1e-3) / 1e9:.0f} GB/s"
      f"   reward sum {float(out.reward.sum()):.6f}")
