"""Time the pieces of one RL-harness chunk (policy, sampling, abr_env_step, observation update) with CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abrsimulator_b200 import synth
from abrsimulator_b200.env import BatchedABREnv
from examples.rl_harness import Policy, _step_outputs, _first_observation

N, V = 524288, 48
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
bitrates, sizes = synth.make_video(V)
bw, tl, ti = synth.make_traces(1024, 2048)
env = BatchedABREnv(bw, sizes, bitrates, N, trace_len=tl, trace_interval=ti)
tid, off = synth.make_sessions(N, 1024, 2048, group=512)
env.reset(tid, off)
policy = Policy(4 + env.A, env.A).to(dev)
out = _step_outputs(env)
obs = _first_observation(env)
obs_rm = obs.t().contiguous()
action = torch.ones(N, dtype=torch.int32, device=dev)


def timeit(name, fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / reps * 1e3:9.1f} us")


with torch.no_grad():
    timeit("policy(obs.t()) feature-major", lambda: policy(obs.t()))
    timeit("policy(obs_rm) row-major", lambda: policy(obs_rm))
    logits = policy(obs_rm)
    timeit("gumbel + argmax", lambda: action.copy_((logits - torch.log(-torch.log(torch.rand_like(logits).clamp_(1e-12, 1.0)))).argmax(dim=1)))
    timeit("argmax only", lambda: action.copy_(logits.argmax(dim=1)))
    timeit("env.step (throughput + next_sizes)", lambda: env.step(action, out=out, want_throughput=True))
    r = out

    def upd_fm():
        obs[0] = r.buffer / 10.0; obs[1] = r.throughput; obs[2] = r.delay / 10.0; obs[3] = action / 6.0; obs[4:] = r.next_sizes.t()

    def upd_rm():
        obs_rm[:, 0] = r.buffer / 10.0; obs_rm[:, 1] = r.throughput; obs_rm[:, 2] = r.delay / 10.0
        obs_rm[:, 3] = action / 6.0; obs_rm[:, 4:] = r.next_sizes
    timeit("obs update feature-major", upd_fm)
    timeit("obs update row-major", upd_rm)
