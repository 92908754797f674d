#!/usr/bin/env python
"""RL rollout harness (BASELINE.json configs[4]): a batched torch policy <-> GPU environment step with every
state tensor kept on the device.

    python examples/rl_harness.py --sessions 524288 --chunks 48

The observation is built on the device from the step outputs (buffer, last throughput, last delay, next-chunk
sizes, last action); the policy is a small MLP; the action goes straight back into `env.step`.  No host
round trip happens inside the episode loop.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from abrsimulator_b200 import synth                      # noqa: E402
from abrsimulator_b200.env import BatchedABREnv, StepResult   # noqa: E402


class Policy(torch.nn.Module):
    def __init__(self, n_obs, n_actions, hidden=64):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Linear(n_obs, hidden), torch.nn.ReLU(),
                                       torch.nn.Linear(hidden, hidden), torch.nn.ReLU(),
                                       torch.nn.Linear(hidden, n_actions))

    def forward(self, obs):
        return self.net(obs)


def run_episode(env, policy, chunks, sample=True, out=None):
    """One episode of `chunks` steps for all sessions; returns (sum of rewards [N], steps)."""
    n, A, dev = env.n, env.A, env.device
    out = out or StepResult(*[torch.empty(n, dtype=torch.float64, device=dev) for _ in range(5)],
                            torch.empty(n, A, dtype=torch.float64, device=dev),
                            torch.empty(n, dtype=torch.uint8, device=dev),
                            torch.empty(n, dtype=torch.float64, device=dev))
    obs = torch.zeros(n, 4 + A, dtype=torch.float32, device=dev)
    obs[:, 4:] = env.state("sizes")[0].float()          # sizes of the first chunk
    total = torch.zeros(n, dtype=torch.float64, device=dev)
    action = torch.full((n,), 1, dtype=torch.int32, device=dev)
    with torch.no_grad():
        for _ in range(chunks):
            logits = policy(obs)
            action = (torch.distributions.Categorical(logits=logits).sample() if sample
                      else logits.argmax(dim=1)).to(torch.int32)
            r = env.step(action, out=out, want_throughput=True)
            total += r.reward
            obs[:, 0] = (r.buffer / 10.0).float()
            obs[:, 1] = r.throughput.float()
            obs[:, 2] = (r.delay / 10.0).float()
            obs[:, 3] = action.float() / A
            obs[:, 4:] = r.next_sizes.float()
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sessions", type=int, default=524288)
    ap.add_argument("--chunks", type=int, default=48)
    ap.add_argument("--episodes", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    bitrates, sizes = synth.make_video(args.chunks)
    bw, tl, ti = synth.make_traces(1024, 2048)
    env = BatchedABREnv(bw, sizes, bitrates, args.sessions, trace_len=tl, trace_interval=ti)
    # sessions sorted by trace: every 256-session tile of the step kernel stages one capacity row in shared memory
    tid, off = synth.make_sessions(args.sessions, 1024, 2048, group=max(256, args.sessions // 1024))
    policy = Policy(4 + env.A, env.A).to(dev)
    for ep in range(args.episodes):
        env.reset(tid, off)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        total = run_episode(env, policy, args.chunks)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"episode {ep}: {args.sessions * args.chunks / dt:.3e} env steps/s "
              f"(policy + step, {args.sessions} sessions x {args.chunks} chunks in {dt * 1e3:.1f} ms), "
              f"mean episode reward {total.mean().item():.3f}")


if __name__ == "__main__":
    main()
