#!/usr/bin/env python
"""RL rollout harness (BASELINE.json configs[4]): a batched torch policy <-> GPU environment step with every
state tensor kept on the device.

    python examples/rl_harness.py --sessions 524288 --chunks 48

The policy is a small MLP; its logits go straight into `env.step_policy` (abr_env_step_policy, SPEC §4.1): the step
kernel draws the action (Gumbel-max over Philox noise), steps the session, adds up the reward and writes the next
observation (buffer, last throughput, last delay, last action, next-chunk sizes) as the feature-major fp32 matrix the
first Linear reads — one library launch per chunk besides the policy's own kernels.  `--torch-glue` keeps the
sampling and the observation in eager PyTorch around `env.step` (twelve more kernels per chunk) for comparison.
No host round trip happens inside the episode loop.
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


def _step_outputs(env):
    n, A, dev = env.n, env.A, env.device
    return StepResult(*[torch.empty(n, dtype=torch.float64, device=dev) for _ in range(5)],
                      torch.empty(n, A, dtype=torch.float64, device=dev),
                      torch.empty(n, dtype=torch.uint8, device=dev),
                      torch.empty(n, dtype=torch.float64, device=dev), None)


OBS_SCALES = (0.1, 1.0, 0.1, 1.0)      # buffer / 10, throughput, delay / 10, sizes


def _one_chunk(env, policy, obs, action, total, out, sample, fused=True):
    """policy(obs) -> action -> env.step -> next observation, all on the device and all in place (so that the
    sequence can be captured into a CUDA graph once and replayed for every chunk).  ``obs`` is feature-major
    [4 + A, N]: every feature row is written coalesced and the first Linear reads it transposed."""
    A = env.A
    logits = policy(obs.t())
    if fused:       # sampling, step, reward sum and observation in the step kernel
        env.step_policy(logits, sample=sample, seed=0x5EED, obs=obs, action_out=action, reward_sum=total,
                        obs_scales=OBS_SCALES)
        return
    if sample:      # Gumbel-max: argmax(logits + G) is a draw from softmax(logits); three elementwise kernels
        u = torch.rand_like(logits).clamp_(1e-12, 1.0)
        logits = logits - torch.log(-torch.log(u))
    action.copy_(logits.argmax(dim=1))
    r = env.step(action, out=out, want_throughput=True)
    total += r.reward
    obs[0] = r.buffer * 0.1
    obs[1] = r.throughput
    obs[2] = r.delay * 0.1
    obs[3] = action / float(A)
    obs[4:] = r.next_sizes.t()


def _first_observation(env):
    obs = torch.zeros(4 + env.A, env.n, dtype=torch.float32, device=env.device)
    obs[4:] = env.state("sizes")[0].float()[:, None]     # sizes of the first chunk
    return obs


class GraphedEpisode:
    """One chunk (policy kernels + abr_env_step + observation update) captured into a CUDA graph once; an episode is
    `chunks` replays.  Buffers are persistent, so the capture is reused across episodes (call after ``env.reset``)."""

    def __init__(self, env, policy, sample=True, fused=True):
        self.env, self.policy, self.sample, self.fused = env, policy, sample, fused
        n, dev = env.n, env.device
        self.out = _step_outputs(env)
        self.obs = _first_observation(env)
        self.total = torch.zeros(n, dtype=torch.float64, device=dev)
        self.action = torch.full((n,), 1, dtype=torch.int32, device=dev)
        self.graph = None

    def _capture(self):
        env = self.env
        # warm up the allocator and cuBLAS on a side stream, restore the state, then capture (capture runs nothing)
        snap = {f: env.state(f).clone() for f in ("seg", "chunk", "last_q", "phase", "pos", "buffer")}
        side = torch.cuda.Stream(device=env.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            _one_chunk(env, self.policy, self.obs, self.action, self.total, self.out, self.sample, self.fused)
        torch.cuda.current_stream().wait_stream(side)
        for f, t in snap.items():
            env.state(f).copy_(t)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            _one_chunk(env, self.policy, self.obs, self.action, self.total, self.out, self.sample, self.fused)

    def run(self, chunks):
        with torch.no_grad():
            if self.graph is None:
                self._capture()
            self.obs.copy_(_first_observation(self.env))
            self.total.zero_()
            for _ in range(chunks):
                self.graph.replay()
        return self.total


def run_episode(env, policy, chunks, sample=True, out=None, use_graph=False, fused=True):
    """One episode of `chunks` steps for all sessions; returns the sum of rewards [N] (``use_graph``: through a
    freshly captured ``GraphedEpisode``; keep one around to reuse the capture across episodes)."""
    if use_graph:
        return GraphedEpisode(env, policy, sample, fused).run(chunks)
    n, dev = env.n, env.device
    out = out or _step_outputs(env)
    obs = _first_observation(env)
    total = torch.zeros(n, dtype=torch.float64, device=dev)
    action = torch.full((n,), 1, dtype=torch.int32, device=dev)
    with torch.no_grad():
        for _ in range(chunks):
            _one_chunk(env, policy, obs, action, total, out, sample, fused)
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sessions", type=int, default=524288)
    ap.add_argument("--chunks", type=int, default=48)
    ap.add_argument("--episodes", type=int, default=3)
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel of every chunk separately")
    ap.add_argument("--torch-glue", action="store_true", help="sampling and observation in eager PyTorch around env.step")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = True        # the stand-in policy may use the tensor cores
    bitrates, sizes = synth.make_video(args.chunks)
    bw, tl, ti = synth.make_traces(1024, 2048)
    env = BatchedABREnv(bw, sizes, bitrates, args.sessions, trace_len=tl, trace_interval=ti)
    # sessions sorted by trace: every 256-session tile of the step kernel stages one capacity row in shared memory
    tid, off = synth.make_sessions(args.sessions, 1024, 2048, group=max(256, args.sessions // 1024))
    policy = Policy(4 + env.A, env.A).to(dev)
    runner = None
    for ep in range(args.episodes):
        env.reset(tid, off)
        if runner is None and not args.no_graph:
            runner = GraphedEpisode(env, policy, fused=not args.torch_glue)      # buffers are sized by the reset; captured on first use
            runner.run(1)
            env.reset(tid, off)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        total = runner.run(args.chunks) if runner else run_episode(env, policy, args.chunks, fused=not args.torch_glue)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"episode {ep}: {args.sessions * args.chunks / dt:.3e} env steps/s "
              f"(policy + step, {args.sessions} sessions x {args.chunks} chunks in {dt * 1e3:.1f} ms), "
              f"mean episode reward {total.mean().item():.3f}")


if __name__ == "__main__":
    main()
