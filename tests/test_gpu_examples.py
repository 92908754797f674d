"""The RL rollout harness (BASELINE.json configs[4]) keeps working: policy <-> env.step with device-resident state."""
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def test_rl_harness_episode_matches_manual_stepping():
    import numpy as np
    from abrsimulator_b200 import synth
    from abrsimulator_b200.env import BatchedABREnv
    from examples.rl_harness import Policy, run_episode
    from oracle import oracle as orc

    N, V = 2048, 12
    bitrates, sizes = synth.make_video(V)
    bw, tl, ti = synth.make_traces(16, 128)
    tid, off = synth.make_sessions(N, 16, 128, group=64)
    env = BatchedABREnv(bw, sizes, bitrates, N, trace_len=tl, trace_interval=ti)
    env.reset(tid, off)
    torch.manual_seed(0)
    policy = Policy(4 + env.A, env.A).cuda()
    total = run_episode(env, policy, V, sample=False)
    assert total.shape == (N,) and bool(torch.isfinite(total).all())
    assert env.error_count() == 0
    # greedy policy is deterministic: replaying the same actions through the oracle gives the same return
    env.reset(tid, off)
    ref = orc.OracleEnv(bw, tl, ti, sizes, bitrates, N)
    ref.reset(tid, off)
    obs = torch.zeros(4 + env.A, N, dtype=torch.float32, device="cuda")      # feature-major, like the harness
    obs[4:] = env.state("sizes")[0].float()[:, None]
    tot_ref = np.zeros(N)
    with torch.no_grad():
        for _ in range(V):
            a = policy(obs.t()).argmax(dim=1).to(torch.int32)
            r = env.step(a, want_throughput=True)
            e = ref.step(a.cpu().numpy())
            tot_ref = tot_ref + e["reward"]
            obs[0] = r.buffer * 0.1
            obs[1] = r.throughput
            obs[2] = r.delay * 0.1
            obs[3] = a / float(env.A)
            obs[4:] = r.next_sizes.t()
    np.testing.assert_allclose(total.cpu().numpy(), tot_ref, rtol=1e-9, atol=1e-9)


def test_rl_harness_cuda_graph_replay_equals_eager():
    """One chunk (policy kernels + abr_env_step + observation update) captured into a CUDA graph and replayed gives
    the same episode as launching every kernel separately (greedy policy: deterministic)."""
    import numpy as np
    from abrsimulator_b200 import synth
    from abrsimulator_b200.env import BatchedABREnv
    from examples.rl_harness import Policy, run_episode

    N, V = 4096, 20
    bitrates, sizes = synth.make_video(V)
    bw, tl, ti = synth.make_traces(16, 128)
    tid, off = synth.make_sessions(N, 16, 128, group=256)
    env = BatchedABREnv(bw, sizes, bitrates, N, trace_len=tl, trace_interval=ti)
    torch.manual_seed(0)
    policy = Policy(4 + env.A, env.A).cuda()
    env.reset(tid, off)
    eager = run_episode(env, policy, V, sample=False).cpu().numpy()
    state_eager = {f: env.state(f).cpu().numpy().copy() for f in ("seg", "chunk", "phase", "pos", "buffer")}
    env.reset(tid, off)
    graphed = run_episode(env, policy, V, sample=False, use_graph=True).cpu().numpy()
    assert np.array_equal(eager, graphed)
    for f, want in state_eager.items():
        assert np.array_equal(env.state(f).cpu().numpy(), want), f
    # sampling runs under capture too (graph-safe Philox offsets) and stays finite
    env.reset(tid, off)
    sampled = run_episode(env, policy, V, sample=True, use_graph=True)
    assert bool(torch.isfinite(sampled).all()) and env.error_count() == 0


def test_rl_harness_fused_step_policy_equals_torch_glue():
    """abr_env_step_policy (sampling / arg max, step, reward sum and observation in the step kernel) gives the same
    greedy episode, bit for bit, as env.step with the arg max and the observation built by eager PyTorch."""
    import numpy as np
    from abrsimulator_b200 import synth
    from abrsimulator_b200.env import BatchedABREnv
    from examples.rl_harness import Policy, run_episode

    N, V = 4096, 30
    bitrates, sizes = synth.make_video(20)                 # the episode crosses an end of video
    bw, tl, ti = synth.make_traces(16, 128)
    tid, off = synth.make_sessions(N, 16, 128, group=256)
    env = BatchedABREnv(bw, sizes, bitrates, N, trace_len=tl, trace_interval=ti)
    torch.manual_seed(0)
    policy = Policy(4 + env.A, env.A).cuda()
    env.reset(tid, off)
    glue = run_episode(env, policy, V, sample=False, fused=False).cpu().numpy()
    state = {f: env.state(f).cpu().numpy().copy() for f in ("seg", "chunk", "last_q", "phase", "pos", "buffer")}
    env.reset(tid, off)
    fused = run_episode(env, policy, V, sample=False, fused=True).cpu().numpy()
    assert np.array_equal(glue, fused)
    for f, want in state.items():
        assert np.array_equal(env.state(f).cpu().numpy(), want), f
    assert env.error_count() == 0


def test_mpc_dropin_example_prints_the_reference_answer(capsys):
    """examples/mpc_dropin.py: the reference's mpc_test.py scenario -> 'Test next bitrate: 2'."""
    from examples.mpc_dropin import main
    assert main() == 2
    assert capsys.readouterr().out.strip() == "Test next bitrate: 2"
