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
    obs = torch.zeros(N, 4 + env.A, dtype=torch.float32, device="cuda")
    obs[:, 4:] = env.state("sizes")[0].float()
    tot_ref = np.zeros(N)
    with torch.no_grad():
        for _ in range(V):
            a = policy(obs).argmax(dim=1).to(torch.int32)
            r = env.step(a, want_throughput=True)
            e = ref.step(a.cpu().numpy())
            tot_ref = tot_ref + e["reward"]
            obs[:, 0] = (r.buffer / 10.0).float()
            obs[:, 1] = r.throughput.float()
            obs[:, 2] = (r.delay / 10.0).float()
            obs[:, 3] = a.float() / env.A
            obs[:, 4:] = r.next_sizes.float()
    np.testing.assert_allclose(total.cpu().numpy(), tot_ref, rtol=1e-9, atol=1e-9)
