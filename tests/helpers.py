"""Shared helpers for the parity tests."""
import numpy as np

from abrsimulator_b200 import synth


def small_world(n_traces=16, T=128, V=48, interval=1.0, seed=0, ladder=synth.LADDER_KBPS, ragged=False):
    bitrates, sizes = synth.make_video(V, ladder=ladder, seed=seed)
    bw, tl, ti = synth.make_traces(n_traces, T, interval=interval, seed=1234 + seed)
    if ragged:   # ragged trace lengths and mixed intervals
        rng = np.random.default_rng(seed + 5)
        tl = rng.integers(1, T + 1, size=n_traces).astype(np.int32)
        tl[0] = 1
        tl[-1] = T
        ti = rng.choice([0.5, 1.0, 2.0], size=n_traces)
    return bitrates, sizes, bw, tl, ti


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)
    b = np.ascontiguousarray(b, dtype=np.float64).view(np.uint64)
    return int((a != b).sum())


def assert_close(gpu, cpu, name, rtol=1e-9, atol=1e-12, exact=True):
    """BASELINE.json tolerance (1e-9 relative, fp64); additionally require bit-identity when `exact`."""
    gpu = np.asarray(gpu, dtype=np.float64)
    cpu = np.asarray(cpu, dtype=np.float64)
    np.testing.assert_allclose(gpu, cpu, rtol=rtol, atol=atol, err_msg=name)
    if exact:
        nd = bits_equal(gpu, cpu)
        assert nd == 0, f"{name}: {nd} of {gpu.size} values are within 1e-9 but not bit-identical"
