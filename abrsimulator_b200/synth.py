"""Synthetic FCC/HSDPA-shaped workloads — numpy only, deterministic.

Units: bandwidth in Mbit/s, sizes in Mbit, bitrates in kbit/s (utility_scale 0.001 -> Mbit/s utility).
"""
from __future__ import annotations

import numpy as np

LADDER_KBPS = (300.0, 750.0, 1200.0, 1850.0, 2850.0, 4300.0)


def make_video(V=48, ladder=LADDER_KBPS, chunk_length=4.0, seed=0):
    """bitrates[V][A] (kbit/s) and VBR sizes[V][A] = bitrate/1000 · chunk_length · U(0.8, 1.2) (Mbit)."""
    rng = np.random.default_rng(seed)
    lad = np.asarray(ladder, dtype=np.float64)
    bitrates = np.tile(lad, (V, 1))
    sizes = bitrates / 1000.0 * chunk_length * rng.uniform(0.8, 1.2, size=bitrates.shape)
    return np.ascontiguousarray(bitrates), np.ascontiguousarray(sizes)


def make_traces(n_traces=1024, T=2048, interval=1.0, seed=1234):
    """Log-normal AR(1) bandwidth: x' = 0.9x + N(0, 0.3²), bw = clip(exp(mu + x), 0.2, 6.0) Mbit/s,
    mu ~ U(ln 0.5, ln 3) per trace.  Returns (trace_bw[n][T], trace_len[n], trace_interval[n])."""
    rng = np.random.default_rng(seed)
    mu = rng.uniform(np.log(0.5), np.log(3.0), size=(n_traces, 1))
    eps = rng.normal(0.0, 0.3, size=(n_traces, T))
    x = np.empty((n_traces, T))
    x[:, 0] = eps[:, 0]
    for t in range(1, T):
        x[:, t] = 0.9 * x[:, t - 1] + eps[:, t]
    bw = np.clip(np.exp(mu + x), 0.2, 6.0)
    return (np.ascontiguousarray(bw), np.full(n_traces, T, np.int32), np.full(n_traces, float(interval)))


def make_sessions(n_sessions, n_traces, T, interval=1.0, seed=42, session_base=0, group=1):
    """trace_id = (global session index // group) mod n_traces; start offset ~ U(0, T·interval).
    ``group`` consecutive sessions share a trace (group = 64 lets every 64-thread block of the fused episode
    kernel stage its trace in shared memory).  Sharding-invariant: session g always gets the same draw
    regardless of which rank owns it."""
    g = np.arange(session_base, session_base + n_sessions, dtype=np.int64)
    trace_id = ((g // group) % n_traces).astype(np.int32)
    # counter-based draw so that shards agree with the unsharded run
    u = ((g * 2654435761 + seed * 40503) % 2**32).astype(np.float64) / 2**32
    return trace_id, np.ascontiguousarray(u * T * interval)
