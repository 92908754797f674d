"""Data model of the reference, in both of its vocabularies.

``Simulator.py:4-42`` and ``mpc_test.py:13-37`` define the same attribute bags with
different fields and constructor signatures.  The classes here
accept either call form and expose both sets of attribute names, so callers of
either file can switch to this package unchanged.
"""
from __future__ import annotations

import numpy as np


class Chunk:
    """``Chunk(bitrates)`` (Simulator.py:4-6) or ``Chunk(bitrates, sizes)`` (mpc_test.py:13-16)."""

    def __init__(self, bitrates, sizes=None):
        self.bitrates = bitrates
        self.sizes = sizes


class MPD:
    """``MPD(video_length, chunk_length, max_buffer, start_up_length, chunks)`` (Simulator.py:11-17)
    or ``MPD(video_length, chunk_length, max_buffer, chunks)`` (mpc_test.py:18-23)."""

    def __init__(self, video_length, chunk_length, max_buffer, *rest, start_up_length=None, chunks=None):
        if len(rest) == 2:
            start_up_length, chunks = rest
        elif len(rest) == 1:
            chunks = rest[0]
        elif len(rest) != 0:
            raise TypeError("MPD takes (video_length, chunk_length, max_buffer, [start_up_length,] chunks)")
        self.video_length = video_length
        self.chunk_length = chunk_length
        self.max_buffer = max_buffer
        self.start_up_length = start_up_length
        self.chunks = chunks

    def tables(self):
        """Flatten to (bitrates[V][A], sizes[V][A]) float64; sizes default to bitrate·chunk_length
        (``target_size``, Simulator.py:156)."""
        if not self.chunks:
            raise ValueError("MPD has no chunks")
        bitrates = np.array([list(c.bitrates) for c in self.chunks], dtype=np.float64)
        if bitrates.ndim != 2:
            raise ValueError("every chunk must offer the same number of bitrates")
        if all(getattr(c, "sizes", None) is not None for c in self.chunks):
            sizes = np.array([list(c.sizes) for c in self.chunks], dtype=np.float64)
        else:
            sizes = bitrates * float(self.chunk_length)
        if sizes.shape != bitrates.shape:
            raise ValueError("sizes and bitrates must have the same shape")
        return bitrates, sizes


class QOEMetric:
    """``QOEMetric(rebuffer_weight, variance_weight, startup_weight[, latency_weight])``
    (Simulator.py:19-24 / mpc_test.py:25-29)."""

    def __init__(self, rebuffer_weight, variance_weight, startup_weight, latency_weight=0.0):
        self.rebuffer_weight = rebuffer_weight
        self.variance_weight = variance_weight
        self.startup_weight = startup_weight
        self.latency_weight = latency_weight


class ChunkInfo:
    """``ChunkInfo(chunk_id, previous_bitrates, previous_bandwidths, buffer_level)`` (Simulator.py:30-35)
    or ``ChunkInfo(chunk_number, previous_bitrate, previous_bandwidths, buffer_level)`` (mpc_test.py:31-37).
    The second argument may be a scalar index or a list of indices."""

    def __init__(self, chunk, previous, previous_bandwidths, buffer_level):
        self.chunk_id = self.chunk_number = chunk
        if isinstance(previous, (list, tuple, np.ndarray)):
            self.previous_bitrates = previous
            self.previous_bitrate = previous[-1] if len(previous) else None
        else:
            self.previous_bitrate = previous
            self.previous_bitrates = [previous]
        self.previous_bandwidths = previous_bandwidths
        self.buffer_level = buffer_level


class NetworkInfo:
    """Square-wave throughput: ``bandwidths[i]`` holds on ``[i·interval, (i+1)·interval)`` (Simulator.py:37-42)."""

    def __init__(self, interval, bandwidths):
        self.interval = interval
        self.bandwidths = bandwidths


def load_network_trace(path):
    """One float per line (Simulator.py:59-65)."""
    with open(path) as f:
        return [float(line) for line in f if line.strip()]


def load_mpd_file(path):
    """One line per chunk, whitespace-separated bitrates (the intent of Simulator.py:68-77, whose
    ``float(line.split())`` cannot run).  Returns a list of ``Chunk``."""
    chunks = []
    with open(path) as f:
        for line in f:
            if line.strip():
                chunks.append(Chunk([float(x) for x in line.split()]))
    return chunks


def pack_traces(network_infos):
    """List of NetworkInfo -> (trace_bw[n][T_max], trace_len[n], trace_interval[n]) padded with 1.0."""
    n = len(network_infos)
    if n == 0:
        raise ValueError("need at least one trace")
    t_max = max(len(ni.bandwidths) for ni in network_infos)
    bw = np.ones((n, t_max), dtype=np.float64)
    length = np.empty(n, dtype=np.int32)
    interval = np.empty(n, dtype=np.float64)
    for i, ni in enumerate(network_infos):
        b = np.asarray(ni.bandwidths, dtype=np.float64)
        bw[i, :len(b)] = b
        length[i] = len(b)
        interval[i] = float(ni.interval)
    return bw, length, interval
