"""``Simulator`` façade: the reference's environment API (``Simulator.py:45-210``) over the GPU kernels.

``Simulator(AbrController, SpeedController)``, ``set_qoe_metric``, ``set_network_info``, ``set_mpd`` and ``run()``
keep their reference signatures.  ``run()`` plays one session (or ``run_batch`` many) at chunk granularity
(SPEC.md §3) and returns the QoE *cost* of ``calculate_qoe`` (``Simulator.py:79-86``):
``rw·rebuffer_time + vw·Σ|Δbitrate| + sw·start_up_time + lw·average_latency`` — in this version start-up time and
latency are not modelled and contribute 0.

Controllers
-----------
* any object with ``get_next_bitrate(chunk_id, previous_bitrates, previous_bandwidths, buffer_level) -> int``
  (the protocol ``Simulator.run`` calls, ``Simulator.py:155``) is driven chunk by chunk through the step kernel;
* the markers ``RandomPolicy(seed)`` / ``BufferBasedPolicy()`` / ``FixedPolicy(actions)`` and this package's
  ``MPCBitrateController`` run fully on the device (fused episode kernel, or decide+step kernels for MPC).
"""
from __future__ import annotations

import numpy as np
import torch

from .datamodel import MPD, NetworkInfo, QOEMetric, load_mpd_file, load_network_trace
from .env import BatchedABREnv
from .mpc import MPCBitrateController
from ._lib import MPC_ROBUST


class RandomPolicy:
    def __init__(self, seed=0):
        self.seed = int(seed)


class BufferBasedPolicy:
    def __init__(self, reservoir=5.0, cushion=10.0):
        self.reservoir, self.cushion = float(reservoir), float(cushion)


class FixedPolicy:
    def __init__(self, actions):
        self.actions = np.asarray(actions, dtype=np.int32)


class Simulator:
    def __init__(self, AbrController=None, SpeedController=None, **params):
        self.qoe_metric = None
        self.mpd = None
        self.network_info = None
        self.abr_controller = AbrController
        self.speed_controller = SpeedController      # playback-speed hook: not modelled yet (always 1x)
        self.params = dict(params)
        self._env = None
        self.last_run = None

    # -- setters (Simulator.py:54-77) --
    def set_qoe_metric(self, qoe_metric):
        self.qoe_metric = qoe_metric
        self._env = None

    def set_network_info(self, interval, networktrace):
        """``networktrace``: path of a one-float-per-line file (Simulator.py:59-65), a sequence of
        bandwidths, a NetworkInfo, or a list of NetworkInfo (one trace per session group)."""
        if isinstance(networktrace, (str, bytes)):
            self.network_info = NetworkInfo(interval, load_network_trace(networktrace))
        elif isinstance(networktrace, NetworkInfo):
            self.network_info = networktrace
        elif len(networktrace) and isinstance(networktrace[0], NetworkInfo):
            self.network_info = list(networktrace)
        else:
            self.network_info = NetworkInfo(interval, list(networktrace))
        self._env = None

    def set_mpd(self, chunk_length, max_buffer, start_up_length, mpdfile):
        """``mpdfile``: path (one line of bitrates per chunk), or a list of ``Chunk``."""
        chunks = load_mpd_file(mpdfile) if isinstance(mpdfile, (str, bytes)) else list(mpdfile)
        self.mpd = MPD(len(chunks), chunk_length, max_buffer, start_up_length, chunks)
        self._env = None

    def get_mpd(self):
        return self.mpd

    def get_qoe_metric(self):
        return self.qoe_metric

    def calculate_qoe(self, rebuffer_time, previous_bitrates, start_up_time, average_latency):
        """Simulator.py:79-86 with the indexing repaired (D3): ladder of each chunk, consecutive indices."""
        q = self.qoe_metric
        variance = 0.0
        for i in range(0, len(previous_bitrates) - 1):
            variance += abs(self.mpd.chunks[i].bitrates[previous_bitrates[i]] -
                            self.mpd.chunks[i + 1].bitrates[previous_bitrates[i + 1]])
        return (q.rebuffer_weight * rebuffer_time + q.variance_weight * variance +
                q.startup_weight * start_up_time + getattr(q, "latency_weight", 0.0) * average_latency)

    # -- environment --
    def _make_env(self, n_sessions, **extra):
        if self.mpd is None or self.network_info is None or self.qoe_metric is None:
            raise RuntimeError("set_qoe_metric, set_network_info and set_mpd must be called before run()")
        kw = dict(utility_scale=1.0, default_quality=-1, auto_reset=0)   # cost is in bitrate units, first chunk free
        kw.update(self.params)
        kw.update(extra)
        return BatchedABREnv.from_objects(self.network_info, self.mpd, self.qoe_metric, max_sessions=n_sessions, **kw)

    def run(self):
        """One session over trace 0 from offset 0; returns the scalar QoE cost."""
        return float(self.run_batch(1)[0])

    def run_batch(self, n_sessions, trace_id=None, start_offset=None, session_base=0):
        """``n_sessions`` independent sessions; returns a numpy vector of QoE costs (one per session)."""
        V = len(self.mpd.chunks)
        n_traces = len(self.network_info) if isinstance(self.network_info, list) else 1
        tid = np.arange(n_sessions, dtype=np.int32) % n_traces if trace_id is None else np.asarray(trace_id, np.int32)
        ctrl = self.abr_controller
        q = self.qoe_metric
        if isinstance(ctrl, (RandomPolicy, BufferBasedPolicy, FixedPolicy)) or ctrl is None:
            extra = {}
            if isinstance(ctrl, BufferBasedPolicy):
                extra = dict(bba_reservoir=ctrl.reservoir, bba_cushion=ctrl.cushion)
            env = self._make_env(n_sessions, **extra)
            policy = "random" if isinstance(ctrl, RandomPolicy) else "fixed" if isinstance(ctrl, FixedPolicy) else "bba"
            out = env.run_host(policy, V, tid, start_offset, seed=getattr(ctrl, "seed", 0), session_base=session_base,
                               actions=getattr(ctrl, "actions", None), want_qoe_cost=True)
            acc = out["acc"]
            self.last_run = dict(rebuffer=acc[1], smooth=acc[3], utility=acc[2], reward=acc[0], sleep=acc[4],
                                 delay=acc[5])
            return out["qoe_cost"]          # rw*rebuffer + vw*smooth, computed on the device
        elif isinstance(ctrl, MPCBitrateController):
            env = self._make_env(n_sessions, track_history=1, track_acc=1, **ctrl.extra_params)
            env.reset(tid, start_offset, session_base)
            env.mpc_episode(V, horizon=ctrl.horizon, mode="robust" if ctrl.mode == MPC_ROBUST else "reference")
            acc = env.session_acc().cpu().numpy()
        else:
            acc = self._run_callback(n_sessions, tid, start_offset)
        self.last_run = dict(rebuffer=acc[1], smooth=acc[3], utility=acc[2], reward=acc[0], sleep=acc[4], delay=acc[5])
        return q.rebuffer_weight * acc[1] + q.variance_weight * acc[3]

    def _run_callback(self, n_sessions, tid, start_offset):
        """Generic controller protocol: one ``get_next_bitrate`` call per session and chunk (Simulator.py:155)."""
        V = len(self.mpd.chunks)
        env = self._make_env(n_sessions, track_acc=1)
        env.reset(tid, start_offset)
        prev_q = [[] for _ in range(n_sessions)]
        prev_bw = [[] for _ in range(n_sessions)]
        buf = np.zeros(n_sessions)
        for k in range(V):
            acts = np.array([self.abr_controller.get_next_bitrate(k, prev_q[s], prev_bw[s], float(buf[s]))
                             for s in range(n_sessions)], dtype=np.int32)
            r = env.step(torch.from_numpy(acts), want_next_sizes=False, want_throughput=True)
            buf = r.buffer.cpu().numpy()
            thr = r.throughput.cpu().numpy()
            for s in range(n_sessions):
                prev_q[s].append(int(acts[s]))
                prev_bw[s].append(float(thr[s]))
        return env.session_acc().cpu().numpy()
