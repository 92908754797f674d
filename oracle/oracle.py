"""ctypes binding of the C oracle (oracle/abr_oracle.c) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.
Arrays are numpy, contiguous; layouts follow abr_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libabr_oracle.so")

NUM_STATS = 11
NUM_ACC = 11
POLICY_FIXED, POLICY_RANDOM, POLICY_BBA = 0, 1, 2


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "chunk_length", "max_buffer", "rtt", "payload", "sleep_quantum", "rebuf_penalty",
        "smooth_penalty", "utility_scale", "bba_reservoir", "bba_cushion", "start_up_length",
        "startup_penalty", "latency_penalty", "latency_tick")] + [
        (n, C.c_int32) for n in ("utility_mode", "default_quality", "auto_reset", "hist_k",
                                 "track_history", "reserved0", "live", "smooth_prev_ladder")]


DEFAULTS = dict(chunk_length=4.0, max_buffer=60.0, rtt=0.08, payload=0.95, sleep_quantum=0.5,
                rebuf_penalty=4.3, smooth_penalty=1.0, utility_scale=0.001, bba_reservoir=5.0,
                bba_cushion=10.0, start_up_length=0.0, startup_penalty=0.0, latency_penalty=0.0, latency_tick=0.01, utility_mode=0,
                default_quality=1, auto_reset=1, hist_k=5, track_history=0, live=0, smooth_prev_ladder=0)


def make_params(**kw) -> OrcParams:
    d = dict(DEFAULTS)
    d.update(kw)
    p = OrcParams()
    for k, v in d.items():
        setattr(p, k, v)
    return p


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("abr_oracle.c", "abr_oracle.h")]
    if (not force and os.path.exists(_SO) and
            (not all(os.path.exists(s) for s in src) or
             os.path.getmtime(_SO) >= max(os.path.getmtime(s) for s in src))):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_env_create.restype = C.c_void_p
        _lib.orc_env_field.restype = C.c_void_p
        _lib.orc_env_error_count.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def utility_table(bitrates, mode=0, scale=1.0):
    b = _f64(bitrates)
    out = np.empty_like(b)
    lib().orc_utility_table(_p(b), C.c_int(b.shape[0]), C.c_int(b.shape[1]), C.c_int(mode), C.c_double(scale), _p(out))
    return out


def philox(c0, c1, c2, c3, k0, k1):
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(c2), C.c_uint32(c3), C.c_uint32(k0),
                            C.c_uint32(k1), out)
    return list(out)


class OracleEnv:
    """N independent sessions, SPEC §2-§4, scalar C loops."""

    def __init__(self, trace_bw, trace_len, trace_interval, sizes, bitrates, N, **params):
        self.trace_bw = _f64(trace_bw)
        self.n_traces, self.T_max = self.trace_bw.shape
        self.trace_len = _i32(trace_len)
        self.trace_interval = _f64(trace_interval)
        self.sizes = _f64(sizes)
        self.bitrates = _f64(bitrates)
        self.V, self.A = self.sizes.shape
        self.N = int(N)
        self.params = make_params(**params)
        self.K = max(1, self.params.hist_k)
        self._h = C.c_void_p(lib().orc_env_create(
            _p(self.trace_bw), _p(self.trace_len), _p(self.trace_interval), C.c_int(self.n_traces),
            C.c_int(self.T_max), _p(self.sizes), _p(self.bitrates), C.c_int(self.V), C.c_int(self.A),
            C.byref(self.params), C.c_int(self.N)))

    def __del__(self):
        try:
            if self._h:
                lib().orc_env_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def reset(self, trace_id, start_offset=None):
        t = _i32(trace_id)
        o = None if start_offset is None else _f64(start_offset)
        lib().orc_env_reset(self._h, _p(t), _p(o))

    def step(self, action, want_next_sizes=True, speed=None, acc=None):
        """One chunk step (SPEC §3, or §7 when live = 1).  ``speed``: [V, N] playback-speed table (live mode);
        ``acc``: [NUM_ACC, N] accumulator table to add this step into."""
        N, A = self.N, self.A
        a = _i32(action)
        v = None if speed is None else _f64(speed)
        out = dict(delay=np.empty(N), sleep=np.empty(N), buffer=np.empty(N), rebuf=np.empty(N),
                   reward=np.empty(N), latency=np.empty(N), eov=np.empty(N, np.uint8), throughput=np.empty(N),
                   next_sizes=np.empty((N, A)) if want_next_sizes else None)
        lib().orc_env_step_live(self._h, _p(a), _p(v), _p(out["delay"]), _p(out["sleep"]), _p(out["buffer"]),
                                _p(out["rebuf"]), _p(out["reward"]), _p(out["latency"]), _p(out["next_sizes"]),
                                _p(out["eov"]), _p(out["throughput"]), _p(acc))
        return out

    def rollout(self, policy, steps, seed=0, session_base=0, actions=None, want_traj=True, speed=None):
        """Fused-episode semantics (SPEC §3+§4; §7 with the ``speed`` table [V, N] when live = 1)."""
        N = self.N
        a_in = None if actions is None else _i32(actions)
        v = None if speed is None else _f64(speed)
        tr = dict(acc=np.empty((NUM_ACC, N)))
        if want_traj:
            for k in ("delay", "sleep", "buffer", "rebuf", "reward", "latency"):
                tr[k] = np.empty((steps, N))
            tr["eov"] = np.empty((steps, N), np.uint8)
            tr["actions"] = np.empty((steps, N), np.int32)
        lib().orc_env_rollout_live(self._h, C.c_int(policy), C.c_uint64(seed), C.c_int64(session_base), C.c_int(steps),
                                   _p(a_in), _p(v), _p(tr.get("delay")), _p(tr.get("sleep")), _p(tr.get("buffer")),
                                   _p(tr.get("rebuf")), _p(tr.get("reward")), _p(tr.get("latency")), _p(tr.get("eov")),
                                   _p(tr.get("actions")), _p(tr["acc"]))
        return tr

    def mpc_decide(self, H, mode, want_seq=False):
        N = self.N
        act = np.empty(N, np.int32)
        bj = np.empty(N)
        seq = np.full((N, H), -1, np.int32) if want_seq else None
        lib().orc_env_mpc_decide(self._h, C.c_int(H), C.c_int(mode), _p(act), _p(bj), _p(seq))
        return (act, bj, seq) if want_seq else (act, bj)

    def field(self, name):
        ids = dict(seg=(0, np.int32, 1), chunk=(1, np.int32, 1), last_q=(2, np.int32, 1), trace_id=(3, np.int32, 1),
                   hist_len=(4, np.int32, 1), done=(5, np.uint8, 1), err_len=(6, np.int32, 1), phase=(10, np.float64, 1),
                   buffer=(11, np.float64, 1), bw_hist=(12, np.float64, self.K), last_pred=(13, np.float64, 1),
                   err_ring=(14, np.float64, self.K), t_now=(16, np.float64, 1), play_time=(17, np.float64, 1),
                   started=(7, np.uint8, 1), pos=(18, np.float64, 1), play_id=(8, np.int32, 1), play_len=(19, np.float64, 1))
        fid, dt, w = ids[name]
        ptr = lib().orc_env_field(self._h, C.c_int(fid))
        n = self.N * w
        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dt, count=n).copy()
        return arr.reshape(self.N, w) if w > 1 else arr

    def errors(self):
        return lib().orc_env_error_count(self._h)


def stats_from_acc(acc):
    acc = _f64(acc)
    out = np.empty(NUM_STATS)
    lib().orc_stats_from_acc(_p(acc), C.c_int(acc.shape[1]), _p(out))
    return out


def mpc_decide(sizes, util, chunk_idx, prev_q, buffer, bw_hist, hist_len, H, mode, params=None,
               last_pred=None, err_ring=None, err_len=None, ses=False, startup=None, n_ts=1, ts_step=0.0):
    """Standalone batched decision (orc_mpc_decide_ex).  Returns dict of numpy arrays.  ``ses``: predictor
    "expsmoothing" (SPEC 5.4); ``n_ts`` > 1: start-up phase with the delay grid jt * ts_step (SPEC 5.3) for the sessions
    with ``startup`` != 0 (None = all)."""
    sizes, util = _f64(sizes), _f64(util)
    V, A = sizes.shape
    chunk_idx, prev_q, hist_len = _i32(chunk_idx), _i32(prev_q), _i32(hist_len)
    buffer, bw_hist = _f64(buffer), _f64(bw_hist)
    N, K = bw_hist.shape
    p = params if params is not None else make_params()
    act = np.empty(N, np.int32)
    bj = np.empty(N)
    seq = np.empty((N, H), np.int32)
    preds = np.empty((N, H))
    nerr = C.c_int32(0)
    ts = np.zeros(N)
    su = None if startup is None else np.ascontiguousarray(startup, dtype=np.uint8)
    lib().orc_mpc_decide_ex(_p(sizes), _p(util), C.c_int(V), C.c_int(A), C.byref(p), C.c_int(N), _p(chunk_idx),
                            _p(prev_q), _p(buffer), _p(bw_hist), _p(hist_len), C.c_int(K), _p(last_pred), _p(err_ring),
                            _p(err_len), C.c_int(H), C.c_int(mode), C.c_int(1 if ses else 0), _p(su), C.c_int(n_ts),
                            C.c_double(ts_step), _p(act), _p(ts), _p(bj), _p(seq), _p(preds), C.byref(nerr))
    return dict(action=act, best_J=bj, best_seq=seq, preds=preds, n_errors=nerr.value, startup_delay=ts)
