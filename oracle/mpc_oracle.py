"""CPU oracle for the MPC decision path — TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
``abrsimulator_b200`` never does.

Pure-Python restatement (binary64, Python floats, loops) of:

* Profile R  (SPEC.md §5.1) — ``/root/reference/mpc.py`` exactly as shipped:
  ``predict_throughput`` mpc.py:81-93, ``calc_wait`` mpc.py:104-109,
  ``next_buffer`` mpc.py:111-118, ``objective`` mpc.py:120-162,
  ``optimize_qoe``/``scipy.optimize.brute`` mpc.py:171-179 (C-order grid, first
  minimum), ``next_bitrate`` mpc.py:181-186.
* Profile N robust MPC (SPEC.md §5.2), the start-up phase (§5.3) and the
  "expsmoothing" predictor (§5.4) — parity unpinned by the reference.

Parity status: Profile R is PINNED — ``tests/test_oracle_golden.py`` checks this
file against fixtures produced by importing the unmodified reference
(``oracle/gen_golden.py`` → ``tests/golden/mpc_ref_golden.json``, 169 scenarios with
full score grids; ``oracle/gen_golden_bulk.py`` → ``mpc_ref_bulk.json``, 10 240 more
decisions) and against the reference's only golden (``mpc_test.py:52-86`` →
"Test next bitrate: 2").
Profile N: parity unpinned (defined by SPEC.md).
"""
from __future__ import annotations

import itertools
import math


class MpcInputError(Exception):
    """Base for the input errors the reference raises as plain Python errors."""


def predict_harmonic_ref(horizon, history):
    """mpc.py:81-93.  Returns (predictions, mutated_history).

    The reference re-sums the whole (growing) list on every iteration; since
    the summation is left-to-right, iteration i+1's sum equals iteration i's
    sum plus 1/p_i bit for bit, which is what is done here.  ``history`` is
    NOT mutated; the list the reference would have left behind is returned.
    """
    hist = list(history)
    if len(hist) == 0:
        raise ZeroDivisionError("empty throughput history (mpc.py:90)")
    s = 0
    for x in hist:
        s += 1 / x                      # ZeroDivisionError on x == 0 (mpc.py:88)
    preds = []
    for _ in range(horizon):
        p = len(hist) / s
        preds.append(p)
        hist.append(p)
        s += 1 / p
    return preds, hist


def _max0(x):
    # Python max(0, x): first argument wins ties, NaN compares false -> 0
    return x if x > 0 else 0.0


def next_buffer_ref(size, buf, bw, chunk_length, max_buffer):
    """mpc.py:104-118 with chunk_size already looked up."""
    t = _max0(buf - size / bw)
    wait = _max0(t + chunk_length - max_buffer)
    return _max0(t + chunk_length - wait)


def objective_ref(R, k, prev_q, buffer_level, preds, bitrates, sizes,
                  chunk_length, max_buffer, vw, rw, utility=None):
    """mpc.py:120-162: J = -QoE of one bitrate sequence (quirks D11, D12 kept).

    ``utility`` is an optional [V][A] table; default identity (mpc.py:95-97).
    """
    H = len(R)
    if k + H > len(sizes):
        raise IndexError("list index out of range (mpc.py:125-128)")
    U = bitrates if utility is None else utility
    seq = [prev_q] + [int(r) for r in R]
    vq = 0
    qv = 0
    rt = 0
    b = buffer_level
    for i in range(H):
        a, ap = seq[i + 1], seq[i]
        vq += U[k + i][a]
        qv += abs(U[k + i][a] - U[k + i][ap])
        rt += max(0, sizes[k + i][a], chunk_length) / preds[i] - b
        if i != H - 1:
            b = next_buffer_ref(sizes[k][a], b, preds[i], chunk_length, max_buffer)
    qoe = vq - vw * qv - rw * rt
    return -qoe


def decide_ref(k, prev_q, buffer_level, history, H, bitrates, sizes,
               chunk_length, max_buffer, vw, rw, utility=None, want_grid=False):
    """mpc.py:181-186.  Returns dict(action, best_seq, best_J, preds, history_after[, J])."""
    A = len(bitrates[0])
    preds, hist_after = predict_harmonic_ref(H, history)
    best_j = None
    best_seq = None
    grid = [] if want_grid else None
    # np.mgrid row-major: first slice slowest == itertools.product order
    for R in itertools.product(range(A), repeat=H):
        j = objective_ref(R, k, prev_q, buffer_level, preds, bitrates, sizes,
                          chunk_length, max_buffer, vw, rw, utility)
        if want_grid:
            grid.append(j)
        if best_j is None or j < best_j:          # argmin keeps the FIRST minimum
            best_j, best_seq = j, R
    out = dict(action=int(best_seq[0]), best_seq=list(best_seq), best_J=best_j,
               preds=preds, history_after=hist_after)
    if want_grid:
        out["J"] = grid
    return out


# ----------------------------------------------------------------------------
# SPEC §5.4: predictor "expsmoothing" (mpc.py:72-79) — parity unpinned (statsmodels is not installed here)
# ----------------------------------------------------------------------------

def predict_ses(horizon, history, alpha=0.5):
    """``SimpleExpSmoothing(data).fit(0.5).predict(n, n+horizon-1)`` (mpc.py:72-79) in closed form.

    Simple exponential smoothing keeps a level ``l_t = alpha*y_t + (1-alpha)*l_{t-1}`` and forecasts it flat:
    every one of the ``horizon`` predictions is ``l_n``.  statsmodels' ``fit`` with a fixed smoothing level estimates the
    initial level ``l_0`` by minimising the sum of squared one-step errors ``sum_t (y_t - l_{t-1})**2``; with ``alpha``
    fixed every ``l_t`` is affine in ``l_0`` (``l_t = la_t + lb_t*l_0``), so that minimisation is a one-dimensional
    linear least-squares problem with the solution computed below (statsmodels reaches the same point with a numerical
    optimiser, to its tolerance).  alpha = 0.5 here (``fit(0.5)``), so ``lb_t = 0.5**t`` exactly."""
    if len(history) == 0:
        raise ZeroDivisionError("empty throughput history")
    la, lb, num, den = 0.0, 1.0, 0.0, 0.0
    for y in history:
        r = y - la
        num = num + lb * r
        den = den + lb * lb
        la = alpha * y + (1.0 - alpha) * la
        lb = (1.0 - alpha) * lb
    l_n = la + lb * (num / den)
    return [l_n] * horizon


def decide_ref_ses(k, prev_q, buffer_level, history, H, bitrates, sizes, chunk_length, max_buffer, vw, rw, utility=None):
    """Profile R's search (objective and enumeration of SPEC §5.1) over the "expsmoothing" predictions."""
    A = len(bitrates[0])
    preds = predict_ses(H, history)
    best_j, best_seq = None, None
    for R in itertools.product(range(A), repeat=H):
        j = objective_ref(R, k, prev_q, buffer_level, preds, bitrates, sizes, chunk_length, max_buffer, vw, rw, utility)
        if best_j is None or j < best_j:
            best_j, best_seq = j, R
    return dict(action=int(best_seq[0]), best_seq=list(best_seq), best_J=best_j, preds=preds)


# ----------------------------------------------------------------------------
# SPEC §5.3: start-up phase (f_st of mpc.py:7-18; the reference's TODO at mpc.py:141) — defined by SPEC.md
# ----------------------------------------------------------------------------

def decide_startup(objective_of_buffer, A, h, buffer_level, sw, n_ts, ts_step):
    """``objective_of_buffer(R, b0) -> J`` is one of the two objectives above with the initial buffer exposed.
    Start-up delay T_s = jt*ts_step is the slowest axis; total cost J + sw*T_s computed as -((-J) - sw*T_s)."""
    best, best_seq, best_ts = None, None, 0.0
    for jt in range(n_ts):
        ts = 0.0 if jt == 0 else jt * ts_step
        b0 = buffer_level if jt == 0 else buffer_level + ts
        bj, bs = None, None
        for R in itertools.product(range(A), repeat=h):
            j = objective_of_buffer(R, b0)
            if bj is None or j < bj:
                bj, bs = j, R
        tot = -((-bj) - sw * ts)
        if best is None or tot < best:
            best, best_seq, best_ts = tot, bs, ts
    return dict(action=int(best_seq[0]), best_seq=list(best_seq), best_J=best, startup_delay=best_ts)


# ----------------------------------------------------------------------------
# Profile N: robust MPC (SPEC.md §5.2)
# ----------------------------------------------------------------------------

class RobustState:
    """Per-session predictor state: last harmonic estimate and error ring."""

    def __init__(self, K=5):
        self.K = K
        self.last_pred = 0.0
        self.errs = []

    def reset(self):
        self.last_pred = 0.0
        self.errs = []


def robust_predict(history, state: RobustState):
    """SPEC §5.2 predictor.  ``history``: all samples so far, oldest first."""
    K = state.K
    hist = list(history)[-K:]
    n = len(hist)
    if n == 0:
        return None
    s = 0.0
    for x in hist:
        s = s + 1 / x
    hm = n / s
    if state.last_pred > 0:
        state.errs.append(abs(state.last_pred - hist[-1]) / hist[-1])
        state.errs = state.errs[-K:]
    max_err = 0.0
    for e in state.errs:
        if e > max_err:
            max_err = e
    state.last_pred = hm
    return hm / (1 + max_err)


def objective_robust(R, k, prev_q, buffer_level, c, U, sizes, chunk_length,
                     max_buffer, vw, rw):
    h = len(R)
    seq = [prev_q] + [int(r) for r in R]
    vq = 0.0
    qv = 0.0
    rt = 0.0
    b = buffer_level
    for i in range(h):
        a, ap = seq[i + 1], seq[i]
        vq = vq + U[k + i][a]
        if ap >= 0:
            qv = qv + abs(U[k + i][a] - U[k + i][ap])
        dl = sizes[k + i][a] / c
        rt = rt + _max0(dl - b)
        if i != h - 1:
            t = _max0(b - dl)
            w = _max0(t + chunk_length - max_buffer)
            b = _max0(t + chunk_length - w)
    return -((vq - vw * qv) - rw * rt)


def decide_robust(k, prev_q, buffer_level, history, state: RobustState, H, U,
                  sizes, chunk_length, max_buffer, vw, rw, default_quality=1):
    """SPEC §5.2.  Returns dict(action, best_seq, best_J, c)."""
    A = len(U[0])
    V = len(U)
    c = robust_predict(history, state)
    if c is None:
        return dict(action=default_quality, best_seq=[], best_J=math.nan, c=math.nan)
    h = min(H, V - k)
    if h <= 0:
        return dict(action=0, best_seq=[], best_J=math.nan, c=c)
    best_j, best_seq = None, None
    for R in itertools.product(range(A), repeat=h):
        j = objective_robust(R, k, prev_q, buffer_level, c, U, sizes,
                             chunk_length, max_buffer, vw, rw)
        if best_j is None or j < best_j:
            best_j, best_seq = j, R
    return dict(action=int(best_seq[0]), best_seq=list(best_seq), best_J=best_j, c=c)
