"""Generate golden vectors from the UNMODIFIED reference — TEST INFRASTRUCTURE.

Runs only in the build container (needs /root/reference, which does not exist
on the GPU box).  Imports ``/root/reference/mpc.py`` under a ``statsmodels``
stub (the default "harmonic" predictor never touches statsmodels, mpc.py:72-79
vs :81-93; the package is not installed here), drives it through
the player protocol of ``mpc_test.py:39-50`` and writes

    tests/golden/mpc_ref_golden.json

Usage:  python oracle/gen_golden.py            (re-creates the fixture)
Nothing from the reference is copied: only its *outputs* are stored.
"""
from __future__ import annotations

import io
import json
import os
import sys
import types
import contextlib

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                   "mpc_ref_golden.json")


def import_reference_mpc():
    for m in ("statsmodels", "statsmodels.tsa", "statsmodels.tsa.holtwinters"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["statsmodels.tsa.holtwinters"].SimpleExpSmoothing = None
    sys.path.insert(0, REF)
    import mpc  # noqa: the reference module, unmodified
    return mpc


# --- minimal player protocol (attribute bags the reference reads; cf. mpc_test.py:13-50) ---
class _Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _Player:
    def __init__(self, mpd, qoe, info):
        self._m, self._q, self._i = mpd, qoe, info

    def get_mpd(self):
        return self._m

    def get_qoe_metric(self):
        return self._q

    def get_next_chunk_info(self):
        return self._i


def make_player(sc):
    chunks = [_Bag(bitrates=list(b), sizes=list(s)) for b, s in zip(sc["bitrates"], sc["sizes"])]
    mpd = _Bag(video_length=len(chunks), chunk_length=sc["chunk_length"],
               max_buffer=sc["max_buffer"], chunks=chunks)
    qoe = _Bag(rebuffer_weight=sc["rw"], variance_weight=sc["vw"], startup_weight=0)
    info = _Bag(chunk_number=sc["k"], previous_bitrate=sc["prev_q"],
                previous_bandwidths=list(sc["history"]), buffer_level=sc["buffer"])
    return _Player(mpd, qoe, info)


def run_reference(mpc, sc, want_grid=False, repeat=1):
    player = make_player(sc)
    abr = mpc.MPCBitrateController(player)
    abr.horizon = sc["H"]
    out = {"actions": [], "hist_len_after": []}
    for _ in range(repeat):
        with contextlib.redirect_stdout(io.StringIO()):
            a = abr.next_bitrate()
        out["actions"].append(int(a))
        out["hist_len_after"].append(len(player._i.previous_bandwidths))
    # full result of one more optimisation on a FRESH player (first call semantics)
    player = make_player(sc)
    abr = mpc.MPCBitrateController(player)
    abr.horizon = sc["H"]
    abr.update_bandwidth_prediction()
    info = player.get_next_chunk_info()
    out["preds"] = [float(p) for p in abr.predicted_bandwidths]
    with contextlib.redirect_stdout(io.StringIO()):
        best = abr.optimize_qoe(info)
    best = [int(x) for x in np.atleast_1d(best)]
    out["best_seq"] = best
    out["best_J"] = float(abr.objective(best, info))
    if want_grid:
        import itertools
        A = len(sc["bitrates"][0])
        out["J"] = [float(abr.objective(list(R), info))
                    for R in itertools.product(range(A), repeat=sc["H"])]
    return out


def scenario_mpc_test():
    """The fixed scenario of mpc_test.py:52-72 (values restated, not imported)."""
    ladder = [1, 2.5, 5, 8]
    return dict(name="mpc_test", bitrates=[ladder] * 60, sizes=[ladder] * 60, chunk_length=1,
                max_buffer=20, rw=1, vw=0, k=20, prev_q=1, history=[2, 2.5, 4, 6, 8],
                buffer=20, H=5)


def random_scenario(rng, idx):
    A = int(rng.choice([2, 3, 4, 6]))
    H = int(rng.integers(2, 6)) if A <= 4 else int(rng.integers(2, 5))   # H=1 raises in the reference, see errors
    V = int(rng.choice([12, 48, 60]))
    L = float(rng.choice([1.0, 2.0, 4.0]))
    ladders = {2: [300, 4300], 3: [300, 1200, 4300], 4: [1, 2.5, 5, 8],
               6: [300, 750, 1200, 1850, 2850, 4300]}
    lad = [float(x) for x in ladders[A]]
    scale = 1.0 if A == 4 else 1e-3
    bitrates = [[b * scale for b in lad] for _ in range(V)]
    style = idx % 3
    if style == 0:       # sizes == bitrates (mpc_test style): lots of exact ties
        sizes = [list(r) for r in bitrates]
    elif style == 1:     # VBR sizes
        sizes = [[b * L * float(rng.uniform(0.8, 1.2)) for b in r] for r in bitrates]
    else:                # CBR sizes, rounded values -> ties
        sizes = [[round(b * L, 2) for b in r] for r in bitrates]
    n_hist = int(rng.integers(1, 9))
    history = [float(np.round(rng.uniform(0.2, 6.0), int(rng.integers(1, 6)))) for _ in range(n_hist)]
    history = [h if h > 0 else 0.5 for h in history]
    k = int(rng.integers(0, V - H + 1))
    return dict(name=f"rand{idx}", bitrates=bitrates, sizes=sizes, chunk_length=L,
                max_buffer=float(rng.choice([10.0, 20.0, 60.0])),
                rw=float(rng.choice([1.0, 4.3, 8.0])), vw=float(rng.choice([0.0, 1.0, 0.5])),
                k=k, prev_q=int(rng.integers(0, A)), history=history,
                buffer=float(np.round(rng.uniform(0.0, 30.0), 2)), H=H)


def main():
    mpc = import_reference_mpc()
    rng = np.random.default_rng(20261018)
    cases = []
    sc = scenario_mpc_test()
    cases.append(dict(scenario=sc, ref=run_reference(mpc, sc, want_grid=True, repeat=3)))
    # survey §4 KAT variants of the shipped scenario
    for name, kw in [("w4.3_1_buf4", dict(rw=4.3, vw=1, buffer=4.0)),
                     ("w4.3_1_buf0.5", dict(rw=4.3, vw=1, buffer=0.5, history=[1.0, 0.8, 1.2])),
                     ("w1_1_buf20", dict(rw=1, vw=1)),
                     ("k55", dict(k=55))]:
        s2 = dict(scenario_mpc_test(), name=name, **kw)
        cases.append(dict(scenario=s2, ref=run_reference(mpc, s2, want_grid=False, repeat=1)))
    for i in range(160):
        sc = random_scenario(rng, i)
        combos = len(sc["bitrates"][0]) ** sc["H"]
        cases.append(dict(scenario=sc, ref=run_reference(mpc, sc, want_grid=(combos <= 256 and i < 60),
                                                         repeat=2 if i % 5 == 0 else 1)))
    # the BASELINE.json ladder at H=5 (7 776 combos), a few sessions
    for i in range(4):
        sc = random_scenario(rng, 1000 + i)
        lad = [0.3, 0.75, 1.2, 1.85, 2.85, 4.3]
        sc.update(bitrates=[lad] * 48, H=5, k=int(rng.integers(0, 44)), chunk_length=4.0, max_buffer=60.0,
                  sizes=[[b * 4.0 * float(rng.uniform(0.8, 1.2)) for b in lad] for _ in range(48)],
                  prev_q=int(rng.integers(0, 6)), name=f"a6h5_{i}")
        cases.append(dict(scenario=sc, ref=run_reference(mpc, sc)))
    # error behaviour of the reference
    errors = []
    for name, kw in [("index_error_k56", dict(k=56)), ("empty_history", dict(history=[])),
                     ("zero_sample", dict(history=[2.0, 0.0, 3.0])),
                     ("horizon_1", dict(H=1))]:   # brute returns a 0-d array -> result[0] fails (mpc.py:186)
        s2 = dict(scenario_mpc_test(), name=name, **kw)
        try:
            run_reference(mpc, s2)
            errors.append(dict(scenario=s2, raises=None))
        except Exception as e:  # noqa
            errors.append(dict(scenario=s2, raises=type(e).__name__))
    # raw predictor / buffer-model KATs
    abr = mpc.MPCBitrateController(make_player(scenario_mpc_test()))
    h1 = [1, 2, 3, 4]
    kat = dict(predict_3_1234=[float(x) for x in abr.predict_throughput(3, h1)], predict_3_1234_len_after=len(h1),
               calc_wait=float(abr.calc_wait(20, 20, 0, 3.468208092485549)),
               next_buffer_a=float(abr.next_buffer(20, 20, 0, 3.468208092485549)),
               next_buffer_b=float(abr.next_buffer(20, 0.3, 3, 3.468208092485549)))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(dict(generator="oracle/gen_golden.py", reference="Elliotshui/ABRSimulator mpc.py (unmodified, "
                       "statsmodels stubbed)", scipy=__import__("scipy").__version__,
                       numpy=np.__version__, cases=cases, errors=errors, kat=kat), f)
    print(f"wrote {OUT}: {len(cases)} cases, {len(errors)} error cases, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
