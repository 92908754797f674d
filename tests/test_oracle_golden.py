"""Pins the CPU oracles (oracle/mpc_oracle.py, oracle/abr_oracle.c) to the reference.

Fixtures in tests/golden/mpc_ref_golden.json were produced by importing the
unmodified /root/reference/mpc.py (oracle/gen_golden.py).  The reference's one
golden is mpc_test.py:52-86 -> "Test next bitrate: 2".
"""
import numpy as np
import pytest

from oracle import mpc_oracle as mo
from oracle import oracle as orc


def _py_decide(sc, want_grid=False, history=None):
    return mo.decide_ref(sc["k"], sc["prev_q"], sc["buffer"], sc["history"] if history is None else history,
                         sc["H"], sc["bitrates"], sc["sizes"], sc["chunk_length"], sc["max_buffer"],
                         sc["vw"], sc["rw"], want_grid=want_grid)


def _c_decide(sc, K=None):
    sizes = np.array(sc["sizes"], float)
    util = np.array(sc["bitrates"], float)          # identity utility (mpc.py:95-97)
    hist = list(sc["history"])
    K = K or max(1, len(hist))
    ring = np.zeros((1, K))
    ring[0, :len(hist)] = hist
    p = orc.make_params(chunk_length=sc["chunk_length"], max_buffer=sc["max_buffer"], rebuf_penalty=sc["rw"],
                        smooth_penalty=sc["vw"], utility_scale=1.0)
    return orc.mpc_decide(sizes, util, [sc["k"]], [sc["prev_q"]], [sc["buffer"]], ring, [len(hist)], sc["H"], 0, p)


def test_reference_golden_mpc_test_scenario(golden):
    """mpc_test.py:81-86 prints 'Test next bitrate: 2'; survey KATs for the same scenario."""
    c = golden["cases"][0]
    assert c["scenario"]["name"] == "mpc_test"
    assert c["ref"]["actions"] == [2, 2, 2]
    assert c["ref"]["hist_len_after"] == [10, 15, 20]        # D10 list mutation
    assert c["ref"]["best_seq"] == [2, 1, 3, 3, 3]
    assert c["ref"]["best_J"] == -117.56833333333331
    r = _py_decide(c["scenario"], want_grid=True)
    assert r["action"] == 2 and r["best_seq"] == [2, 1, 3, 3, 3]
    assert r["best_J"] == -117.56833333333331
    assert r["preds"] == c["ref"]["preds"]
    assert r["J"] == c["ref"]["J"]                            # all 1 024 scores bit-identical
    assert len(r["history_after"]) == 10
    rc = _c_decide(c["scenario"])
    assert rc["action"][0] == 2 and list(rc["best_seq"][0]) == [2, 1, 3, 3, 3]
    assert rc["best_J"][0] == -117.56833333333331
    assert list(rc["preds"][0]) == c["ref"]["preds"]


def test_kats(golden):
    k = golden["kat"]
    preds, after = mo.predict_harmonic_ref(3, [1, 2, 3, 4])
    assert preds == k["predict_3_1234"] and len(after) == k["predict_3_1234_len_after"] == 7
    assert preds == [1.9200000000000004, 1.9200000000000004, 1.9200000000000006]
    bw = 3.468208092485549
    assert mo.next_buffer_ref(1, 20, bw, 1, 20) == k["next_buffer_a"] == 20.0
    assert mo.next_buffer_ref(8, 0.3, bw, 1, 20) == k["next_buffer_b"] == 1.0
    assert k["calc_wait"] == 0.711666666666666


def test_python_oracle_matches_reference_everywhere(golden):
    n_grid = 0
    for c in golden["cases"]:
        sc, ref = c["scenario"], c["ref"]
        r = _py_decide(sc, want_grid="J" in ref)
        assert r["preds"] == ref["preds"], sc["name"]
        assert r["best_seq"] == ref["best_seq"], sc["name"]
        assert r["best_J"] == ref["best_J"], sc["name"]
        assert r["action"] == ref["actions"][0], sc["name"]
        if "J" in ref:
            assert r["J"] == ref["J"], sc["name"]
            n_grid += 1
        # repeated calls on the same player see the polluted history (D10)
        hist = list(sc["history"])
        for i, a in enumerate(ref["actions"]):
            rr = _py_decide(sc, history=hist)
            assert rr["action"] == a, (sc["name"], i)
            hist = rr["history_after"]
            assert len(hist) == ref["hist_len_after"][i]
    assert n_grid >= 20 and len(golden["cases"]) >= 160


def test_c_oracle_matches_reference_everywhere(golden):
    for c in golden["cases"]:
        sc, ref = c["scenario"], c["ref"]
        r = _c_decide(sc)
        assert r["n_errors"] == 0
        assert list(r["preds"][0]) == ref["preds"], sc["name"]
        assert list(r["best_seq"][0]) == ref["best_seq"], sc["name"]
        assert r["best_J"][0] == ref["best_J"], sc["name"]
        assert r["action"][0] == ref["actions"][0], sc["name"]
        # ring storage with a larger capacity must not change anything
        r2 = _c_decide(sc, K=len(sc["history"]) + 3)
        assert r2["action"][0] == ref["actions"][0] and r2["best_J"][0] == ref["best_J"]


def test_reference_error_behaviour(golden):
    by = {e["scenario"]["name"]: e for e in golden["errors"]}
    assert by["index_error_k56"]["raises"] == "IndexError"
    assert by["empty_history"]["raises"] == "ZeroDivisionError"
    assert by["zero_sample"]["raises"] == "ZeroDivisionError"
    assert by["horizon_1"]["raises"] == "IndexError"       # brute returns a 0-d array at H=1 (mpc.py:186)
    with pytest.raises(IndexError):
        _py_decide(by["index_error_k56"]["scenario"])
    with pytest.raises(ZeroDivisionError):
        _py_decide(by["empty_history"]["scenario"])
    with pytest.raises(ZeroDivisionError):
        _py_decide(by["zero_sample"]["scenario"])
    for name in ("index_error_k56", "empty_history", "zero_sample"):
        r = _c_decide(by[name]["scenario"])
        assert r["action"][0] == -1 and r["n_errors"] == 1


# ---- SPEC §5.4 / §5.3: the two parts of mpc.py the default path never reaches ----
def test_ses_closed_form_is_the_least_squares_initial_level():
    """mpc.py:72-79 calls statsmodels' SimpleExpSmoothing(data).fit(0.5): fixed smoothing level, initial level estimated
    by minimising the sum of squared one-step errors.  statsmodels is not installed here (parity unpinned); this checks
    that the closed form of SPEC §5.4 is the minimiser a numerical optimiser finds for that same objective, and the
    textbook identities of simple exponential smoothing."""
    from scipy.optimize import minimize_scalar
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 5, 8, 20, 60):
        y = rng.uniform(0.2, 6.0, size=n)

        def level_and_sse(l0):
            lvl, sse = l0, 0.0
            for v in y:
                sse += (v - lvl) ** 2
                lvl = 0.5 * v + 0.5 * lvl
            return lvl, sse

        pred = mo.predict_ses(4, list(y))
        assert len(pred) == 4 and len(set(pred)) == 1                      # flat forecast
        best = minimize_scalar(lambda l0: level_and_sse(l0)[1], bracket=(0.0, 6.0), tol=1e-14)
        assert abs(level_and_sse(best.x)[0] - pred[0]) <= 1e-6 * max(1.0, abs(pred[0])), n
        # perturbing the closed-form initial level never lowers the objective
        la, lb, num, den = 0.0, 1.0, 0.0, 0.0
        for v in y:
            num, den = num + lb * (v - la), den + lb * lb
            la, lb = 0.5 * v + 0.5 * la, 0.5 * lb
        l0 = num / den
        s0 = level_and_sse(l0)[1]
        assert all(level_and_sse(l0 + d)[1] >= s0 - 1e-12 for d in (-1e-3, 1e-3, -0.5, 0.5))
    assert mo.predict_ses(3, [2.5]) == [2.5, 2.5, 2.5]                       # one sample: the level is the sample
    with pytest.raises(ZeroDivisionError):
        mo.predict_ses(3, [])


def test_c_oracle_ses_and_startup_equal_python():
    """C restatement of SPEC §5.3 (start-up delay as a decision variable) and §5.4 (expsmoothing predictor) against the
    pure-Python one, bit for bit; n_ts = 1 is the plain decision."""
    rng = np.random.default_rng(21)
    V, A, H, K = 16, 4, 3, 12
    bitrates = np.tile(np.array([300.0, 1200.0, 2850.0, 4300.0]) / 1000.0, (V, 1))
    sizes = bitrates * 4.0 * rng.uniform(0.8, 1.2, size=(V, A))
    nonzero_ts = 0
    for trial in range(30):
        n = int(rng.integers(1, K + 1))
        hist = [float(x) for x in rng.uniform(0.3, 5.0, size=n)]
        k, prev_q, buf = int(rng.integers(0, V - H + 1)), int(rng.integers(0, A)), float(rng.uniform(0, 6))
        sw = float(rng.choice([0.1, 0.5, 1.0, 3.0]))
        P = orc.make_params(chunk_length=4.0, max_buffer=30.0, rebuf_penalty=4.3, smooth_penalty=1.0, startup_penalty=sw,
                            utility_scale=1.0)
        ring = np.zeros((1, K)); ring[0, :n] = hist
        # §5.4
        c = orc.mpc_decide(sizes, bitrates, [k], [prev_q], [buf], ring, [n], H, 0, P, ses=True)
        r = mo.decide_ref_ses(k, prev_q, buf, hist, H, bitrates.tolist(), sizes.tolist(), 4.0, 30.0, 1.0, 4.3)
        assert c["n_errors"] == 0 and c["action"][0] == r["action"] and c["best_J"][0] == r["best_J"]
        assert list(c["preds"][0]) == r["preds"] and list(c["best_seq"][0]) == r["best_seq"]
        # §5.3 over both objectives
        for mode in (0, 1):
            if mode == 0:
                preds, _ = mo.predict_harmonic_ref(H, hist)
                obj = lambda R, b0: mo.objective_ref(R, k, prev_q, b0, preds, bitrates.tolist(), sizes.tolist(), 4.0, 30.0, 1.0, 4.3)
            else:
                cc = mo.robust_predict(hist, mo.RobustState(K))
                obj = lambda R, b0: mo.objective_robust(R, k, prev_q, b0, cc, bitrates.tolist(), sizes.tolist(), 4.0, 30.0, 1.0, 4.3)
            r = mo.decide_startup(obj, A, H, buf, sw, 9, 0.75)
            Pm = orc.make_params(chunk_length=4.0, max_buffer=30.0, rebuf_penalty=4.3, smooth_penalty=1.0,
                                 startup_penalty=sw, utility_scale=1.0, hist_k=K)
            c = orc.mpc_decide(sizes, bitrates, [k], [prev_q], [buf], ring, [n], H, mode, Pm, n_ts=9, ts_step=0.75)
            assert c["action"][0] == r["action"] and c["best_J"][0] == r["best_J"], (trial, mode)
            assert c["startup_delay"][0] == r["startup_delay"] and list(c["best_seq"][0]) == r["best_seq"]
            nonzero_ts += r["startup_delay"] > 0
            # sessions outside the start-up phase and n_ts = 1 take the plain decision
            plain = orc.mpc_decide(sizes, bitrates, [k], [prev_q], [buf], ring, [n], H, mode, Pm)
            off = orc.mpc_decide(sizes, bitrates, [k], [prev_q], [buf], ring, [n], H, mode, Pm, startup=[0], n_ts=9, ts_step=0.75)
            assert off["action"][0] == plain["action"][0] and off["best_J"][0] == plain["best_J"][0]
            assert off["startup_delay"][0] == 0.0
    assert nonzero_ts >= 5        # the grid is exercised: waiting pays when the start-up weight is small


def _bulk():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpc_ref_bulk.json")) as f:
        return json.load(f)


def test_oracles_match_10k_decisions_of_the_reference_itself():
    """tests/golden/mpc_ref_bulk.json: 10 240 decisions taken by the unmodified /root/reference/mpc.py (oracle/
    gen_golden_bulk.py).  Scenario i is rebuilt from its index; the C oracle must return the reference's best sequence
    (scipy.optimize.brute's first minimum), objective value and first prediction bit for bit on every one, the
    pure-Python restatement on every 16th."""
    from oracle.gen_golden_bulk import bulk_scenario
    doc = _bulk()
    assert doc["n"] >= 10000
    for i in range(doc["n"]):
        sc = bulk_scenario(i)
        want_seq = [int(ch) for ch in doc["best_seq"][i]]
        want_j, want_p0 = float.fromhex(doc["best_J"][i]), float.fromhex(doc["pred0"][i])
        r = _c_decide(sc)
        assert r["n_errors"] == 0
        assert list(r["best_seq"][0]) == want_seq, i
        assert r["best_J"][0] == want_j and r["preds"][0][0] == want_p0, i
        assert r["action"][0] == want_seq[0]
        if i % 16 == 0:
            rp = _py_decide(sc)
            assert rp["best_seq"] == want_seq and rp["best_J"] == want_j and rp["preds"][0] == want_p0, i
