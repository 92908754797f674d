"""GPU parity of the MPC decision kernel, through the C-ABI (ctypes) — run with -m gpu on a B200.

Bar: bitrate sequences identical to the reference (golden fixtures generated from the unmodified mpc.py)
and to the C oracle; scores bit-identical (the kernel performs the reference's operations in the
reference's order), which is stricter than BASELINE.json's "identical except at near-ties (gap < 1e-9)".
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from abrsimulator_b200 import _lib
from abrsimulator_b200.datamodel import Chunk, ChunkInfo, MPD, QOEMetric
from abrsimulator_b200.mpc import MPCBitrateController, decide_batch
from oracle import oracle as orc
from helpers import bits_equal


class Player:
    """The player protocol of mpc_test.py:39-50."""

    def __init__(self, sc):
        chunks = [Chunk(list(b), list(s)) for b, s in zip(sc["bitrates"], sc["sizes"])]
        self.mpd = MPD(len(chunks), sc["chunk_length"], sc["max_buffer"], chunks)
        self.qoe = QOEMetric(sc["rw"], sc["vw"], 0)
        self.info = ChunkInfo(sc["k"], sc["prev_q"], list(sc["history"]), sc["buffer"])

    def get_mpd(self):
        return self.mpd

    def get_qoe_metric(self):
        return self.qoe

    def get_next_chunk_info(self):
        return self.info


def make_abr(sc, **kw):
    abr = MPCBitrateController(Player(sc), **kw)
    abr.horizon = sc["H"]
    return abr


def test_reference_golden_next_bitrate_is_2(golden):
    """mpc_test.py:81-86 prints 'Test next bitrate: 2'."""
    sc = golden["cases"][0]["scenario"]
    abr = make_abr(sc)
    assert abr.next_bitrate() == 2
    assert len(abr.player.info.previous_bandwidths) == 10      # D10
    assert abr.next_bitrate() == 2 and abr.next_bitrate() == 2
    assert len(abr.player.info.previous_bandwidths) == 20
    abr = make_abr(sc)
    abr.update_bandwidth_prediction()
    assert abr.predicted_bandwidths == golden["cases"][0]["ref"]["preds"]
    best = abr.optimize_qoe(abr.player.get_next_chunk_info())
    assert best.dtype == np.float64 and list(best) == [2.0, 1.0, 3.0, 3.0, 3.0]
    assert abr.objective([2, 1, 3, 3, 3], abr.player.info) == -117.56833333333331
    assert abr.objective([1., 2., 3., 0., 1.], abr.player.info) == -110.29499999999999
    assert abr.default_bitrate_utility(2.5) == 2.5


def test_all_golden_cases_through_facade(golden):
    n_grid = 0
    for c in golden["cases"]:
        sc, ref = c["scenario"], c["ref"]
        abr = make_abr(sc)
        acts = [abr.next_bitrate() for _ in ref["actions"]]
        assert acts == ref["actions"], sc["name"]
        assert len(abr.player.info.previous_bandwidths) == ref["hist_len_after"][-1], sc["name"]
        abr = make_abr(sc)
        act, seq, bj, preds = abr._decide(sc["k"], sc["prev_q"], sc["history"], sc["buffer"], sc["H"])
        assert list(preds) == ref["preds"], sc["name"]
        assert list(seq) == ref["best_seq"], sc["name"]
        assert bj == ref["best_J"], sc["name"]
        if "J" in ref:
            grid = abr.score_grid(abr.player.info)
            assert bits_equal(grid, np.array(ref["J"])) == 0, sc["name"]
            assert int(np.argmin(grid)) == int(np.argmin(np.array(ref["J"])))
            n_grid += 1
    assert n_grid >= 20


def test_kernel_matches_10k_decisions_of_the_reference_itself():
    """tests/golden/mpc_ref_bulk.json (oracle/gen_golden_bulk.py): 10 240 decisions of the unmodified reference
    controller; every one through the drop-in controller's host path (abr_mpc_decide_host): identical best sequence,
    bit-identical objective value and prediction."""
    import json
    import os
    from oracle.gen_golden_bulk import bulk_scenario
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpc_ref_bulk.json")) as f:
        doc = json.load(f)
    assert doc["n"] >= 10000
    bad = []
    for i in range(doc["n"]):
        sc = bulk_scenario(i)
        abr = make_abr(sc)
        act, seq, bj, preds = abr._decide(sc["k"], sc["prev_q"], sc["history"], sc["buffer"], sc["H"])
        want_seq = [int(ch) for ch in doc["best_seq"][i]]
        if (list(seq) != want_seq or bj != float.fromhex(doc["best_J"][i]) or preds[0] != float.fromhex(doc["pred0"][i])
                or act != want_seq[0]):
            bad.append((i, list(seq), want_seq, bj, float.fromhex(doc["best_J"][i])))
    assert not bad, (len(bad), bad[:5])


def test_reference_error_behaviour(golden):
    by = {e["scenario"]["name"]: e["scenario"] for e in golden["errors"]}
    with pytest.raises(IndexError):
        make_abr(by["index_error_k56"]).next_bitrate()
    with pytest.raises(ZeroDivisionError):
        make_abr(by["empty_history"]).next_bitrate()
    with pytest.raises(ZeroDivisionError):
        make_abr(by["zero_sample"]).next_bitrate()
    # H = 1 raises in the reference only because brute returns a 0-d array (mpc.py:186); supported here
    assert make_abr(by["horizon_1"]).next_bitrate() in (0, 1, 2, 3)


def _random_batch(rng, N, V, A, K, ties):
    if ties:
        lad = np.sort(rng.choice(np.arange(1, 40) * 0.25, size=A, replace=False))
        bitrates = np.tile(lad, (V, 1))
        sizes = bitrates.copy()
    else:
        lad = np.sort(rng.uniform(0.2, 5.0, size=A))
        bitrates = np.tile(lad, (V, 1)) * rng.uniform(0.9, 1.1, size=(V, 1))
        sizes = bitrates * 4.0 * rng.uniform(0.8, 1.2, size=(V, A))
    hist_len = rng.integers(1, 2 * K + 1, size=N).astype(np.int32)
    bw_hist = np.round(rng.uniform(0.2, 6.0, size=(N, K)), 2 if ties else 9)
    return dict(bitrates=bitrates, sizes=sizes, chunk=None, prev_q=rng.integers(0, A, size=N).astype(np.int32),
                buffer=np.round(rng.uniform(0, 40, size=N), 1 if ties else 9), bw_hist=bw_hist, hist_len=hist_len)


def _run_both(b, V, A, K, H, mode, params_kw, flags=0, robust_state=False, N=None, startup=None, n_ts=1, ts_step=0.0):
    dev = torch.device("cuda")
    p_gpu = _lib.default_params(**params_kw)
    p_cpu = orc.make_params(**params_kw)
    util = orc.utility_table(b["bitrates"], params_kw.get("utility_mode", 0), params_kw.get("utility_scale", 0.001))
    N = len(b["prev_q"])
    st_cpu = st_gpu = (None, None, None)
    if robust_state:
        rng = np.random.default_rng(99)
        lp = np.where(rng.random(N) < 0.8, rng.uniform(0.2, 6.0, N), 0.0)
        er = rng.uniform(0, 0.5, size=(N, K))
        el = rng.integers(0, 2 * K, size=N).astype(np.int32)
        st_cpu = (lp.copy(), er.copy(), el.copy())
        st_gpu = tuple(torch.from_numpy(x.copy()).to(dev) for x in (lp, er, el))
    exp = orc.mpc_decide(b["sizes"], util, b["chunk"], b["prev_q"], b["buffer"], b["bw_hist"], b["hist_len"], H, mode,
                         p_cpu, *st_cpu, ses=bool(flags & _lib.MPC_PRED_SES), startup=startup, n_ts=n_ts, ts_step=ts_step)
    t = lambda x, dt: torch.from_numpy(np.ascontiguousarray(x)).to(dev, dt)
    got = decide_batch(t(b["sizes"], torch.float64), t(util, torch.float64), t(b["chunk"], torch.int32),
                       t(b["prev_q"], torch.int32), t(b["buffer"], torch.float64), t(b["bw_hist"], torch.float64),
                       t(b["hist_len"], torch.int32), H, mode, flags, p_gpu, *st_gpu,
                       startup=None if startup is None else t(startup, torch.uint8), n_ts=n_ts, ts_step=ts_step)
    torch.cuda.synchronize()
    return exp, got, st_cpu, st_gpu


@pytest.mark.parametrize("A,H", [(2, 1), (2, 5), (3, 4), (4, 5), (5, 3), (6, 2), (6, 4), (6, 5), (7, 3), (8, 3), (6, 1)])
@pytest.mark.parametrize("ties", [False, True])
def test_mode0_batch_matches_oracle(A, H, ties):
    rng = np.random.default_rng(1000 * A + H + (7 if ties else 0))
    N, V, K = (2048 if A ** H <= 1300 else 384), 48, 6
    b = _random_batch(rng, N, V, A, K, ties)
    b["chunk"] = rng.integers(0, V - H + 1, size=N).astype(np.int32)
    kw = dict(chunk_length=4.0 if not ties else 1.0, max_buffer=20.0, rebuf_penalty=4.3 if not ties else 1.0,
              smooth_penalty=1.0, utility_scale=1.0)
    exp, got, _, _ = _run_both(b, V, A, K, H, 0, kw)
    assert exp["n_errors"] == 0 and int(got["errors"].item()) == 0
    assert np.array_equal(got["action"].cpu().numpy(), exp["action"])
    assert np.array_equal(got["best_seq"].cpu().numpy(), exp["best_seq"])
    assert bits_equal(got["best_j"].cpu().numpy(), exp["best_J"]) == 0
    assert bits_equal(got["preds"].cpu().numpy(), exp["preds"]) == 0


@pytest.mark.parametrize("A,H", [(2, 4), (4, 5), (6, 3), (6, 5), (7, 2), (8, 3)])
def test_mode1_robust_batch_matches_oracle(A, H):
    rng = np.random.default_rng(2000 * A + H)
    N, V, K = (2048 if A ** H <= 1300 else 384), 48, 5
    b = _random_batch(rng, N, V, A, K, False)
    b["chunk"] = rng.integers(0, V, size=N).astype(np.int32)          # includes truncated horizons
    b["hist_len"][:16] = 0                                              # empty history -> default quality
    b["prev_q"][16:32] = -1                                             # no previous chunk
    kw = dict(chunk_length=4.0, max_buffer=30.0, hist_k=K)
    exp, got, st_cpu, st_gpu = _run_both(b, V, A, K, H, 1, kw, robust_state=True)
    assert np.array_equal(got["action"].cpu().numpy(), exp["action"])
    assert np.array_equal(got["best_seq"].cpu().numpy(), exp["best_seq"])
    ok = ~np.isnan(exp["best_J"])
    assert bits_equal(got["best_j"].cpu().numpy()[ok], exp["best_J"][ok]) == 0
    # predictor state was updated identically
    assert bits_equal(st_gpu[0].cpu().numpy(), st_cpu[0]) == 0
    assert np.array_equal(st_gpu[2].cpu().numpy(), st_cpu[2])
    m = np.minimum(st_cpu[2], K)
    er_g, er_c = st_gpu[1].cpu().numpy(), st_cpu[1]
    for s in range(0, len(m), 37):
        assert bits_equal(er_g[s, :m[s]], er_c[s, :m[s]]) == 0


@pytest.mark.parametrize("A,H", [(2, 5), (4, 3), (6, 5), (6, 1), (8, 2)])
def test_expsmoothing_predictor_matches_oracle(A, H):
    """SPEC §5.4 (mpc.py:72-79): flat forecast of simple exponential smoothing, least-squares initial level."""
    rng = np.random.default_rng(3000 * A + H)
    N, V, K = (1024 if A ** H <= 1300 else 256), 48, 9
    b = _random_batch(rng, N, V, A, K, False)
    b["chunk"] = rng.integers(0, V - H + 1, size=N).astype(np.int32)
    b["hist_len"][:8] = 1                                              # one sample: the level is the sample
    kw = dict(chunk_length=4.0, max_buffer=20.0, utility_scale=1.0)
    exp, got, _, _ = _run_both(b, V, A, K, H, 0, kw, flags=_lib.MPC_PRED_SES)
    assert exp["n_errors"] == 0 and int(got["errors"].item()) == 0
    assert bits_equal(got["preds"].cpu().numpy(), exp["preds"]) == 0
    assert np.array_equal(got["action"].cpu().numpy(), exp["action"])
    assert np.array_equal(got["best_seq"].cpu().numpy(), exp["best_seq"])
    assert bits_equal(got["best_j"].cpu().numpy(), exp["best_J"]) == 0
    p = got["preds"].cpu().numpy()
    assert np.all(p == p[:, :1])                                       # flat
    first = b["bw_hist"][:8, 0]
    assert np.array_equal(p[:8, 0], first)
    # the robust mode has its own predictor
    with pytest.raises(_lib.AbrError):
        _run_both(b, V, A, K, H, 1, kw, flags=_lib.MPC_PRED_SES)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("A,H", [(3, 4), (6, 5), (6, 1), (6, 7)])
def test_startup_phase_decision_matches_oracle(mode, A, H):
    """SPEC §5.3 (f_st of mpc.py:7-18): the start-up delay on a grid as a second decision variable; sessions outside the
    start-up phase take the plain decision; warp-per-session and block-per-session (horizon 7) kernels."""
    rng = np.random.default_rng(4000 * A + 10 * H + mode)
    N, V, K = (512 if A ** H <= 1300 else 96 if H < 7 else 4), 48, 5
    b = _random_batch(rng, N, V, A, K, False)
    b["chunk"] = rng.integers(0, V - H + 1, size=N).astype(np.int32)
    b["buffer"] = np.round(rng.uniform(0, 6, size=N), 9)              # start-up: little buffered
    startup = (rng.random(N) < 0.7).astype(np.uint8)
    kw = dict(chunk_length=4.0, max_buffer=30.0, startup_penalty=0.5, hist_k=K)
    n_ts, step = (6, 0.75) if H < 7 else (3, 1.5)
    exp, got, _, _ = _run_both(b, V, A, K, H, mode, kw, startup=startup, n_ts=n_ts, ts_step=step)
    assert exp["n_errors"] == 0 and int(got["errors"].item()) == 0
    assert np.array_equal(got["action"].cpu().numpy(), exp["action"])
    assert np.array_equal(got["best_seq"].cpu().numpy(), exp["best_seq"])
    assert bits_equal(got["best_j"].cpu().numpy(), exp["best_J"]) == 0
    ts = got["startup_delay"].cpu().numpy()
    assert bits_equal(ts, exp["startup_delay"]) == 0
    assert np.all(ts[startup == 0] == 0.0) and ((ts > 0).any() or N < 96)
    # startup = None means every session; n_ts = 1 is the plain decision bit for bit
    exp_all, got_all, _, _ = _run_both(b, V, A, K, H, mode, kw, n_ts=n_ts, ts_step=step)
    assert bits_equal(got_all["startup_delay"].cpu().numpy(), exp_all["startup_delay"]) == 0
    assert np.array_equal(got_all["action"].cpu().numpy(), exp_all["action"])
    _, plain, _, _ = _run_both(b, V, A, K, H, mode, kw)
    off = startup == 0
    assert np.array_equal(got["action"].cpu().numpy()[off], plain["action"].cpu().numpy()[off])
    assert bits_equal(got["best_j"].cpu().numpy()[off], plain["best_j"].cpu().numpy()[off]) == 0


def test_facade_expsmoothing_and_startup():
    """predict_throughput(method="expsmoothing") (mpc.py:72-79) and next_bitrate_startup() (mpc.py:7-18) of the drop-in
    controller, against the Python restatement of SPEC §5.3 / §5.4."""
    from oracle import mpc_oracle as mo
    from abrsimulator_b200.datamodel import Chunk, MPD, QOEMetric, ChunkInfo
    V = 20
    ladder = [300.0, 750.0, 1200.0, 1850.0, 2850.0, 4300.0]
    chunks = [Chunk(list(ladder), [b * 4.0 for b in ladder]) for _ in range(V)]
    mpd = MPD(V, 4.0, 30.0, 8.0, chunks)
    qoe = QOEMetric(4.3, 1.0, 1.0, 0.0)             # rebuffer, variance, startup, latency weights
    hist = [1800.0, 2500.0, 900.0, 3100.0, 2200.0]

    class Player:
        def __init__(self):
            self.info = ChunkInfo(2, 1, list(hist), 1.0)

        def get_mpd(self):
            return mpd

        def get_qoe_metric(self):
            return qoe

        def get_next_chunk_info(self):
            return self.info

    ctl = MPCBitrateController(Player(), horizon=4, strict_history=False)
    p = ctl.predict_throughput(4, list(hist), method="expsmoothing")
    assert isinstance(p, np.ndarray) and list(p) == mo.predict_ses(4, hist)
    with pytest.raises(ValueError):
        ctl.predict_throughput(4, list(hist), method="arima")
    # a controller that decides with that predictor
    ctl_s = MPCBitrateController(Player(), horizon=4, predictor="expsmoothing")
    bitrates = [list(ladder)] * V
    sizes = [[b * 4.0 for b in ladder]] * V
    r = mo.decide_ref_ses(2, 1, 1.0, hist, 4, bitrates, sizes, 4.0, 30.0, 1.0, 4.3)
    assert ctl_s.next_bitrate() == r["action"]
    assert ctl_s.predicted_bandwidths == r["preds"]
    assert ctl_s.player.info.previous_bandwidths == hist             # the expsmoothing branch does not mutate (D10 is harmonic's)
    # start-up branch
    act, ts = ctl.next_bitrate_startup(n_ts=12, ts_step=0.5)
    preds, _ = mo.predict_harmonic_ref(4, hist)
    obj = lambda R, b0: mo.objective_ref(R, 2, 1, b0, preds, bitrates, sizes, 4.0, 30.0, 1.0, 4.3)
    want = mo.decide_startup(obj, 6, 4, 1.0, 1.0, 12, 0.5)
    assert (act, ts) == (want["action"], want["startup_delay"]) and ts > 0.0


def test_horizon7_block_per_session():
    """BASELINE config 4 shape: A=6, H=7 (279 936 sequences), few sessions -> one block per session."""
    rng = np.random.default_rng(77)
    N, V, K, A, H = 6, 48, 5, 6, 7
    b = _random_batch(rng, N, V, A, K, False)
    b["chunk"] = rng.integers(0, V - H + 1, size=N).astype(np.int32)
    for mode in (0, 1):
        exp, got, _, _ = _run_both(b, V, A, K, H, mode, dict(chunk_length=4.0, max_buffer=60.0))
        assert np.array_equal(got["action"].cpu().numpy(), exp["action"])
        assert np.array_equal(got["best_seq"].cpu().numpy(), exp["best_seq"])
        assert bits_equal(got["best_j"].cpu().numpy(), exp["best_J"]) == 0


def test_error_flags_and_lenient_flags():
    rng = np.random.default_rng(5)
    N, V, K, A, H = 64, 12, 4, 4, 3
    b = _random_batch(rng, N, V, A, K, False)
    b["chunk"] = rng.integers(0, V - H + 1, size=N).astype(np.int32)
    b["chunk"][0] = V - 1            # k + H > V      -> IndexError in the reference
    b["hist_len"][1] = 0             # empty history  -> ZeroDivisionError
    b["bw_hist"][2, 0] = 0.0         # zero sample    -> ZeroDivisionError
    b["hist_len"][2] = max(b["hist_len"][2], 1)
    b["prev_q"][3] = A               # out-of-range index
    exp, got, _, _ = _run_both(b, V, A, K, H, 0, dict(max_buffer=20.0))
    act = got["action"].cpu().numpy()
    assert list(act[:4]) == [-1, -1, -1, -1] and exp["n_errors"] == 4 and int(got["errors"].item()) == 4
    assert np.array_equal(act, exp["action"])
    # lenient flags: truncate the horizon / default quality instead of flagging
    _, got2, _, _ = _run_both(b, V, A, K, H, 0, dict(max_buffer=20.0), flags=_lib.MPC_TRUNCATE | _lib.MPC_EMPTY_DEFAULT)
    act2 = got2["action"].cpu().numpy()
    assert act2[0] >= 0 and act2[1] == 1 and act2[2] == -1 and act2[3] == -1
    assert int(got2["errors"].item()) == 2


def test_shape_limits_are_rejected():
    dev = torch.device("cuda")
    z = lambda *s, dt=torch.float64: torch.zeros(*s, dtype=dt, device=dev)
    with pytest.raises(_lib.AbrError):
        decide_batch(z(4, 4), z(4, 4), z(1, dt=torch.int32), z(1, dt=torch.int32), z(1), z(1, 5),
                     z(1, dt=torch.int32), 9, 0)


@pytest.mark.parametrize("A,H", [(2, 4), (3, 5), (4, 5), (6, 4), (6, 5), (5, 6), (8, 4), (6, 2), (6, 3)])
@pytest.mark.parametrize("ties,vw", [(False, 1.0), (True, 1.0), (False, 0.6), (True, 2.0)])
def test_branch_and_bound_equals_exhaustive_enumeration(A, H, ties, vw):
    """Robust mode skips the partial sequences whose bound already loses (SPEC §5.5): the decision, the whole best
    sequence (first minimum in C order — ladders with exact ties included) and the objective value are those of the
    exhaustive enumeration (ABR_MPC_EXHAUSTIVE) and of the oracle, for the compacted form (A^(H-2) <= 256 prefixes),
    the plain form (larger shapes, H < 4) and penalties other than 1."""
    rng = np.random.default_rng(31 * A + H + (5 if ties else 0) + int(10 * vw))
    N, V, K = (3072 if A ** H <= 8000 else 512), 48, 5
    b = _random_batch(rng, N, V, A, K, ties)
    b["chunk"] = rng.integers(0, V, size=N).astype(np.int32)
    b["prev_q"][:64] = -1
    kw = dict(chunk_length=4.0 if not ties else 1.0, max_buffer=30.0, hist_k=K, smooth_penalty=vw,
              rebuf_penalty=4.3 if not ties else 1.0, utility_scale=1.0)
    exp, got, _, _ = _run_both(b, V, A, K, H, 1, kw)
    _, got_x, _, _ = _run_both(b, V, A, K, H, 1, kw, flags=_lib.MPC_EXHAUSTIVE)
    for g in (got, got_x):
        assert np.array_equal(g["action"].cpu().numpy(), exp["action"])
        assert np.array_equal(g["best_seq"].cpu().numpy(), exp["best_seq"])
        ok = ~np.isnan(exp["best_J"])
        assert bits_equal(g["best_j"].cpu().numpy()[ok], exp["best_J"][ok]) == 0


def test_branch_and_bound_is_switched_off_for_negative_penalties():
    """The bounds need non-negative penalties; with a negative one the search enumerates and still matches the oracle."""
    rng = np.random.default_rng(77)
    N, V, A, K, H = 1024, 48, 6, 5, 5
    b = _random_batch(rng, N, V, A, K, False)
    b["chunk"] = rng.integers(0, V - H, size=N).astype(np.int32)
    kw = dict(chunk_length=4.0, max_buffer=30.0, hist_k=K, smooth_penalty=-0.5, rebuf_penalty=4.3, utility_scale=1.0)
    exp, got, _, _ = _run_both(b, V, A, K, H, 1, kw)
    assert np.array_equal(got["best_seq"].cpu().numpy(), exp["best_seq"])
    assert bits_equal(got["best_j"].cpu().numpy(), exp["best_J"]) == 0


def test_env_mpc_decide_exhaustive_flag():
    from abrsimulator_b200 import synth
    from abrsimulator_b200.env import BatchedABREnv
    N = 8192
    bitrates, sizes = synth.make_video(48)
    bw, tl, ti = synth.make_traces(64, 256)
    tid, off = synth.make_sessions(N, 64, 256, group=64)
    env = BatchedABREnv(bw, sizes, bitrates, N, trace_len=tl, trace_interval=ti, track_history=1)
    env.reset(tid, off)
    env.rollout("random", 11, seed=3, want=())
    snap = {f: env.state(f).clone() for f in ("last_pred", "err_ring", "err_len")}     # a decision updates the predictor
    a, ja = env.mpc_decide(5, "robust", want_score=True)
    for f, x in snap.items():
        env.state(f).copy_(x)
    b, jb = env.mpc_decide(5, "robust", want_score=True, exhaustive=True)
    assert torch.equal(a, b) and torch.equal(ja, jb)
