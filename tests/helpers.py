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


def load_step_golden():
    """tests/golden/step_spec_golden.json (oracle/gen_step_golden.py): cases with inputs rebuilt and outputs as float64."""
    import json
    import os
    from oracle.gen_step_golden import world, KEYS
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_spec_golden.json")
    with open(path) as f:
        doc = json.load(f)
    cases = []
    for c in doc["cases"]:
        bitrates, sizes, bw, tl, ti = world(c["seed"])
        unhex = lambda rows: np.array([[float.fromhex(x) for x in r] for r in rows])
        cases.append(dict(name=c["name"], params=c["params"], bitrates=bitrates, sizes=sizes, bw=bw, tl=tl, ti=ti,
                          trace_id=np.array(c["trace_id"], np.int32),
                          start_offset=np.array([float.fromhex(x) for x in c["start_offset"]]),
                          actions=np.array(c["actions"], np.int32), speeds=c["speeds"],
                          outputs={k: unhex(c["outputs"][k]) for k in KEYS}, eov=np.array(c["outputs"]["eov"], np.uint8),
                          final=dict(seg=np.array(c["final"]["seg"], np.int32), chunk=np.array(c["final"]["chunk"], np.int32),
                                     phase=np.array([float.fromhex(x) for x in c["final"]["phase"]]),
                                     pos=np.array([float.fromhex(x) for x in c["final"]["pos"]]),
                                     buffer=np.array([float.fromhex(x) for x in c["final"]["buffer"]]),
                                     play_id=np.array(c["final"]["play_id"], np.int32),
                                     play_len=np.array([float.fromhex(x) for x in c["final"]["play_len"]]),
                                     play_time=np.array([float.fromhex(x) for x in c["final"]["play_time"]]))))
    return cases


def speed_table(speeds, V, N):
    """The fixture's playback speeds as the [V, N] table of SPEC §7: content chunk k plays at speeds[k mod len]."""
    if speeds is None:
        return None
    return np.ascontiguousarray(np.repeat(np.array([speeds[k % len(speeds)] for k in range(V)], np.float64)[:, None], N, 1))


# ---- tests/golden/sim_ref_tick_golden.json: the reference's own tick loop (oracle/make_ref_simulator.py) ----
def load_ref_tick_golden():
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sim_ref_tick_golden.json")
    with open(path) as f:
        return json.load(f)


def ref_tick_params(sc, tick):
    """Parameters under which SPEC §7 is the closed form of the reference's loop (Simulator.py:135-210): no RTT, no
    payload factor, sizes = bitrate * chunk_length (:156), cost in bitrate units with each chunk's own ladder (:81-82),
    no smoothness term for the first chunk, one session, no restart."""
    w = sc["weights"]          # QOEMetric(rebuffer_weight, variance_weight, startup_weight, latency_weight)
    return dict(chunk_length=sc["chunk_length"], max_buffer=sc["max_buffer"], rtt=0.0, payload=1.0, rebuf_penalty=w[0],
                smooth_penalty=w[1], utility_scale=1.0, start_up_length=sc["start_up_length"], startup_penalty=w[2],
                latency_penalty=w[3], latency_tick=tick, default_quality=-1, auto_reset=0, live=1, smooth_prev_ladder=1)


def ref_tick_world(sc):
    """(bitrates[V][A], sizes[V][A], bw[1][T], speed[V][1]) of a scenario."""
    br = np.array(sc["bitrates"], np.float64)
    bw = np.array(sc["bandwidths"], np.float64)[None, :]
    speed = np.array([sc["speeds"][k % len(sc["speeds"])] for k in range(sc["V"])], np.float64)[:, None]
    return br, np.ascontiguousarray(br * sc["chunk_length"]), bw, np.ascontiguousarray(speed)


def check_against_ref_tick(sc, name, tick, t, rebuf, startup, play_time, smooth, area, played, speed_calls):
    """Cumulative per-chunk timers of a closed-form run (arrays of V) against the reference loop run with `tick`.
    The loop quantises every event (a download ends, playback starts, the buffer runs dry, the live edge arrives) to
    its tick, and a shifted event moves the ones after it, so the bound grows with the number of chunks:
    2 ticks per chunk + 2.  It is the same bound, in ticks, for the reference's own 0.01 s tick and for 0.001 s —
    the loop converges to the closed form at first order."""
    ref = sc[name]
    pc = ref["per_chunk"]
    V = sc["V"]
    bound = (2 * V + 2) * tick
    dev = dict(t=np.abs(t - np.array(pc["t"])).max(), rebuffer=np.abs(rebuf - np.array(pc["rebuffer_time"])).max(),
               startup=np.abs(startup - np.array(pc["start_up_time"])).max(),
               play_time=np.abs(play_time - np.array(pc["play_time"])).max())
    for k, d in dev.items():
        assert d <= bound, (sc["index"], name, k, d, bound)
    w = sc["weights"]
    avg = area / (tick * played) if played > 0 else 0.0       # Simulator.py:179-180, see SPEC §7
    ref_avg = ref["final"]["average_latency"]
    assert abs(avg - ref_avg) <= 0.01 * (tick / 0.01) * max(ref_avg, 1.0), (sc["index"], name, avg, ref_avg)
    cost = w[0] * rebuf[-1] + w[1] * smooth + w[2] * startup[-1] + w[3] * avg
    # the variance term of calculate_qoe (Simulator.py:80-82, each chunk's own ladder) carries no discretisation error
    b, q = sc["bitrates"], sc["actions"]
    variance = 0.0
    for i in range(V - 1):
        variance += abs(b[i][q[i]] - b[i + 1][q[i + 1]])
    assert abs(smooth - variance) <= 1e-12 * max(variance, 1.0), (sc["index"], name, smooth, variance)
    assert abs(cost - ref["qoe"]) <= w[0] * bound + w[2] * bound + w[3] * 0.01 * (tick / 0.01) * max(ref_avg, 1.0) + \
        1e-9 * abs(ref["qoe"]), (sc["index"], name, cost, ref["qoe"])
    assert abs(speed_calls - ref["speed_calls"]) <= 1, (sc["index"], name, speed_calls, ref["speed_calls"])
    dev["avg_rel"] = abs(avg - ref_avg) / max(ref_avg, 1.0)
    return dev
