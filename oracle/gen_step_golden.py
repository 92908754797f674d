"""Generate the chunk-step SPEC fixture — TEST INFRASTRUCTURE.

    python oracle/gen_step_golden.py        ->  tests/golden/step_spec_golden.json

The reference's environment loop (Simulator.py:93-210) does not run, so there is nothing to
import: this fixture is produced by the repo's own pure-Python restatement of SPEC.md §2-§4 and §7
(oracle/step_oracle.py) and pins *the SPEC* — a change of the arithmetic contract must regenerate it on purpose.
(What pins the step's dynamics to the reference is the other fixture, tests/golden/sim_ref_tick_golden.json, produced
by the reference's own repaired tick loop: oracle/make_ref_simulator.py.)  Values are stored as IEEE-754 hex strings,
so the C oracle and the CUDA kernels are compared bit for bit.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from abrsimulator_b200 import synth          # noqa: E402
from oracle import step_oracle as so         # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "step_spec_golden.json")
KEYS = ("delay", "sleep", "buffer", "rebuf", "reward", "throughput", "latency")

PARAMS = dict(chunk_length=4.0, max_buffer=24.0, rtt=0.08, payload=0.95, sleep_quantum=0.5, rebuf_penalty=4.3,
              smooth_penalty=1.0, utility_scale=0.001, bba_reservoir=5.0, bba_cushion=10.0, start_up_length=0.0,
              startup_penalty=0.0, latency_penalty=0.0, latency_tick=0.01, utility_mode=0, default_quality=1, auto_reset=1,
              hist_k=5, track_history=1, live=0, smooth_prev_ladder=0)


def world(seed):
    """Small ragged world: 5 traces (lengths 1, 7, 40, 64, 64; intervals 0.5 / 1 / 2 / 0.3 / 1), 10-chunk video."""
    bitrates, sizes = synth.make_video(10, seed=seed)
    if seed >= 3:   # a ladder that differs from chunk to chunk (the smoothness term then depends on whose ladder is read)
        bitrates = np.ascontiguousarray(bitrates * np.random.default_rng(300 + seed).uniform(0.8, 1.2, size=bitrates.shape))
    bw, _, _ = synth.make_traces(5, 64, seed=77 + seed)
    tl = np.array([1, 7, 40, 64, 64], np.int32)
    ti = np.array([0.5, 1.0, 2.0, 0.3, 1.0])
    return bitrates, sizes, bw, tl, ti


def run_case(name, params, seed, n_sessions=20, steps=25, speeds=None):
    bitrates, sizes, bw, tl, ti = world(seed)
    rng = np.random.default_rng(1000 + seed)
    tid = rng.integers(0, 5, size=n_sessions)
    off = rng.uniform(0.0, 150.0, size=n_sessions)
    off[:3] = [0.0, float(ti[tid[1]]) * 3, 1e-9]
    actions = rng.integers(0, bitrates.shape[1], size=(steps, n_sessions))
    util = (bitrates * params["utility_scale"]).tolist()
    out = {k: [] for k in KEYS}
    out["eov"] = []
    sessions = [so.Session(bw[tid[s], :tl[tid[s]]].tolist(), float(ti[tid[s]]), sizes.tolist(), util, params,
                           float(off[s])) for s in range(n_sessions)]
    V = sizes.shape[0]
    for sess in sessions:      # live mode: content chunk k plays at speeds[k mod len(speeds)] in every session
        sess.speed = None if speeds is None else [float(speeds[k % len(speeds)]) for k in range(V)]
    for t in range(steps):
        row = {k: [] for k in KEYS}
        eov = []
        for s, sess in enumerate(sessions):
            r = sess.step(int(actions[t, s]))
            for k in KEYS:
                row[k].append(float(r.get(k, 0.0)).hex())
            eov.append(int(r["eov"]))
        for k in KEYS:
            out[k].append(row[k])
        out["eov"].append(eov)
    final = dict(seg=[s.seg for s in sessions], phase=[float(s.phi).hex() for s in sessions],
                 pos=[float(s.pos).hex() for s in sessions], buffer=[float(s.buffer).hex() for s in sessions], chunk=[s.chunk for s in sessions],
                 play_id=[s.play_id for s in sessions], play_len=[float(s.play_len).hex() for s in sessions],
                 play_time=[float(s.play_time).hex() for s in sessions])
    return dict(name=name, params=params, seed=seed, trace_id=tid.tolist(), start_offset=[float(x).hex() for x in off],
                actions=actions.tolist(), speeds=speeds, outputs=out, final=final)


def main():
    cases = [run_case("on_demand", dict(PARAMS), 0),
             run_case("non_pow2_interval_and_quantum", dict(PARAMS, sleep_quantum=0.3, max_buffer=12.0), 1),
             run_case("live", dict(PARAMS, live=1, start_up_length=8.0, max_buffer=16.0, latency_penalty=0.05,
                                   startup_penalty=1.0), 2, speeds=[1.0, 1.25, 0.75, 1.5]),
             run_case("live_own_ladders", dict(PARAMS, live=1, start_up_length=4.0, max_buffer=12.0, latency_penalty=0.05,
                                               startup_penalty=1.0, smooth_prev_ladder=1), 3, speeds=[0.8, 1.0, 1.3]),
             run_case("on_demand_own_ladders", dict(PARAMS, smooth_prev_ladder=1), 4)]
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(dict(generator="oracle/gen_step_golden.py (oracle/step_oracle.py, SPEC.md §2-§4, §7)",
                       pins="the SPEC bit for bit; the reference-derived pin of the dynamics is sim_ref_tick_golden.json", cases=cases), f)
    print(f"wrote {OUT}: {len(cases)} cases")


if __name__ == "__main__":
    main()
