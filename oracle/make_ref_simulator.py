"""Run the reference's OWN tick loop — TEST INFRASTRUCTURE.

    python oracle/make_ref_simulator.py            ->  tests/golden/sim_ref_tick_golden.json

``/root/reference/Simulator.py`` does not run as shipped.  This script reads that file where
it lies, applies the textual repairs of SURVEY.md §3.2 *programmatically* (nothing is copied into the repository: the
repaired text exists only in this process's memory), executes ``Simulator.run()`` — the reference's own per-tick
loop, Simulator.py:135-210 — over scripted ABR / speed controllers, and commits what it produced as a fixture.  The
chunk-step kernels are closed forms of that loop; the fixture lets ``tests/`` check them against the reference's code
itself, within the loop's own discretisation.  Every scenario is run twice: with the reference's tick (0.01 s,
``reference``) and with a tick of 0.001 s (``reference_fine``, one more textual edit: the literal on Simulator.py:133),
so that the tests can also check that the loop *converges* to the closed form as the tick shrinks.

Repairs (each is an edit of the reference text, located by its content, so a changed reference fails loudly):
  D1  the ``return self.calculate_qoe(...)`` at :210 is dedented out of the ``while`` body;
  D2  ``download_pause`` / ``play_pause`` (:144-145, :148-149) get the missing ``else: ... = False``;
  D3  ``self.mpd.chunks.bitrates[...]`` (:82, :156) index the chunk list first;
  D6  the bandwidth index (:158-159) wraps around at the end of the trace.
One probe line is added after ``chunk_id += 1`` (:166): it appends the loop's local timers to ``self.ref_log`` and
changes nothing the loop computes.  ``tick`` replaces the literal of ``dt = 0.01`` (:133) for the convergence runs.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ABR_REFERENCE_DIR", "/root/reference")
GOLDEN = os.path.join(ROOT, "tests", "golden", "sim_ref_tick_golden.json")


def _replace_once(text, old, new, what):
    if text.count(old) != 1:
        raise RuntimeError(f"repair {what}: expected exactly one occurrence of {old!r} in the reference, found {text.count(old)}")
    return text.replace(old, new)


def repaired_source(wrap_trace=True, tick=None) -> str:
    src = open(os.path.join(REF, "Simulator.py")).read()
    if tick is not None:
        src = _replace_once(src, "        dt = 0.01\n", f"        dt = {tick!r}\n", "tick")
    # D1: the return statement sits inside the while body (12 spaces) -> function level (8 spaces)
    src = _replace_once(src, "\n            return self.calculate_qoe(rebuffer_time, previous_bitrates, start_up_time, average_latency)",
                        "\n        return self.calculate_qoe(rebuffer_time, previous_bitrates, start_up_time, average_latency)", "D1")
    # D2: the pause flags are only ever set
    src = _replace_once(src, "                download_pause = True\n",
                        "                download_pause = True\n            else:\n                download_pause = False\n", "D2a")
    src = _replace_once(src, "                play_pause = True\n",
                        "                play_pause = True\n            else:\n                play_pause = False\n", "D2b")
    # D3: chunks is a list of Chunk
    src = _replace_once(src, "self.mpd.chunks.bitrates[current_bitrate] * self.mpd.chunk_length",
                        "self.mpd.chunks[chunk_id].bitrates[current_bitrate] * self.mpd.chunk_length", "D3a")
    src = _replace_once(src, "abs(self.mpd.chunks.bitrates[previous_bitrates[i]] - self.mpd.chunks.bitrates[previous_bitrates[i + 1]])",
                        "abs(self.mpd.chunks[i].bitrates[previous_bitrates[i]] - self.mpd.chunks[i + 1].bitrates[previous_bitrates[i + 1]])",
                        "D3b")
    if wrap_trace:   # D6
        src = _replace_once(src, "bandwidth = self.network_info.bandwidths[bandwidth_idx]",
                            "bandwidth = self.network_info.bandwidths[bandwidth_idx % len(self.network_info.bandwidths)]", "D6")
    # probe: the loop's timers at the moment a chunk completes
    src = _replace_once(src, "                    chunk_id += 1\n",
                        "                    chunk_id += 1\n"
                        "                    self.ref_log.append(dict(t=global_time, download_time=download_time, "
                        "rebuffer_time=rebuffer_time, start_up_time=start_up_time, average_latency=average_latency, "
                        "play_time=play_time, buffer_level=buffer_level, play_id=play_id))\n", "probe")
    return src


def load_repaired(wrap_trace=True, tick=None):
    """The repaired reference module, compiled from memory (no copy of the reference's text is written anywhere)."""
    import types
    mod = types.ModuleType("Simulator_repaired")
    exec(compile(repaired_source(wrap_trace, tick), os.path.join(REF, "Simulator.py") + " (repaired in memory)", "exec"),
         mod.__dict__)
    return mod


class ScriptedAbr:
    """``get_next_bitrate`` of the controller protocol (Simulator.py:155): a fixed action per chunk."""

    def __init__(self, actions):
        self.actions = list(actions)
        self.calls = []

    def get_next_bitrate(self, chunk_id, previous_bitrates, previous_bandwidths, buffer_level):
        self.calls.append((chunk_id, float(buffer_level)))
        return int(self.actions[chunk_id])


class ScriptedSpeed:
    """``get_next_speed`` (Simulator.py:177): the k-th call returns speeds[k mod len]."""

    def __init__(self, speeds):
        self.speeds = list(speeds)
        self.n = 0

    def get_next_speed(self):
        v = self.speeds[self.n % len(self.speeds)]
        self.n += 1
        return float(v)


LADDERS = ([300.0, 750.0, 1200.0, 1850.0, 2850.0, 4300.0], [1.0, 2.5, 5.0, 8.0], [400.0, 1000.0, 2500.0, 6000.0, 9000.0])


def scenario(i):
    """Scenario i -> plain dict of inputs (everything the test needs to rebuild it)."""
    rng = np.random.default_rng(5000 + i)
    ladder = LADDERS[i % len(LADDERS)]
    V = int(rng.integers(6, 25))
    L = float(rng.choice([1.0, 2.0, 4.0]))
    interval = float(rng.choice([0.5, 1.0, 2.0]))
    T = int(rng.integers(40, 400))
    mean = ladder[len(ladder) // 2] * float(rng.uniform(0.5, 2.5))
    bw = np.clip(mean * np.exp(rng.normal(0.0, 0.6, size=T)), ladder[0] * 0.3, ladder[-1] * 3.0)
    per_chunk = (i % 4 == 3)      # a ladder that varies from chunk to chunk (exercises the D3 indexing of calculate_qoe)
    bitrates = [[b * (float(rng.uniform(0.8, 1.2)) if per_chunk else 1.0) for b in ladder] for _ in range(V)]
    max_buffer = float(rng.choice([2.0, 3.0, 5.0, 8.0])) * L
    start_up = float(rng.choice([1.0, 2.0])) * L
    speeds = [[1.0], [1.0], [1.25], [1.0, 1.25, 0.8], [0.75, 1.5]][i % 5]
    actions = rng.integers(0, len(ladder), size=V).tolist()
    weights = [float(rng.choice([1.0, 4.3])), float(rng.choice([0.0, 1.0])), float(rng.choice([0.0, 1.0, 2.0])),
               float(rng.choice([0.0, 0.05]))]
    return dict(index=i, V=V, chunk_length=L, interval=interval, bandwidths=[float(x) for x in bw], bitrates=bitrates,
                max_buffer=max_buffer, start_up_length=start_up, speeds=speeds, actions=actions, weights=weights)


def run_reference(mod, sc):
    abr, spd = ScriptedAbr(sc["actions"]), ScriptedSpeed(sc["speeds"])
    sim = mod.Simulator(abr, spd)
    sim.set_qoe_metric(mod.QOEMetric(*sc["weights"]))
    sim.network_info = mod.NetworkInfo(sc["interval"], list(sc["bandwidths"]))
    sim.mpd = mod.MPD(sc["V"], sc["chunk_length"], sc["max_buffer"], sc["start_up_length"],
                      [mod.Chunk(list(b)) for b in sc["bitrates"]])
    sim.ref_log = []
    qoe = sim.run()
    log = sim.ref_log
    last = log[-1]
    return dict(qoe=float(qoe), speed_calls=spd.n, abr_buffer=[b for _, b in abr.calls],
                per_chunk={k: [float(r[k]) for r in log] for k in
                           ("t", "download_time", "rebuffer_time", "start_up_time", "average_latency", "play_time",
                            "buffer_level")},
                play_id=[int(r["play_id"]) for r in log],
                final=dict(rebuffer_time=float(last["rebuffer_time"]), start_up_time=float(last["start_up_time"]),
                           average_latency=float(last["average_latency"])))


def main(n=60):
    mod = load_repaired(wrap_trace=True)
    fine = load_repaired(wrap_trace=True, tick=0.001)
    cases = []
    for i in range(n):
        sc = scenario(i)
        sc["reference"] = run_reference(mod, sc)
        f = run_reference(fine, sc)
        sc["reference_fine"] = dict(qoe=f["qoe"], final=f["final"], speed_calls=f["speed_calls"],
                                    per_chunk={k: f["per_chunk"][k] for k in ("t", "rebuffer_time", "start_up_time", "play_time")})
        cases.append(sc)
    with open(GOLDEN, "w") as f:
        json.dump(dict(generator="oracle/make_ref_simulator.py: /root/reference/Simulator.py with repairs D1, D2, D3, D6 applied "
                                 "programmatically, executed here (tick dt = 0.01 s; reference_fine: dt = 0.001 s)",
                       dt=0.01, dt_fine=0.001, cases=cases), f)
    print(f"wrote {GOLDEN}: {len(cases)} scenarios from the reference's own tick loop")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 60)
