"""Time the UNMODIFIED reference controller (mpc.py + scipy.optimize.brute) on this machine — TEST INFRASTRUCTURE.

    python oracle/time_reference_mpc.py [decisions_per_process]   ->  profiles/ref_mpc_cpu_baseline.json

The reference is pure Python and lives only in the build container (/root/reference does not exist on the GPU box), so
bench.py's CPU legs run the port in oracle/ there.  This script is the build-box measurement of the real thing
(BASELINE.md §4.1): the reference's own ``MPCBitrateController.next_bitrate()`` on bench.py's MPC
workload (48-chunk video, 6 bitrates 300-4300 kbps, horizon 5 = 7 776 sequences per decision, five history samples),
on one core and on every core of this machine, next to the port bench.py times (oracle/mpc_oracle.py) on the same
inputs — the committed record lets a reader scale bench.py's port figure to the reference's.
"""
from __future__ import annotations

import contextlib
import io
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "profiles", "ref_mpc_cpu_baseline.json")
V, A, H = 48, 6, 5


def _inputs(seed, count):
    from abrsimulator_b200 import synth
    bitrates, sizes = synth.make_video(V)
    rng = np.random.default_rng(1000 + seed)
    jobs = [dict(hist=[float(x) for x in rng.uniform(0.2, 6.0, size=5)], k=int(rng.integers(0, V - H)),
                 prev_q=int(rng.integers(0, A)), buffer=float(rng.uniform(0, 30))) for _ in range(count)]
    return (bitrates * 0.001).tolist(), sizes.tolist(), jobs


def _reference_worker(job):
    seed, count = job
    from oracle.gen_golden import import_reference_mpc, make_player
    mpc = import_reference_mpc()
    util, sizes, jobs = _inputs(seed, count)
    t0 = time.perf_counter()
    acts = []
    for j in jobs:
        sc = dict(bitrates=util, sizes=sizes, chunk_length=4.0, max_buffer=60.0, rw=4.3, vw=1.0, k=j["k"],
                  prev_q=j["prev_q"], history=list(j["hist"]), buffer=j["buffer"], H=H)
        abr = mpc.MPCBitrateController(make_player(sc))
        abr.horizon = H
        with contextlib.redirect_stdout(io.StringIO()):
            acts.append(int(abr.next_bitrate()))
    return count, time.perf_counter() - t0, acts


def _port_worker(job):
    seed, count = job
    from oracle import mpc_oracle as mo
    util, sizes, jobs = _inputs(seed, count)
    t0 = time.perf_counter()
    acts = []
    for j in jobs:
        r = mo.decide_ref(j["k"], j["prev_q"], j["buffer"], j["hist"], H, util, sizes, 4.0, 60.0, 1.0, 4.3)
        acts.append(r["action"])
    return count, time.perf_counter() - t0, acts


def _pool(worker, procs, count):
    jobs = [(i, count) for i in range(procs)]
    if procs == 1:
        res = [worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(worker, jobs)
    n = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return n / busy, n, busy, [r[2] for r in res]


def main(count=24):
    cores = os.cpu_count() or 1
    r1, n1, b1, a_ref = _pool(_reference_worker, 1, count)
    rN, nN, bN, _ = _pool(_reference_worker, cores, count)
    p1, m1, c1, a_port = _pool(_port_worker, 1, count)
    rec = dict(script="oracle/time_reference_mpc.py", machine="build container (no GPU)", cores=cores,
               workload=f"robust-MPC-shaped inputs of bench.py: V={V}, A={A}, horizon {H} ({A ** H} sequences per decision), "
                        "5 history samples; reference semantics (mode 0)",
               reference_mpc_py=dict(impl="/root/reference/mpc.py unmodified + scipy.optimize.brute "
                                          f"(scipy {__import__('scipy').__version__})",
                                     one_core=dict(decisions_per_s=r1, decisions=n1, seconds=b1),
                                     all_cores=dict(decisions_per_s=rN, processes=cores, decisions=nN, seconds=bN)),
               port_mpc_oracle_py=dict(impl="oracle/mpc_oracle.py decide_ref (what bench.py's cpu_baseline times on the GPU box)",
                                       one_core=dict(decisions_per_s=p1, decisions=m1, seconds=c1)),
               port_over_reference=p1 / r1,
               same_actions=bool(a_ref[0] == a_port[0]))
    with open(OUT, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 24)
