"""10 240 more decisions of the UNMODIFIED reference controller — TEST INFRASTRUCTURE.

    python oracle/gen_golden_bulk.py [n]      ->  tests/golden/mpc_ref_bulk.json

Companion of oracle/gen_golden.py (same import of /root/reference/mpc.py under the statsmodels stub, same player
protocol, same scenario generator).  Scenario i is ``random_scenario(default_rng(777000 + i), i)``: the fixture stores
only the index and what the reference answered — best sequence (``scipy.optimize.brute``'s first minimum), its
objective value and the first prediction, as IEEE-754 hex — so the tests rebuild the inputs from the index and compare
the oracle (CPU) and the kernels (GPU) with the reference itself on every one of them.
Runs only in the build container (needs /root/reference); a few minutes on one core.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import time

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from oracle.gen_golden import import_reference_mpc, make_player, random_scenario   # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "mpc_ref_bulk.json")
SEED0 = 777000


def bulk_scenario(i):
    return random_scenario(np.random.default_rng(SEED0 + i), i)


def main(n=10240):
    mpc = import_reference_mpc()
    seqs, js, p0 = [], [], []
    t0 = time.perf_counter()
    for i in range(n):
        sc = bulk_scenario(i)
        player = make_player(sc)
        abr = mpc.MPCBitrateController(player)
        abr.horizon = sc["H"]
        abr.update_bandwidth_prediction()
        info = player.get_next_chunk_info()
        with contextlib.redirect_stdout(io.StringIO()):
            best = abr.optimize_qoe(info)
        best = [int(x) for x in np.atleast_1d(best)]
        seqs.append("".join(str(b) for b in best))
        js.append(float(abr.objective(best, info)).hex())
        p0.append(float(abr.predicted_bandwidths[0]).hex())
    dt = time.perf_counter() - t0
    with open(OUT, "w") as f:
        json.dump(dict(generator="oracle/gen_golden_bulk.py", reference="Elliotshui/ABRSimulator mpc.py (unmodified, "
                       "statsmodels stubbed)", scipy=__import__("scipy").__version__, numpy=np.__version__,
                       seed0=SEED0, n=n, best_seq=seqs, best_J=js, pred0=p0), f, separators=(",", ":"))
    print(f"wrote {OUT}: {n} decisions of the reference in {dt:.0f} s, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 10240)
