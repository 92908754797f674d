"""The reference's own test scenario (mpc_test.py:52-86) through the drop-in controller.

``python examples/mpc_dropin.py`` prints ``Test next bitrate: 2`` — what the reference prints for the same player.
The scenario (ladder, sizes, history, buffer, horizon) is read from the fixture generated from the unmodified
reference (tests/golden/mpc_ref_golden.json, case "mpc_test"); the controller runs the search on the GPU through
libabr_b200's C-ABI (abr_mpc_decide_host).  Only the import line differs from reference-side code:

    import abrsimulator_b200.mpc as mpc        # was: import mpc
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import abrsimulator_b200.mpc as mpc
from abrsimulator_b200 import Chunk, ChunkInfo, MPD, QOEMetric


class Player:
    """The three getters MPCBitrateController asks its player for (mpc.py:56-57,166,184)."""

    def __init__(self, mpd, qoe_metric, chunk_info):
        self._mpd, self._qoe, self._info = mpd, qoe_metric, chunk_info

    def get_mpd(self):
        return self._mpd

    def get_qoe_metric(self):
        return self._qoe

    def get_next_chunk_info(self):
        return self._info


def reference_scenario():
    with open(os.path.join(ROOT, "tests", "golden", "mpc_ref_golden.json")) as f:
        sc = json.load(f)["cases"][0]["scenario"]
    chunks = [Chunk(list(b), list(s)) for b, s in zip(sc["bitrates"], sc["sizes"])]
    mpd = MPD(len(chunks), sc["chunk_length"], sc["max_buffer"], chunks)
    player = Player(mpd, QOEMetric(sc["rw"], sc["vw"], 0), ChunkInfo(sc["k"], sc["prev_q"], list(sc["history"]), sc["buffer"]))
    return player, sc["H"]


def main():
    player, horizon = reference_scenario()
    abr = mpc.MPCBitrateController(player)
    abr.horizon = horizon
    choice = abr.next_bitrate()
    print("Test next bitrate: {}".format(choice))
    return choice


if __name__ == "__main__":
    main()
