"""Differential soak: the randomized-worlds tests of tests/test_gpu_env.py with many more seeds than the suite runs.
usage: python profiles/soak_fuzz.py [seconds]   (default 120; prints the number of worlds that matched the oracle)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_env as t   # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
t0 = time.time()
seed = 100
worlds = 0
while time.time() - t0 < budget:
    t._fuzz_fast_worlds(seed, 25)
    t.test_randomized_worlds_match_oracle(seed)
    worlds += 50
    seed += 1
print(f"{worlds} random worlds ({seed - 100} seeds from 100) bit-identical to the C oracle in {time.time() - t0:.0f} s")
