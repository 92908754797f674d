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
                                     buffer=np.array([float.fromhex(x) for x in c["final"]["buffer"]]))))
    return cases
