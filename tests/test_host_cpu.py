"""CPU-only: host logic, C-ABI surface, sharding and the gloo statistics reduction (world size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from abrsimulator_b200 import _build, _lib
    lib = _lib.load()                                  # builds with nvcc when the .so is missing
    assert os.path.exists(_build.LIB)
    hdr = open(os.path.join(ROOT, "include", "abr_b200.h")).read()
    declared = set(re.findall(r"\b(abr_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.abr_version() == 200
    p = _lib.default_params(chunk_length=2.0)
    assert p.chunk_length == 2.0 and p.max_buffer == 60.0 and p.hist_k == 5 and p.rebuf_penalty == 4.3
    assert ctypes.sizeof(_lib.AbrParams) == 14 * 8 + 8 * 4
    with pytest.raises(TypeError):
        _lib.default_params(nonsense=1)


def test_validation_happens_before_any_device_work():
    """Argument errors are reported without a GPU; compute without a GPU fails loudly (no CPU fallback)."""
    import torch
    from abrsimulator_b200 import _lib
    lib = _lib.load()
    p = _lib.default_params()
    h = ctypes.c_void_p()
    bw = np.ones((1, 4)); tl = np.array([4], np.int32); ti = np.ones(1); sz = np.ones((2, 3))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.abr_env_create(P(bw), P(tl), P(ti), 1, 4, P(sz), P(sz), 2, 17, ctypes.byref(p), 4, ctypes.byref(h))
    assert rc == 3 and b"A" in lib.abr_last_error()
    bad = bw.copy(); bad[0, 2] = 0.0
    rc = lib.abr_env_create(P(bad), P(tl), P(ti), 1, 4, P(sz), P(sz), 2, 3, ctypes.byref(p), 4, ctypes.byref(h))
    assert rc == 1 and b"bandwidth" in lib.abr_last_error()
    if not torch.cuda.is_available():
        rc = lib.abr_env_create(P(bw), P(tl), P(ti), 1, 4, P(sz), P(sz), 2, 3, ctypes.byref(p), 4, ctypes.byref(h))
        assert rc == 2 and b"no CPU fallback" in lib.abr_last_error()
        from abrsimulator_b200.env import BatchedABREnv
        with pytest.raises(RuntimeError):
            BatchedABREnv(bw, sz, sz, 4)


def test_datamodel_both_vocabularies(tmp_path):
    from abrsimulator_b200 import Chunk, MPD, QOEMetric, ChunkInfo, NetworkInfo, load_network_trace, load_mpd_file
    from abrsimulator_b200.datamodel import pack_traces
    m1 = MPD(2, 4.0, 60.0, 8.0, [Chunk([1, 2]), Chunk([1, 2])])                 # Simulator.py:11-17
    m2 = MPD(2, 1, 20, [Chunk([1, 2.5], [1, 2.5]), Chunk([1, 2.5], [2, 5])])    # mpc_test.py:18-23
    assert m1.start_up_length == 8.0 and m2.start_up_length is None
    b, s = m1.tables()
    assert s.tolist() == [[4.0, 8.0], [4.0, 8.0]]                                # bitrate * chunk_length
    assert m2.tables()[1].tolist() == [[1, 2.5], [2, 5]]
    q = QOEMetric(1, 2, 3)
    assert q.latency_weight == 0.0 and QOEMetric(1, 2, 3, 4).latency_weight == 4
    c1 = ChunkInfo(20, 1, [2, 3], 20)
    c2 = ChunkInfo(5, [0, 2, 1], [2, 3], 7.5)
    assert c1.chunk_number == c1.chunk_id == 20 and c1.previous_bitrate == 1
    assert c2.previous_bitrate == 1 and c2.previous_bitrates == [0, 2, 1]
    f = tmp_path / "trace.txt"
    f.write_text("1.5\n2.5\n\n3\n")
    assert load_network_trace(str(f)) == [1.5, 2.5, 3.0]
    g = tmp_path / "mpd.txt"
    g.write_text("300 750 1200\n300 750 1200\n")
    ch = load_mpd_file(str(g))
    assert len(ch) == 2 and ch[0].bitrates == [300.0, 750.0, 1200.0]
    bw, tl, ti = pack_traces([NetworkInfo(1.0, [1, 2, 3]), NetworkInfo(0.5, [4])])
    assert bw.shape == (2, 3) and tl.tolist() == [3, 1] and ti.tolist() == [1.0, 0.5]


def test_shard_range_and_synth_are_sharding_invariant():
    from abrsimulator_b200.distributed import shard_range
    from abrsimulator_b200 import synth
    for n, w in ((10, 3), (1 << 20, 8), (7, 8), (65536, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    tid, off = synth.make_sessions(1000, 64, 128)
    tid2, off2 = synth.make_sessions(400, 64, 128, session_base=600)
    assert np.array_equal(tid[600:], tid2) and np.array_equal(off[600:], off2)
    assert off.min() >= 0 and off.max() < 128
    b, s = synth.make_video()
    assert b.shape == (48, 6) and np.all(s > 0) and b[0].tolist() == list(synth.LADDER_KBPS)
    bw, tl, ti = synth.make_traces(8, 64)
    assert bw.min() >= 0.2 and bw.max() <= 6.0 and tl.tolist() == [64] * 8


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["ABR_ROOT"])
from abrsimulator_b200.distributed import allreduce_stats, shard_range, max_over_ranks
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(1001, r, w)
# per-rank "statistics" of its shard: sum of session indices etc.
idx = torch.arange(lo, hi, dtype=torch.float64)
stats = torch.stack([idx.sum(), (idx * 0.1).sum(), torch.tensor(float(hi - lo), dtype=torch.float64)])
tot = allreduce_stats(stats)
full = torch.arange(0, 1001, dtype=torch.float64)
assert tot[0].item() == full.sum().item() and tot[2].item() == 1001.0
assert abs(tot[1].item() - (full * 0.1).sum().item()) < 1e-9
assert max_over_ranks(float(r + 1)) == float(w)
# deterministic: a second reduction is bit-identical
assert torch.equal(tot, allreduce_stats(stats))
dist.barrier()
dist.destroy_process_group()
print("ok", r)
"""


def test_gloo_world_size_2_stats_reduction(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    port = 29000 + os.getpid() % 2000
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   ABR_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out


def test_bind_to_gpu_cpus_is_a_no_op_without_a_gpu():
    """No NVML device in the CPU container: the helper reports None and leaves the affinity alone."""
    import os
    from abrsimulator_b200.distributed import bind_to_gpu_cpus
    before = os.sched_getaffinity(0)
    got = bind_to_gpu_cpus(0)
    after = os.sched_getaffinity(0)
    if got is None:
        assert after == before
    else:                                   # a GPU box: the process now sits on a non-empty subset of what it had
        assert set(got) == after and after <= before and after
        os.sched_setaffinity(0, before)


def test_committed_bench_lines_keep_the_driver_contract():
    """The JSON lines `bench.py` printed on the B200 box (profiles/r2_bench_*.json) carry every key the driver and the
    judge read: the base contract, `roofline`, `cpu_baseline`, `e2e`, `gpu_launches`, `clocks` and the same-run parity."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def last_line(name):
        with open(os.path.join(root, "profiles", name)) as f:
            return json.loads([ln for ln in f.read().splitlines() if ln.startswith("{")][-1])

    line = last_line("r2_bench_line.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "parity"):
        assert k in line, k
    assert line["n_gpus"] == 1 and line["dtype"] == "f64" and line["vs_baseline"] is None and "workload" in line["config"]
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["traffic"] is None or r["traffic"] > 0
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(line["cpu_baseline"])
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < line["value"]
    assert line["gpu_launches"] == line["steps"]                 # one launch per timed step
    assert line["parity"]["ok"] and line["parity"]["non_identical_values"] == 0
    assert line["mpc"]["parity"]["ok"] and line["mpc"]["parity"]["disagreements"] == 0
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    ref = last_line("r2_bench_reference_arm.json")
    assert ref["impl"] == "reference" and ref["metric"] == line["metric"] and ref["unit"] == line["unit"]
    assert ref["config"]["workload"] == line["config"]["workload"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0
    assert ref["cpu_baseline"]["value"] == ref["value"]
