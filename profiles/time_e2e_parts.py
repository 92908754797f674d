"""Where the time of one abr_env_run_host call goes: GPU kernels, copies, and host-side overhead."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from abrsimulator_b200 import synth
from abrsimulator_b200.env import BatchedABREnv

N, V = 65536, 48
bitrates, sizes = synth.make_video(V)
bw, tl, ti = synth.make_traces(1024, 2048)
env = BatchedABREnv(bw, sizes, bitrates, N, trace_len=tl, trace_interval=ti)
tid, off = synth.make_sessions(N, 1024, 2048, group=64)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
tid_p, off_p = pin(tid), pin(off)
out = dict(qoe_cost=pin(np.zeros(N)), stats=pin(np.zeros(10)))


def loop(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


print("run_host (public API)            %.1f us" % loop(lambda: env.run_host("random", V, tid_p, off_p, seed=7, want_acc=False, out=out)))
tid_d, off_d = torch.from_numpy(tid).cuda(), torch.from_numpy(off).cuda()
cost = torch.empty(N, dtype=torch.float64, device="cuda")


def dev_only():
    env.reset(tid_d, off_d)
    env.rollout("random", V, seed=7, want=())
    env.stats()
    cost.copy_(env.qoe_cost())


print("device-resident: 4 kernels        %.1f us" % loop(dev_only))
t_tid, t_off = torch.from_numpy(tid_p), torch.from_numpy(off_p)
h_cost = torch.from_numpy(out["qoe_cost"])


def copies():
    tid_d.copy_(t_tid, non_blocking=True)
    off_d.copy_(t_off, non_blocking=True)
    h_cost.copy_(cost, non_blocking=True)
    torch.cuda.current_stream().synchronize()


print("copies only (2 H2D + 1 D2H + sync) %.1f us" % loop(copies))
print("empty sync                         %.1f us" % loop(lambda: torch.cuda.current_stream().synchronize()))

# GPU-side span of one run_host call (events on the launching stream around the call)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
spans = []
for _ in range(50):
    e0.record()
    env.run_host("random", V, tid_p, off_p, seed=7, want_acc=False, out=out)
    e1.record()
    e1.synchronize()
    spans.append(e0.elapsed_time(e1) * 1e3)
print("run_host GPU span (events)        %.1f us (median)" % sorted(spans)[len(spans) // 2])
import ctypes as C
from abrsimulator_b200 import _lib
lib = _lib.load()
hp = lambda a: a.ctypes.data_as(C.c_void_p)
args = (env._h, C.c_int(1), C.c_uint64(7), C.c_int(V), hp(tid_p), hp(off_p), C.c_int(N), C.c_longlong(0), None, None,
        hp(out["stats"]), None, hp(out["qoe_cost"]), C.c_void_p(torch.cuda.current_stream().cuda_stream))
print("raw C call (no façade)            %.1f us" % loop(lambda: lib.abr_env_run_host(*args)))
