"""Phase timeline of the thread-per-env kernel from in-kernel %globaltimer stamps (run on the GPU box with a
-DST_TPE_TRACE=1 build:  tools/build_variant.sh trace -DST_TPE_TRACE=1;
ST_B200_LIB=$PWD/gym_simpletetris_b200/libst_trace.so python tools/tpe_trace.py [workload] [n]).
Eight back-to-back launches (one CUDA graph, rotating over replicas like bench.py) are traced; the report shows the
in-kernel phases of each launch and the gaps between consecutive launches."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import gym_simpletetris_b200 as st
from gym_simpletetris_b200 import native

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
wl = bench.WORKLOADS[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 else wl["n"]
dev = torch.device("cuda:0")
os.environ.setdefault("ST_B200_RAM_PATH", "thread")
lib = native.lib()
R = max(2, -(-int(2.5 * bench.L2_BYTES) // int(bench.algorithmic_bytes(wl["kw"]) * n)))
R = min(R, 96)
g = torch.Generator(device=dev).manual_seed(0)
envs = []
for r in range(R):
    env = st.VecEnv(n, device=dev, seed=r, **wl["kw"])
    env.reset()
    env.step_many(torch.randint(0, 7, (100, n), dtype=torch.uint8, device=dev, generator=g))
    envs.append(env)
acts = torch.randint(0, 7, (64, n), dtype=torch.uint8, device=dev, generator=g)
KW = 1 << 15
trace = torch.zeros(8 * KW * 16, dtype=torch.int64, device=dev)
lib.st_debug_set_tpe_trace.argtypes = [C.c_void_p]
assert lib.st_debug_set_tpe_trace(trace.data_ptr()) == 0
NAMES = ["start(after griddep wait)", "records arrived", "engine done", "overlay done", "staging free",
         "obs issued", "info+records stored", "exit"]
stream = torch.cuda.Stream(device=dev)
torch.cuda.synchronize()
with torch.cuda.stream(stream):
    for u in range(8):  # warm-up; the trace build numbers launches from 0, so 8 launches = one lap of the slabs
        envs[u % R].step(acts[u])
    stream.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=stream):
        for u in range(8):
            envs[(8 + u) % R].step(acts[8 + u])
    for rep in range(3):
        gr.replay()
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    trace.zero_()
    stream.synchronize()
    e0.record(stream)
    for rep in range(4):
        gr.replay()
    e1.record(stream)
    stream.synchronize()
print(f"{name} n={n} replicas={R}: {e0.elapsed_time(e1) * 1e3 / 32:.2f} us per step over 4 graph replays of 8 launches")
t = trace.cpu().numpy().reshape(8, KW, 16)
prev_exit = None
for u in range(8):
    tu = t[u]
    tu = tu[tu[:, 0] > 0]
    base = tu[:, 0].min()
    entry = tu[:, 9].min()
    line = f"launch {u}: warps {len(tu)}  first entry {(entry - base) / 1e3:+.2f}  "
    if prev_exit is not None:
        line += f"previous last exit {(prev_exit - base) / 1e3:+.2f}  "
    line += "medians: " + " ".join(f"{np.median(tu[:, k] - base) / 1e3:.2f}" for k in range(8))
    line += f"  last exit {(tu[:, 7].max() - base) / 1e3:.2f} us"
    print(line)
    prev_exit = tu[:, 7].max()
tu = t[7]
tu = tu[tu[:, 0] > 0]
base = tu[:, 0].min()
print("last launch, relative to its first start:")
for k in range(8):
    v = (tu[:, k] - base) / 1e3
    print(f"  {NAMES[k]:28s} p5 {np.percentile(v, 5):6.2f}  median {np.median(v):6.2f}  p95 {np.percentile(v, 95):6.2f}  max {v.max():6.2f} us")
d = np.diff(tu[:, :8], axis=1) / 1e3
print("  phase durations (median us): " + "  ".join(f"{NAMES[k + 1].split()[0]} {np.median(d[:, k]):.2f}" for k in range(7)))
