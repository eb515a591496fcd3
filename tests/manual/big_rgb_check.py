"""BASELINE config C5a on ONE GPU: 2^20 envs with rgb 84x84x3 float32 observations (88.8 GB) — index arithmetic
at full size, spot-checked against the oracle (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import gym_simpletetris_b200 as st
from oracle.oracle import OracleEnv

n, T = 1 << 20, 12
env = st.VecEnv(n, obs_type="rgb", device="cuda:0", seed=11)
env.reset()
g = torch.Generator(device="cuda").manual_seed(3)
sample = [0, 1, 7, 8, n // 2 - 1, n // 2, n - 9, n - 8, n - 1]
acts = []
for t in range(T):
    a = torch.randint(0, 7, (n,), dtype=torch.uint8, device="cuda", generator=g)
    acts.append(a[sample].cpu().numpy())
    obs, r, d, info = env.step(a)
torch.cuda.synchronize()
a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a0.record(); env.step(a); a1.record(); torch.cuda.synchronize()
ms = a0.elapsed_time(a1)
print(f"2^20 rgb envs: {ms:.2f} ms per step = {n / ms / 1e3:.1f} M steps/s, {n * 84878 / ms / 1e6:.0f} GB/s")
for k, e in enumerate(sample):
    o = OracleEnv(obs_type="rgb", seed=11, env_id=e)
    o.reset()
    for t in range(T):
        want, rr, dd, _ = o.step(int(acts[t][k]))
        if dd:
            want = o.reset()
    # one more step was taken for the timing
    want, rr, dd, _ = o.step(int(a[e].item()))
    if dd:
        want = o.reset()
    assert np.array_equal(obs[e].cpu().numpy(), want), e
print("spot check vs oracle ok:", sample)
