"""The program compute-sanitizer is pointed at (tools/sanitize.sh): __graft_entry__.smoke() plus ragged-tail image
cases (a last CTA with fewer than 8 envs, masked reset, T steps per launch, terminal observations, render), on both
ram kernels.  Small on purpose: memcheck / racecheck slow kernels down by 10-100x."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as ge
from gym_simpletetris_b200 import VecEnv

ge.smoke()
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
for path in ("warp", "thread"):
    os.environ["ST_B200_RAM_PATH"] = path
    for kw, n in ((dict(obs_type="grayscale", extend_dims=True), 13), (dict(obs_type="rgb"), 11),
                  (dict(obs_type="rgb", width=7, height=9), 19), (dict(), 37), (dict(width=20, height=40), 45),
                  (dict(width=5, height=7, lock_delay=2, step_reset=True, penalise_holes=True), 67)):
        env = VecEnv(n, device=dev, seed=2, terminal_obs=True, **kw)
        env.reset()
        acts = torch.randint(0, 7, (24, n), dtype=torch.uint8, device=dev, generator=g)
        for t in range(8):
            env.step(acts[t])
        env.step_many(acts[8:24])
        mask = torch.zeros(n, dtype=torch.bool, device=dev)
        mask[::3] = True
        env.reset(mask=mask)
        env.step(acts[0])
        env.render()
        assert env.poll_errors() == 0
        env.close()
torch.cuda.synchronize()
print("sanitizer case ok")
