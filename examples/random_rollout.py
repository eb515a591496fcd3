"""Random-policy rollout on the GPU: the reference's README loop, batched.

    python examples/random_rollout.py [num_envs] [obs_type]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout
import gym_simpletetris_b200 as st  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
obs_type = sys.argv[2] if len(sys.argv) > 2 else "ram"
env = st.VecEnv(n, obs_type=obs_type, reward_step=True, advanced_clears=True, device="cuda:0", seed=0)
obs = env.reset()
returns = torch.zeros(n, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
T = 500
for _ in range(T):
    actions = torch.randint(0, 7, (n,), dtype=torch.uint8, device="cuda")  # a policy network would go here
    obs, reward, done, info = env.step(actions)                            # everything stays on the GPU
    returns += reward
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{n} envs x {T} steps ({obs_type}): {n * T / dt / 1e6:.1f} M env-steps/s including the policy stub")
print("episodes:", env.episode_stats(), " mean return per env:", float(returns.mean()))
print("info keys:", list(info.keys()), " obs:", tuple(obs.shape), obs.dtype)
