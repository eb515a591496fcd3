"""A device-resident policy driving the envs with no host work per step: the policy forward,
the action choice and the env step are captured into ONE CUDA graph; the host replays it.

    python examples/graphed_policy_loop.py [num_envs] [steps]

The step launch goes through the C ABI on torch's current stream, so it is captured like any torch op.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout
import gym_simpletetris_b200 as st  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
dev = torch.device("cuda:0")
env = st.VecEnv(n, device=dev, seed=0, reward_step=True, advanced_clears=True)
obs = env.reset()                                   # float32 [n, 10, 20], always the same tensor: env.obs
policy = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(200, 64), torch.nn.ReLU(), torch.nn.Linear(64, 7)).to(dev)
returns = torch.zeros(n, device=dev)
episodes = torch.zeros((), dtype=torch.int64, device=dev)


def one_step():
    with torch.no_grad():
        logits = policy(env.obs)
        gumbel = -torch.log(-torch.log(torch.rand_like(logits).clamp_(1e-10, 1.0)))  # sampling without a host sync
        actions = (logits + gumbel).argmax(dim=1).to(torch.uint8)
        _, reward, done, _ = env.step(actions)      # writes env.obs / env.reward / env.done in place
        returns.add_(reward)
        episodes.add_(done.sum())


side = torch.cuda.Stream(dev)
with torch.cuda.stream(side):
    for _ in range(3):                              # warm-up outside the graph (cuBLAS handles, allocator)
        one_step()
    side.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        one_step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(T):
        graph.replay()
    torch.cuda.synchronize(dev)
dt = time.perf_counter() - t0
print(f"{n} envs x {T} graphed policy+step iterations: {n * T / dt / 1e6:.1f} M env-steps/s, "
      f"{dt / T * 1e6:.1f} us per iteration, {int(episodes)} episodes, mean return {float(returns.mean()):.1f}")
assert env.poll_errors() == 0
