"""Where do the microseconds of a 4096-env step go?  (run on the GPU box)"""
import statistics
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_simpletetris_b200 as st

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kw = dict(reward_step=True, advanced_clears=True)
env = st.VecEnv(n, device=dev, seed=0, **kw)
env.reset()
g = torch.Generator(device=dev).manual_seed(0)
acts = torch.randint(0, 7, (400, n), dtype=torch.uint8, device=dev, generator=g)
env.step_many(acts[:200])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tiny = torch.zeros(32, device=dev)


def timed(fn, reps=100, do_flush=False):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for i in range(reps):
        if do_flush:
            flush.zero_()
        ev[i][0].record()
        fn(i)
        ev[i][1].record()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) * 1e3 for a, b in ev]
    return statistics.median(ts), min(ts), statistics.mean(ts)


print("tiny torch op, events around one launch (us): med/min/mean", timed(lambda i: tiny.add_(1)))
print("tiny torch op after L2 flush                       ", timed(lambda i: tiny.add_(1), do_flush=True))
print("env.step warm L2                                   ", timed(lambda i: env.step(acts[200 + i])))
print("env.step after L2 flush                            ", timed(lambda i: env.step(acts[200 + i]), do_flush=True))
# back-to-back steps, one event pair around 100 launches
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for i in range(100):
    env.step(acts[200 + i])
b.record(); torch.cuda.synchronize()
print("100 back-to-back env.step launches: us/step        ", a.elapsed_time(b) * 10)
# CUDA graph of 100 steps
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for i in range(3):
        env.step(acts[i])
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        for i in range(100):
            env.step(acts[200 + i])
torch.cuda.synchronize()
a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
print("CUDA graph of 100 steps: us/step                   ", a.elapsed_time(b) * 10)
a.record(); env.step_many(acts[200:300]); b.record(); torch.cuda.synchronize()
print("step_many T=100 (one launch): us/step              ", a.elapsed_time(b) * 10)
