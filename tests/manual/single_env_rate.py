"""Steps/s of the single-env NumPy facade (N=1: launch-latency bound) next to the CPU oracle (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import gym_simpletetris_b200 as st
from oracle.oracle import OracleEnv

for kw in (dict(), dict(obs_type="grayscale"), dict(obs_type="rgb")):
    acts = np.random.RandomState(0).randint(0, 7, 20000)
    for name, env in (("b200 facade", st.make("SimpleTetris-v0", **kw)), ("cpu oracle ", OracleEnv(**kw))):
        env.reset()
        T = 5000
        t0 = time.perf_counter()
        for a in acts[:T]:
            obs, r, d, info = env.step(int(a))
            if d:
                env.reset()
        dt = time.perf_counter() - t0
        print(f"{name} {kw.get('obs_type', 'ram'):9s} {T / dt:10.0f} steps/s  ({dt / T * 1e6:.1f} us/step)")
