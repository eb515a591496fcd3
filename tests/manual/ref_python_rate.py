"""Steps/s of the UNMODIFIED Python reference under the import shim (build container only): single env, and one
process per core without IPC (an upper bound for a gym AsyncVectorEnv, which is not installed here)."""
import multiprocessing as mp
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

CONFIGS = {"C1 default ram": dict(), "C2 ram step+adv": dict(reward_step=True, advanced_clears=True),
           "C4 grayscale": dict(obs_type="grayscale", extend_dims=True, high_scoring=True), "C5a rgb": dict(obs_type="rgb")}


def run(kw, T, seed=0):
    from oracle.ref_shim import make_reference_env

    warnings.simplefilter("ignore")
    env = make_reference_env(**kw)
    acts = np.random.RandomState(seed).randint(0, 7, T)
    env.reset()
    t0 = time.perf_counter()
    for a in acts:
        _, _, d, _ = env.step(int(a))
        if d:
            env.reset()
    return T / (time.perf_counter() - t0)


def worker(args):
    return run(*args)


if __name__ == "__main__":
    ncpu = os.cpu_count()
    for name, kw in CONFIGS.items():
        T = 20000 if kw.get("obs_type", "ram") == "ram" else 5000
        single = max(run(kw, T) for _ in range(3))
        with mp.Pool(ncpu) as pool:
            t0 = time.perf_counter()
            pool.map(worker, [(kw, T, s) for s in range(ncpu)])
            multi = ncpu * T / (time.perf_counter() - t0)
        print(f"{name:18s} single env {single:9.0f} steps/s   {ncpu} processes {multi:9.0f} steps/s")
