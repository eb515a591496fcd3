"""TEST / BENCH INFRASTRUCTURE ONLY — steps/s of the UNMODIFIED Python reference on this box's host cores.

SURVEY.md section 8(d), "CPU reference timing": the loop of the reference's README.md:43-51 (random action ->
`env.step` (tetris_env.py:397-403) -> `env.reset()` (tetris_env.py:405-411) on done), run

  (i)   as a single Python env, best of 3;
  (ii)  as a vector env with one worker process per host core.  gym's `AsyncVectorEnv` is the comparator BASELINE.json
        names, but `gym` is not installed in this image, so this is a `multiprocessing` stand-in with the same
        protocol: one process per env, `step` commands over a pipe, auto-reset in the worker, observations returned
        to the parent through shared memory;
  (iii) as the no-IPC sum of independent workers (an upper bound for any such vector env).

The reference file is loaded by `oracle/ref_shim.py` from `/root/reference` (build container) or from the pip
install under `baseline/_ref` (which travels to the GPU box).  Only `bench.py`'s CPU legs and tests use this.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time
import warnings

import numpy as np


def _make(kw):
    from oracle.ref_shim import make_reference_env

    warnings.simplefilter("ignore")
    env = make_reference_env(**kw)
    env.reset()
    return env


def single_env_rate(kw, budget_s=2.0, seed=0, repeats=3):
    """README.md:43-51 on one env: best of `repeats` runs of ~budget_s/repeats seconds each.  Returns (steps/s, steps)."""
    env = _make(kw)
    acts = np.random.RandomState(seed).randint(0, 7, 1 << 16).tolist()
    best, total = 0.0, 0
    for _ in range(repeats):
        n, t0 = 0, time.perf_counter()
        deadline = t0 + budget_s / repeats
        while True:
            for a in acts[n & 0xFFFF:(n & 0xFFFF) + 256]:
                _, _, d, _ = env.step(a)
                if d:
                    env.reset()
            n += 256
            t1 = time.perf_counter()
            if t1 >= deadline:
                break
        best = max(best, n / (t1 - t0))
        total += n
    return best, total


def _free_worker(kw, budget_s, seed, q):
    rate, n = single_env_rate(kw, budget_s, seed, repeats=1)
    q.put((rate, n))


def sum_of_workers_rate(kw, workers, budget_s=2.0):
    """`workers` independent processes, no IPC on the step path: sum of their own rates."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_free_worker, args=(kw, budget_s, s, q), daemon=True) for s in range(workers)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120 + budget_s) for _ in ps]
    for p in ps:
        p.join(timeout=30)
    return sum(r for r, _ in res), sum(n for _, n in res)


def _vector_worker(kw, conn, shm_name, index, obs_shape):
    from multiprocessing import shared_memory

    shm = shared_memory.SharedMemory(name=shm_name)
    try:
        nelem = int(np.prod(obs_shape))
        view = np.ndarray(obs_shape, dtype=np.float32, buffer=shm.buf, offset=index * nelem * 4)
        env = _make(kw)
        view[...] = env.reset()
        conn.send(("ready", 0.0, False))
        while True:
            cmd, arg = conn.recv()
            if cmd == "step":
                obs, r, d, _info = env.step(arg)
                if d:
                    obs = env.reset()  # gym<=0.25 vector auto-reset: the returned observation is the reset one
                view[...] = obs
                conn.send(("ok", float(r), bool(d)))
            else:
                break
    finally:
        del view
        shm.close()
        conn.close()


def per_core_vector_rate(kw, workers, budget_s=3.0, seed=0):
    """AsyncVectorEnv stand-in: `workers` processes x one env, lock-step vector steps.  Returns (env-steps/s, vector steps)."""
    from multiprocessing import shared_memory

    probe = _make(kw)
    obs_shape = tuple(np.asarray(probe.reset()).shape)
    del probe
    nelem = int(np.prod(obs_shape))
    shm = shared_memory.SharedMemory(create=True, size=max(1, workers * nelem * 4))
    ctx = mp.get_context("spawn")
    pipes, procs = [], []
    try:
        for i in range(workers):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_vector_worker, args=(kw, b, shm.name, i, obs_shape), daemon=True)
            p.start()
            b.close()
            pipes.append(a)
            procs.append(p)
        for c in pipes:
            if not c.poll(180):
                raise RuntimeError("reference vector worker did not start")
            c.recv()
        obs = np.ndarray((workers,) + obs_shape, dtype=np.float32, buffer=shm.buf)
        rs = np.random.RandomState(seed)
        acts = rs.randint(0, 7, (512, workers)).tolist()
        reward = np.zeros(workers, np.float64)
        done = np.zeros(workers, bool)

        def vstep(row):  # step_async + step_wait
            for c, a in zip(pipes, row):
                c.send(("step", a))
            for i, c in enumerate(pipes):
                _, reward[i], done[i] = c.recv()
            return obs, reward, done

        for t in range(8):
            vstep(acts[t])
        n, t0 = 0, time.perf_counter()
        while True:
            vstep(acts[n & 511])
            n += 1
            if (n & 15) == 0 and time.perf_counter() - t0 >= budget_s:
                break
        dt = time.perf_counter() - t0
        return workers * n / dt, n
    finally:
        for c in pipes:
            try:
                c.send(("close", None))
            except Exception:  # noqa: BLE001
                pass
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
        shm.close()
        shm.unlink()


def reference_python_report(kw, budget_s=12.0, workers=None):
    """The `cpu_baseline.reference_python` object of bench.py for one set of env kwargs."""
    from oracle.ref_shim import REFERENCE_FILE, reference_available

    if not reference_available():
        return {"unavailable": "reference not found under /root/reference or baseline/_ref"}
    workers = workers or os.cpu_count() or 1
    b = budget_s / 4.0
    single, n1 = single_env_rate(kw, b)
    vec, nv = per_core_vector_rate(kw, workers, b * 1.5)
    free, nf = sum_of_workers_rate(kw, workers, b)
    return {
        "unit": "env-steps/s", "cores": workers, "host_cores": os.cpu_count(),
        "single_env": round(single, 1),
        "per_core_vector": round(vec, 1),
        "per_core_no_ipc_sum": round(free, 1),
        "kind": "reference",
        "source": REFERENCE_FILE(),
        "sample": f"unmodified tetris_env.py under the gym/pygame import shim, README.md:43-51 loop: single env "
                  f"{n1} steps (best of 3); {workers} worker processes x 1 env, {nv} lock-step vector steps, obs via "
                  f"shared memory (multiprocessing stand-in for gym AsyncVectorEnv, gym is not installed); "
                  f"{workers} independent processes without IPC, {nf} steps",
    }


if __name__ == "__main__":
    import json
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    CONFIGS = {"C1": dict(), "C2": dict(reward_step=True, advanced_clears=True),
               "C4": dict(obs_type="grayscale", extend_dims=True, high_scoring=True), "C5a": dict(obs_type="rgb")}
    for name, kw in CONFIGS.items():
        print(name, json.dumps(reference_python_report(kw, budget_s=8.0)))
