#!/usr/bin/env python
"""bench.py — env-steps/sec of the batched SimpleTetris step path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--modes all|none|C3,C4,...]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one `st_step` launch over one batch of envs (action -> obs, reward, done, info, with
in-kernel auto-reset).  Headline workload (BASELINE.json configs[1], "C2"): 4096 envs per GPU, 20x10 board,
ram observations, reward_step + advanced_clears, uniform random actions already resident in HBM.
Timing: the K steps are K kernel launches captured into CUDA graphs; the captured sequence is replayed `reps` times
inside one CUDA-event pair on the launching stream until the timed region is at least 50 ms (a 20-launch region of
the headline workload would be 0.14 ms: launch jitter, not throughput), bracketed by barrier + synchronize on both
sides, max over ranks; `ms_per_step` = region / (steps x reps).  Inputs are larger than L2: the steps rotate over
enough replicas of the batch that every step's lines come from HBM, whatever --steps is.  `e2e` is the same metric
through the host-buffer API (`HostVecEnv.step_async/step_wait` = `st_host_step_async/st_host_wait`, and the
synchronous `st_host_step`): pinned host actions in, obs/reward/done/info out, every step.
`modes` carries the other BASELINE.json configs (C3, C4, C5a, C5b) measured the same way, each with its own
roofline, plus the T-steps-per-launch form (`C2_T32`, `C3_T32`: `st_step_many`, observations written every step).
`cpu_baseline` is the CPU oracle (a C port of the reference algorithm, kind "port") on this box's host cores and,
in `cpu_baseline.reference_python`, the UNMODIFIED Python reference (single env, and one worker process per core).
`--impl reference` times the CPU arm alone, on all host threads, on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# ---- workloads (BASELINE.json configs) ---------------------------------------------------------------
WORKLOADS = {
    "C2": dict(n=4096, kw=dict(reward_step=True, advanced_clears=True),
               desc="C2: 4096 envs/GPU, 20x10, ram obs, reward_step+advanced_clears"),
    "C3": dict(n=65536, kw=dict(penalise_height_increase=True, penalise_holes_increase=True, lock_delay=3,
                                step_reset=True),
               desc="C3: 65536 envs/GPU, 20x10, ram obs, penalise_height_increase+penalise_holes_increase, "
                    "lock_delay=3, step_reset"),
    "C4": dict(n=262144, kw=dict(obs_type="grayscale", extend_dims=True, high_scoring=True),
               desc="C4: 262144 envs/GPU, grayscale 84x84x1 obs, high_scoring"),
    "C5a": dict(n=131072, kw=dict(obs_type="rgb"),
                desc="C5a: 131072 envs/GPU (the per-GPU share of 2^20 envs on 8 GPUs), rgb 84x84x3 obs"),
    "C5b": dict(n=65536, kw=dict(width=20, height=40),
                desc="C5b: 65536 envs/GPU, wide board 40x20 (H=40, W=20), ram obs"),
}
# uint8-observation extension (same values, a quarter of the observation bytes): reported separately, with its own
# algorithmic bytes, and only when asked for with --modes ...,C4_u8,C5a_u8 or --modes all+u8
U8_WORKLOADS = {
    "C4_u8": dict(n=262144, kw=dict(obs_type="grayscale", extend_dims=True, high_scoring=True, obs_dtype="uint8"),
                  desc="C4 with uint8 observations (extension, not the float32 parity mode): 262144 envs/GPU, grayscale"),
    "C5a_u8": dict(n=131072, kw=dict(obs_type="rgb", obs_dtype="uint8"),
                   desc="C5a with uint8 observations (extension, not the float32 parity mode): 131072 envs/GPU, rgb"),
}
WORKLOADS.update(U8_WORKLOADS)
# T steps per launch: the same workloads through st_step_many, reported separately
T_MODES = {"C2_T32": ("C2", 32), "C3_T32": ("C3", 32)}
HEADLINE = "C2"
L2_BYTES = 126 << 20


def algorithmic_bytes(kw, T=1):
    """obs_bytes (float32) + 2 * state_bytes + 6 (action 1, reward 4, done 1) per env-step.
    With T steps per launch the env record is read and written once per launch: 2 * state_bytes / T per env-step."""
    W, H = kw.get("width", 10), kw.get("height", 20)
    ot = kw.get("obs_type", "ram")
    esz = 1 if kw.get("obs_dtype") == "uint8" else 4
    obs = esz * (W * H if ot == "ram" else 84 * 84 * (3 if ot == "rgb" else 1))
    state = 60 + W * (4 if H <= 31 else 8)  # 15 words + one column word (two above 31 rows) per board column
    return obs + 6 + (2 * state if T == 1 else 2.0 * state / T)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ---- clocks sampling during the timed regions -----------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _smi(self):  # fallback when pynvml is not importable: one nvidia-smi query per sample
        import subprocess

        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(int(f[0]))
                self.max_mhz = int(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.02)

    def start(self):
        self._thr = threading.Thread(target=self._run if self.nv else self._smi, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        busy = [s for s in self.samples if self.max_mhz and s > 0.4 * self.max_mhz] or self.samples
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- the GPU arm ---------------------------------------------------------------------------------------
MIN_REGION_MS = 50.0


def ncu_traffic_bytes(name):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE timed-region launch of this workload's step kernel, read from
    the newest `profiles/r*_<workload>_step_kernel_ncu_full.txt` (an `ncu --set full` summary, tools/ncu_summary.py).
    ncu flushes caches before the single replayed launch and stops at kernel end, so writes still sitting in the
    126 MB L2 are not counted: small ram workloads read their state from DRAM but their observations stay in L2."""
    import glob
    import re

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_{name}_step_kernel_ncu_full.txt")),
                   key=lambda f: int(re.search(r"r(\d+)_", os.path.basename(f)).group(1)))
    if not files:
        return None, None
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    with open(files[-1]) as f:
        for ln in f:
            m = re.match(r"dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", ln)
            if m:
                tot += float(m.group(2)) * scale.get(m.group(3), 1)
    return (int(tot) if tot else None), os.path.relpath(files[-1], ROOT)


def ncu_issue_metrics(name):
    """Issue-side figures of the same ncu capture (ram modes are judged by instruction issue, not bytes):
    % of cycles a scheduler issued while the SM was active, ALU / LSU pipe utilisation and dynamic warp-instructions
    of ONE launch.  None when the round has no profile of this workload."""
    import glob
    import re

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_{name}_step_kernel_ncu_full.txt")),
                   key=lambda f: int(re.search(r"r(\d+)_", os.path.basename(f)).group(1)))
    if not files:
        return None
    want = {"smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
            "smsp__inst_executed.sum": "warp_instructions_per_launch"}
    out = {}
    with open(files[-1]) as f:
        for ln in f:
            parts = ln.split()
            if len(parts) >= 2 and parts[0] in want and want[parts[0]] not in out:
                out[want[parts[0]]] = float(parts[1])
    out["source"] = os.path.relpath(files[-1], ROOT)
    return out if len(out) > 1 else None


def replica_plan(steps, per_step_bytes):
    """(R, L): R replicas of the batch, and a captured sequence of L launches (a multiple of `steps`, R divides L,
    launch u runs on replica u % R) so that every replica is revisited only after R - 1 other launches, i.e. after
    >= 2.5 x 126 MB of other traffic — for any --steps."""
    R = max(1, -(-int(2.5 * L2_BYTES) // per_step_bytes))
    L = steps * -(-R // steps)
    R = next(d for d in range(R, L + 1) if L % d == 0)
    return R, L


def time_workload(name, steps, warmup, rank, world, dist, burn_in=200, T=1, n_override=None, tag=None):
    """Device-timed steps of one workload on this rank's GPU.

    L2 rule: the working set is made larger than L2 by rotating over R independent replicas of the batch
    (launch u runs on replica u % R; R * bytes-per-launch >= 2.5 x the 126 MB L2), so every step's state and
    observation lines come from / go to HBM.  The launches are captured into CUDA graphs (<= 500 launches each)
    so that the device, not the Python launch loop, is what is timed; one CUDA-event pair brackets `reps` replays
    of the captured sequence (reps chosen so that the region is >= 50 ms).  T > 1: every launch is one
    `st_step_many` call of T steps (observations, rewards, dones and info written for every step).
    """
    import ctypes as C

    import torch

    import gym_simpletetris_b200 as st
    from gym_simpletetris_b200 import native
    wl = WORKLOADS[name]
    n, kw = n_override or wl["n"], dict(wl["kw"])
    dev = torch.device("cuda", torch.cuda.current_device())
    B = algorithmic_bytes(kw, T)
    if kw.get("obs_dtype") == "uint8":
        kw["obs_dtype"] = torch.uint8
    R, L = replica_plan(steps, int(B * n * T))
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    image = kw.get("obs_type", "ram") != "ram"
    lib = native.lib()
    envs, roll = [], []
    for r in range(R):
        env = st.VecEnv(n, device=dev, seed=r, env_id_base=rank * n, **kw)
        env.reset()
        burn = torch.randint(0, 7, (burn_in, n), dtype=torch.uint8, device=dev, generator=g)
        if image:  # steady-state boards without writing burn_in images: step the same state through a ram twin
            twin = st.VecEnv(n, device=dev, seed=r, env_id_base=rank * n,
                             **{**kw, "obs_type": "ram", "extend_dims": False})
            twin.state.copy_(env.state)
            twin.step_many(burn)
            env.state.copy_(twin.state)
            del twin
        else:
            env.step_many(burn)
        envs.append(env)
        if T > 1:  # rollout buffers of one st_step_many launch: [T, n, ...]
            roll.append(dict(obs=torch.empty((T,) + tuple(env.obs.shape), dtype=env.obs.dtype, device=dev),
                             reward=torch.empty((T, n), dtype=torch.float32, device=dev),
                             done=torch.empty((T, n), dtype=torch.uint8, device=dev),
                             info=torch.empty((T, n, native.ST_INFO_WORDS), dtype=torch.int32, device=dev)))
    n_act = max(warmup, 3) + L
    actions = torch.randint(0, 7, (n_act, T, n), dtype=torch.uint8, device=dev, generator=g)
    stream = torch.cuda.Stream(device=dev)

    def launch(u, a):
        env = envs[u % R]
        if T == 1:
            env.step(a[0])
            return
        rb = roll[u % R]
        native.check(lib.st_step_many(C.byref(env.cfg), env.state.data_ptr(), a.data_ptr(), T, rb["obs"].data_ptr(),
                                      n * env.obs_elems, rb["reward"].data_ptr(), rb["done"].data_ptr(),
                                      rb["info"].data_ptr(), n * native.ST_INFO_WORDS, C.byref(env._aux_many()), n,
                                      torch.cuda.current_stream(dev).cuda_stream), "st_step_many")

    torch.cuda.synchronize(dev)  # set-up ran on the default stream; the side stream does not wait for it
    graphs = []
    with torch.cuda.stream(stream):
        for t in range(warmup):
            launch(t, actions[t])
        stream.synchronize()
        t = 0
        while t < L:
            chunk = min(500, L - t)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=stream):
                for u in range(t, t + chunk):
                    launch(u, actions[warmup + u])
            graphs.append(gr)
            t += chunk
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def region(reps):
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            e0.record(stream)
            for _ in range(reps):
                for gr in graphs:
                    gr.replay()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            return float(e0.elapsed_time(e1))

        probe_ms = region(1)  # untimed pass over the whole sequence: warms every replica, sizes the timed region
        if world > 1:
            tt = torch.tensor([probe_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            probe_ms = float(tt.item())
        reps = max(1, int(-(-MIN_REGION_MS // max(probe_ms, 1e-3))))
        for _ in range(4):  # the probe pass runs cold (and slower): grow reps until the region really is long enough
            rank_ms = region(reps)
            ms = rank_ms
            if world > 1:
                tt = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms = float(tt.item())
            if ms >= MIN_REGION_MS:
                break
            reps = int(reps * 1.25 * MIN_REGION_MS / max(ms, 1e-3)) + 1
    launches = L * reps
    total_ms = rank_ms
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    stats = {"episodes": 0}
    for env in envs:
        assert env.poll_errors() == 0
        stats["episodes"] += env.episode_stats(reduce=True)["episodes"]
    kernel_name = lib.st_step_kernel_name(C.byref(envs[0].cfg), n).decode()
    peak, peak_src = measured_peak()
    kernel_ms = rank_ms / launches
    achieved = B * n * T / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic_bytes(tag or name)
    res = {
        "workload": wl["desc"] if n_override is None else wl["desc"].replace(f"{wl['n']} envs/GPU", f"{n} envs/GPU"),
        "envs_per_gpu": n, "steps_per_launch": T, "steps": steps, "reps": launches // steps,
        "timed_launches": launches, "timed_region_ms": round(total_ms, 3), "replicas": R,
        "value": n * T * world * launches / (total_ms * 1e-3), "ms_per_step": total_ms / (launches * T),
        "launches": launches, "episodes_all_ranks": stats["episodes"],
        "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic,
                     "traffic_note": "ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch, from "
                                     f"{traffic_src} (L2-resident writes not included)" if traffic else None,
                     "algorithmic_bytes_per_launch": int(B * n * T), "peak_source": peak_src,
                     "kernel": kernel_name,
                     "algorithmic_bytes_per_env_step": B, "env_steps_per_launch": n * T,
                     "avg_launch_ms": round(kernel_ms, 6)},
    }
    issue = ncu_issue_metrics(tag or name) if not image else None
    if issue:  # ram modes: the north star asks for integer-pipe / issue utilisation beside the byte roofline
        if "warp_instructions_per_launch" in issue:
            issue["warp_instructions_per_env_step"] = round(issue["warp_instructions_per_launch"] / (n * T), 1)
        issue["note"] = ("ncu --set full capture of one launch (cold, serialised): issue slots busy while the SM is active; "
                         "ram modes are bound by instruction issue / latency, not by bytes")
        res["roofline"]["issue"] = issue
    del envs, actions, graphs, roll
    torch.cuda.empty_cache()
    return res


def write_only_ceiling_gbs():
    """Practical ceiling of a store-only kernel: GB/s of a plain 4 GiB fill, best of 5."""
    import torch

    x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")
    best = 0.0
    for i in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        x.fill_(128.0)
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            best = max(best, x.numel() * 4 / (a.elapsed_time(b) * 1e-3) / 1e9)
    del x
    torch.cuda.empty_cache()
    return round(best, 1)


def pcie_d2h_peak_gbs(dist=None):
    """GB/s of a plain 256 MiB device -> pinned-host cudaMemcpyAsync on this GPU's link (best of 5): the ceiling of `e2e`.
    With `dist` (N > 1, called by every rank): all ranks copy AT THE SAME TIME and the rates of one round are summed —
    what the host's memory system delivers when every link is busy; returns the best round's sum / N, the mean per
    link, which is the ceiling of the N-GPU `e2e`."""
    import torch

    d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    best = 0.0
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()
        a.record()
        h.copy_(d, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        rate = d.numel() / (a.elapsed_time(b) * 1e-3) / 1e9
        if dist is not None:
            t = torch.tensor([rate], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            rate = float(t.item()) / dist.get_world_size()
        if i >= 1:
            best = max(best, rate)
    del d, h
    torch.cuda.empty_cache()
    return round(best, 2)


def time_e2e(name, steps, warmup, rank, world, dist, min_seconds=0.25):
    """Same metric end to end through the public host-buffer API (`HostVecEnv`, i.e. the `st_host_*` C ABI): every step
    takes that step's actions from host memory and delivers obs / reward / done / info to host memory.  Two forms are
    timed: the pipelined `step_async` / `step_wait` pair (gym vector-env protocol; the kernel of step t+1 runs while
    the results of step t cross PCIe; every result is waited for and touched inside the timed region) and the
    synchronous `step`.  The headline `value` is the pipelined one."""
    import torch

    from gym_simpletetris_b200 import native
    from gym_simpletetris_b200.host_env import HostVecEnv

    wl = WORKLOADS[name]
    n, kw = wl["n"], wl["kw"]
    env = HostVecEnv(n, device=torch.cuda.current_device(), seed=0, env_id_base=rank * n, **kw)
    env.reset()
    nact = 64
    acts = np.random.RandomState(7 + rank).randint(0, 7, (nact, n)).astype(np.uint8)
    d2h = n * (env.obs_elems * 4 + 4 + 1 + 4 * native.ST_INFO_WORDS)

    def run_sync(k):
        chk = 0.0
        for t in range(k):
            obs, rew, done, info = env.step(acts[t % nact])
            chk += float(rew[0]) + float(obs.flat[-1])
        return chk

    def run_pipe(k):
        chk = 0.0
        env.step_async(acts[0])
        for t in range(1, k):
            env.step_async(acts[t % nact])       # step t is enqueued ...
            obs, rew, done, info = env.step_wait()  # ... while the results of step t-1 arrive
            chk += float(rew[0]) + float(obs.flat[-1])
        obs, rew, done, info = env.step_wait()
        return chk + float(rew[0]) + float(obs.flat[-1])

    out = {}
    for label, fn in (("pipelined", run_pipe), ("sync", run_sync)):
        fn(max(warmup, 3))
        k = max(steps, 8)
        while True:  # at least min_seconds of wall time in the timed region
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            fn(k)
            dt = time.perf_counter() - t0
            if world > 1:  # max over ranks; also makes every rank take the same decision below
                tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            if dt >= min_seconds or k >= 1 << 16:
                break
            k = int(k * max(2.0, 1.2 * min_seconds / max(dt, 1e-6)))
        out[label] = {"value": n * world * k / dt, "ms_per_step": dt / k * 1e3, "steps": k,
                      "d2h_gbs_per_gpu": round(d2h * k / dt / 1e9, 2)}
    assert env.poll_errors() == 0
    env.close()
    best = "pipelined" if out["pipelined"]["value"] >= out["sync"]["value"] else "sync"
    return {"value": out[best]["value"], "unit": "env-steps/s", "h2d_bytes_per_step": n,
            "d2h_bytes_per_step": d2h, "ms_per_step": out[best]["ms_per_step"], "steps": out[best]["steps"],
            "form": best, "pipelined": out["pipelined"], "sync": out["sync"],
            "d2h_gbs_per_gpu": out[best]["d2h_gbs_per_gpu"],
            "api": "HostVecEnv.step_async/step_wait (st_host_step_async/st_host_wait) and HostVecEnv.step "
                   "(st_host_step): C ABI, host buffers in and out every step"}


# ---- the CPU arm (oracle port of the reference algorithm) ---------------------------------------------------
def cpu_vec_env(name, nthreads, n=None):
    from oracle.oracle import OracleVecEnv

    wl = WORKLOADS[name]
    env = OracleVecEnv(n or wl["n"], seed=0, nthreads=nthreads, **wl["kw"])
    env.reset()
    return env


def time_cpu(name, budget_s, nthreads, n=None):
    """Bounded sample: whole vector steps of the workload until ~budget_s seconds of wall time are used."""
    from oracle.oracle import max_threads

    env = cpu_vec_env(name, nthreads, n)
    rs = np.random.RandomState(3)
    chunk, done_steps, t_used = 4, 0, 0.0
    env.step_many(rs.randint(0, 7, (2, env.n)).astype(np.uint8))  # warm caches
    while t_used < budget_s:
        a = rs.randint(0, 7, (chunk, env.n)).astype(np.uint8)
        t0 = time.perf_counter()
        env.step_many(a)
        t_used += time.perf_counter() - t0
        done_steps += chunk
        chunk = min(chunk * 2, 4096)
    threads = nthreads if nthreads > 0 else max_threads()
    return {"value": env.n * done_steps / t_used, "unit": "env-steps/s", "cores": threads, "kind": "port",
            "sample": f"{env.n} envs x {done_steps} vector steps of the {name} workload in {t_used:.1f} s, "
                      f"oracle/st_oracle.c (C port of the reference algorithm), OpenMP {threads} threads",
            "host_cores": os.cpu_count()}


def reference_python(name, budget_s):
    """The unmodified Python reference on this box's host cores (SURVEY.md 8(d), row d'): single env and one process per core."""
    try:
        from oracle.ref_python_bench import reference_python_report

        return reference_python_report(WORKLOADS[name]["kw"], budget_s=budget_s)
    except Exception as e:  # noqa: BLE001  (a missing baseline/_ref must not cost the GPU numbers)
        return {"unavailable": f"{type(e).__name__}: {e}"}


def cpu_legs(name, cpu_seconds, py_seconds, n=None):
    """`cpu_baseline`: the C port on all host threads (bounded sample), its single-thread rate, and the Python reference."""
    cpu = time_cpu(name, cpu_seconds, 0, n)
    cpu["single_thread"] = time_cpu(name, min(3.0, cpu_seconds), 1)["value"]
    if py_seconds > 0:
        cpu["reference_python"] = reference_python(name, py_seconds)
    return cpu


def headline_config(name):
    """`config` of the JSON line — identical in both arms (`--impl b200` and `--impl reference`)."""
    wl = WORKLOADS[name]
    return {"workload": wl["desc"], "envs_per_gpu": wl["n"]}


def reference_arm(args):
    """`--impl reference`: the CPU arm alone — the C port of the reference algorithm on all host threads over the
    headline workload (n x gpus envs), timed for at least 1 s of whole vector steps (K is a lower bound: 20 vector
    steps are 2 ms, which measures OpenMP start-up), plus the unmodified Python reference beside it."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return  # other ranks exit 0 without work
    from oracle.oracle import max_threads

    wl = WORKLOADS[args.workload]
    n = wl["n"] * args.gpus
    env = cpu_vec_env(args.workload, 0, n)
    rs = np.random.RandomState(3)
    acts = rs.randint(0, 7, (256, n)).astype(np.uint8)
    for t in range(args.warmup):
        env.step(acts[t % 256])
    done_steps, dt = 0, 0.0
    while done_steps < args.steps or dt < 1.0:
        t0 = time.perf_counter()
        for t in range(args.steps):
            env.step(acts[(done_steps + t) % 256])
        dt += time.perf_counter() - t0
        done_steps += args.steps
    v = n * done_steps / dt
    threads = max_threads()
    cpu = {"value": v, "unit": "env-steps/s", "cores": threads, "kind": "port",
           "sample": f"{n} envs x {done_steps} vector steps in {dt:.2f} s, oracle/st_oracle.c (C port of the reference "
                     f"algorithm), OpenMP {threads} threads", "host_cores": os.cpu_count()}
    if args.py_seconds > 0:
        cpu["reference_python"] = reference_python(args.workload, args.py_seconds)
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / done_steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": headline_config(args.workload),
        "method": {"envs": n, "timed_steps": done_steps, "timed_region_s": round(dt, 3),
                   "note": "CPU arm: oracle/st_oracle.c, a C restatement of the reference's dense float64 algorithm, "
                           "OpenMP over envs; the reference itself is pure Python and ~100x slower per core "
                           "(cpu_baseline.reference_python, timed in this same run)"},
        "cpu_baseline": cpu,
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def ensure_built(rank):
    """A checkout without the in-tree .so (it is git-ignored): rank 0 compiles it (nvcc, ~30 s), the others wait."""
    so = os.path.join(ROOT, "gym_simpletetris_b200", "libsimpletetris_b200.so")
    if os.path.exists(so):
        return
    if rank == 0:
        import __graft_entry__

        __graft_entry__.build()
    else:
        for _ in range(600):
            if os.path.exists(so):
                return
            time.sleep(0.5)
        raise SystemExit("bench.py: libsimpletetris_b200.so was not built")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--modes", default="all", help="extra workloads to report in `modes`: all | none | C3,C4,...")
    ap.add_argument("--mode-steps", type=int, default=30)
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="0 skips the cpu_baseline leg")
    ap.add_argument("--py-seconds", type=float, default=12.0, help="budget of the Python-reference timing; 0 skips it")
    ap.add_argument("--workload", default=HEADLINE, choices=list(WORKLOADS),
                    help="headline workload (default C2 = BASELINE.json configs[1]); others are for profiling")
    ap.add_argument("--steps-per-launch", type=int, default=1, help="profiling runs: headline through st_step_many")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return reference_arm(args)

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    ensure_built(rank)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)

    sampler = ClockSampler(local)
    sampler.start()
    head = time_workload(args.workload, args.steps, args.warmup, rank, world, dist, T=args.steps_per_launch)
    e2e = None if args.no_e2e else time_e2e(args.workload, min(args.steps, 200), args.warmup, rank, world, dist)
    modes = {}
    base = [k for k in WORKLOADS if k != args.workload and k not in U8_WORKLOADS]
    names = [] if args.modes == "none" else (base + list(T_MODES) if args.modes == "all" else
                                              base + list(T_MODES) + list(U8_WORKLOADS) if args.modes == "all+u8"
                                              else args.modes.split(","))
    for nm in names:
        if nm in T_MODES:  # T steps per launch (st_step_many), observations written every step
            wl_name, T = T_MODES[nm]
            r = time_workload(wl_name, max(3, args.mode_steps // 4), 3, rank, world, dist, T=T, tag=nm)
            r["workload"] += f"; {T} steps per launch (st_step_many), obs/reward/done/info written for every step"
            modes[nm] = r
            continue
        r = time_workload(nm, args.mode_steps, max(3, min(args.warmup, 5)), rank, world, dist)
        small = WORKLOADS[nm]["n"] * algorithmic_bytes(WORKLOADS[nm]["kw"]) < (2 << 30) and nm not in U8_WORKLOADS
        r["e2e"] = time_e2e(nm, 5, 3, rank, world, dist, min_seconds=0.1) if small else None
        modes[nm] = r
        if nm == "C4" and world > 1:  # C4 is 262144 envs IN TOTAL, N/G per GPU (strong scaling)
            r = time_workload(nm, args.mode_steps, 3, rank, world, dist, n_override=WORKLOADS[nm]["n"] // world)
            r["scaling"] = "strong"
            modes["C4_strong"] = r
    fill_gbs = write_only_ceiling_gbs() if rank == 0 else None
    pcie_all = pcie_d2h_peak_gbs(dist) if (world > 1 and not args.no_e2e) else None  # every rank at the same time
    if world > 1:
        dist.barrier()
    pcie = pcie_d2h_peak_gbs() if (rank == 0 and not args.no_e2e) else None         # rank 0 alone
    clocks = sampler.stop()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cpu = None
    if args.cpu_seconds > 0:  # rank 0, after the other ranks have finished their GPU work (they exit; no spinning peers)
        if world > 1:
            time.sleep(2.0)
        cpu = cpu_legs(args.workload, args.cpu_seconds, args.py_seconds)
    if e2e is not None:
        e2e["pcie_d2h_peak_gbs"] = pcie
        e2e["pcie_frac"] = round(e2e["d2h_gbs_per_gpu"] / pcie, 3) if pcie else None
        if pcie_all:  # N > 1: the host delivers less per link when all links copy at once; that is the ceiling here
            e2e["pcie_d2h_all_ranks_busy_gbs"] = pcie_all
            e2e["pcie_frac_all_ranks_busy"] = round(e2e["d2h_gbs_per_gpu"] / pcie_all, 3)
    line = {
        "metric": "env-steps/sec", "value": head["value"], "unit": "env-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": headline_config(args.workload),
        "method": {"reps": head["reps"], "timed_launches": head["timed_launches"],
                   "timed_region_ms": head["timed_region_ms"], "steps_per_launch": head["steps_per_launch"],
                   "actions": "uniform iid over 0..6, uint8, resident in HBM; boards in steady state (200 burn-in steps)",
                   "obs": "float32, as the reference returns", "info": "written every step", "auto_reset": True,
                   "l2": f"inputs larger than L2: launch u runs on replica u % {head['replicas']} of the batch "
                         f"({head['replicas']} x {head['envs_per_gpu']} envs, >= 2.5 x 126 MB per rotation)",
                   "timing": f"the K = {args.steps} step-kernel launches are captured into CUDA graphs (padded to a "
                             f"multiple of the replica count) and replayed until the region is >= {MIN_REGION_MS:.0f} ms, one "
                             "CUDA-event pair on the launching stream around the region, barrier + synchronize on "
                             "both sides; max over ranks; ms_per_step = region / timed_launches"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": head["launches"], "roofline": head["roofline"],
        "cpu_baseline": cpu, "modes": modes,
        "hbm_write_only_fill_gbs": fill_gbs,  # plain 4 GiB torch fill_ on this GPU: ceiling of a store-only kernel
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
