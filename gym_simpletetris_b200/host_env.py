"""`HostVecEnv`: N SimpleTetris instances for NumPy callers — the place a gym `AsyncVectorEnv` over the reference
(`TetrisEnv`, tetris_env.py:338-433, one process per env) would take in a training loop.

Same constructor kwargs as the reference; results are NumPy arrays in host memory.  `step(actions)` is the
synchronous call (tetris_env.py:397-403, batched, gym<=0.25 auto-reset).  `step_async(actions)` /
`step_wait()` are the gym vector-env names for the pipelined form: up to two steps may be in flight, and the
step kernel of call t+1 runs while the results of call t cross PCIe (`st_host_step_async` / `st_host_wait`
of include/simpletetris_b200.h).  The arrays `step_wait` returns are views of page-locked slots owned by the
library; they stay valid until the second `step_async` after that wait (copy them to keep them longer).
Needs no torch.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import native
from .vec_env import INFO_COLS, INFO_KEYS, obs_shape


class HostVecEnv:
    def __init__(self, num_envs, width=10, height=20, obs_type="ram", extend_dims=False, render_mode="rgb_array",
                 reward_step=False, penalise_height=False, penalise_height_increase=False, advanced_clears=False,
                 high_scoring=False, penalise_holes=False, penalise_holes_increase=False, lock_delay=0,
                 step_reset=False, *, device=0, seed=0, env_id_base=0, auto_reset=True, zero_copy=None):
        self._L = native.lib()
        self.num_envs = int(num_envs)
        self.width, self.height, self.obs_type, self.extend_dims = width, height, obs_type, extend_dims
        self.render_mode = render_mode
        self.cfg = native.make_config(
            width=width, height=height, obs_type=obs_type, extend_dims=extend_dims, lock_delay=lock_delay,
            step_reset=step_reset, reward_step=reward_step, penalise_height=penalise_height,
            penalise_height_increase=penalise_height_increase, advanced_clears=advanced_clears,
            high_scoring=high_scoring, penalise_holes=penalise_holes,
            penalise_holes_increase=penalise_holes_increase, auto_reset=auto_reset, device=int(device), seed=seed,
            env_id_base=env_id_base)
        self._h = self._L.st_host_create(C.byref(self.cfg), self.num_envs)
        if not self._h:
            raise RuntimeError("st_host_create failed: " + self._L.st_last_error().decode())
        if zero_copy is not None:
            native.check(self._L.st_host_set_zero_copy(self._h, int(zero_copy)), "st_host_set_zero_copy")
        self.single_observation_shape = obs_shape(width, height, obs_type if obs_type in native.OBS_TYPES else "rgb",
                                                  extend_dims)
        self.obs_elems = int(self._L.st_obs_elems(C.byref(self.cfg)))
        n = self.num_envs
        # synchronous path: page-locked caller buffers (the kernel or the copy engine writes them in place)
        nbytes = n * (self.obs_elems * 4 + native.ST_INFO_WORDS * 4 + 4 + 1 + 1) + 64
        self._pinned = self._L.st_host_alloc_pinned(nbytes)
        if not self._pinned:
            raise RuntimeError("st_host_alloc_pinned failed: " + self._L.st_last_error().decode())
        buf = (C.c_uint8 * nbytes).from_address(self._pinned)
        o = 0
        self.obs = np.frombuffer(buf, np.float32, n * self.obs_elems, o).reshape((n,) + self.single_observation_shape)
        o += n * self.obs_elems * 4
        self.info_buf = np.frombuffer(buf, np.int32, n * native.ST_INFO_WORDS, o).reshape(n, native.ST_INFO_WORDS)
        o += n * native.ST_INFO_WORDS * 4
        self.reward = np.frombuffer(buf, np.float32, n, o)
        o += n * 4
        self._done = np.frombuffer(buf, np.uint8, n, o)
        o += n
        self._act = np.frombuffer(buf, np.uint8, n, o)
        self._in_flight = 0

    # ---- the reference API, batched, NumPy ----
    def _info(self, buf):
        d = {k: buf[:, INFO_COLS[k]] for k in INFO_KEYS if k != "statistics"}
        d["statistics"] = buf[:, 8:15]
        return d

    def reset(self):
        """TetrisEnv.reset (tetris_env.py:405-411) for every env; returns obs [N, ...] float32."""
        native.check(self._L.st_host_reset(self._h, None, self.obs.ctypes.data), "st_host_reset")
        self._in_flight = 0
        return self.obs

    def _stage(self, actions):
        a = np.asarray(actions)
        if a.shape != (self.num_envs,):
            raise ValueError(f"actions must have shape ({self.num_envs},), got {a.shape}")
        np.copyto(self._act, a, casting="unsafe")
        return self._act.ctypes.data

    def step(self, actions):
        """(obs [N,...] f32, reward [N] f32, done [N] bool, info dict of arrays); buffers are reused by the next call."""
        ap = self._stage(actions)
        native.check(self._L.st_host_step(self._h, ap, self.obs.ctypes.data, self.reward.ctypes.data,
                                          self._done.ctypes.data, self.info_buf.ctypes.data), "st_host_step")
        self._in_flight = 0
        return self.obs, self.reward, self._done.view(np.bool_), self._info(self.info_buf)

    def step_async(self, actions):
        native.check(self._L.st_host_step_async(self._h, self._stage(actions)), "st_host_step_async")
        self._in_flight += 1

    def step_wait(self):
        o, r, d, i = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        native.check(self._L.st_host_wait(self._h, C.byref(o), C.byref(r), C.byref(d), C.byref(i)), "st_host_wait")
        self._in_flight -= 1
        n = self.num_envs

        def view(ptr, ctype, count, dtype, shape):
            return np.frombuffer((ctype * count).from_address(ptr.value), dtype=dtype).reshape(shape)

        obs = view(o, C.c_float, n * self.obs_elems, np.float32, (n,) + self.single_observation_shape)
        reward = view(r, C.c_float, n, np.float32, (n,))
        done = view(d, C.c_uint8, n, np.bool_, (n,))
        info = view(i, C.c_int32, n * native.ST_INFO_WORDS, np.int32, (n, native.ST_INFO_WORDS))
        return obs, reward, done, self._info(info)

    def poll_errors(self) -> int:
        err = C.c_int32(0)
        native.check(self._L.st_host_poll(self._h, C.byref(err), None), "st_host_poll")
        return int(err.value)

    def close(self):
        if getattr(self, "_h", None):
            self._L.st_host_destroy(self._h)
            self._h = None
        if getattr(self, "_pinned", None):
            self.obs = self.info_buf = self.reward = self._done = self._act = None
            self._L.st_host_free_pinned(self._pinned)
            self._pinned = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
