"""TEST INFRASTRUCTURE ONLY — ctypes front end of `oracle/st_oracle.c`.

`OracleEnv` mirrors the reference `TetrisEnv` (tetris_env.py:338-433) closely
enough that parity tests read like tests of the reference: `reset()`,
`step(a) -> (obs, reward, done, info)`, `info` keys of tetris_env.py:232-241.
The product package never imports this module (see st_oracle.c header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libst_oracle.so")
SHAPE_NAMES = ["T", "J", "L", "Z", "S", "I", "O"]  # tetris_env.py:19
OBS_TYPES = {"ram": 0, "grayscale": 1, "rgb": 2}
FLAG_NAMES = ["reward_step", "penalise_height", "penalise_height_increase", "advanced_clears",
              "high_scoring", "penalise_holes", "penalise_holes_increase"]

_lib = None


def build_oracle(force: bool = False) -> str:
    """Compile st_oracle.c -> libst_oracle.so (gcc from PATH; OpenMP if it links)."""
    src = os.path.join(_HERE, "st_oracle.c")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    base = ["gcc", "-O2", "-fPIC", "-fvisibility=hidden", "-std=c11", "-shared", "-o", _SO, src]
    for extra in (["-fopenmp"], []):
        r = subprocess.run(base + extra, capture_output=True, text=True)
        if r.returncode == 0:
            return _SO
    raise RuntimeError("oracle build failed:\n" + r.stderr)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build_oracle()
        L = C.CDLL(_SO)
        L.or_create.restype = C.c_void_p
        L.or_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_uint64, C.c_int64]
        L.or_destroy.argtypes = [C.c_void_p]
        L.or_set_queue.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.or_error.argtypes = [C.c_void_p]
        L.or_reset.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.or_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.or_info.argtypes = [C.c_void_p, C.c_void_p]
        L.or_get_piece.argtypes = [C.c_void_p, C.c_void_p]
        L.or_set_piece.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.or_get_board.argtypes = [C.c_void_p, C.c_void_p]
        L.or_render.argtypes = [C.c_void_p, C.c_void_p]
        L.or_set_board.argtypes = [C.c_void_p, C.c_void_p]
        L.or_get_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.or_set_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.or_convert_grayscale.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.or_obs_elems.argtypes = [C.c_int, C.c_int, C.c_int]
        L.or_rollout.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int,
                                 C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.or_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _flags(kw) -> "C.Array":
    return (C.c_int * 7)(*[int(bool(kw.get(k, False))) for k in FLAG_NAMES])


def obs_shape(width, height, obs_type, extend_dims):
    """Observation shapes of tetris_env.py:381-392."""
    if obs_type == "ram":
        return (width, height, 1) if extend_dims else (width, height)
    if obs_type == "grayscale":
        return (84, 84, 1) if extend_dims else (84, 84)
    return (84, 84, 3)


class OracleEnv:
    def __init__(self, width=10, height=20, obs_type="ram", extend_dims=False, render_mode="rgb_array",
                 reward_step=False, penalise_height=False, penalise_height_increase=False,
                 advanced_clears=False, high_scoring=False, penalise_holes=False,
                 penalise_holes_increase=False, lock_delay=0, step_reset=False,
                 seed=0, env_id=0, pieces=None):
        self.width, self.height, self.obs_type, self.extend_dims = width, height, obs_type, extend_dims
        self._ot = OBS_TYPES.get(obs_type, 2)  # unknown obs_type falls through to rgb (tetris_env.py:432-433)
        kw = dict(reward_step=reward_step, penalise_height=penalise_height,
                  penalise_height_increase=penalise_height_increase, advanced_clears=advanced_clears,
                  high_scoring=high_scoring, penalise_holes=penalise_holes,
                  penalise_holes_increase=penalise_holes_increase)
        self._L = lib()
        self._h = self._L.or_create(width, height, int(lock_delay), int(bool(step_reset)), _flags(kw),
                                    int(seed) & (2**64 - 1), int(env_id))
        self._shape = obs_shape(width, height, obs_type if obs_type in OBS_TYPES else "rgb", extend_dims)
        self._queue = None
        if pieces is not None:
            self.set_pieces(pieces)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.or_destroy(self._h)
            self._h = None

    def set_pieces(self, pieces):
        ids = [SHAPE_NAMES.index(p) if isinstance(p, str) else int(p) for p in pieces]
        self._queue = np.asarray(ids, dtype=np.uint8)
        self._L.or_set_queue(self._h, self._queue.ctypes.data, len(self._queue))

    def _new_obs(self):
        return np.empty(self._shape, dtype=np.float32)

    def reset(self, return_info=False):
        obs = self._new_obs()
        self._L.or_reset(self._h, self._ot, obs.ctypes.data)
        return (obs, self.info()) if return_info else obs

    def step(self, action):
        obs = self._new_obs()
        r, d = C.c_double(), C.c_int()
        rc = self._L.or_step(self._h, int(action), self._ot, obs.ctypes.data, C.byref(r), C.byref(d))
        if rc != 0:
            raise RuntimeError("oracle: step() before reset()")
        return obs, r.value, bool(d.value), self.info()

    def info(self):
        a = np.zeros(13, dtype=np.int32)
        self._L.or_info(self._h, a.ctypes.data)
        return {"time": int(a[0]), "current_piece": SHAPE_NAMES[a[1]] if a[1] >= 0 else None,
                "score": int(a[2]), "lines_cleared": int(a[3]), "holes": int(a[4]), "deaths": int(a[5]),
                "statistics": {n: int(a[6 + i]) for i, n in enumerate(SHAPE_NAMES)}}

    @property
    def error(self):
        return self._L.or_error(self._h)

    # ---- debug state (mirrors engine attributes the reference exposes) ----
    @property
    def board(self):
        b = np.zeros((self.width, self.height), dtype=np.float64)
        self._L.or_get_board(self._h, b.ctypes.data)
        return b

    @board.setter
    def board(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        assert b.shape == (self.width, self.height)
        self._L.or_set_board(self._h, b.ctypes.data)

    def render(self):
        """TetrisEngine.render (tetris_env.py:317-321): the board with the active piece drawn."""
        b = np.zeros((self.width, self.height), dtype=np.float64)
        if self._L.or_render(self._h, b.ctypes.data) != 0:
            raise TypeError("'NoneType' object is not iterable")  # shape is None before the first reset
        return b

    def __repr__(self):
        """TetrisEngine.__repr__ (tetris_env.py:329-335)."""
        state = self.render()
        s = "o" + "-" * self.width + "o\n"
        s += "\n".join(["|" + "".join(["X" if j else " " for j in i]) + "|" for i in state.T])
        s += "\no" + "-" * self.width + "o"
        return s

    def human_frame(self, window_size=512):
        """The array TetrisEnv.render('human') hands to pygame (tetris_env.py:444-447): the TRANSPOSED board through
        convert_grayscale at window_size, then convert_grayscale_rgb."""
        g = convert_grayscale(np.ascontiguousarray(self.render().T), window_size)
        return np.repeat(g[:, :, None], 3, axis=2)

    def piece(self):
        """(id, rot, x, y, lock_delay_counter, piece_height)"""
        a = np.zeros(6, dtype=np.int32)
        self._L.or_get_piece(self._h, a.ctypes.data)
        return tuple(int(v) for v in a)

    def set_piece(self, pid, rot=0, x=None, y=0):
        pid = SHAPE_NAMES.index(pid) if isinstance(pid, str) else int(pid)
        self._L.or_set_piece(self._h, pid, rot, self.width // 2 if x is None else x, y)

    def counters(self):
        a = np.zeros(14, dtype=np.int32)
        self._L.or_get_counters(self._h, a.ctypes.data)
        return a

    def set_counters(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        assert a.shape == (14,)
        self._L.or_set_counters(self._h, a.ctypes.data)


def convert_grayscale(board, size):
    """C restatement of tetris_env.py:76-114 for a (W,H) board."""
    b = np.ascontiguousarray(board, dtype=np.float64)
    out = np.zeros((size, size), dtype=np.uint8)
    lib().or_convert_grayscale(b.ctypes.data, b.shape[0], b.shape[1], size, out.ctypes.data)
    return out


def rollout(num_envs, actions, *, width=10, height=20, obs_type="ram", lock_delay=0, step_reset=False,
            seed=0, env_id_base=0, auto_reset=True, want_info=False, want_obs=True, nthreads=0, **flags):
    """Fresh envs -> reset -> T steps with gym<=0.25 auto-reset (or_rollout).  actions: uint8 [T, N].
    Returns dict(obs [N,...] of the last step, reward [T,N] f32, done [T,N] u8, info [T,N,13] i32|None)."""
    actions = np.ascontiguousarray(actions, dtype=np.uint8)
    T, n = actions.shape
    assert n == num_envs
    ot = OBS_TYPES[obs_type]
    L = lib()
    elems = L.or_obs_elems(width, height, ot)
    obs = np.zeros((n, elems), dtype=np.float32) if want_obs else None
    reward = np.zeros((T, n), dtype=np.float32)
    done = np.zeros((T, n), dtype=np.uint8)
    info = np.zeros((T, n, 13), dtype=np.int32) if want_info else None
    err = L.or_rollout(None, width, height, int(lock_delay), int(bool(step_reset)), _flags(flags), ot,
                       int(seed) & (2**64 - 1), int(env_id_base), n, T, actions.ctypes.data,
                       obs.ctypes.data if want_obs else None, reward.ctypes.data, done.ctypes.data,
                       info.ctypes.data if want_info else None, int(bool(auto_reset)), int(nthreads))
    return {"obs": obs, "reward": reward, "done": done, "info": info, "error": err}


class OracleVecEnv:
    """N persistent oracle envs with VecEnv semantics (auto-reset), stepped by or_rollout over OpenMP threads.
    CPU baseline / reference arm of bench.py, and a checker in tests."""

    def __init__(self, num_envs, *, width=10, height=20, obs_type="ram", lock_delay=0, step_reset=False,
                 seed=0, env_id_base=0, auto_reset=True, nthreads=0, extend_dims=False, **flags):
        self._L = lib()
        self.n, self.width, self.height = int(num_envs), width, height
        self._ot = OBS_TYPES[obs_type]
        self._args = (width, height, int(lock_delay), int(bool(step_reset)), _flags(flags), self._ot,
                      int(seed) & (2**64 - 1), int(env_id_base))
        self._auto, self._nt = int(bool(auto_reset)), int(nthreads)
        self._envs = (C.c_void_p * self.n)(*[
            self._L.or_create(width, height, int(lock_delay), int(bool(step_reset)), _flags(flags),
                              int(seed) & (2**64 - 1), int(env_id_base) + i) for i in range(self.n)])
        self.elems = self._L.or_obs_elems(width, height, self._ot)
        self.obs = np.zeros((self.n, self.elems), dtype=np.float32)

    def __del__(self):
        for h in getattr(self, "_envs", []):
            if h:
                self._L.or_destroy(h)
        self._envs = []

    def reset(self):
        for i, h in enumerate(self._envs):
            self._L.or_reset(h, self._ot, self.obs[i].ctypes.data)
        return self.obs

    def step_many(self, actions, want_info=False):
        """actions uint8 [T, N] -> (obs of last step [N, elems], reward [T,N], done [T,N], info [T,N,13]|None)"""
        actions = np.ascontiguousarray(actions, dtype=np.uint8)
        T = actions.shape[0]
        reward = np.zeros((T, self.n), dtype=np.float32)
        done = np.zeros((T, self.n), dtype=np.uint8)
        info = np.zeros((T, self.n, 13), dtype=np.int32) if want_info else None
        a = self._args
        self._L.or_rollout(self._envs, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], self.n, T,
                           actions.ctypes.data, self.obs.ctypes.data, reward.ctypes.data, done.ctypes.data,
                           info.ctypes.data if want_info else None, self._auto, self._nt)
        return self.obs, reward, done, info

    def step(self, actions, want_info=False):
        obs, r, d, i = self.step_many(np.asarray(actions, dtype=np.uint8).reshape(1, self.n), want_info)
        return obs, r[0], d[0], (i[0] if want_info else None)


def max_threads() -> int:
    return int(lib().or_max_threads())
