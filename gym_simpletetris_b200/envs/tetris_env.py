"""`TetrisEnv`: the reference's single-env class (tetris_env.py:338-467) as an N=1 view of the CUDA path.

Same constructor kwargs, spaces, `step`/`reset` tuple shapes, reward values and info keys as the
reference; NumPy in, NumPy out.  It owns a `st_host_*` handle (include/simpletetris_b200.h), so it needs
neither torch nor a caller-side device allocator.  `env.engine` is a read/write view of the device
state with the attribute names of the reference's `TetrisEngine` (tetris_env.py:125-181).
"""
from __future__ import annotations

import ctypes as C
import random

import numpy as np

from .. import native
from ..spaces import Box, Discrete

shape_names = ["T", "J", "L", "Z", "S", "I", "O"]  # tetris_env.py:19
shapes = {  # tetris_env.py:10-18 (kept for callers that read `engine.shape`)
    "T": [(0, 0), (-1, 0), (1, 0), (0, -1)], "J": [(0, 0), (-1, 0), (0, -1), (0, -2)],
    "L": [(0, 0), (1, 0), (0, -1), (0, -2)], "Z": [(0, 0), (-1, 0), (0, -1), (1, -1)],
    "S": [(0, 0), (-1, -1), (0, -1), (1, 0)], "I": [(0, 0), (0, -1), (0, -2), (0, -3)],
    "O": [(0, 0), (0, -1), (-1, 0), (-1, -1)],
}

try:
    import gym as _gym  # type: ignore

    _Base = _gym.Env
except Exception:  # noqa: BLE001
    try:
        import gymnasium as _gym  # type: ignore

        _Base = _gym.Env
    except Exception:  # noqa: BLE001
        _Base = object


class EngineView:
    """Attribute view of the one env's device state (names of tetris_env.py:138-181)."""

    def __init__(self, env):
        self._env = env
        self.width, self.height = env.width, env.height

    def _scalars(self):
        s = np.zeros((1, native.ST_UNPACKED_WORDS), dtype=np.int32)
        native.check(self._env._L.st_host_get_state(self._env._h, None, s.ctypes.data), "st_host_get_state")
        return s[0]

    def _set_scalar(self, idx, value):
        s = self._scalars().reshape(1, -1).copy()
        s[0, idx] = value
        native.check(self._env._L.st_host_set_state(self._env._h, None, s.ctypes.data), "st_host_set_state")

    @property
    def board(self):
        b = np.zeros((1, self.width, self.height), dtype=np.uint8)
        native.check(self._env._L.st_host_get_state(self._env._h, b.ctypes.data, None), "st_host_get_state")
        return b[0].astype(np.float64)

    @board.setter
    def board(self, value):
        b = np.ascontiguousarray(np.asarray(value) != 0, dtype=np.uint8).reshape(1, self.width, self.height)
        native.check(self._env._L.st_host_set_state(self._env._h, b.ctypes.data, None), "st_host_set_state")

    shape_name = property(lambda self: shape_names[self._scalars()[0]] if self._scalars()[0] < 7 else None)
    anchor = property(lambda self: (int(self._scalars()[2]), int(self._scalars()[3])) if self._scalars()[0] < 7 else None)
    _lock_delay = property(lambda self: int(self._scalars()[4]), lambda self, v: self._set_scalar(4, v))
    time = property(lambda self: int(self._scalars()[5]), lambda self, v: self._set_scalar(5, v))
    score = property(lambda self: int(self._scalars()[6]), lambda self, v: self._set_scalar(6, v))
    lines_cleared = property(lambda self: int(self._scalars()[7]), lambda self, v: self._set_scalar(7, v))
    holes = property(lambda self: int(self._scalars()[8]), lambda self, v: self._set_scalar(8, v))
    piece_height = property(lambda self: int(self._scalars()[9]), lambda self, v: self._set_scalar(9, v))
    n_deaths = property(lambda self: int(self._scalars()[10]), lambda self, v: self._set_scalar(10, v))

    @property
    def shape(self):
        s = self._scalars()
        if s[0] >= 7:
            return None
        cells = list(shapes[shape_names[s[0]]])
        for _ in range(int(s[1])):  # rotated(cclk=False), tetris_env.py:22-26
            cells = [(j, -i) for i, j in cells]
        return cells

    @property
    def shape_counts(self):
        s = self._scalars()
        return {n: int(s[11 + i]) for i, n in enumerate(shape_names)}

    def set_piece(self, name, rot=0, x=None, y=0):
        """Replace the active piece (test hook; the reference's tests would assign engine.shape/anchor)."""
        s = self._scalars().reshape(1, -1).copy()
        s[0, 0] = shape_names.index(name) if isinstance(name, str) else int(name)
        s[0, 1], s[0, 2], s[0, 3] = rot, self.width // 2 if x is None else x, y
        native.check(self._env._L.st_host_set_state(self._env._h, None, s.ctypes.data), "st_host_set_state")
        self._env._was_reset = True  # there is a piece now

    def set_pieces(self, pieces):
        """Inject the lifetime piece sequence (the `_choose_shape` replacement used by parity tests)."""
        q = np.asarray([shape_names.index(p) if isinstance(p, str) else int(p) for p in pieces], dtype=np.uint8)
        native.check(self._env._L.st_host_set_piece_queue(self._env._h, q.ctypes.data, len(q)),
                     "st_host_set_piece_queue")
        self._env._queue_set = True

    def get_info(self):
        return self._env._get_info()

    def render(self):
        """TetrisEngine.render (tetris_env.py:317-321): copy of the board with the active piece drawn."""
        s = self._scalars()
        if s[0] >= 7:
            raise TypeError("'NoneType' object is not iterable")  # shape is None before the first reset (ref:170-172)
        state = self.board
        for i, j in self.shape:  # _set_piece(True), tetris_env.py:323-327: in-board cells only
            x, y = i + int(s[2]), j + int(s[3])
            if 0 <= x < self.width and 0 <= y < self.height:
                state[x, y] = 1.0
        return state

    def __repr__(self):
        """The ASCII dump of tetris_env.py:329-335, byte for byte: board with the piece drawn, one text row per y."""
        state = self.render()
        s = "o" + "-" * self.width + "o\n"
        s += "\n".join(["|" + "".join(["X" if j else " " for j in i]) + "|" for i in state.T])
        s += "\no" + "-" * self.width + "o"
        return s


class TetrisEnv(_Base):
    metadata = {"render.modes": ["human", "rgb_array"], "render_fps": 8}  # tetris_env.py:339

    def __init__(self, width=10, height=20, obs_type="ram", extend_dims=False, render_mode="rgb_array",
                 reward_step=False, penalise_height=False, penalise_height_increase=False, advanced_clears=False,
                 high_scoring=False, penalise_holes=False, penalise_holes_increase=False, lock_delay=0,
                 step_reset=False, *, device=0, seed=None, env_id=0):
        self.width, self.height, self.obs_type, self.extend_dims = width, height, obs_type, extend_dims
        self.render_mode = render_mode
        self.window_size = 512
        self.window = None
        self.clock = None
        if seed is None:  # the reference draws pieces from the global `random` (tetris_env.py:2,187);
            seed = random.getrandbits(64)  # so `random.seed(s)` before construction fixes the piece stream
        self._L = native.lib()
        self._cfg = native.make_config(
            width=width, height=height, obs_type=obs_type, extend_dims=extend_dims, lock_delay=lock_delay,
            step_reset=step_reset, reward_step=reward_step, penalise_height=penalise_height,
            penalise_height_increase=penalise_height_increase, advanced_clears=advanced_clears,
            high_scoring=high_scoring, penalise_holes=penalise_holes,
            penalise_holes_increase=penalise_holes_increase, auto_reset=False, device=device, seed=seed,
            env_id_base=env_id)
        self._h = self._L.st_host_create(C.byref(self._cfg), 1)
        if not self._h:
            raise RuntimeError("st_host_create failed: " + self._L.st_last_error().decode())
        self.action_space = Discrete(7)  # tetris_env.py:377
        if obs_type == "ram":  # tetris_env.py:381-392 (declared range 0..1 kept, images are 0/128/190)
            shp = (width, height, 1) if extend_dims else (width, height)
        elif obs_type == "grayscale":
            shp = (84, 84, 1) if extend_dims else (84, 84)
        else:
            shp = (84, 84, 3)
        if obs_type in ("ram", "grayscale", "rgb"):
            self.observation_space = Box(0, 1, shape=shp, dtype=np.float32)
        self._obs_shape = shp
        # One page-locked block for everything a step moves: the kernel reads the action from it and writes
        # obs / reward / done / info into it directly (no staging copies), so a step is one launch + one sync.
        nobs = int(np.prod(shp))
        nbytes = nobs * 4 + native.ST_INFO_WORDS * 4 + 16
        self._pinned = self._L.st_host_alloc_pinned(nbytes)
        if not self._pinned:
            raise RuntimeError("st_host_alloc_pinned failed: " + self._L.st_last_error().decode())
        buf = (C.c_uint8 * nbytes).from_address(self._pinned)
        self._obs_buf = np.frombuffer(buf, dtype=np.float32, count=nobs, offset=0).reshape(shp)
        self._info_buf = np.frombuffer(buf, dtype=np.int32, count=native.ST_INFO_WORDS, offset=nobs * 4).reshape(1, -1)
        tail = nobs * 4 + native.ST_INFO_WORDS * 4
        self._reward_buf = np.frombuffer(buf, dtype=np.float32, count=1, offset=tail)
        self._done_buf = np.frombuffer(buf, dtype=np.uint8, count=1, offset=tail + 4)
        self._act_buf = np.frombuffer(buf, dtype=np.uint8, count=1, offset=tail + 8)
        self._ptr = {k: int(getattr(self, "_" + k + "_buf").ctypes.data) for k in ("obs", "info", "reward", "done", "act")}
        self._was_reset = False
        self._queue_set = False
        self.engine = EngineView(self)

    def _get_info(self):
        """engine.get_info() (tetris_env.py:232-241) of the current device state."""
        s = self.engine._scalars()
        return {"time": int(s[5]), "current_piece": shape_names[s[0]] if s[0] < 7 else None, "score": int(s[6]),
                "lines_cleared": int(s[7]), "holes": int(s[8]), "deaths": int(s[10]),
                "statistics": {n: int(s[11 + i]) for i, n in enumerate(shape_names)}}

    def step(self, action):
        if isinstance(action, (bool, np.bool_)) or int(action) != action or not 0 <= int(action) <= 6:
            raise KeyError(action)  # value_action_map lookup, tetris_env.py:245
        if not self._was_reset:
            raise TypeError("step() called before reset(): the engine has no piece (tetris_env.py:170-172)")
        self._act_buf[0] = int(action)
        p = self._ptr
        rc = self._L.st_host_step(self._h, p["act"], p["obs"], p["reward"], p["done"], p["info"])
        if rc:
            native.check(rc, "st_host_step")
        if self._queue_set:  # only an injected piece queue can raise a device-side error here
            err = C.c_int32(0)
            native.check(self._L.st_host_poll(self._h, C.byref(err), None), "st_host_poll")
            if err.value & 1:
                raise IndexError("injected piece queue exhausted")
        obs = self._obs_buf.copy()  # the reference returns a fresh array every step (tetris_env.py:400)
        r = float(self._reward_buf[0])
        done = self._done_buf
        i = self._info_buf[0]
        info = {"time": int(i[2]), "current_piece": shape_names[i[0]], "score": int(i[3]),
                "lines_cleared": int(i[4]), "holes": int(i[5]), "deaths": int(i[7]),
                "statistics": {n: int(i[8 + k]) for k, n in enumerate(shape_names)}}
        return obs, (int(r) if r.is_integer() else r), bool(done[0]), info

    def reset(self, return_info=False):
        obs = np.empty(self._obs_shape, dtype=np.float32)
        native.check(self._L.st_host_reset(self._h, None, obs.ctypes.data), "st_host_reset")
        self._was_reset = True
        return (obs, self._get_info()) if return_info else obs

    def _observation(self):
        """engine.render() through _observation (tetris_env.py:413-433): board with the piece drawn."""
        obs = np.empty(self._obs_shape, dtype=np.float32)
        native.check(self._L.st_host_observe(self._h, 1, obs.ctypes.data), "st_host_observe")
        return obs

    def _render_frame(self, size):
        out = np.empty((size, size, 3), dtype=np.uint8)
        native.check(self._L.st_host_render(self._h, 1, size, out.ctypes.data), "st_host_render")
        return out

    def render(self, mode="human"):
        """mode='rgb_array': uint8 (160, 160, 3) image of the board with the active piece (tetris_env.py:458-462),
        written by the `st_render` kernel.  mode='human' (tetris_env.py:437-457): the same writer at window_size,
        shown in a pygame window; pygame is imported here, on first use, so the step path never needs it."""
        if mode == "rgb_array":
            return self._render_frame(160)
        if mode == "human":
            try:
                import pygame
            except ImportError as e:
                raise ImportError("render('human') needs pygame (pygame>=2.1.0, as the reference's setup.py:20 "
                                  "requires); render('rgb_array') does not") from e
            if self.window is None:
                pygame.init()
                pygame.display.init()
                self.window = pygame.display.set_mode((self.window_size, self.window_size))
            if self.clock is None:
                self.clock = pygame.time.Clock()
            # ref:444-447 renders the TRANSPOSED board: convert_grayscale(board.T, S) == convert_grayscale(board, S).T,
            # which is what pygame's [x][y] surface arrays want
            obs = np.ascontiguousarray(self._render_frame(self.window_size).transpose(1, 0, 2))
            pygame.pixelcopy.array_to_surface(self.window, obs)
            canvas = pygame.surfarray.make_surface(obs)
            self.window.blit(canvas, canvas.get_rect())
            pygame.event.pump()
            pygame.display.update()
            self.clock.tick(self.metadata["render_fps"])
            return None
        return None  # the reference defers to gym.Env.render, which does nothing for unknown modes

    def close(self):
        if getattr(self, "_h", None):
            self._L.st_host_destroy(self._h)
            self._h = None
        if getattr(self, "_pinned", None):
            self._obs_buf = self._info_buf = self._reward_buf = self._done_buf = self._act_buf = None
            self._L.st_host_free_pinned(self._pinned)
            self._pinned = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class TetrisEnvV26(TetrisEnv):
    """The same env behind the gym >= 0.26 / gymnasium calling convention:
    `reset(seed=None, options=None) -> (obs, info)` and `step(a) -> (obs, reward, terminated, truncated, info)`.
    The reference only speaks the old API (tetris_env.py:397-411); nothing else changes."""

    def reset(self, *, seed=None, options=None):  # noqa: ARG002
        if seed is not None:
            native.check(self._L.st_host_set_seed(self._h, int(seed) & (2 ** 64 - 1)), "st_host_set_seed")
        obs, info = TetrisEnv.reset(self, return_info=True)
        return obs, info

    def step(self, action):
        obs, reward, done, info = TetrisEnv.step(self, action)
        return obs, reward, done, False, info  # Tetris never truncates: episodes end by topping out (ref:277-281)
