"""`VecEnv`: N SimpleTetris instances stepped by one CUDA launch.

Same constructor kwargs as the reference `TetrisEnv` (tetris_env.py:343-357); the step/reset results
are the batched form of tetris_env.py:397-411, with gym<=0.25 vector auto-reset: when an env is done,
the returned observation is its reset observation and reward/done/info are the terminal step's.
All buffers are torch tensors on the env's device; the work happens in `libsimpletetris_b200.so`
(hand-written sm_100a kernels) through the C ABI of include/simpletetris_b200.h.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import native
from .native import ST_INFO_WORDS, ST_STATS_WORDS, ST_UNPACKED_WORDS, StAux

SHAPE_NAMES = ["T", "J", "L", "Z", "S", "I", "O"]  # piece ids 0..6, tetris_env.py:19
# columns of the info tensor (int32 [N, 15]); get_info keys of tetris_env.py:232-241
INFO_COLS = {"current_piece": 0, "lock_delay_counter": 1, "time": 2, "score": 3, "lines_cleared": 4, "holes": 5,
             "piece_height": 6, "deaths": 7}
INFO_KEYS = ("time", "current_piece", "score", "lines_cleared", "holes", "deaths", "statistics")


def obs_shape(width, height, obs_type, extend_dims):
    """Per-env observation shape, tetris_env.py:381-392."""
    if obs_type == "ram":
        return (width, height, 1) if extend_dims else (width, height)
    if obs_type == "grayscale":
        return (84, 84, 1) if extend_dims else (84, 84)
    return (84, 84, 3)


class VecEnv:
    def __init__(self, num_envs, width=10, height=20, obs_type="ram", extend_dims=False, render_mode="rgb_array",
                 reward_step=False, penalise_height=False, penalise_height_increase=False, advanced_clears=False,
                 high_scoring=False, penalise_holes=False, penalise_holes_increase=False, lock_delay=0,
                 step_reset=False, *, device="cuda", seed=0, env_id_base=0, auto_reset=True, with_info=True,
                 obs_dtype=torch.float32, terminal_obs=False):
        self._L = native.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("gym_simpletetris_b200.VecEnv needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"VecEnv device must be a CUDA device, got {device!r}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self.width, self.height, self.obs_type, self.extend_dims = width, height, obs_type, extend_dims
        self.render_mode = render_mode
        self.seed, self.env_id_base = int(seed), int(env_id_base)
        if obs_dtype not in (torch.float32, torch.uint8):
            raise ValueError("obs_dtype must be torch.float32 (the reference's dtype) or torch.uint8")
        self.obs_dtype = obs_dtype  # uint8: same values, a quarter of the observation bytes (not the parity mode)
        self.cfg = native.make_config(
            width=width, height=height, obs_type=obs_type, extend_dims=extend_dims, lock_delay=lock_delay,
            step_reset=step_reset, reward_step=reward_step, penalise_height=penalise_height,
            penalise_height_increase=penalise_height_increase, advanced_clears=advanced_clears,
            high_scoring=high_scoring, penalise_holes=penalise_holes,
            penalise_holes_increase=penalise_holes_increase, auto_reset=auto_reset, device=self.device.index,
            seed=seed, env_id_base=env_id_base, obs_u8=obs_dtype == torch.uint8)
        stride = self._L.st_state_stride(C.byref(self.cfg))
        if stride <= 0:
            raise ValueError("unsupported geometry: width must be 1..32 and height 1..63")
        self.state_stride = int(stride)
        self.obs_elems = int(self._L.st_obs_elems(C.byref(self.cfg)))
        self.single_observation_shape = obs_shape(width, height, obs_type if obs_type in native.OBS_TYPES else "rgb",
                                                  extend_dims)
        n, dev = self.num_envs, self.device
        self.state = self._zeros(n * self.state_stride, torch.uint8)
        self.obs = self._zeros((n,) + self.single_observation_shape, obs_dtype)
        self.reward = self._zeros(n, torch.float32)
        self.done = self._zeros(n, torch.bool)
        self.info_buf = self._zeros((n, ST_INFO_WORDS), torch.int32) if with_info else None
        self.err = self._zeros(1, torch.int32)
        self.stats = self._zeros(ST_STATS_WORDS, torch.int64)
        self._queue = None
        # gym<=0.25 vector envs report the last observation of a finished episode in info["terminal_observation"]
        # (the returned obs is already the reset one); optional because it is a second observation-sized buffer
        self.term_obs = self._zeros(tuple(self.obs.shape), obs_dtype) if (terminal_obs and auto_reset) else None
        self._aux = StAux(None, 0, 0, self.err.data_ptr(), self.stats.data_ptr(),
                          self.term_obs.data_ptr() if self.term_obs is not None else None)
        # per-step host cost matters for small batches: everything constant across steps is bound once
        self._cfg_ref, self._aux_ref = C.byref(self.cfg), C.byref(self._aux)
        self._info_views = self._info(self.info_buf)
        self._ptrs = (self.state.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(),
                      self.info_buf.data_ptr() if self.info_buf is not None else None)
        self._st_step = self._L.st_step
        self._closed = False
        self._check(self._L.st_init(C.byref(self.cfg), self.state.data_ptr(), n, self._stream()), "st_init")

    # ---- plumbing ----
    def _zeros(self, shape, dtype):
        """Every persistent device buffer the kernels write is allocated here, zero-filled (tests override `_empty` to
        put guard bands around all of them)."""
        return self._empty(shape, dtype).zero_()

    def _empty(self, shape, dtype):
        """Per-call outputs (rollout buffers, observe / render results): the kernels write every byte of them."""
        return torch.empty(shape, dtype=dtype, device=self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _check(rc, what):
        native.check(rc, what)

    def _actions(self, actions, shape):
        if (torch.is_tensor(actions) and actions.dtype == torch.uint8 and actions.device == self.device
                and actions.shape == shape and actions.is_contiguous()):
            return actions
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions))
        if actions.dtype != torch.uint8:
            actions = actions.to(torch.uint8)
        actions = actions.to(self.device, non_blocking=True).contiguous()
        if tuple(actions.shape) != tuple(shape):
            raise ValueError(f"actions must have shape {shape}, got {tuple(actions.shape)}")
        return actions

    def _info(self, buf=None, terminal=True):
        buf = self.info_buf if buf is None else buf
        if buf is None:
            return {}
        d = {k: buf[..., INFO_COLS[k]] for k in INFO_KEYS if k != "statistics"}
        d["statistics"] = buf[..., 8:15]
        if terminal and getattr(self, "term_obs", None) is not None and buf is self.info_buf:
            d["terminal_observation"] = self.term_obs  # rows are valid where `done` is set in the same step
        return d

    def _live(self):
        if self._closed:  # the reference deletes its engine in close() (tetris_env.py:466-467): later calls raise
            raise RuntimeError("VecEnv is closed")

    # ---- the reference API, batched ----
    def reset(self, mask=None):
        """TetrisEnv.reset (tetris_env.py:405-411) for all envs, or those with mask[e] != 0.  Returns obs [N,...]."""
        self._live()
        mptr = None
        if mask is not None:
            mask = torch.as_tensor(mask).to(self.device).to(torch.uint8).contiguous()
            mptr = mask.data_ptr()
        self._check(self._L.st_reset(C.byref(self.cfg), self.state.data_ptr(), mptr, self.obs.data_ptr(),
                                     C.byref(self._aux), self.num_envs, self._stream()), "st_reset")
        return self.obs

    def step(self, actions):
        """TetrisEnv.step (tetris_env.py:397-403) for N envs: (obs [N,...] f32, reward [N] f32, done [N] bool, info).
        The returned tensors are the env's own buffers and are overwritten by the next call."""
        if self._closed:
            self._live()
        a = self._actions(actions, (self.num_envs,))
        state, obs, reward, done, info = self._ptrs
        rc = self._st_step(self._cfg_ref, state, a.data_ptr(), obs, reward, done, info, self._aux_ref, self.num_envs,
                           torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            self._check(rc, "st_step")
        return self.obs, self.reward, self.done, self._info_views

    def step_many(self, actions, rollout_obs=False, rollout_info=False):
        """T steps in one launch.  actions [T, N].  Returns (obs, reward [T,N], done [T,N], info): obs is
        [T,N,...] if rollout_obs else the last step's [N,...]; same for info."""
        self._live()
        T = int(actions.shape[0])
        a = self._actions(actions, (T, self.num_envs))
        n, dev = self.num_envs, self.device
        reward = self._empty((T, n), torch.float32)
        done = self._empty((T, n), torch.bool)
        obs = self._empty((T,) + tuple(self.obs.shape), self.obs_dtype) if rollout_obs else self.obs
        info = self.info_buf
        if rollout_info and info is not None:
            info = self._empty((T, n, ST_INFO_WORDS), torch.int32)
        self._check(self._L.st_step_many(
            C.byref(self.cfg), self.state.data_ptr(), a.data_ptr(), T, obs.data_ptr(),
            n * self.obs_elems if rollout_obs else 0, reward.data_ptr(), done.data_ptr(),
            info.data_ptr() if info is not None else None, n * ST_INFO_WORDS if rollout_info else 0,
            C.byref(self._aux_many()), n, self._stream()), "st_step_many")
        # the terminal-observation buffer holds one step: step_many does not fill it, so it is not reported here
        return obs, reward, done, self._info(info, terminal=False)

    def capture_step(self):
        """CUDA-graph the one-step launch: returns `GraphedStep`, whose call copies the
        given actions into a static device buffer and replays the graph — no per-step ctypes/launch cost on the
        host, so a device-resident policy can drive small batches at kernel rate."""
        return GraphedStep(self)

    def _aux_many(self):
        """StAux for st_step_many: no terminal-observation buffer (it is sized for one step; use step() for those)."""
        a = self._aux
        return StAux(a.piece_queue, a.queue_len, 0, a.error_flag, a.stats, None)

    def observe(self, draw_piece=True):
        """_observation(engine.render()) (tetris_env.py:317-321, 413-433) of the current state, no step."""
        self._live()
        out = self._empty(tuple(self.obs.shape), self.obs_dtype)
        self._check(self._L.st_observe(C.byref(self.cfg), self.state.data_ptr(), int(bool(draw_piece)),
                                       out.data_ptr(), self.num_envs, self._stream()), "st_observe")
        return out

    def render(self, size=160, draw_piece=True):
        """Batched TetrisEnv.render('rgb_array') (tetris_env.py:458-462): uint8 [N, size, size, 3]."""
        self._live()
        out = self._empty((self.num_envs, size, size, 3), torch.uint8)
        self._check(self._L.st_render(C.byref(self.cfg), self.state.data_ptr(), int(bool(draw_piece)), int(size),
                                      out.data_ptr(), self.num_envs, self._stream()), "st_render")
        return out

    def close(self):
        """TetrisEnv.close (tetris_env.py:466-467).  The cached raw device pointers go with the tensors: any later
        step / reset / graph replay raises instead of launching onto memory the allocator has handed to someone else."""
        if self._closed:
            return
        if self.state is not None and self.state.is_cuda:
            torch.cuda.synchronize(self.device)
        self._closed = True
        self._ptrs = (None,) * 5
        self._aux = self._aux_ref = self._info_views = None
        self.state = self.obs = self.reward = self.done = self.info_buf = self.term_obs = self._queue = None

    # ---- piece source / state injection / diagnostics ----
    def set_piece_queue(self, queue):
        """queue [N, Q] of piece ids 0..6 (or None): the k-th piece of env e's lifetime becomes queue[e, k]
        instead of a draw of _choose_shape (tetris_env.py:183-191)."""
        self._live()
        if queue is None:
            self._queue = None
            self._aux.piece_queue, self._aux.queue_len = None, 0
            return
        q = torch.as_tensor(np.asarray(queue)).to(torch.uint8)
        if q.dim() == 1:
            q = q.unsqueeze(0).expand(self.num_envs, -1)
        if q.shape[0] != self.num_envs:
            raise ValueError("queue must be [N, Q]")
        self._queue = q.contiguous().to(self.device)
        self._aux.piece_queue, self._aux.queue_len = self._queue.data_ptr(), int(self._queue.shape[1])

    def get_state(self):
        """(boards uint8 [N,W,H], scalars int32 [N,18]): id, rot, x, y, lock-delay counter, time, score,
        lines_cleared, holes, piece_height, deaths, shape_counts[7]."""
        self._live()
        boards = torch.empty((self.num_envs, self.width, self.height), dtype=torch.uint8, device=self.device)
        scalars = torch.empty((self.num_envs, ST_UNPACKED_WORDS), dtype=torch.int32, device=self.device)
        self._check(self._L.st_get_state(C.byref(self.cfg), self.state.data_ptr(), boards.data_ptr(),
                                         scalars.data_ptr(), self.num_envs, self._stream()), "st_get_state")
        return boards, scalars

    def set_state(self, boards=None, scalars=None):
        self._live()
        b = s = None
        if boards is not None:
            b = torch.as_tensor(np.asarray(boards.cpu() if torch.is_tensor(boards) else boards)).to(torch.uint8)
            b = b.to(self.device).contiguous()
            assert tuple(b.shape) == (self.num_envs, self.width, self.height)
        if scalars is not None:
            s = torch.as_tensor(np.asarray(scalars.cpu() if torch.is_tensor(scalars) else scalars)).to(torch.int32)
            s = s.to(self.device).contiguous()
            assert tuple(s.shape) == (self.num_envs, ST_UNPACKED_WORDS)
        self._check(self._L.st_set_state(C.byref(self.cfg), self.state.data_ptr(),
                                         b.data_ptr() if b is not None else None,
                                         s.data_ptr() if s is not None else None, self.num_envs, self._stream()),
                    "st_set_state")

    # ---- checkpoint / resume: the whole simulator state is one tensor of env records + four counters ----
    def state_dict(self):
        return {"state": self.state.clone(), "stats": self.stats.clone(), "seed": self.seed,
                "env_id_base": self.env_id_base, "num_envs": self.num_envs, "stride": self.state_stride}

    def load_state_dict(self, sd):
        if sd["num_envs"] != self.num_envs or sd["stride"] != self.state_stride:
            raise ValueError("state_dict is for a different batch size or board geometry")
        if sd["seed"] != self.seed or sd["env_id_base"] != self.env_id_base:
            raise ValueError("state_dict was saved with a different seed / env_id_base: piece streams would differ")
        self.state.copy_(sd["state"])
        self.stats.copy_(sd["stats"])

    def poll_errors(self) -> int:
        """Sticky ST_ERR_* bits raised by the kernels since the last poll (synchronises)."""
        e = int(self.err.item())
        if e:
            self.err.zero_()
        return e

    def episode_stats(self, reduce=True):
        """dict(episodes, length_sum, lines_sum, score_sum) accumulated at every done; summed over all ranks
        with one all_reduce when torch.distributed is initialised (the only collective, off the step path)."""
        from .sharding import all_reduce_sum

        s = self.stats.clone()
        if reduce:
            s = all_reduce_sum(s)
        e, ln, li, sc = (int(v) for v in s.tolist())
        return {"episodes": e, "length_sum": ln, "lines_sum": li, "score_sum": sc}


class GraphedStep:
    """`g = env.capture_step(); obs, reward, done, info = g(actions)` — one CUDA-graph replay per step.
    `g.actions` is the static uint8 [N] device buffer a policy kernel can write into directly (then call `g()`)."""

    def __init__(self, env: VecEnv):
        self.env = env
        self.actions = torch.full((env.num_envs,), 6, dtype=torch.uint8, device=env.device)
        side = torch.cuda.Stream(device=env.device)
        side.wait_stream(torch.cuda.current_stream(env.device))
        with torch.cuda.stream(side):
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self.out = env.step(self.actions)
        torch.cuda.current_stream(env.device).wait_stream(side)

    def __call__(self, actions=None):
        self.env._live()  # a replay after close() would write through pointers the allocator has recycled
        if actions is not None:
            self.actions.copy_(actions, non_blocking=True)
        self.graph.replay()
        return self.out
