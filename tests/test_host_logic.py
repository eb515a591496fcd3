"""Host-side pieces that need no GPU: spaces stand-ins, observation shapes, config marshalling, env-id routing."""
import ctypes as C

import numpy as np
import pytest

import gym_simpletetris_b200 as st
from gym_simpletetris_b200 import native
from gym_simpletetris_b200.spaces import Box, Discrete
from gym_simpletetris_b200.vec_env import INFO_COLS, obs_shape


def test_obs_shapes_follow_reference():  # tetris_env.py:381-392
    assert obs_shape(10, 20, "ram", False) == (10, 20)
    assert obs_shape(10, 20, "ram", True) == (10, 20, 1)
    assert obs_shape(7, 9, "grayscale", False) == (84, 84)
    assert obs_shape(7, 9, "grayscale", True) == (84, 84, 1)
    assert obs_shape(7, 9, "rgb", True) == (84, 84, 3)


def test_spaces():
    a = Discrete(7)
    assert a.n == 7 and all(0 <= a.sample() < 7 for _ in range(50))
    b = Box(0, 1, shape=(10, 20), dtype=np.float32)
    assert tuple(b.shape) == (10, 20) and np.dtype(b.dtype) == np.float32


def test_config_marshalling_round_trip():
    cfg = native.make_config(width=12, height=24, obs_type="grayscale", extend_dims=True, lock_delay=-3,
                             step_reset=1, reward_step=True, penalise_height=False, penalise_height_increase=True,
                             advanced_clears=0, high_scoring=1, penalise_holes=False, penalise_holes_increase=True,
                             auto_reset=True, device=3, seed=2 ** 64 - 5, env_id_base=10 ** 12)
    assert (cfg.width, cfg.height, cfg.obs_type, cfg.extend_dims, cfg.lock_delay) == (12, 24, 1, 1, -3)
    assert (cfg.reward_step, cfg.penalise_height_increase, cfg.high_scoring, cfg.penalise_holes_increase) == (1, 1, 1, 1)
    assert (cfg.device, cfg.seed, cfg.env_id_base) == (3, 2 ** 64 - 5, 10 ** 12)
    L = native.lib()
    assert L.st_state_stride(C.byref(cfg)) == 60 + 24 * 2
    assert L.st_step_kernel_name(C.byref(cfg), 4096).decode().startswith("st_main_kernel<grayscale")
    ram = native.make_config(**{**{f[0]: getattr(cfg, f[0]) for f in cfg._fields_}, "obs_type": "ram"})
    assert L.st_step_kernel_name(C.byref(ram), 4096).decode() == "st_main_kernel<ram,STEP>"
    assert L.st_step_kernel_name(C.byref(ram), 1 << 20).decode() == "st_step_tpe_kernel"


def test_info_columns_match_header_order():
    # info row: piece id, lock-delay counter, time, score, lines_cleared, holes, piece_height, deaths, counts[7]
    assert INFO_COLS == {"current_piece": 0, "lock_delay_counter": 1, "time": 2, "score": 3, "lines_cleared": 4,
                         "holes": 5, "piece_height": 6, "deaths": 7}
    assert native.ST_INFO_WORDS == native.ST_STATE_WORDS == 15


def test_package_surface():
    assert st.ENV_ID == "SimpleTetris-v0" and st.ENTRY_POINT.endswith(":TetrisEnv")
    for name in ("VecEnv", "TetrisEnv", "TetrisEnvV26", "make", "shard_bounds", "make_sharded_vec_env"):
        assert hasattr(st, name)
    assert st.TetrisEnv.metadata == {"render.modes": ["human", "rgb_array"], "render_fps": 8}  # tetris_env.py:339
    with pytest.raises(KeyError):
        st.make("CartPole-v1")
