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
    assert L.st_step_kernel_name(C.byref(ram), 4096).decode() == "st_step_cols_kernel"   # column lanes: small batches
    assert L.st_step_kernel_name(C.byref(ram), 1 << 20).decode() == "st_step_tpe_kernel"  # thread per env: large ones
    wide = native.make_config(**{**{f[0]: getattr(ram, f[0]) for f in ram._fields_}, "obs_type": "ram", "width": 28})
    assert L.st_step_kernel_name(C.byref(wide), 4096).decode() == "st_main_kernel<ram,STEP>"  # > 24 columns: row lanes


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


def test_gym_registration_path():
    """gym_simpletetris/__init__.py:3-6: `register(id='SimpleTetris-v0', entry_point=...)` — run against stand-in
    `gym` / `gymnasium` registration modules (neither package is installed in this image)."""
    import importlib
    import sys
    import types

    calls = []

    def fake(modname, version):
        top = types.ModuleType(modname)
        top.__version__ = version
        envs = types.ModuleType(modname + ".envs")
        reg = types.ModuleType(modname + ".envs.registration")
        reg.register = lambda id, entry_point, **kw: calls.append((modname, id, entry_point))  # noqa: A002
        top.envs, envs.registration = envs, reg
        return {modname: top, modname + ".envs": envs, modname + ".envs.registration": reg}

    saved = {k: sys.modules.get(k) for k in ("gym", "gym.envs", "gym.envs.registration", "gymnasium", "gymnasium.envs",
                                             "gymnasium.envs.registration")}
    try:
        sys.modules.update(fake("gym", "0.21.0"))
        sys.modules.update(fake("gymnasium", "0.29.1"))
        st._register()
        assert ("gym", "SimpleTetris-v0", "gym_simpletetris_b200.envs:TetrisEnv") in calls  # 4-tuple API for gym <= 0.25
        assert ("gymnasium", "SimpleTetris-v0", "gym_simpletetris_b200.envs:TetrisEnvV26") in calls
        calls.clear()
        sys.modules.update(fake("gym", "0.26.2"))
        st._register()
        assert ("gym", "SimpleTetris-v0", "gym_simpletetris_b200.envs:TetrisEnvV26") in calls
        # the entry point strings resolve to the classes
        for _, _, entry in calls:
            mod, cls = entry.split(":")
            assert getattr(importlib.import_module(mod), cls) in (st.TetrisEnv, st.TetrisEnvV26)
        # a registry that refuses (id already registered) must not break the import
        sys.modules["gym.envs.registration"].register = lambda **kw: (_ for _ in ()).throw(RuntimeError("dup"))
        st._register()
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_host_vec_env_has_no_cpu_path():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        st.HostVecEnv(8)
