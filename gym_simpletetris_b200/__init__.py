"""gym_simpletetris_b200 — B200-native batched SimpleTetris behind the gym-simpletetris API.

    import gym_simpletetris_b200 as st
    env = st.make("SimpleTetris-v0", obs_type="grayscale")        # single env, NumPy, reference API
    vec = st.VecEnv(262144, obs_type="rgb", device="cuda:0")      # batched, torch tensors

When `gym` (or `gymnasium`) is importable the id 'SimpleTetris-v0' is registered exactly like
gym_simpletetris/__init__.py:3-6 does, so `gym.make('SimpleTetris-v0', **kwargs)` works unchanged.
"""
from . import native  # noqa: F401
from .envs import TetrisEnv, TetrisEnvV26  # noqa: F401
from .host_env import HostVecEnv  # noqa: F401
from .sharding import all_reduce_sum, make_sharded_vec_env, shard_bounds  # noqa: F401
from .vec_env import SHAPE_NAMES, VecEnv  # noqa: F401

ENV_ID = "SimpleTetris-v0"
ENTRY_POINT = "gym_simpletetris_b200.envs:TetrisEnv"
__version__ = "0.1.0"


def make(env_id=ENV_ID, **kwargs):
    """`gym.make` stand-in for images without gym: only knows 'SimpleTetris-v0'."""
    if env_id != ENV_ID:
        raise KeyError(f"unknown env id {env_id!r}; this package provides {ENV_ID!r}")
    return TetrisEnv(**kwargs)


def _register():
    # gym <= 0.25 speaks the reference's 4-tuple API; gymnasium (and gym >= 0.26) the 5-tuple one
    for modname, entry in (("gym", ENTRY_POINT), ("gymnasium", ENTRY_POINT + "V26")):
        try:
            mod = __import__(modname + ".envs.registration", fromlist=["register"])
            if modname == "gym" and tuple(int(v) for v in __import__("gym").__version__.split(".")[:2]) >= (0, 26):
                entry = ENTRY_POINT + "V26"
            mod.register(id=ENV_ID, entry_point=entry)
        except Exception:  # noqa: BLE001  (not installed, or id already registered)
            pass


_register()
