"""TEST INFRASTRUCTURE ONLY — loader for the UNMODIFIED reference file.

Loads `gym_simpletetris/envs/tetris_env.py` by path under an import shim (the
reference imports `gym`, `gym.spaces` and `pygame`, none of which exist in
this image, and uses the removed alias `np.float` at tetris_env.py:140).
The file is executed where it lies: `/root/reference/...` in the build
container, else the pip install of the unmodified reference under
`baseline/_ref/` (git-ignored; made by `__graft_entry__.build()`, it travels
to the GPU box so that `bench.py` can time the reference's own Python path
there).  No reference source is part of this repository.

Users: `tests/golden/make_golden.py`, the `not gpu` tests that pin the C
oracle (`oracle/st_oracle.c`), and `bench.py`'s CPU legs
(`oracle/ref_python_bench.py`).  `-m gpu` tests never load it.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
import warnings

_CANDIDATES = (
    "/root/reference/gym_simpletetris/envs/tetris_env.py",
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "gym_simpletetris",
                 "envs", "tetris_env.py"),
)


def REFERENCE_FILE() -> str:
    """Path of the unmodified reference file that will be executed ('' if there is none)."""
    for p in _CANDIDATES:
        if os.path.isfile(p):
            return p
    return ""

_module = None


def reference_available() -> bool:
    return bool(REFERENCE_FILE())


class _Env:  # stand-in for gym.Env (tetris_env.py:338)
    metadata = {}

    def render(self, mode="human"):
        raise NotImplementedError


class _Discrete:  # stand-in for gym.spaces.Discrete (tetris_env.py:377)
    def __init__(self, n):
        self.n = n


class _Box:  # stand-in for gym.spaces.Box (tetris_env.py:381-392)
    def __init__(self, low, high, shape=None, dtype=None):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


def load_reference():
    """Return the reference module (executed from its own path, unmodified)."""
    global _module
    if _module is not None:
        return _module
    if not reference_available():
        raise FileNotFoundError(_CANDIDATES[0])
    import numpy as np

    if not hasattr(np, "float"):
        np.float = float  # tetris_env.py:140 uses the alias removed in numpy 1.24
    saved = {k: sys.modules.get(k) for k in ("gym", "gym.spaces", "pygame")}
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")
    gym.Env = _Env
    spaces.Discrete = _Discrete
    spaces.Box = _Box
    gym.spaces = spaces
    pygame = types.ModuleType("pygame")
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "pygame": pygame})
    try:
        spec = importlib.util.spec_from_file_location("_ref_tetris_env", REFERENCE_FILE())
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _module = mod
    return mod


def make_reference_env(pieces=None, **kwargs):
    """Reference `TetrisEnv(**kwargs)`; if `pieces` (iterable of letters or ids
    0..6 in the order of tetris_env.py:19) is given, `_choose_shape`
    (tetris_env.py:183-191, called at :198) is replaced by that sequence."""
    mod = load_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = mod.TetrisEnv(**kwargs)
    if pieces is not None:
        it = iter(pieces)

        def _next():
            p = next(it)
            return p if isinstance(p, str) else mod.shape_names[int(p)]

        env.engine._choose_shape = _next
    return env
