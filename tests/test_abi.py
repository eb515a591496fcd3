"""The C-ABI library loads and exports every symbol include/simpletetris_b200.h declares (no compute)."""
import ctypes as C
import os
import re

import pytest

import gym_simpletetris_b200 as st
from gym_simpletetris_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "simpletetris_b200.h")).read()
    return sorted(set(re.findall(r"^ST_API [^;(]*?\b(st_\w+)\(", text, flags=re.M)))


def test_every_declared_symbol_is_exported_and_bound():
    names = header_symbols()
    assert len(names) >= 20
    L = C.CDLL(native.SO_PATH)
    for n in names:
        assert hasattr(L, n), n
    assert sorted(native.SYMBOLS) == names
    assert native.lib().st_abi_version() == 1


def test_sizes_follow_survey_accounting():
    L = native.lib()
    mk = lambda **kw: native.make_config(**{**dict(
        width=10, height=20, obs_type="ram", extend_dims=False, lock_delay=0, step_reset=False, reward_step=False,
        penalise_height=False, penalise_height_increase=False, advanced_clears=False, high_scoring=False,
        penalise_holes=False, penalise_holes_increase=False, auto_reset=True, device=0, seed=0, env_id_base=0), **kw})
    c = mk()
    assert L.st_state_stride(C.byref(c)) == 100 and L.st_obs_elems(C.byref(c)) == 200
    c = mk(width=20, height=40)
    assert L.st_state_stride(C.byref(c)) == 220 and L.st_obs_elems(C.byref(c)) == 800
    assert L.st_obs_elems(C.byref(mk(obs_type="grayscale"))) == 84 * 84
    assert L.st_obs_elems(C.byref(mk(obs_type="rgb"))) == 84 * 84 * 3
    assert L.st_obs_elems(C.byref(mk(obs_type="bogus"))) == 84 * 84 * 3  # falls through to rgb like ref:432-433
    for bad in (dict(width=0), dict(width=33), dict(height=0), dict(height=64)):
        assert L.st_state_stride(C.byref(mk(**bad))) == -1


def test_struct_layout_matches_header():
    assert C.sizeof(native.StConfig) == 16 * 4 + 16
    assert C.sizeof(native.StAux) == 8 + 8 + 8 + 8 + 8


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        st.VecEnv(8)
    with pytest.raises(RuntimeError):
        st.make("SimpleTetris-v0")
    with pytest.raises(KeyError):
        st.make("Pong-v0")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gym_simpletetris_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f
