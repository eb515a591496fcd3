"""ctypes binding of include/simpletetris_b200.h (the drop-in C ABI).

The product path has NO fallback: if `libsimpletetris_b200.so` is missing or lacks a symbol the
import of this module's `lib()` raises, and so does every env built on it.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("ST_B200_LIB") or os.path.join(_HERE, "libsimpletetris_b200.so")  # override: kernel experiments

ST_STATE_WORDS = 15
ST_INFO_WORDS = 15
ST_UNPACKED_WORDS = 18
ST_STATS_WORDS = 4
OBS_TYPES = {"ram": 0, "grayscale": 1, "rgb": 2}
ERR_BITS = {1: "piece queue exhausted", 2: "action outside 0..6", 4: "step() before reset()"}


class StConfig(C.Structure):
    _fields_ = [(k, C.c_int32) for k in (
        "width", "height", "obs_type", "extend_dims", "lock_delay", "step_reset", "reward_step",
        "penalise_height", "penalise_height_increase", "advanced_clears", "high_scoring", "penalise_holes",
        "penalise_holes_increase", "auto_reset", "device", "obs_u8")] + [
        ("seed", C.c_uint64), ("env_id_base", C.c_int64)]


class StAux(C.Structure):
    _fields_ = [("piece_queue", C.c_void_p), ("queue_len", C.c_int32), ("reserved", C.c_int32),
                ("error_flag", C.c_void_p), ("stats", C.c_void_p), ("terminal_obs", C.c_void_p)]


# every symbol include/simpletetris_b200.h declares: name -> (restype, argtypes)
_P, _I32, _I64 = C.c_void_p, C.c_int32, C.c_int64
_CFG, _AUX = C.POINTER(StConfig), C.POINTER(StAux)
SYMBOLS = {
    "st_state_stride": (_I64, [_CFG]),
    "st_obs_elems": (_I64, [_CFG]),
    "st_init": (C.c_int, [_CFG, _P, _I64, _P]),
    "st_reset": (C.c_int, [_CFG, _P, _P, _P, _AUX, _I64, _P]),
    "st_step": (C.c_int, [_CFG, _P, _P, _P, _P, _P, _P, _AUX, _I64, _P]),
    "st_step_many": (C.c_int, [_CFG, _P, _P, _I32, _P, _I64, _P, _P, _P, _I64, _AUX, _I64, _P]),
    "st_observe": (C.c_int, [_CFG, _P, _I32, _P, _I64, _P]),
    "st_render": (C.c_int, [_CFG, _P, _I32, _I32, _P, _I64, _P]),
    "st_get_state": (C.c_int, [_CFG, _P, _P, _P, _I64, _P]),
    "st_set_state": (C.c_int, [_CFG, _P, _P, _P, _I64, _P]),
    "st_host_create": (_P, [_CFG, _I64]),
    "st_host_destroy": (None, [_P]),
    "st_host_set_piece_queue": (C.c_int, [_P, _P, _I32]),
    "st_host_reset": (C.c_int, [_P, _P, _P]),
    "st_host_step": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "st_host_observe": (C.c_int, [_P, _I32, _P]),
    "st_host_step_async": (C.c_int, [_P, _P]),
    "st_host_wait": (C.c_int, [_P, _P, _P, _P, _P]),
    "st_host_set_zero_copy": (C.c_int, [_P, _I32]),
    "st_host_set_seed": (C.c_int, [_P, C.c_uint64]),
    "st_host_alloc_pinned": (_P, [C.c_size_t]),
    "st_host_free_pinned": (None, [_P]),
    "st_host_render": (C.c_int, [_P, _I32, _I32, _P]),
    "st_host_get_state": (C.c_int, [_P, _P, _P]),
    "st_host_set_state": (C.c_int, [_P, _P, _P]),
    "st_host_poll": (C.c_int, [_P, _P, _P]),
    "st_last_error": (C.c_char_p, []),
    "st_step_kernel_name": (C.c_char_p, [_CFG, _I64]),
    "st_abi_version": (C.c_int, []),
    "st_launch_count": (C.c_uint64, []),
}

_lib = None


def lib():
    """Load the CUDA library (once).  Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} not found: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "gym_simpletetris_b200 has no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the ABI symbol is missing
            fn.restype, fn.argtypes = res, args
        if L.st_abi_version() != 1:
            raise RuntimeError("libsimpletetris_b200.so: ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int, what: str = "simpletetris_b200"):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().st_last_error().decode()}")


def make_config(*, width, height, obs_type, extend_dims, lock_delay, step_reset, reward_step, penalise_height,
                penalise_height_increase, advanced_clears, high_scoring, penalise_holes, penalise_holes_increase,
                auto_reset, device, seed, env_id_base, obs_u8=False) -> StConfig:
    # an unknown obs_type falls through to the rgb branch in the reference (tetris_env.py:432-433)
    ot = OBS_TYPES.get(obs_type, 2)
    return StConfig(int(width), int(height), ot, int(bool(extend_dims)), int(lock_delay), int(bool(step_reset)),
                    int(bool(reward_step)), int(bool(penalise_height)), int(bool(penalise_height_increase)),
                    int(bool(advanced_clears)), int(bool(high_scoring)), int(bool(penalise_holes)),
                    int(bool(penalise_holes_increase)), int(bool(auto_reset)), int(device), int(bool(obs_u8)),
                    int(seed) & (2 ** 64 - 1), int(env_id_base))


def launch_count() -> int:
    return int(lib().st_launch_count())
