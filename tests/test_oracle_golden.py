"""Pins oracle/st_oracle.c against fixtures generated from the unmodified reference."""
import numpy as np
import pytest

from _golden import digest, golden
from oracle.oracle import SHAPE_NAMES, OracleEnv

ROLL = golden("rollouts.npz")
SCEN = golden("scenarios.npz")


def info_row(info):
    return [info["time"], SHAPE_NAMES.index(info["current_piece"]), info["score"], info["lines_cleared"],
            info["holes"], info["deaths"]] + [info["statistics"][n] for n in SHAPE_NAMES]


@pytest.mark.parametrize("key", ROLL.keys())
def test_rollout_matches_reference(key):
    g = lambda f: ROLL.get(key, f)
    env = OracleEnv(pieces=g("pieces"), **ROLL.kwargs(key))
    resets = [digest(env.reset())]
    for t, a in enumerate(g("actions")):
        obs, r, d, info = env.step(int(a))
        assert r == g("reward")[t], (key, t)
        assert d == bool(g("done")[t]), (key, t)
        assert info_row(info) == g("info")[t].tolist(), (key, t)
        assert digest(obs) == g("digest")[t], (key, t)
        if d:
            resets.append(digest(env.reset()))
    assert resets == g("reset_digest").tolist()
    assert env.error == 0


@pytest.mark.parametrize("key", SCEN.keys())
def test_scenario_matches_reference(key):
    g = lambda f: SCEN.get(key, f)
    env = OracleEnv(pieces=g("pieces"), **SCEN.kwargs(key))
    assert digest(env.reset()) == g("reset_digest")[0]
    env.board = g("board").astype(np.float64)
    for t, a in enumerate(g("actions")):
        obs, r, d, info = env.step(int(a))
        assert r == g("reward")[t], (key, t)
        assert d == bool(g("done")[t]), (key, t)
        assert info_row(info) == g("info")[t].tolist(), (key, t)
        assert digest(obs) == g("digest")[t], (key, t)
        pid, rot, x, y, ld, ph = env.piece()
        assert [x, y, ld, ph] == g("anchor")[t].tolist(), (key, t)
    assert np.array_equal(env.board.astype(np.uint8), g("final_board"))


def test_survey_b2_reward_table():
    """SURVEY.md B.2 numbers, written out (independent of the .npz)."""
    table = {
        "default": [0, 100, 200, 300, 400], "reward_step": [1, 101, 201, 301, 401],
        "penalise_height": [-5, 96, 197, 298, 399], "penalise_height_increase": [-50, 60, 170, 280, 390],
        "advanced_clears": [0, 100, 250, 750, 3000], "high_scoring": [0, 1000, 2000, 3000, 4000],
        "penalise_holes": [-45, 60, 165, 270, 375], "penalise_holes_increase": [-45, 60, 165, 270, 375],
        "all7": [-49, 57, 213, 719, 2975], "C2": [1, 101, 251, 751, 3001],
    }
    for name, want in table.items():
        got = [SCEN.get(f"b2_{name}_k{k}", "reward")[0] for k in range(5)]
        assert got == want, name
    # C3 locks after the 3-step delay
    assert [SCEN.get(f"b2_C3_k{k}", "reward")[3] for k in range(5)] == [-95, 20, 135, 250, 365]


def test_survey_b4_five_I():
    g = lambda f: SCEN.get("b4_five_I", f)
    assert g("done").tolist()[:7] == [0, 0, 0, 0, 1, 1, 1]
    assert g("reward").tolist()[4:7] == [-100, -100, -100]
    assert g("info")[4:7, 5].tolist() == [1, 2, 3]  # deaths


DBG = golden("debug.npz")


def u8_digest(a):
    import hashlib

    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint8).tobytes()).digest()[:8], dtype=np.uint64)[0]


@pytest.mark.parametrize("key", DBG.keys())
def test_repr_and_human_frame_match_reference(key):
    """tetris_env.py:329-335 (`__repr__`) and :437-457 (`render('human')` frame) after reset and after every step."""
    g = lambda f: DBG.get(key, f)
    want_repr = bytes(g("repr")).decode().split("\x00")
    env = OracleEnv(pieces=g("pieces"), **DBG.kwargs(key))
    env.reset()
    assert repr(env) == want_repr[0]
    assert np.array_equal(env.human_frame(), g("human_first"))
    for t, a in enumerate(g("actions")):
        _, _, d, _ = env.step(int(a))
        assert d == bool(g("done")[t])
        if d:
            env.reset()
        assert repr(env) == want_repr[t + 1], (key, t)
        assert u8_digest(env.human_frame()) == g("human_digest")[t + 1], (key, t)
