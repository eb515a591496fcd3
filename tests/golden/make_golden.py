"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Run:  python tests/golden/make_golden.py
The reference (/root/reference/gym_simpletetris/envs/tetris_env.py) is executed
in place through oracle/ref_shim.py; nothing of it is copied.  Two fixtures:

  rollouts.npz   per case of tests/_cases.py: `random.seed(seed)` so the
                 reference's own `_choose_shape` (tetris_env.py:183-191) picks
                 the pieces (recorded, so that the oracle / CUDA kernel can be
                 fed the same sequence), frozen-stream random actions, manual
                 reset on done.  Per step: reward (f64), done, the 13 info ints
                 (tetris_env.py:232-241) and a 64-bit digest of the float32
                 observation bytes; plus the digest of every reset observation.
  render.npz     digests of `render('rgb_array')` (uint8 160x160x3, tetris_env.py:458-462) after reset and
                 after every step of short rollouts (`--render-only` regenerates just this file).
  debug.npz      `repr(engine)` strings (tetris_env.py:329-335) and digests of the `render('human')` frames
                 (tetris_env.py:437-457, pygame replaced by a recording stand-in); `--debug-only` regenerates it.
  scenarios.npz  injected-board known answers: the reward table of SURVEY.md
                 B.2 for every flag combination used there, lock-delay traces
                 (B.3) and the edge cases of B.4, each as (initial board,
                 piece queue, actions) -> per-step (reward, done, info, digest,
                 anchor x/y, lock-delay counter).
"""
import hashlib
import itertools
import json
import os
import random
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.ref_shim import load_reference, make_reference_env  # noqa: E402
from _cases import CASES, SHAPE_NAMES, actions_for  # noqa: E402

warnings.simplefilter("ignore")


def digest(obs) -> np.uint64:
    a = np.ascontiguousarray(obs, dtype=np.float32)
    return np.frombuffer(hashlib.sha256(a.tobytes()).digest()[:8], dtype=np.uint64)[0]


def info_row(info):
    return [info["time"], SHAPE_NAMES.index(info["current_piece"]), info["score"], info["lines_cleared"],
            info["holes"], info["deaths"]] + [info["statistics"][n] for n in SHAPE_NAMES]


def record_pieces(env):
    """Wrap the reference's own chooser so that its picks are logged."""
    log = []
    orig = env.engine._choose_shape

    def wrapped():
        s = orig()
        log.append(SHAPE_NAMES.index(s))
        return s

    env.engine._choose_shape = wrapped
    return log


def rollout(kwargs, seed, T):
    random.seed(seed)
    env = make_reference_env(**kwargs)
    pieces = record_pieces(env)
    actions = actions_for(seed, T)
    reset_digests = [digest(env.reset())]
    rew, don, inf, dig = [], [], [], []
    for a in actions:
        obs, r, d, info = env.step(int(a))
        rew.append(float(r)); don.append(bool(d)); inf.append(info_row(info)); dig.append(digest(obs))
        if d:
            reset_digests.append(digest(env.reset()))
    return dict(actions=actions, pieces=np.asarray(pieces, np.uint8), reward=np.asarray(rew, np.float64),
                done=np.asarray(don, np.uint8), info=np.asarray(inf, np.int32),
                digest=np.asarray(dig, np.uint64), reset_digest=np.asarray(reset_digests, np.uint64))


def scenario(kwargs, pieces, actions, board=None, step_after_done=True):
    """reset -> (inject board) -> actions.  No reset on done (steps after done are part of the answer)."""
    env = make_reference_env(pieces=list(pieces) + ["O"] * 64, **kwargs)
    obs0 = env.reset()
    if board is not None:
        env.engine.board = np.array(board, dtype=np.float64)
    rew, don, inf, dig, anc = [], [], [], [], []
    for a in actions:
        obs, r, d, info = env.step(int(a))
        rew.append(float(r)); don.append(bool(d)); inf.append(info_row(info)); dig.append(digest(obs))
        anc.append([int(env.engine.anchor[0]), int(env.engine.anchor[1]), int(env.engine._lock_delay),
                    int(env.engine.piece_height)])
    W, H = kwargs.get("width", 10), kwargs.get("height", 20)
    return dict(board=np.zeros((W, H), np.uint8) if board is None else np.asarray(board, np.uint8),
                pieces=np.asarray([SHAPE_NAMES.index(p) for p in pieces] + [6] * 64, np.uint8),
                actions=np.asarray(actions, np.uint8), reward=np.asarray(rew, np.float64),
                done=np.asarray(don, np.uint8), info=np.asarray(inf, np.int32),
                digest=np.asarray(dig, np.uint64), anchor=np.asarray(anc, np.int32),
                reset_digest=np.asarray([digest(obs0)], np.uint64),
                final_board=np.asarray(env.engine.board, np.uint8))


FLAGSETS = {
    "default": {}, "reward_step": dict(reward_step=True), "penalise_height": dict(penalise_height=True),
    "penalise_height_increase": dict(penalise_height_increase=True), "advanced_clears": dict(advanced_clears=True),
    "high_scoring": dict(high_scoring=True), "penalise_holes": dict(penalise_holes=True),
    "penalise_holes_increase": dict(penalise_holes_increase=True),
    "adv_high": dict(advanced_clears=True, high_scoring=True),
    "both_height": dict(penalise_height=True, penalise_height_increase=True),
    "both_holes": dict(penalise_holes=True, penalise_holes_increase=True),
    "all7": dict(reward_step=True, penalise_height=True, penalise_height_increase=True, advanced_clears=True,
                 high_scoring=True, penalise_holes=True, penalise_holes_increase=True),
    "C2": dict(reward_step=True, advanced_clears=True),
    "C3": dict(penalise_height_increase=True, penalise_holes_increase=True, lock_delay=3, step_reset=True),
}


def b2_board(k, W=10, H=20):
    """bottom k rows full except column 5, plus a floating cell board[0,10]."""
    b = np.zeros((W, H), np.uint8)
    for y in range(H - k, H):
        b[:, y] = 1
        b[5, y] = 0
    b[0, 10] = 1
    return b


def make_scenarios():
    sc = {}
    for name, fl in FLAGSETS.items():
        for k in range(5):
            acts = [2] if fl.get("lock_delay", 0) == 0 else [2, 6, 6, 6]
            sc[f"b2_{name}_k{k}"] = (fl, scenario(fl, ["I"], acts + [6, 6], board=b2_board(k)))
    # B.3 lock delay: all O, idle until lock
    fl = dict(lock_delay=3)
    sc["b3_idle_ld3"] = (fl, scenario(fl, ["O", "O"], [6] * 26))
    sc["b3_slide_ld3"] = (fl, scenario(fl, ["O", "O"], [6] * 19 + [0, 0, 0, 0, 6, 6]))
    ledge_actions = [2, 1, 1, 1, 1, 6, 6, 6, 6, 6]
    for sr in (False, True):
        fl = dict(lock_delay=3, step_reset=sr)
        # I laid flat on the floor: rotate then hard drop (rot r1 puts cells at x-3..x) -> then O
        sc[f"b3_ledge_sr{int(sr)}"] = (fl, scenario(fl, ["I", "O", "O"], [4, 6, 6, 6] + [2, 6, 6, 6] + ledge_actions * 2))
    fl = dict(lock_delay=-5)
    sc["b3_ldneg"] = (fl, scenario(fl, ["O", "O", "T"], [2, 2, 6, 6, 2]))
    # B.4 edge cases
    sc["b4_five_I"] = ({}, scenario({}, ["I"] * 8, [2] * 5 + [6, 6, 0, 2]))
    b = np.zeros((10, 20), np.uint8); b[1:, 1] = 1
    fl = dict(lock_delay=50)
    sc["b4_Z_wall_ld50"] = (fl, scenario(fl, ["Z", "T"], [1] * 8 + [0] * 8 + [4, 5, 3], board=b))
    sc["b4_Z_wall_ld0"] = ({}, scenario({}, ["Z", "T"], [1, 6, 6], board=b))
    b = np.zeros((10, 20), np.uint8); b[:9, 1] = 1
    sc["b4_S_wall_ld50"] = (fl, scenario(fl, ["S", "T"], [0] * 8 + [1] * 3 + [5, 4], board=b))
    fl = dict(width=7, height=9)
    sc["b4_w7h9"] = (fl, scenario(fl, ["T", "L", "J", "S"], [6, 0, 0, 0, 0, 4, 4, 2, 1, 1, 1, 1, 5, 2, 2, 2]))
    # multi-line clears with floating remainder and every clear-count on a 4-wide board
    for k in range(5):
        b = np.zeros((4, 20), np.uint8)
        for y in range(20 - k, 20):
            b[:, y] = 1; b[2, y] = 0
        b[0, 20 - k - 1] = 1; b[3, 12] = 1
        for name in ("default", "advanced_clears", "high_scoring", "all7", "C3"):
            fl = dict(FLAGSETS[name], width=4, height=20)
            acts = [2] if fl.get("lock_delay", 0) == 0 else [2, 6, 6, 6]
            sc[f"w4_{name}_k{k}"] = (fl, scenario(fl, ["I", "O", "T"], acts + [2, 6, 6, 2], board=b))
    # every piece x every action prefix on a cluttered board: rotation table + wall/floor tests
    rs = np.random.RandomState(7)
    for pid, (w, h) in itertools.product(range(7), [(10, 20), (5, 8)]):
        b = (rs.rand(w, h) < 0.35).astype(np.uint8)
        b[:, : h // 2] = 0
        b[w // 2, :] = 0
        fl = dict(width=w, height=h, lock_delay=2)
        acts = rs.randint(0, 7, 60).tolist()
        sc[f"clutter_{SHAPE_NAMES[pid]}_{w}x{h}"] = (fl, scenario(fl, [SHAPE_NAMES[pid]] * 30, acts, board=b))
    return sc


RENDER_CASES = {
    "default_10x20": dict(), "wide_20x40": dict(width=20, height=40), "odd_7x9": dict(width=7, height=9),
    "narrow_4x20_ld2": dict(width=4, height=20, lock_delay=2), "sq_16x16": dict(width=16, height=16),
}


def render_rollout(kwargs, seed, T):
    """TetrisEnv.render('rgb_array') (tetris_env.py:458-462) after reset and after every step."""
    random.seed(seed)
    env = make_reference_env(**kwargs)
    pieces = record_pieces(env)
    actions = actions_for(seed + 50, T)
    env.reset()
    first = env.render(mode="rgb_array")
    assert first.dtype == np.uint8 and first.shape == (160, 160, 3)
    dig, don = [np.frombuffer(hashlib.sha256(first.tobytes()).digest()[:8], dtype=np.uint64)[0]], []
    for a in actions:
        _, _, d, _ = env.step(int(a))
        don.append(bool(d))
        if d:
            env.reset()
        img = env.render(mode="rgb_array")
        dig.append(np.frombuffer(hashlib.sha256(np.ascontiguousarray(img).tobytes()).digest()[:8], dtype=np.uint64)[0])
    return dict(actions=actions, pieces=np.asarray(pieces, np.uint8), done=np.asarray(don, np.uint8),
                digest=np.asarray(dig, np.uint64), first=first)


def make_render():
    out, meta = {}, {}
    for name, kw in RENDER_CASES.items():
        r = render_rollout(kw, 3, 400)
        meta[name] = dict(kwargs=kw)
        for k, v in r.items():
            out[f"{name}/{k}"] = v
    out["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "render.npz"), **out)
    print("render cases:", len(meta))


def make_debug():
    """debug.npz: `repr(engine)` (tetris_env.py:329-335) and the frame of `render('human')` (tetris_env.py:437-457,
    captured through a recording pygame stand-in) after reset and after every step of short rollouts."""
    from _fake_pygame import make_fake_pygame

    ref = load_reference()
    out, meta = {}, {}
    for name, kw in RENDER_CASES.items():
        random.seed(11)
        env = make_reference_env(**kw)
        pieces = record_pieces(env)
        actions = actions_for(61, 120)
        fake = make_fake_pygame()
        ref.pygame = fake
        env.reset()
        reprs, don = [repr(env.engine)], []
        env.render(mode="human")
        for a in actions:
            _, _, d, _ = env.step(int(a))
            don.append(bool(d))
            if d:
                env.reset()
            reprs.append(repr(env.engine))
            env.render(mode="human")
        assert len(fake.frames) == len(actions) + 1 and fake.frames[0].shape == (512, 512, 3)
        hd = [np.frombuffer(hashlib.sha256(np.ascontiguousarray(f, dtype=np.uint8).tobytes()).digest()[:8],
                            dtype=np.uint64)[0] for f in fake.frames]
        meta[name] = dict(kwargs=kw)
        out[f"{name}/actions"] = actions
        out[f"{name}/pieces"] = np.asarray(pieces, np.uint8)
        out[f"{name}/done"] = np.asarray(don, np.uint8)
        out[f"{name}/repr"] = np.frombuffer("\x00".join(reprs).encode(), dtype=np.uint8)
        out[f"{name}/human_digest"] = np.asarray(hd, np.uint64)
        out[f"{name}/human_first"] = np.ascontiguousarray(fake.frames[0], dtype=np.uint8)
        out[f"{name}/calls"] = np.frombuffer(json.dumps([c if isinstance(c, str) else list(c) for c in fake.calls[:12]]).encode(), dtype=np.uint8)
    out["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "debug.npz"), **out)
    print("debug cases:", len(meta))


def main():
    ref = load_reference()
    assert ref.shape_names == SHAPE_NAMES
    if "--render-only" in sys.argv:
        return make_render()
    if "--debug-only" in sys.argv:
        return make_debug()
    out = {}
    meta = {}
    for name, kw in CASES.items():
        image = kw.get("obs_type", "ram") != "ram"
        T = 1500 if image else 4000
        for seed in (0, 1):
            key = f"{name}__s{seed}"
            r = rollout(kw, seed, T)
            meta[key] = dict(kwargs=kw, seed=seed, T=T)
            for k, v in r.items():
                out[f"{key}/{k}"] = v
            print(key, "episodes", int(r["done"].sum()), "lines", int(r["info"][:, 3].max()),
                  "sum_reward", float(r["reward"].sum()))
    out["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "rollouts.npz"), **out)

    out, meta = {}, {}
    for key, (fl, r) in make_scenarios().items():
        meta[key] = dict(kwargs=fl)
        for k, v in r.items():
            out[f"{key}/{k}"] = v
    out["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "scenarios.npz"), **out)
    print("scenarios:", len(meta))
    make_render()
    make_debug()


if __name__ == "__main__":
    main()
