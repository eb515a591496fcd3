"""Property tests of the oracle (hypothesis): its np.repeat/np.insert image construction against the closed form of
SURVEY.md A.5 written independently in NumPy, and its lock bookkeeping (holes, height, line clears) against NumPy
restatements of tetris_env.py:205-220, 287 on random boards.  These also run on the GPU box, where the reference
itself is absent."""
import numpy as np
from hypothesis import given, settings, strategies as st_

from oracle.oracle import OracleEnv, convert_grayscale


def closed_form(board, size):
    """pixel (row, col) of the size x size image of a (W, H) board."""
    W, H = board.shape
    gap = size // 100 + 1
    bs = (size - 2 * gap) // max(W, H) - gap
    pitch = bs + gap
    inner_v, inner_h = gap + pitch * H, gap + pitch * W
    top, left = (size - inner_v) // 2, (size - inner_h) // 2
    img = np.zeros((size, size), np.uint8)
    rr = np.arange(size)[:, None] - top
    cc = np.arange(size)[None, :] - left
    inside = (rr >= 0) & (rr < inner_v) & (cc >= 0) & (cc < inner_h)
    grid = (rr % pitch < gap) | (cc % pitch < gap)
    y = np.clip(rr // pitch, 0, H - 1)
    x = np.clip(cc // pitch, 0, W - 1)
    filled = board[x, y] != 0
    img[inside] = 128
    img[inside & ~grid & filled] = 190
    return img


@settings(max_examples=120, deadline=None)
@given(st_.integers(1, 32), st_.integers(1, 63), st_.sampled_from([84, 160, 96, 257]), st_.integers(0, 2 ** 31 - 1))
def test_convert_grayscale_equals_closed_form(W, H, size, seed):
    gap = size // 100 + 1
    if (size - 2 * gap) // max(W, H) - gap < 0:
        return  # the reference raises (negative np.repeat count)
    board = (np.random.RandomState(seed).rand(W, H) < 0.4).astype(np.float64)
    assert np.array_equal(convert_grayscale(board, size), closed_form(board, size))


@settings(max_examples=150, deadline=None)
@given(st_.integers(2, 16), st_.integers(4, 40), st_.floats(0.05, 0.95), st_.integers(0, 2 ** 31 - 1))
def test_lock_bookkeeping_on_random_boards(W, H, density, seed):
    """Drop a vertical I onto a random board (row 0..3 kept empty) and compare the step's info with NumPy."""
    rs = np.random.RandomState(seed)
    board = (rs.rand(W, H) < density).astype(np.float64)
    board[:, :4] = 0
    full = rs.rand(H) < 0.2          # make some rows completely full except the drop column
    col = W // 2
    board[:, full] = 1
    board[:, :4] = 0
    board[col, :] = 0                # free drop column: the I lands on the floor
    env = OracleEnv(width=W, height=H, penalise_height=True, penalise_holes=True, pieces=["I", "O", "O"])
    env.reset()
    env.board = board
    obs, reward, done, info = env.step(2)
    # NumPy restatement: lock I into rows H-4..H-1 of the column, clear full rows, count holes / non-empty rows
    b = board.copy()
    b[col, H - 4:] = 1
    can_clear = np.all(b != 0, axis=0)
    k = int(can_clear.sum())
    nb = np.zeros_like(b)
    keep = [i for i in range(H) if not can_clear[i]]
    if keep:
        nb[:, H - len(keep):] = b[:, keep]
    holes = int(np.count_nonzero(nb.cumsum(axis=1) * (nb == 0)))
    height = int(np.any(nb != 0, axis=0).sum())
    assert info["lines_cleared"] == k and info["holes"] == holes
    if np.any(nb[:, 0] != 0):
        assert done and reward == -100
    else:
        assert not done and reward == 100 * k - height - 5 * holes
        assert np.array_equal(env.board, nb)
