"""Lock-step oracle vs the unmodified reference, full observations compared.
Runs only where /root/reference exists (the build container)."""
import random
import warnings

import numpy as np
import pytest

from _cases import CASES, SHAPE_NAMES, actions_for
from oracle.oracle import OracleEnv, convert_grayscale
from oracle.ref_shim import load_reference, make_reference_env, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference not mounted")


@pytest.mark.parametrize("name", list(CASES))
def test_lockstep(name):
    warnings.simplefilter("ignore")
    kw = CASES[name]
    random.seed(1234)
    ref = make_reference_env(**kw)
    log = []
    orig = ref.engine._choose_shape

    def chooser():
        s = orig()
        log.append(SHAPE_NAMES.index(s))
        return s

    ref.engine._choose_shape = chooser
    T = 400 if kw.get("obs_type", "ram") != "ram" else 1500
    # pre-run the reference to learn its piece sequence, then replay both
    acts = actions_for(99, T)
    trace = []
    obs = ref.reset()
    trace.append(("reset", obs))
    for a in acts:
        o, r, d, info = ref.step(int(a))
        info = dict(info, statistics=dict(info["statistics"]))
        trace.append(("step", o, r, d, info))
        if d:
            trace.append(("reset", ref.reset()))
    env = OracleEnv(pieces=log, **kw)
    it = iter(trace)
    kind, o = next(it)
    assert np.array_equal(env.reset(), o)
    ai = 0
    for rec in it:
        if rec[0] == "reset":
            o2 = env.reset()
            assert o2.dtype == rec[1].dtype and o2.shape == rec[1].shape
            assert np.array_equal(o2, rec[1])
            continue
        _, o, r, d, info = rec
        o2, r2, d2, info2 = env.step(int(acts[ai]))
        ai += 1
        assert o2.shape == o.shape and o2.dtype == o.dtype
        assert np.array_equal(o2, o), (name, ai)
        assert r2 == r and d2 == d and info2 == info, (name, ai)


@pytest.mark.parametrize("wh", [(10, 20), (20, 40), (7, 9), (16, 16), (4, 41), (1, 1), (32, 63), (5, 82), (3, 45)])
@pytest.mark.parametrize("size", [84, 160])
def test_convert_grayscale(wh, size):
    ref = load_reference()
    gap = size // 100 + 1
    if (size - 2 * gap) // max(wh) - gap < 0:
        pytest.skip("negative block size: the reference itself raises here")
    rs = np.random.RandomState(wh[0] * 100 + wh[1])
    b = (rs.rand(*wh) < 0.4).astype(np.float64)
    assert np.array_equal(convert_grayscale(b, size), ref.convert_grayscale(b, size))
