"""A recording stand-in for pygame (not installed in this image): just the calls TetrisEnv.render('human') makes
(tetris_env.py:437-457).  `frames` collects every array handed to `surfarray.make_surface`."""
import types

import numpy as np


def make_fake_pygame():
    pg = types.ModuleType("pygame")
    pg.frames, pg.calls = [], []

    class _Surface:
        def __init__(self, arr=None):
            self.arr = arr

        def blit(self, canvas, rect):
            pg.calls.append("blit")

        def get_rect(self):
            return (0, 0) + (self.arr.shape[:2] if self.arr is not None else (0, 0))

    class _Clock:
        def tick(self, fps):
            pg.calls.append(("tick", fps))

    pg.init = lambda: pg.calls.append("init")
    pg.display = types.SimpleNamespace(
        init=lambda: pg.calls.append("display.init"),
        set_mode=lambda size: (pg.calls.append(("set_mode", tuple(size))), _Surface())[1],
        update=lambda: pg.calls.append("update"))
    pg.time = types.SimpleNamespace(Clock=_Clock)
    pg.event = types.SimpleNamespace(pump=lambda: pg.calls.append("pump"))

    def array_to_surface(surface, arr):
        assert arr.ndim == 3 and arr.shape[2] == 3
        pg.calls.append("array_to_surface")

    def make_surface(arr):
        pg.frames.append(np.array(arr, copy=True))
        return _Surface(arr)

    pg.pixelcopy = types.SimpleNamespace(array_to_surface=array_to_surface)
    pg.surfarray = types.SimpleNamespace(make_surface=make_surface)
    return pg
