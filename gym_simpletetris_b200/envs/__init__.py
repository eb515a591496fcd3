from .tetris_env import TetrisEnv, TetrisEnvV26  # noqa: F401  (mirrors gym_simpletetris/envs/__init__.py:1)
