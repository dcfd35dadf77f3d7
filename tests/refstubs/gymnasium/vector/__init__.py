from .vector_env import AutoresetMode, VectorEnv  # noqa: F401
