"""envpool_b200 -- B200-native batched env-step engine behind EnvPool's Python API.

    import envpool_b200 as envpool
    env = envpool.make("CartPole-v1", env_type="gymnasium", num_envs=65536)
    obs, info = env.reset()
    obs, rew, term, trunc, info = env.step(actions)          # numpy in / numpy out
    out = env.step_device(actions_cuda)                      # torch CUDA in / zero-copy out

Env state lives on the GPU as structure-of-arrays and is advanced by hand-written sm_100a
kernels (envpool_b200/csrc); this package is the host-side mirror of the reference's
envpool/python + envpool/registration.py.  There is no CPU fallback: if the compiled
extension modules are missing, the first make()/make_spec()/list_all_envs() call raises
(build them with `python -m envpool_b200._build`).
"""
from typing import Any, List

from . import registration as _registration

__version__ = "0.1.0"


def _ensure_registered() -> None:
    """Import the family packages (and their compiled pybind modules) exactly once; kept
    out of import time so that `python -m envpool_b200._build` can run on a fresh tree."""
    from . import entry  # noqa: F401


def register(*args: Any, **kwargs: Any) -> None:
    _registration.register(*args, **kwargs)


def make(task_id: str, env_type: str, **kwargs: Any):
    _ensure_registered()
    return _registration.make(task_id, env_type, **kwargs)


def make_dm(task_id: str, **kwargs: Any):
    _ensure_registered()
    return _registration.make_dm(task_id, **kwargs)


def make_gym(task_id: str, **kwargs: Any):
    _ensure_registered()
    return _registration.make_gym(task_id, **kwargs)


def make_gymnasium(task_id: str, **kwargs: Any):
    _ensure_registered()
    return _registration.make_gymnasium(task_id, **kwargs)


def make_spec(task_id: str, **kwargs: Any):
    _ensure_registered()
    return _registration.make_spec(task_id, **kwargs)


def list_all_envs() -> List[str]:
    _ensure_registered()
    return _registration.list_all_envs()


__all__ = ["register", "make", "make_dm", "make_gym", "make_gymnasium", "make_spec",
           "list_all_envs"]
