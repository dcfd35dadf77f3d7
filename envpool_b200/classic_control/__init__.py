"""classic_control family: binds the engine's pybind11 classes (`_XxxEnvSpec` / `_XxxEnvPool`, csrc/py_module.cc)
to the Python adapters and exports, per env, `XxxEnvSpec`, `XxxDMEnvPool` and
`XxxGymnasiumEnvPool` -- the names envpool/classic_control/__init__.py exports, so that
`registration.py` import paths stay interchangeable with the reference's."""
from ..python.api import py_env
from . import classic_control_envpool as _ext

ENVS = ("CartPole", "Pendulum", "MountainCar", "MountainCarContinuous", "Acrobot")

__all__ = []
for _env in ENVS:
    _classes = py_env(getattr(_ext, f"_{_env}EnvSpec"), getattr(_ext, f"_{_env}EnvPool"))
    for _suffix, _cls in zip(("EnvSpec", "DMEnvPool", "GymnasiumEnvPool"), _classes):
        globals()[_env + _suffix] = _cls
        __all__.append(_env + _suffix)
del _env, _classes, _suffix, _cls
