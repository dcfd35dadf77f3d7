"""toy_text family: binds the engine's pybind11 classes (`_XxxEnvSpec` / `_XxxEnvPool`, csrc/py_module.cc)
to the Python adapters and exports, per env, `XxxEnvSpec`, `XxxDMEnvPool` and
`XxxGymnasiumEnvPool` -- the names envpool/toy_text/__init__.py exports, so that
`registration.py` import paths stay interchangeable with the reference's."""
from ..python.api import py_env
from . import toy_text_envpool as _ext

ENVS = ("Catch", "FrozenLake", "Taxi", "NChain", "CliffWalking", "Blackjack")

__all__ = []
for _env in ENVS:
    _classes = py_env(getattr(_ext, f"_{_env}EnvSpec"), getattr(_ext, f"_{_env}EnvPool"))
    for _suffix, _cls in zip(("EnvSpec", "DMEnvPool", "GymnasiumEnvPool"), _classes):
        globals()[_env + _suffix] = _cls
        __all__.append(_env + _suffix)
del _env, _classes, _suffix, _cls
