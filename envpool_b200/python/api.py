"""py_env(): mirror of envpool/python/api.py:22-41."""
from typing import Tuple, Type

from .dm_envpool import DMEnvPoolMeta
from .env_spec import EnvSpecMeta
from .gymnasium_envpool import GymnasiumEnvPoolMeta


def py_env(envspec: Type, envpool: Type) -> Tuple[Type, Type, Type]:
    """Wrap a pybind (_XxxEnvSpec, _XxxEnvPool) pair into (Spec, DMEnvPool,
    GymnasiumEnvPool) classes."""
    spec_name = envspec.__name__[1:]
    pool_name = envpool.__name__[1:]
    return (
        EnvSpecMeta(spec_name, (envspec,), {}),
        DMEnvPoolMeta(pool_name.replace("EnvPool", "DMEnvPool"), (envpool,), {}),
        GymnasiumEnvPoolMeta(pool_name.replace("EnvPool", "GymnasiumEnvPool"), (envpool,), {}),
    )
