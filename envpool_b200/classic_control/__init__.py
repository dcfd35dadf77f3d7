"""Classic control env in envpool_b200 (mirror of envpool/classic_control/__init__.py)."""
from ..python.api import py_env
from .classic_control_envpool import (_AcrobotEnvPool, _AcrobotEnvSpec, _CartPoleEnvPool,
                                      _CartPoleEnvSpec, _MountainCarContinuousEnvPool,
                                      _MountainCarContinuousEnvSpec, _MountainCarEnvPool,
                                      _MountainCarEnvSpec, _PendulumEnvPool,
                                      _PendulumEnvSpec)

CartPoleEnvSpec, CartPoleDMEnvPool, CartPoleGymnasiumEnvPool = py_env(
    _CartPoleEnvSpec, _CartPoleEnvPool)
PendulumEnvSpec, PendulumDMEnvPool, PendulumGymnasiumEnvPool = py_env(
    _PendulumEnvSpec, _PendulumEnvPool)
MountainCarEnvSpec, MountainCarDMEnvPool, MountainCarGymnasiumEnvPool = py_env(
    _MountainCarEnvSpec, _MountainCarEnvPool)
(MountainCarContinuousEnvSpec, MountainCarContinuousDMEnvPool,
 MountainCarContinuousGymnasiumEnvPool) = py_env(_MountainCarContinuousEnvSpec,
                                                 _MountainCarContinuousEnvPool)
AcrobotEnvSpec, AcrobotDMEnvPool, AcrobotGymnasiumEnvPool = py_env(
    _AcrobotEnvSpec, _AcrobotEnvPool)

__all__ = [
    "CartPoleEnvSpec", "CartPoleDMEnvPool", "CartPoleGymnasiumEnvPool",
    "PendulumEnvSpec", "PendulumDMEnvPool", "PendulumGymnasiumEnvPool",
    "MountainCarEnvSpec", "MountainCarDMEnvPool", "MountainCarGymnasiumEnvPool",
    "MountainCarContinuousEnvSpec", "MountainCarContinuousDMEnvPool",
    "MountainCarContinuousGymnasiumEnvPool",
    "AcrobotEnvSpec", "AcrobotDMEnvPool", "AcrobotGymnasiumEnvPool",
]
