"""Toy text env in envpool_b200 (mirror of envpool/toy_text/__init__.py)."""
from ..python.api import py_env
from .toy_text_envpool import (_BlackjackEnvPool, _BlackjackEnvSpec, _CatchEnvPool,
                               _CatchEnvSpec, _CliffWalkingEnvPool, _CliffWalkingEnvSpec,
                               _FrozenLakeEnvPool, _FrozenLakeEnvSpec, _NChainEnvPool,
                               _NChainEnvSpec, _TaxiEnvPool, _TaxiEnvSpec)

CatchEnvSpec, CatchDMEnvPool, CatchGymnasiumEnvPool = py_env(_CatchEnvSpec, _CatchEnvPool)
FrozenLakeEnvSpec, FrozenLakeDMEnvPool, FrozenLakeGymnasiumEnvPool = py_env(
    _FrozenLakeEnvSpec, _FrozenLakeEnvPool)
TaxiEnvSpec, TaxiDMEnvPool, TaxiGymnasiumEnvPool = py_env(_TaxiEnvSpec, _TaxiEnvPool)
NChainEnvSpec, NChainDMEnvPool, NChainGymnasiumEnvPool = py_env(_NChainEnvSpec, _NChainEnvPool)
CliffWalkingEnvSpec, CliffWalkingDMEnvPool, CliffWalkingGymnasiumEnvPool = py_env(
    _CliffWalkingEnvSpec, _CliffWalkingEnvPool)
BlackjackEnvSpec, BlackjackDMEnvPool, BlackjackGymnasiumEnvPool = py_env(
    _BlackjackEnvSpec, _BlackjackEnvPool)

__all__ = [
    "CatchEnvSpec", "CatchDMEnvPool", "CatchGymnasiumEnvPool",
    "FrozenLakeEnvSpec", "FrozenLakeDMEnvPool", "FrozenLakeGymnasiumEnvPool",
    "TaxiEnvSpec", "TaxiDMEnvPool", "TaxiGymnasiumEnvPool",
    "NChainEnvSpec", "NChainDMEnvPool", "NChainGymnasiumEnvPool",
    "CliffWalkingEnvSpec", "CliffWalkingDMEnvPool", "CliffWalkingGymnasiumEnvPool",
    "BlackjackEnvSpec", "BlackjackDMEnvPool", "BlackjackGymnasiumEnvPool",
]
