"""MuJoCo gym env in envpool_b200: HalfCheetah, the MuJoCo task on the accelerated path
(mirror of envpool/mujoco/gym/__init__.py for that task)."""
from ...python.api import py_env
from ..mujoco_gym_envpool import _GymHalfCheetahEnvPool, _GymHalfCheetahEnvSpec

(GymHalfCheetahEnvSpec, GymHalfCheetahDMEnvPool,
 GymHalfCheetahGymnasiumEnvPool) = py_env(_GymHalfCheetahEnvSpec, _GymHalfCheetahEnvPool)

__all__ = ["GymHalfCheetahEnvSpec", "GymHalfCheetahDMEnvPool",
           "GymHalfCheetahGymnasiumEnvPool"]
