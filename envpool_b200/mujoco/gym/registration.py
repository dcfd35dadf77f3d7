"""MuJoCo gym registration: HalfCheetah v3/v4/v5 with the kwargs of
envpool/mujoco/gym/registration.py:22,34-92."""
from ...registration import register

for _version in ("v3", "v4", "v5"):
    _extra = {"gymnasium_v5_render_camera": True} if _version == "v5" else {}
    register(task_id=f"HalfCheetah-{_version}", import_path="envpool_b200.mujoco.gym",
             spec_cls="GymHalfCheetahEnvSpec", dm_cls="GymHalfCheetahDMEnvPool",
             gymnasium_cls="GymHalfCheetahGymnasiumEnvPool",
             post_constraint=(_version == "v5"), max_episode_steps=1000, **_extra)
