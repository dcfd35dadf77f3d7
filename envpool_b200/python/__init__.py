"""Host-side Python layer: mirror of the reference's envpool/python package
(api.py, envpool.py, env_spec.py, data.py, dm_envpool.py, gymnasium_envpool.py) for the
accelerated path -- same names, same call semantics, no third-party hard dependencies."""
from .api import py_env

__all__ = ["py_env"]
