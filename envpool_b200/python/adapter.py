"""Shared construction of the two env-type adapters (dm_env style, gymnasium style).

Both adapters are the engine's pybind pool class + `EnvPoolMixin` (reset/step/send/recv on
top of `_send`/`_recv`/`_reset`) + a small mixin with the spec accessors; they differ only in
how the flat list of state columns that `_recv()` returns is folded into what the user gets
back.  `build_adapter` does the common part once; the metaclasses in dm_envpool.py /
gymnasium_envpool.py supply the fold.  (The reference builds the same two classes in
envpool/python/dm_envpool.py:43-103 and envpool/python/gymnasium_envpool.py:160-239.)
"""
from __future__ import annotations

from abc import ABCMeta
from typing import Callable, List

from .env_spec import check_key_duplication
from .envpool import EnvPoolMixin

XLA_MESSAGE = ("XLA is unavailable in envpool_b200: env state and outputs already live on the "
               "GPU -- use step_device() / reset_device() for device-resident consumers.")


def build_adapter(meta, name: str, pool_cls: type, accessor_mixin: type, attrs: dict,
                  make_fold: Callable[[List[str]], Callable]) -> type:
    """Create class `name` = (pool_cls, accessor_mixin, EnvPoolMixin) with

    * `_to(state_values, reset, return_info)`: the fold built by `make_fold(state_keys)`,
    * `xla()`: raises (outside the accelerated path),
    * `__init__(spec, **engine_kwargs)`: constructs the pybind pool and keeps the spec.
    """
    check_key_duplication(name, "state", pool_cls._state_keys)
    check_key_duplication(name, "action", pool_cls._action_keys)
    fold = make_fold(list(pool_cls._state_keys))

    def _to(self, state_values, reset, return_info):
        return fold(state_values, reset)

    def xla(self):
        raise RuntimeError(XLA_MESSAGE)

    attrs = dict(attrs, _to=_to, xla=xla)
    cls = ABCMeta.__new__(meta, name, (pool_cls, accessor_mixin, EnvPoolMixin), attrs)

    def __init__(self, spec, **engine_kwargs):
        super(cls, self).__init__(spec, **engine_kwargs)
        self.spec = spec

    cls.__init__ = __init__
    return cls
