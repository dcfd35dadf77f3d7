"""dm_env-style adapter: mirror of envpool/python/dm_envpool.py."""
from __future__ import annotations

from abc import ABCMeta
from typing import Any, List

import numpy as np

from .data import dm_structure, fill_tree, make_timestep, to_namedtuple
from .env_spec import check_key_duplication
from .envpool import EnvPoolMixin


class DMEnvPoolMixin:
    def observation_spec(self):
        if not hasattr(self, "_dm_observation_spec"):
            self._dm_observation_spec = self.spec.observation_spec()
        return self._dm_observation_spec

    def action_spec(self):
        if not hasattr(self, "_dm_action_spec"):
            self._dm_action_spec = self.spec.action_spec()
        return self._dm_action_spec


class DMEnvPoolMeta(ABCMeta):
    """envpool/python/dm_envpool.py:43-103"""

    def __new__(cls, name, parents, attrs):
        base = parents[0]

        def _xla(self):
            raise RuntimeError("XLA is unavailable in envpool_b200; use step_device().")

        attrs["xla"] = _xla
        parents = (base, DMEnvPoolMixin, EnvPoolMixin)
        check_key_duplication(name, "state", base._state_keys)
        check_key_duplication(name, "action", base._action_keys)
        tree = dm_structure("State", base._state_keys)

        def _to_dm(self, state_values: List[np.ndarray], reset: bool, return_info: bool):
            state = to_namedtuple("State", fill_tree(tree, state_values))
            return make_timestep(step_type=state.step_type, reward=state.reward,
                                 discount=state.discount, observation=state.State)

        attrs["_to"] = _to_dm
        subcls = super().__new__(cls, name, parents, attrs)

        def init(self, spec, **engine_kwargs):
            super(subcls, self).__init__(spec, **engine_kwargs)
            self.spec = spec

        setattr(subcls, "__init__", init)
        return subcls
