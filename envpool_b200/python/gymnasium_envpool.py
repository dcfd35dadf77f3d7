"""gymnasium-style adapter: mirror of envpool/python/gymnasium_envpool.py."""
from __future__ import annotations

from abc import ABCMeta
from typing import Any, List

import numpy as np

from .data import fill_tree, gym_structure
from .env_spec import check_key_duplication
from .envpool import EnvPoolMixin


class GymnasiumEnvPoolMixin:
    @property
    def observation_space(self):
        if not hasattr(self, "_gym_observation_space"):
            self._gym_observation_space = self.spec.observation_space
        return self._gym_observation_space

    @property
    def action_space(self):
        if not hasattr(self, "_gym_action_space"):
            self._gym_action_space = self.spec.action_space
        return self._gym_action_space

    @property
    def single_observation_space(self):
        return self.observation_space

    @property
    def single_action_space(self):
        return self.action_space

    @property
    def num_envs(self):
        return self.config["num_envs"]


class GymnasiumEnvPoolMeta(ABCMeta):
    """envpool/python/gymnasium_envpool.py:160-239"""

    def __new__(cls, name, parents, attrs):
        base = parents[0]

        def _xla(self):
            raise RuntimeError("XLA is unavailable in envpool_b200; use step_device().")

        attrs["xla"] = _xla
        parents = (base, GymnasiumEnvPoolMixin, EnvPoolMixin)
        check_key_duplication(name, "state", base._state_keys)
        check_key_duplication(name, "action", base._action_keys)
        tree = gym_structure(base._state_keys)

        def _to_gymnasium(self, state_values: List[np.ndarray], reset: bool,
                          return_info: bool):
            state = fill_tree(tree, state_values)
            info = state["info"]
            info["elapsed_step"] = state["elapsed_step"]
            obs = state["obs"]
            while isinstance(obs, dict) and len(obs) == 1:
                obs = next(iter(obs.values()))
            if reset:
                return obs, info
            done, trunc = state["done"], state["trunc"]
            terminated = done & ~trunc  # gymnasium_envpool.py:226
            return obs, state["reward"], terminated, trunc, info

        attrs["_to"] = _to_gymnasium
        subcls = super().__new__(cls, name, parents, attrs)

        def init(self, spec, **engine_kwargs):
            super(subcls, self).__init__(spec, **engine_kwargs)
            self.spec = spec

        setattr(subcls, "__init__", init)
        return subcls
