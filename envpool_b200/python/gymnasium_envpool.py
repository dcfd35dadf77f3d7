"""gymnasium-style adapter (the role of envpool/python/gymnasium_envpool.py): `reset()` gives
`(obs, info)`, `step()` / `recv()` give `(obs, reward, terminated, truncated, info)`."""
from __future__ import annotations

from abc import ABCMeta

from .adapter import build_adapter
from .data import fill_tree, gym_structure


class GymnasiumEnvPoolMixin:
    """Space accessors in gymnasium's property form, computed once per pool."""

    @property
    def observation_space(self):
        try:
            return self._gym_observation_space
        except AttributeError:
            self._gym_observation_space = self.spec.observation_space
            return self._gym_observation_space

    @property
    def action_space(self):
        try:
            return self._gym_action_space
        except AttributeError:
            self._gym_action_space = self.spec.action_space
            return self._gym_action_space

    # vector-env vocabulary: every sub-env has the same spaces
    single_observation_space = observation_space
    single_action_space = action_space

    @property
    def num_envs(self):
        return self.config["num_envs"]


def _five_tuple_fold(state_keys):
    tree = gym_structure(state_keys)

    def fold(state_values, reset):
        state = fill_tree(tree, state_values)
        info = state["info"]
        info["elapsed_step"] = state["elapsed_step"]
        obs = state["obs"]
        while isinstance(obs, dict) and len(obs) == 1:  # single "obs" key: unwrap
            obs = next(iter(obs.values()))
        if reset:
            return obs, info
        truncated = state["trunc"]
        # the engine's `done` covers both endings; gymnasium splits them
        # (envpool/python/gymnasium_envpool.py:226)
        terminated = state["done"] & ~truncated
        return obs, state["reward"], terminated, truncated, info

    return fold


class GymnasiumEnvPoolMeta(ABCMeta):
    def __new__(meta, name, parents, attrs):
        return build_adapter(meta, name, parents[0], GymnasiumEnvPoolMixin, attrs,
                             _five_tuple_fold)
