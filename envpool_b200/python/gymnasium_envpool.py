"""gymnasium-style adapter (the role of envpool/python/gymnasium_envpool.py): `reset()` gives
`(obs, info)`, `step()` / `recv()` give `(obs, reward, terminated, truncated, info)`."""
from __future__ import annotations

import warnings
from abc import ABCMeta

import numpy as np

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

    # what gymnasium wrappers / SB3 probe on a vector env
    # (envpool/python/gymnasium_envpool.py:79-97)
    metadata = {"render_modes": ["rgb_array", "human"], "autoreset_mode": "next_step"}
    is_vector_env = True
    render_mode = None

    def reset(self, env_id=None, *, seed=None, options=None):
        """gymnasium's keyword form (envpool/python/gymnasium_envpool.py:129-153): `seed`
        cannot re-seed a pool (seeds are fixed by make()) and is ignored with a warning;
        options may only carry `reset_mask`, a bool[num_envs] choosing the envs to reset."""
        if seed is not None:
            warnings.warn("envpool_b200 fixes every env's seed at make(); reset(seed=...) is "
                          "ignored -- pass seed to make() instead.", stacklevel=2)
        if options:
            extra = sorted(set(options) - {"reset_mask"})
            if extra:
                raise ValueError(f"unsupported reset options: {extra}")
            mask = options.get("reset_mask")
            if mask is not None:
                if env_id is not None:
                    raise ValueError("give either env_id or options['reset_mask'], not both")
                mask = np.asarray(mask, dtype=np.bool_)
                n = self.config["num_envs"]
                if mask.shape != (n,):
                    raise ValueError(f"reset_mask must have shape ({n},), got {mask.shape}")
                if not mask.any():
                    raise ValueError("reset_mask selects no environment")
                env_id = np.flatnonzero(mask).astype(np.int32)
        return super().reset(env_id)

    def close(self, **kwargs):
        return super().close()


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
