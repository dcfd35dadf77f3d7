"""dm_env-style adapter (the role of envpool/python/dm_envpool.py): `reset()` / `step()` /
`recv()` hand back a `TimeStep(step_type, reward, discount, observation)` whose observation
is a namedtuple tree of the obs/info columns."""
from __future__ import annotations

from abc import ABCMeta

from .adapter import build_adapter
from .data import dm_structure, fill_tree, make_timestep, to_namedtuple


class DMEnvPoolMixin:
    """Spec accessors in dm_env's method form, computed once per pool."""

    def observation_spec(self):
        try:
            return self._dm_observation_spec
        except AttributeError:
            self._dm_observation_spec = self.spec.observation_spec()
            return self._dm_observation_spec

    def action_spec(self):
        try:
            return self._dm_action_spec
        except AttributeError:
            self._dm_action_spec = self.spec.action_spec()
            return self._dm_action_spec


def _timestep_fold(state_keys):
    tree = dm_structure("State", state_keys)

    def fold(state_values, reset):
        state = to_namedtuple("State", fill_tree(tree, state_values))
        return make_timestep(step_type=state.step_type, reward=state.reward,
                             discount=state.discount, observation=state.State)

    return fold


class DMEnvPoolMeta(ABCMeta):
    def __new__(meta, name, parents, attrs):
        return build_adapter(meta, name, parents[0], DMEnvPoolMixin, attrs, _timestep_fold)
