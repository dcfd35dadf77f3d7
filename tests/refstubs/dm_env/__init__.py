"""Stand-in for the dm_env names the reference's Python layer touches."""
import enum
from typing import Any, NamedTuple

from . import specs  # noqa: F401


class StepType(enum.IntEnum):
    FIRST = 0
    MID = 1
    LAST = 2


class TimeStep(NamedTuple):
    step_type: Any
    reward: Any
    discount: Any
    observation: Any

    def first(self):
        return self.step_type == StepType.FIRST

    def mid(self):
        return self.step_type == StepType.MID

    def last(self):
        return self.step_type == StepType.LAST


class Environment:
    """dm_env.Environment: abstract base (reset / step / observation_spec / action_spec)."""

    def close(self):
        pass
