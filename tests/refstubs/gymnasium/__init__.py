"""Stand-in for the gymnasium names the reference's Python layer touches."""
from . import spaces, vector  # noqa: F401
from .spaces import Space  # noqa: F401


class Env:
    metadata: dict = {}
    render_mode = None

    def close(self):
        pass
