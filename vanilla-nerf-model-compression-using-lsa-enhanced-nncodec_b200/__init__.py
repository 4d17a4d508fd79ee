"""B200-native NeRF ray-rendering hot path behind the reference's call surface."""
from . import _lib  # noqa: F401
