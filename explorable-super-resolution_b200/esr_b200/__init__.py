"""esr_b200 — native runtime of the B200 hot path (ctypes binding + tensor-level ops)."""
from . import lib  # noqa: F401
