"""gpyreg_b200 -- B200-native (sm_100a) implementation of GPyReg's GP hot path."""
from . import _lib  # noqa: F401
from .engine import Engine, GpbError, PosteriorBatch, get_engine  # noqa: F401
