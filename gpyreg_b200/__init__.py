"""gpyreg_b200 -- B200-native (sm_100a) implementation of GPyReg's GP hot path, behind
GPyReg's own API: ``GP`` (update / fit / predict / nlZ), and the covariance / mean / noise
``compute(...)`` plugin surface.  NumPy float64 in and out; the numerics run in hand-written
CUDA through the C ABI of include/gpyreg_b200.h.  There is no CPU fallback."""
from . import _lib  # noqa: F401
from . import (covariance_functions, isotropic_covariance_functions, mean_functions,  # noqa: F401
               noise_functions)
from .engine import Engine, GpbError, PosteriorBatch, get_engine  # noqa: F401
from .gaussian_process import GP, Posterior  # noqa: F401
from .slice_sample import SliceSampler  # noqa: F401
from .spec import ModelSpec  # noqa: F401
