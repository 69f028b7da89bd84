"""Device context, reductions and small helpers: the namespace the reference exposes as svirl.parallel."""
from . import reduction as _reduction
from . import startup as _startup
from . import utils as _utils

Startup, Reduction, Utils = _startup.Startup, _reduction.Reduction, _utils.Utils
__all__ = ["Startup", "Reduction", "Utils"]
