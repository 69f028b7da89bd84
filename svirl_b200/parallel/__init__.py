from .startup import Startup
from .utils import Utils
from .reduction import Reduction
