from .td import TD
from .cg import CG
from .solvers import Solvers
