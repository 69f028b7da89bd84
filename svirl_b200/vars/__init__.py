from .fixed_vortices import FixedVortices
from .vars import Vars
from .params import Params
