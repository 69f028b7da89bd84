from .observables import Observables
from .vortex_detector import VortexDetector
