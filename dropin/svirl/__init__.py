"""Drop-in name: with this directory on PYTHONPATH, ``import svirl`` gives svirl_b200.

The package proper is called svirl_b200 so that it can live beside the reference (the oracle and the B-ref baseline
import the real ``svirl``).  A user who switches over puts ``<repo>/dropin`` in front of PYTHONPATH (or installs this
shim) and keeps every ``from svirl import GLSolver`` / ``from svirl.storage import GArray`` unchanged; the reference's
own acceptance scripts run that way, unmodified (tests/test_gpu_dropin.py).  Every ``svirl.*`` module name is bound to
the SAME module object as ``svirl_b200.*``: no second copy of any class or of the process-global config."""
import importlib
import sys

import svirl_b200 as _impl

for _sub in ("config", "storage", "parallel", "solvers", "observables", "vars", "mesh", "scale"):
    importlib.import_module("svirl_b200." + _sub)
for _name, _mod in list(sys.modules.items()):
    if _name == "svirl_b200" or _name.startswith("svirl_b200."):
        sys.modules["svirl" + _name[len("svirl_b200"):]] = _mod
