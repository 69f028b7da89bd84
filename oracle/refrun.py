"""Import the UNMODIFIED reference package on the CPU -- TEST INFRASTRUCTURE ONLY.

``import_reference()`` puts oracle/fake_pycuda (CPU stand-in for pyCUDA) and
/root/reference on sys.path and returns the reference's ``svirl`` module.  Works only
in the build container (where /root/reference exists); used by oracle/make_golden.py and
by the oracle-vs-reference tests, which skip when the reference is absent.
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "svirl"))


def import_reference():
    if not available():
        raise RuntimeError("reference not present")
    for p in (HERE, os.path.join(HERE, "fake_pycuda"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    import svirl  # the reference, not this repo's package
    assert os.path.abspath(svirl.__file__).startswith(REF), svirl.__file__
    return svirl


def launch_counts():
    import pycuda
    return pycuda.LAUNCH_COUNTS
