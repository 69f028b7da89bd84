"""GPU: the reference's own acceptance scripts (/root/reference/tests/at_*.py, tests/run_at.sh:10-17) run UNMODIFIED
against this library under the drop-in name ``svirl`` (dropin/svirl).  The scripts are staged verbatim, outside the
repository history, by baseline/install_ref.py (baseline/_ref/ref_tests); each prints one '... passed (n/n)' or
'>FAILED<' line per check."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT, has_cuda

SCRIPTS = os.path.join(ROOT, "baseline", "_ref", "ref_tests")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device"),
              pytest.mark.skipif(not os.path.isdir(SCRIPTS), reason="reference acceptance scripts not staged")]


@pytest.mark.parametrize("script", ["at_reduction.py", "at_destructor.py", "at_cg_coef_psi.py", "at_cg_coef.py",
                                    "at_cg_jacobians.py"])
def test_reference_acceptance_script_unmodified(script):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT]))
    out = subprocess.run([sys.executable, script], cwd=SCRIPTS, env=env, capture_output=True, text=True, timeout=1500)
    text = out.stdout + out.stderr
    assert out.returncode == 0, text[-3000:]
    assert " passed " in out.stdout and ">FAILED<" not in out.stdout and "FAILED" not in out.stdout, text[-3000:]


def test_precision_script_needs_matplotlib():
    """at_precision.py imports matplotlib and svirl.plotter at the top; the image has no matplotlib (neither for the
    reference), so it cannot run here.  Its numerical content (fp32 vs fp64 energy within 10 % after a long run) is
    restated in tests/test_gpu_properties.py."""
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        pytest.skip("matplotlib is not installed in this image")
