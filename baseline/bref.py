"""B-ref: the UNMODIFIED reference (Python package + CUDA kernels) on the B200 -- BASELINE INFRASTRUCTURE ONLY.

SURVEY.md section 2.2 / 8d(i): "the bar is the reference's own kernels JIT-compiled for sm_100 on the same B200".
The reference package is installed, untouched, in baseline/_ref (baseline/install_ref.py); pyCUDA, which this image
lacks, is replaced by baseline/gpu_pycuda (the dozen calls the reference makes, on the CUDA driver API).  What runs
is therefore the reference's own solver code: one thread per node, block 128, per Jacobi sweep a fill + a launch + a
blocking 4-byte read-back (svirl/solvers/td.py:164-202, 274-311), the as-written CG iteration
(svirl/solvers/cg.py:477-544) and its SciPy / NumPy line search.

Used by bench.py (`gpu_baseline` in the JSON line), tests/test_gpu_bref.py (three-way parity: reference fixtures <->
reference on the GPU <-> this library) and tools/cfg4_adjudicate.py.  Nothing under svirl_b200/ imports this.
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "svirl"))


def import_reference():
    """-> the reference's `svirl` module, running on the GPU through baseline/gpu_pycuda."""
    if not available():
        raise RuntimeError("baseline/_ref is empty: run baseline/install_ref.py in the build container")
    for p in (os.path.join(HERE, "gpu_pycuda"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    if not hasattr(np, "bool"):          # removed in numpy 1.24..1.26, back (as an alias) in 2.x
        np.bool = np.bool_
    import pycuda
    assert os.path.abspath(pycuda.__file__).startswith(HERE), "another pycuda is shadowing baseline/gpu_pycuda"
    import svirl
    assert os.path.abspath(svirl.__file__).startswith(REF), svirl.__file__
    return svirl


def launch_counts():
    import pycuda
    return pycuda.LAUNCH_COUNTS


def synchronize():
    import pycuda.driver as cuda
    cuda.synchronize()


def make_solver(**kw):
    return import_reference().GLSolver(**kw)


def time_td(gl, nsteps, warmup=0, dt=0.1):
    """Wall time of `nsteps` reference TDGL steps (every sweep ends in a blocking read-back, so wall time is
    device time + the reference's own launch/sync overhead, which is part of what it costs).
    -> (seconds, psi sweeps, A sweeps)"""
    if warmup:
        gl.solve.td(dt=dt, Nt=warmup)
    synchronize()
    c = launch_counts()
    p0, a0 = c.get("iterate_order_parameter_jacobi_step", 0), c.get("iterate_vector_potential_jacobi_step", 0)
    t0 = time.perf_counter()
    gl.solve.td(dt=dt, Nt=nsteps)
    synchronize()
    el = time.perf_counter() - t0
    return el, c.get("iterate_order_parameter_jacobi_step", 0) - p0, c.get("iterate_vector_potential_jacobi_step", 0) - a0


def time_cg(gl, niter, warmup=0):
    """Wall time of `niter` reference CG iterations (convergence test disabled). -> (seconds, energies)"""
    gl.solve._init_cg()
    cg = gl.solve._cg
    cg._CG__convergence_rtol = -1.0
    if warmup:
        gl.solve.cg(n_iter=warmup)
    synchronize()
    t0 = time.perf_counter()
    gl.solve.cg(n_iter=niter)
    synchronize()
    el = time.perf_counter() - t0
    return el, [float(e) for e in cg.cg_energies]


def fields(gl):
    """(psi [Nx,Ny], a [Nx-1,Ny], b [Nx,Ny-1]) host copies of the reference solver's state."""
    psi = np.array(gl.vars.order_parameter)
    a, b = gl.vars.vector_potential
    return psi, np.array(a), np.array(b)
