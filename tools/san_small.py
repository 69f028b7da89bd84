#!/usr/bin/env python
"""Small kappa = inf runs for compute-sanitizer (racecheck / memcheck) over the fp32 patch kernel, the fp64 column
kernel, the pipelined solves (k_psi_gate + gated launch) and the chained launches:
    compute-sanitizer --tool racecheck python tools/san_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svirl_b200 import GLSolver  # noqa: E402

for dtype, (Nx, Ny) in ((np.float32, (150, 70)), (np.float64, (131, 71))):
    mt = np.ones((Nx - 1, Ny - 1), dtype=bool)
    mt[30:50, 20:30] = False
    gl = GLSolver(Nx=Nx, Ny=Ny, dx=0.5, dy=0.5, dtype=dtype, gl_parameter=np.inf, homogeneous_external_field=0.1,
                  random_seed=3, material_tiling=mt, order_parameter_Langevin_coefficient=0.02)
    gl.par.set_option("graphs", 0)          # the tile kernels, not the single-cluster small-grid kernel
    gl.solve.td(dt=0.1, Nt=4)
    print(dtype.__name__, "sweeps", gl.solve._td.sweeps_order_parameter, "gate hits", gl.par.stat("spec_hit"),
          "finite", bool(np.isfinite(gl.vars.order_parameter).all()), flush=True)
    gl.par.close()
