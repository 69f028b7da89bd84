#!/usr/bin/env python
"""Diagnostic: TDGL + CG at a given size, printing energy / max|psi| / sweep counts per step,
to find where large grids go wrong.  python tools/diag_scale.py N [steps] [cg_iters]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

n = int(sys.argv[1])
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cgit = int(sys.argv[3]) if len(sys.argv) > 3 else 3
kappa = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
wl = dict(name="diag", Nx=n, Ny=n, dtype=np.float64, kappa=kappa, sigma=10.0, H=0.1, tiling=False, eps_field=False)
t0 = time.time()
gl = bench.make_solver(wl)
print("solver built in %.1f s" % (time.time() - t0), flush=True)
td = None
print("E0 = %.12g" % gl.observables.free_energy, flush=True)
for s in range(steps):
    gl.solve.td(dt=0.1, Nt=1)
    td = gl.solve._td
    print("step %d  sweeps %d/%d  E = %.12g" % (s, td.sweeps_order_parameter, td.sweeps_vector_potential,
                                              gl.observables.free_energy), flush=True)
if cgit:
    gl.solve.cg(n_iter=cgit)
    print("cg energies", gl.solve._cg.cg_energies, flush=True)
if n <= 8192:
    psi = gl.vars.order_parameter
    print("max|psi| %.6g  finite %s" % (np.abs(psi).max(), np.isfinite(psi).all()))
