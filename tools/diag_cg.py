#!/usr/bin/env python
"""Diagnostic: print the 17 line-search coefficients and the BFGS result of the first CG
iterations at a given size (after 20 TDGL steps).  python tools/diag_cg.py N [cg_fused]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

n = int(sys.argv[1])
fused = int(sys.argv[2]) if len(sys.argv) > 2 else 1
wl = dict(name="diag", Nx=n, Ny=n, dtype=np.float64, kappa=2.0, sigma=10.0, H=0.1, tiling=False, eps_field=False)
gl = bench.make_solver(wl)
gl.par.set_option("cg_fused", fused)
gl.solve.td(dt=0.1, Nt=20)
print("E after TD %.10g" % gl.observables.free_energy, flush=True)
gl.solve._init_cg()
cg = gl.solve._cg
orig = cg._cg_alpha_min
np.set_printoptions(precision=6, linewidth=200)


def logged(*a, **k):
    c = cg._CG__c
    r = orig(*a, **k)
    P = np.polynomial.polynomial
    print("c/N =\n", c / (n * n), "\n alpha", r, "poly(alpha)/N %.6g" % (P.polyval2d(r[0], r[1], c) / (n * n)), flush=True)
    return r


cg._cg_alpha_min = logged
gl.solve.cg(n_iter=3)
print("energies/N", [float(e) / (n * n) for e in cg.cg_energies])
