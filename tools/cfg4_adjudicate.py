#!/usr/bin/env python
"""BASELINE configs[3] (CG minimiser, finite kappa, fp64): does the UNMODIFIED reference's line search run away on
large grids, or is that a bug of this library?  ("settle cfg4".)

Runs, on one GPU, the same seeded problem through
  (1) B-ref: the unmodified reference package with its own CUDA kernels (baseline/bref.py), and
  (2) svirl_b200 (reference line search, rescue counted separately),
20 TDGL steps then n CG iterations, and records per iteration the energy and the (alpha_psi, alpha_A) the host line
search returned.  Output: one JSON file (default gpurun_out/cfg4_adjudicate_<N>.json).

    python tools/cfg4_adjudicate.py --n 8192 --iters 12
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))


def record_alphas(cg, name):
    out = []
    orig = getattr(cg, name)

    def wrap(*a, **k):
        r = orig(*a, **k)
        out.append([float(x) for x in np.atleast_1d(r)])
        return r
    setattr(cg, name, wrap)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--td-steps", type=int, default=20)
    ap.add_argument("--kappa", type=float, default=2.0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-ref", action="store_true")
    args = ap.parse_args()
    N = args.n
    kw = dict(Nx=N, Ny=N, dx=0.5, dy=0.5, dtype=np.float64, gl_parameter=args.kappa, normal_conductivity=10.0,
              homogeneous_external_field=0.1, random_seed=1234)
    res = {"config": {k: (v if not isinstance(v, type) else v.__name__) for k, v in kw.items()},
           "td_steps": args.td_steps, "iters": args.iters}

    if not args.skip_ref:
        import bref
        t0 = time.perf_counter()
        ref = bref.make_solver(**kw)
        ref.solve.td(dt=0.1, Nt=args.td_steps)
        ref.solve._init_cg()
        ref.solve._cg._CG__convergence_rtol = -1.0
        al = record_alphas(ref.solve._cg, "_cg_alpha_min")
        with np.errstate(all="ignore"):
            try:
                ref.solve.cg(n_iter=args.iters)
                err = None
            except Exception as e:           # a NaN state can make SciPy / polyroots raise
                err = "%s: %s" % (type(e).__name__, e)
        res["reference"] = {"energies": [float(e) for e in ref.solve._cg.cg_energies], "alphas": al, "error": err,
                            "seconds": time.perf_counter() - t0,
                            "psi_abs_max_end": float(np.nanmax(np.abs(bref.fields(ref)[0])))}
        print("reference:", res["reference"]["energies"], flush=True)
        del ref

    from svirl_b200 import GLSolver
    t0 = time.perf_counter()
    gl = GLSolver(**kw)
    gl.solve.td(dt=0.1, Nt=args.td_steps)
    gl.solve._init_cg()
    gl.solve._cg._CG__convergence_rtol = -1.0
    al = record_alphas(gl.solve._cg, "_cg_alpha_min_guarded")
    with np.errstate(all="ignore"):
        gl.solve.cg(n_iter=args.iters)
    res["ours"] = {"energies": [float(e) for e in gl.solve._cg.cg_energies], "alphas": al,
                   "line_search_rescues": int(gl.solve._cg.line_search_rescues), "seconds": time.perf_counter() - t0}
    print("ours:", res["ours"]["energies"], flush=True)

    if "reference" in res:
        Er, Eo = np.array(res["reference"]["energies"]), np.array(res["ours"]["energies"])
        n = min(len(Er), len(Eo))
        ok = np.isfinite(Er[:n]) & (np.abs(Er[:n]) < 1e30)
        k = int(np.argmin(ok)) if not ok.all() else n        # iterations before the reference goes non-finite / huge
        rel = np.abs(Eo[:k] - Er[:k]) / np.maximum(np.abs(Er[:k]), 1e-300)
        res["verdict"] = {"reference_finite_iterations": k, "reference_energy_decreasing": bool(np.all(np.diff(Er[:k]) < 0)),
                          "max_rel_energy_diff_while_reference_finite": float(rel.max()) if k else None,
                          "rel_energy_diff": [float(x) for x in rel]}
        print("verdict:", res["verdict"], flush=True)
    out = args.out or os.path.join(ROOT, "gpurun_out", "cfg4_adjudicate_%d.json" % N)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
