"""Multi-GPU check (run under torchrun on a box with >= 2 GPUs; see tests/test_gpu_slab.py):
TDGL on row slabs must reproduce the single-GPU trajectory BITWISE (MAX is exact), with the
same sweep counts, for psi-only and finite-kappa runs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def run(kw, Nt, slab):
    from svirl_b200 import GLSolver
    gl = GLSolver(slab=slab, **kw)
    gl.solve.td(dt=0.1, Nt=Nt)
    td = gl.solve._td
    psi = gl.vars._psi.get_d_obj().get()
    ab = gl.vars._vp.get_d_obj().get()
    out = [gl.unflatten_array(psi), ab, td.sweeps_order_parameter, td.sweeps_vector_potential, gl.cfg.slab, None]
    if os.environ.get("SLAB_CG"):          # CG iterations on slabs (option cg_slabs), compared by check()
        if slab is not None:
            gl.par.set_option("cg_slabs", 1)
        gl.solve.cg(n_iter=int(os.environ["SLAB_CG"]))
        out[5] = (np.array(gl.solve._cg.cg_energies, dtype=np.float64), gl.unflatten_array(gl.vars._psi.get_d_obj().get()))
    gl.par.close()
    return out


def check(local, verbose=True):
    """Runs inside an initialised NCCL process group; -> True on every rank iff all ranks agree bit for bit."""
    rank = dist.get_rank()
    Nx, Ny = 300, int(os.environ.get("SLAB_NY", "401"))      # SLAB_NY=201 on 4 ranks: 50-row slabs (edge case)
    rs = np.random.RandomState(3)
    mt = rs.rand(Nx - 1, Ny - 1) > 0.1
    ok = True
    for name, kw, Nt in (
            ("kinf_f32", dict(dtype=np.float32, material_tiling=mt), 12),
            ("k2_f64", dict(dtype=np.float64, gl_parameter=2.0, normal_conductivity=10.0, material_tiling=mt), 8)):
        kw = dict(Nx=Nx, Ny=Ny, dx=0.5, dy=0.5, homogeneous_external_field=0.1, random_seed=5, device_id=local, **kw)
        Nt = int(os.environ.get("SLAB_NT", Nt))
        psi_s, ab_s, ns, na, slab, cg_s = run(kw, Nt, "auto")
        psi_1, ab_1, ns1, na1, _, cg_1 = run(kw, Nt, None)      # every rank also runs the whole grid alone
        j0, j1 = slab
        same_psi = np.array_equal(psi_s[:, j0:j1], psi_1[:, j0:j1])
        Na = (Nx - 1) * Ny
        a_s, a_1 = ab_s[:Na].reshape(Ny, Nx - 1), ab_1[:Na].reshape(Ny, Nx - 1)
        b_s, b_1 = ab_s[Na:].reshape(Ny - 1, Nx), ab_1[Na:].reshape(Ny - 1, Nx)
        same_a = np.array_equal(a_s[j0:j1], a_1[j0:j1]) and np.array_equal(b_s[j0:min(j1, Ny - 1)], b_1[j0:min(j1, Ny - 1)])
        good = same_psi and same_a and (ns, na) == (ns1, na1)
        if cg_s is not None:
            # CG on slabs: the sums go over all ranks in rank order, so every rank must hold the SAME energies bit for
            # bit (the replicated host line search depends on it) and they equal the single-GPU ones up to the different
            # summation order (1e-16 per sum, amplified by the BFGS line search over the iterations)
            Es, E1 = cg_s[0], cg_1[0]
            gathered = [None] * dist.get_world_size()
            dist.all_gather_object(gathered, Es.tobytes())
            same_bits = all(x == gathered[0] for x in gathered)
            close = bool(np.allclose(Es, E1, rtol=1e-9 if kw["dtype"] is np.float64 else 1e-3))
            close_psi = float(np.abs(cg_s[1][:, j0:j1] - cg_1[1][:, j0:j1]).max())
            if verbose:
                print("rank %d %s: CG energies on slabs %s | identical on all ranks %s | vs single GPU close %s (psi diff %.2e)"
                      % (rank, name, Es.tolist(), same_bits, close, close_psi), flush=True)
            good = good and same_bits and close and close_psi < (1e-7 if kw["dtype"] is np.float64 else 1e-2)
        if verbose:
            print("rank %d %s: rows [%d,%d) psi bitwise %s, A bitwise %s, sweeps %d/%d vs %d/%d -> %s"
                  % (rank, name, j0, j1, same_psi, same_a, ns, na, ns1, na1, "OK" if good else "MISMATCH"), flush=True)
        ok = ok and good
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t.item()) == 1


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = check(local)
    dist.destroy_process_group()
    if rank == 0:
        print("SLAB CHECK", "PASSED" if ok else "FAILED", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
