"""Modified nonlinear conjugate-gradient minimiser (API of svirl/solvers/cg.py:12-558).

Device work per iteration is two library calls and two passes over HBM (svl_cg_pass_b: direction
update + line-search coefficients; svl_cg_pass_a: variable update + free energy + the Jacobians of
the NEXT iteration + its PR+ beta); with an external potential or on row slabs the three-pass pair
svl_cg_begin / svl_cg_end is used instead.  The line search itself stays on the host and is the
reference's numpy/scipy call, unchanged, because the trajectory depends on it bit by bit
(cg.py:227-235, 378-419).

State that persists across cg() calls exactly as in the reference (quirk Q6): beta_psi,
beta_A and, for finite kappa, the search directions."""
import ctypes as C

import numpy as np
import scipy.optimize

import svirl_b200.config as cfg
from svirl_b200 import _lib
from svirl_b200.storage.arrays import DeviceArray


def _h(x):
    return x.handle if hasattr(x, 'handle') else None


def _horner2d(x, y, rows):
    """numpy.polynomial.polynomial.polyval2d(x, y, c) for Python floats x, y and c given as nested lists: Horner in x
    down the rows (a vector over the columns, starting from c[-1] + x*0 like polyval), then Horner in y; the same
    IEEE operations in the same order, hence the same bits."""
    n, m = len(rows), len(rows[0])
    z = x * 0.0
    r = [v + z for v in rows[n - 1]]
    for i in range(n - 2, -1, -1):
        ci = rows[i]
        r = [ci[k] + r[k] * x for k in range(m)]
    v = r[m - 1] + y * 0.0
    for k in range(m - 2, -1, -1):
        v = r[k] + v * y
    return v


class CG(object):

    def __init__(self, par, mesh, _vars, params, observables):
        self.par = par
        self.mesh = mesh
        self.vars = _vars
        self.params = params
        self.fixed_vortices = self.params.fixed_vortices
        self.observables = observables
        self.__convergence_rtol = cfg.convergence_rtol
        solveA = self.params.solveA
        self.__reduction_vector_length = 17 if solveA else 5
        Z = lambda kind: DeviceArray.zeros(par, kind)
        self.__gdir_psi, self.__gjac_psi, self.__gjac_psi_prev = Z(_lib.NODE_C), Z(_lib.NODE_C), Z(_lib.NODE_C)
        self._beta_psi = 0.0
        self._beta_A = 0.0
        if solveA:
            self.__gdir_A, self.__gjac_A, self.__gjac_A_prev = Z(_lib.EDGE), Z(_lib.EDGE), Z(_lib.EDGE)
        self.__c = np.zeros((5, 5), dtype=cfg.dtype) if solveA else np.zeros(5, dtype=cfg.dtype)
        self.cg_energies = []
        self.line_search_rescues = 0      # times the reference's BFGS call ran away and was redone on c / N

    # ---- argument helpers
    def _state(self):
        p = self.params
        self.vars._psi.push()
        if self.vars._vp is not None:
            self.vars._vp.push()
        eps = float(np.asarray(p.linear_coefficient_scalar_h()).reshape(-1)[0])
        return dict(k2=float(p.gl_parameter_squared_h()), eps=eps, epsf=_h(p.linear_coefficient_h()),
                    H=float(p.homogeneous_external_field), psi=self.vars.order_parameter_h().handle,
                    abei=_h(p._external_irregular_for_kernels()), ab=_h(self.vars.vector_potential_h()))

    # ---- kernel-level private API used by the reference's tests
    @property
    def _free_energy_jacobian_psi(self):
        """dG/dRe(psi) + i dG/dIm(psi) as a flat device array (cg.py:112-146)."""
        s = self._state()
        _lib.call("svl_jacobian_psi", self.par.ctx, s['k2'], s['eps'], s['epsf'], s['H'], s['psi'], s['abei'],
                  s['ab'], self.__gjac_psi.handle)
        return self.__gjac_psi

    @property
    def _free_energy_jacobian_A(self):
        """[dG/da, dG/db] as a flat device array (cg.py:149-181); None for infinite kappa."""
        if not self.params.solveA:
            return None
        s = self._state()
        _lib.call("svl_jacobian_A", self.par.ctx, s['k2'], s['H'], s['psi'], s['abei'], s['ab'],
                  self.__gjac_A.handle)
        return self.__gjac_A

    def _free_energy_conjgrad_coef_psi(self, gdir_psi):
        """c0..c4 of G(psi + alpha*dpsi) (cg.py:184-224)."""
        s = self._state()
        out = (C.c_double * 5)()
        _lib.call("svl_cg_coef_psi", self.par.ctx, s['k2'], s['eps'], s['H'], s['psi'], gdir_psi.handle,
                  s['abei'], s['ab'], out)
        self.__c[:] = np.array(out[:], dtype=cfg.dtype)     # broadcasts into every row of a 5x5 c, as in the reference
        return self.__c

    def _free_energy_conjgrad_coef(self, gdir_psi, gdir_A):
        """c[i,j] of G(psi + a_psi*dpsi, A + a_A*dA), 4th order in a_A (cg.py:325-375)."""
        s = self._state()
        out = (C.c_double * 17)()
        _lib.call("svl_cg_coef", self.par.ctx, s['k2'], s['eps'], s['H'], s['psi'], gdir_psi.handle,
                  s['abei'], s['ab'], gdir_A.handle, out)
        self._store_c17(np.array(out[:], dtype=cfg.dtype))
        return self.__c

    def _store_c17(self, r):
        c = self.__c
        c[0, :], c[1, :], c[2, :] = r[0:5], r[5:10], r[10:15]
        c[3, 0], c[4, 0] = r[15], r[16]

    # ---- host line searches: the reference's calls, verbatim in behaviour
    def _cg_alpha_psi_min(self):
        c = self.__c
        am = np.polynomial.polynomial.polyroots([c[1], 2.0 * c[2], 3.0 * c[3], 4.0 * c[4]])
        am = am[np.isclose(am.imag, 0)].real
        am = am[am >= 0]
        return np.min(am)

    def _cg_alpha_min(self, alpha0=[0.0, 0.0], tol=1e-8):
        """The reference's call: scipy.optimize.minimize(BFGS, tol=1e-8) on polyval2d(alpha, c) with the polyder
        gradients (cg.py:378-419).  The objective and its gradient are evaluated by `_horner2d`, which performs
        polyval2d's floating-point operations in polyval2d's order on Python floats: every value SciPy sees is
        bit-identical (tests/test_linesearch_host.py compares x, nit and nfev with the verbatim numpy callables),
        so the iterates are the reference's; it only skips numpy's per-call overhead (~0.4 ms per line search)."""
        c = self.__c
        P = np.polynomial.polynomial
        C0, C1, CC = P.polyder(c, axis=0).tolist(), P.polyder(c, axis=1).tolist(), np.asarray(c, dtype=np.float64).tolist()
        if np.asarray(c).dtype != np.float64:          # fp32 solver: keep numpy's own promotion rules
            return self._cg_alpha_min_numpy(alpha0, tol)

        def f(alpha):
            return np.float64(_horner2d(float(alpha[0]), float(alpha[1]), CC))

        def j(alpha):
            x, y = float(alpha[0]), float(alpha[1])
            return np.array([_horner2d(x, y, C0), _horner2d(x, y, C1)])

        r = scipy.optimize.minimize(f, x0=np.array(alpha0), jac=j, method='BFGS', tol=tol)
        return r.x

    def _cg_alpha_min_numpy(self, alpha0=[0.0, 0.0], tol=1e-8):
        """The same minimisation with numpy's polyval2d as the callables, exactly as written in cg.py:378-419."""
        c = self.__c
        P = np.polynomial.polynomial
        cj0 = P.polyder(c, axis=0)
        cj1 = P.polyder(c, axis=1)

        def f(alpha):
            return P.polyval2d(alpha[0], alpha[1], c)

        def j(alpha):
            alpha_psi, alpha_A = alpha
            return np.array([P.polyval2d(alpha_psi, alpha_A, cj0), P.polyval2d(alpha_psi, alpha_A, cj1)])

        r = scipy.optimize.minimize(f, x0=np.array(alpha0), jac=j, method='BFGS', tol=tol)
        return r.x

    def _cg_alpha_min_native(self):
        """Opt-in (cfg.cg_line_search = 'native'): the library's damped-Newton search on c / max|c|."""
        c = self.__c
        flat = (C.c_double * 17)(*([float(v) for v in c[0, :]] + [float(v) for v in c[1, :]] + [float(v) for v in c[2, :]]
                                   + [float(c[3, 0]), float(c[4, 0])]))
        out = (C.c_double * 2)()
        _lib.call("svl_cg_line_search", flat, 1, out, None)
        return np.array([out[0], out[1]])

    def _cg_alpha_min_scaled(self):
        """The same BFGS minimisation on c / max|c| (an O(1) objective)."""
        c = self.__c
        scale = float(np.max(np.abs(c)))
        if not np.isfinite(scale) or scale == 0.0:
            return self._cg_alpha_min()
        self.__c = c / scale
        try:
            return self._cg_alpha_min()
        finally:
            self.__c = c

    def _cg_alpha_min_guarded(self):
        """The reference's BFGS call, plus a rescue for the case where it runs away.

        SciPy's BFGS is not invariant under scaling of the objective: with the coefficients of a
        large grid (they grow with the number of nodes; observed from 8192^2 upwards) its line
        search can leave the basin of the 4th-order polynomial, which is unbounded below in
        alpha_A, and return |alpha| ~ 1e35 -- the state is lost (NaN) from then on, in the
        reference exactly as here.  Whenever the reference result is unusable (not finite, huge,
        or no decrease) the same minimisation is repeated on the normalised polynomial
        c / max|c|, which converges to the local minimum next to alpha = 0.  Runs in which the
        reference call works are unaffected, so trajectory parity holds wherever the reference
        itself survives.  ``cfg.cg_line_search = 'normalized'`` (not a reference option) uses the
        normalised polynomial in every iteration."""
        mode = getattr(cfg, 'cg_line_search', 'reference')
        if mode == 'normalized':
            return self._cg_alpha_min_scaled()
        if mode == 'native':
            return self._cg_alpha_min_native()
        c = self.__c
        P = np.polynomial.polynomial
        with np.errstate(all='ignore'):
            a = self._cg_alpha_min()
            ok = np.all(np.isfinite(a)) and np.max(np.abs(a)) < 1.0e6
            ok = ok and P.polyval2d(a[0], a[1], c) <= c[0, 0]
        if ok:
            return a
        self.line_search_rescues += 1
        return self._cg_alpha_min_scaled()

    # ---- the two minimisation loops
    def _line_search(self, cbuf, solveA):
        r = np.array(cbuf[:], dtype=cfg.dtype)
        if solveA:
            self._store_c17(r)
            return self._cg_alpha_min_guarded()
        self.__c[:] = r
        return self._cg_alpha_psi_min(), 0.0

    def _iterate_two_pass(self, n_iter, solveA, s):
        """One iteration = svl_cg_pass_b (d <- beta d - g, coefficients), host line search, svl_cg_pass_a (update,
        energy, and -- unless this is the last iteration -- the gradient and PR+ beta of the next one).  Same
        arithmetic and the same order of operations on the state as svirl/solvers/cg.py:477-544; the gradient of
        iteration i+1 is merely evaluated in the pass that produces its state."""
        ctx = self.par.ctx
        dA = self.__gdir_A if solveA else None
        gA = self.__gjac_A if solveA else None
        beta = (C.c_double * 2)(self._beta_psi, self._beta_A)
        cbuf = (C.c_double * (17 if solveA else 5))()
        E = C.c_double()
        # gradient at the initial state; beta stays what the previous cg() call left (quirk Q6)
        _lib.call("svl_cg_pass_a", ctx, int(solveA), 0, 1, 0, s['k2'], s['eps'], s['epsf'], s['H'], s['psi'], None, s['ab'],
                  self.__gdir_psi.handle, _h(dA), 0.0, 0.0, self.__gjac_psi.handle, _h(gA), beta, C.byref(E))
        for i in range(n_iter):
            _lib.call("svl_cg_pass_b", ctx, int(solveA), s['k2'], 0.0 if s['epsf'] is not None else s['eps'], s['H'],
                      s['psi'], None, s['ab'], self.__gjac_psi.handle, _h(gA), self.__gdir_psi.handle, _h(dA), cbuf)
            alpha_psi, alpha_A = self._line_search(cbuf, solveA)
            last = i == n_iter - 1
            kept = (beta[0], beta[1])
            _lib.call("svl_cg_pass_a", ctx, int(solveA), 1, int(not last), 1, s['k2'], s['eps'], s['epsf'], s['H'], s['psi'],
                      None, s['ab'], self.__gdir_psi.handle, _h(dA), float(cfg.dtype(alpha_psi)), float(cfg.dtype(alpha_A)),
                      self.__gjac_psi.handle, _h(gA), beta, C.byref(E))
            self.cg_energies.append(cfg.dtype(E.value))
            if i > 0 and np.abs(self.cg_energies[i] / self.cg_energies[i - 1] - 1.0) < self.__convergence_rtol:
                beta[0], beta[1] = kept          # the reference stops before it would compute the next beta (quirk Q6)
                break
        return beta

    def _iterate(self, n_iter, solveA):
        s = self._state()
        ctx = self.par.ctx
        self.cg_energies = []
        if not solveA:
            self.__gdir_psi.fill(0.0)            # the kappa=inf loop restarts from steepest descent (cg.py:264)
        if s['abei'] is None and int(self.par.stat("cg_fused")) >= 2 and not int(self.par.stat("slab_on")) and n_iter > 0:
            beta = self._iterate_two_pass(n_iter, solveA, s)
            self._beta_psi, self._beta_A = beta[0], beta[1]
            self.vars._psi.need_dtoh_sync()
            if solveA:
                self.vars._vp.need_dtoh_sync()
            return
        nul = None
        gA, gAp, dA = ((self.__gjac_A, self.__gjac_A_prev, self.__gdir_A) if solveA else (nul, nul, nul))
        beta = (C.c_double * 2)(self._beta_psi, self._beta_A)
        ncoef = 17 if solveA else 5
        cbuf = (C.c_double * ncoef)()
        E = C.c_double()
        for i in range(n_iter):
            _lib.call("svl_cg_begin", ctx, int(solveA), int(i > 0), s['k2'], s['eps'], s['epsf'], s['H'], s['psi'],
                      s['abei'], s['ab'], self.__gjac_psi.handle, self.__gjac_psi_prev.handle,
                      self.__gdir_psi.handle, _h(gA), _h(gAp), _h(dA), beta, cbuf)
            alpha_psi, alpha_A = self._line_search(cbuf, solveA)
            _lib.call("svl_cg_end", ctx, int(solveA), s['k2'], s['eps'], s['epsf'], s['H'], s['psi'], s['abei'],
                      s['ab'], self.__gdir_psi.handle, _h(dA), float(cfg.dtype(alpha_psi)),
                      float(cfg.dtype(alpha_A)), C.byref(E))
            # "save previous gradient": swap storage instead of the reference's device copy
            self.__gjac_psi.swap(self.__gjac_psi_prev)
            if solveA:
                self.__gjac_A.swap(self.__gjac_A_prev)
            self.cg_energies.append(cfg.dtype(E.value))
            if i > 0 and np.abs(self.cg_energies[i] / self.cg_energies[i - 1] - 1.0) < self.__convergence_rtol:
                break
        self._beta_psi, self._beta_A = beta[0], beta[1]
        self.vars._psi.need_dtoh_sync()
        if solveA:
            self.vars._vp.need_dtoh_sync()

    def _solve(self, n_iter=1000):
        self._iterate(n_iter, bool(self.params.solveA))
