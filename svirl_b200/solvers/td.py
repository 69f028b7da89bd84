"""TDGL time stepping (API of svirl/solvers/td.py:13-378).

The reference drives every Jacobi sweep from Python (zero-fill + kernel + blocking 4-byte
read-back per sweep).  Here the whole loop -- Nt x [psi-solve; A-solve] -- runs inside the
library (svl_td_run): sweeps are launched up to the count predicted from the previous step,
the per-sweep residual maxima are read back once per solve, and the exact reference stop
sweep is recovered (see svirl_b200/csrc/td.cu)."""
import ctypes as C

import numpy as np

import svirl_b200.config as cfg
from svirl_b200 import _lib


def _h(x):
    return x.handle if hasattr(x, 'handle') else None


class TD(object):

    def __init__(self, par, mesh, _vars, params, observables):
        self.par = par
        self.mesh = mesh
        self.vars = _vars
        self.params = params
        self.fixed_vortices = self.params.fixed_vortices
        self.observables = observables
        self.solveA = self.params.solveA
        self._random_t = np.uint32(1)
        if cfg.random_seed is not None:
            self._random_t = np.uint32(cfg.random_seed)
        # stop criteria are clamped to what the precision can resolve (td.py:51-66); cfg is
        # mutated in place like in the reference
        floor = 1e-6 if cfg.dtype is np.float32 else 1e-12
        cfg.stop_criterion_order_parameter = cfg.dtype(max(cfg.stop_criterion_order_parameter, floor))
        cfg.stop_criterion_vector_potential = cfg.dtype(max(cfg.stop_criterion_vector_potential, floor))
        self.sweeps_order_parameter = 0      # executed Jacobi sweeps (new: needed for bytes accounting)
        self.sweeps_vector_potential = 0
        self.td_energies = []

    def _set_iterator_options(self, iterator_type, Nt=None, dt=None, T=None, mandatory_definition=True):
        assert iterator_type in ['order_parameter', 'vector_potential']
        if Nt is not None and T is not None and dt is None:
            dt = float(T) / Nt
        elif Nt is not None and T is None and dt is not None:
            T = float(dt) * Nt
        elif Nt is None and T is not None and dt is not None:
            Nt = int(np.round(T / dt))
        elif Nt is not None and T is not None and dt is not None:
            assert np.isclose(T, dt * Nt)
        if mandatory_definition:
            assert isinstance(dt, (np.floating, float, np.integer, int)) and dt >= 0.0
            assert isinstance(Nt, (np.integer, int)) and Nt >= 0
        vals = (np.int32(Nt) if Nt is not None else None, cfg.dtype(dt) if dt is not None else None,
                cfg.dtype(T) if T is not None else None)
        if iterator_type == 'order_parameter':
            self.Nt, self.dt, self.T = vals
        else:
            self.NtA, self.dtA, self.TA = vals

    def _warn_order_parameter(self):
        d = self.dt / min(cfg.dx, cfg.dy) ** 2
        if d > 1.0 and not getattr(self, '_warned_psi', False):
            print('Warning (order parameter):  dt/min(dx,dy)^2 = %g is too large' % d)
        self._warned_psi = d > 1.0

    def _warn_vector_potential(self):
        k, s, h = self.params.gl_parameter, self.params.normal_conductivity, min(cfg.dx, cfg.dy)
        d = self.dtA * k ** 2 / (s * h ** 2)
        if d > 1.0 and not getattr(self, '_warned_A0', False):
            print('Warning (vector potential):  dt*kappa^2/(sigma*min(dx,dy)^2) = %g is too large' % d)
        self._warned_A0 = d > 1.0
        d = self.dtA / (s * h)
        if d > 1.0 and not getattr(self, '_warned_A1', False):
            print('Warning (vector potential):  dt/(sigma*min(dx,dy)) = %g is too large' % d)
        self._warned_A1 = d > 1.0

    def _eps_args(self):
        p = self.params
        return float(np.asarray(p.linear_coefficient_scalar_h()).reshape(-1)[0]), _h(p.linear_coefficient_h())

    def _run_fixed_vortices(self, Nt, dt, do_psi, do_A, pre_once):
        """Time stepping with fixed vortices and / or phase lock: the irregular potential is folded
        into the regular one around each solve, step by step, exactly in the reference's order and with
        its element counts (svirl/solvers/td.py:120-155, 207-216, 252-325; SURVEY quirk Q5: the fold
        covers the first Nx*Ny entries of the packed edge array only, and after the A-solve the
        irregular potential absorbs the CHANGE of the regular one on those entries)."""
        p, fv = self.params, self.fixed_vortices
        ctx = self.par.ctx
        self.vars._psi.push()
        self.vars._vp.push()
        eps, epsf = self._eps_args()
        psi, ab = self.vars.order_parameter_h(), self.vars.vector_potential_h()
        vpi = fv._vpi.get_d_obj() if fv._vpi is not None else None
        if fv._vpi is not None:
            fv._vpi.push()
        lock = fv._phase_lock_ns
        n = C.c_int()
        N = int(cfg.N)
        k2, rho, H = float(p.gl_parameter_squared_h()), float(p._rho), float(p.homogeneous_external_field)

        def fold(x, y, sign):
            _lib.call("svl_edge_axpy_flat", ctx, x.handle, y.handle, float(sign), N)

        if do_psi and pre_once and vpi is not None:
            fold(ab, vpi, +1.0)                      # eqn='order_parameter': added once, removed every step (td.py:227-231)
        for _ in range(int(Nt)):
            if do_psi:
                if not pre_once and vpi is not None:
                    fold(ab, vpi, +1.0)
                _lib.call("svl_td_psi_solve", ctx, float(dt), eps, epsf, ab.handle, psi.handle,
                          float(p.order_parameter_Langevin_coefficient), int(self._random_t),
                          float(cfg.stop_criterion_order_parameter), C.byref(n))
                self._random_t = np.uint32(int(self._random_t) + 1)
                self.sweeps_order_parameter += n.value
                if lock is not None:
                    _lib.call("svl_phase_lock", ctx, psi.handle, lock.get_d_obj().handle, int(lock.size))
                if vpi is not None:
                    fold(ab, vpi, -1.0)
            if do_A and self.solveA:
                if vpi is not None:
                    fold(vpi, ab, +1.0)
                    _lib.call("svl_td_a_solve_ph", ctx, float(self.dt), k2, rho, H, psi.handle, vpi.handle, ab.handle,
                              float(p.vector_potential_Langevin_coefficient), int(self._random_t),
                              float(cfg.stop_criterion_vector_potential), C.byref(n))
                else:
                    _lib.call("svl_td_a_solve", ctx, float(self.dt), k2, rho, H, psi.handle, ab.handle,
                              float(p.vector_potential_Langevin_coefficient), int(self._random_t),
                              float(cfg.stop_criterion_vector_potential), C.byref(n))
                self._random_t = np.uint32(int(self._random_t) + 1)
                self.sweeps_vector_potential += n.value
                if vpi is not None:
                    fold(vpi, ab, -1.0)
        self.vars._psi.need_dtoh_sync()
        if do_A and self.solveA:
            self.vars._vp.need_dtoh_sync()           # the psi-only path leaves the host copy of A as it was (td.py:216)
        # the reference never marks the irregular potential as changed on the device, so its host copy (what
        # irregular_vector_potential and the vortex detector read) keeps the values set by the user while the
        # device copy drifts; same here

    def _run(self, Nt, dt, do_psi=True, do_A=True, pre_once=False):
        fv = self.fixed_vortices
        if fv._vpi is not None or fv._phase_lock_ns is not None:
            return self._run_fixed_vortices(Nt, dt, do_psi, do_A, pre_once)
        p = self.params
        self.vars._psi.push()
        self.vars._vp.push()
        eps, epsf = self._eps_args()
        psi, ab = self.vars.order_parameter_h(), self.vars.vector_potential_h()
        rt = C.c_uint32(int(self._random_t))
        ctx = self.par.ctx
        if do_psi:
            sweeps = (C.c_longlong * 2)(0, 0)
            _lib.call("svl_td_run", ctx, int(Nt), float(dt), int(bool(do_A and self.solveA)), eps, epsf,
                      float(p.gl_parameter_squared_h()), float(p._rho), float(p.homogeneous_external_field),
                      psi.handle, ab.handle, float(p.order_parameter_Langevin_coefficient),
                      float(p.vector_potential_Langevin_coefficient), C.byref(rt),
                      float(cfg.stop_criterion_order_parameter), float(cfg.stop_criterion_vector_potential), sweeps)
            self.sweeps_order_parameter += sweeps[0]
            self.sweeps_vector_potential += sweeps[1]
        else:
            n = C.c_int()
            for _ in range(int(Nt)):
                _lib.call("svl_td_a_solve", ctx, float(dt), float(p.gl_parameter_squared_h()), float(p._rho),
                          float(p.homogeneous_external_field), psi.handle, ab.handle,
                          float(p.vector_potential_Langevin_coefficient), rt.value,
                          float(cfg.stop_criterion_vector_potential), C.byref(n))
                rt.value += 1
                self.sweeps_vector_potential += n.value
        self._random_t = np.uint32(rt.value)
        self.vars._psi.need_dtoh_sync()
        self.vars._vp.need_dtoh_sync()

    def _solve(self, dt=None, Nt=None, T=None, eqn=None):
        self.solveA = self.params.solveA
        if eqn == "order_parameter":
            self._set_iterator_options('order_parameter', dt=dt, Nt=Nt, T=T)
            self._warn_order_parameter()
            self._run(self.Nt, self.dt, do_psi=True, do_A=False, pre_once=True)
        elif eqn == "vector_potential":
            if not self.solveA:
                return
            self._set_iterator_options('vector_potential', dt=dt, Nt=Nt, T=T)
            self._warn_vector_potential()
            # the reference passes self.dt (not dtA) to the A kernel (td.py:278, quirk Q3) and fails
            # with AttributeError when no psi step has defined it yet
            self._run(self.NtA, self.dt, do_psi=False, do_A=True)
        else:
            self._set_iterator_options('order_parameter', dt=dt, Nt=Nt, T=T)
            self._set_iterator_options('vector_potential', dt=dt, Nt=Nt, T=T)
            self._warn_order_parameter()
            if self.solveA:
                self._warn_vector_potential()
            self.td_energies = []
            self._run(self.Nt, self.dt, do_psi=True, do_A=True)
