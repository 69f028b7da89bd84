"""Solver front-end: ``gl.solve.td(...)`` and ``gl.solve.cg(...)``
(API of svirl/solvers/solvers.py:11-115)."""
import numpy as np

from .cg import CG
from .td import TD


class Solvers(object):

    def __init__(self, par, mesh, _vars, params, observables):
        self.par = par
        self.mesh = mesh
        self.vars = _vars
        self.params = params
        self.observables = observables
        self._td = None
        self._cg = None
        self.__detected = False

    def td(self, dt=0.1, Nt=1000, T=None, eqn=None):
        """Integrate the time-dependent GL equations for Nt steps of size dt (or up to time T).
        eqn: None (psi and, for finite kappa, A), 'order_parameter' or 'vector_potential'."""
        assert isinstance(Nt, (np.integer, int))
        assert isinstance(dt, (np.floating, float, np.integer, int))
        if eqn is not None:
            assert eqn in ['order_parameter', 'vector_potential']
        if T is not None:
            assert isinstance(T, (np.floating, float, np.integer, int))
        self._init_td()
        self._td._solve(dt=dt, Nt=Nt, T=T, eqn=eqn)
        self.vortices_detected = False

    def cg(self, n_iter=1000):
        """Minimise the free energy with the modified nonlinear conjugate-gradient method."""
        self._init_cg()
        self._cg._solve(n_iter)
        self.vortices_detected = False

    @property
    def vortices_detected(self):
        return self.__detected

    @vortices_detected.setter
    def vortices_detected(self, status):
        assert isinstance(status, bool)
        self.__detected = status

    def _init_td(self):
        if self._td is None:
            self._td = TD(self.par, self.mesh, self.vars, self.params, self.observables)

    def _init_cg(self):
        if self._cg is None:
            self._cg = CG(self.par, self.mesh, self.vars, self.params, self.observables)
