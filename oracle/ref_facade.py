"""TEST INFRASTRUCTURE (oracle side) -- run the reference's OWN controller code on a pluggable QP core.

`make_reference_controller(model, core)` instantiates the unmodified
``warm_start_hmpc.controller.HybridModelPredictiveController`` (imported from /root/reference) with
only ``_build_mip`` and ``_update_mu`` overridden: ``_build_mip`` returns `QPFacade`, an object with
the BoundedQP surface that ``controller.py`` / ``subproblem_solution.py`` touch
(bounded_qp.py:159-332: set/get_constraint_rhs, optimize, primal/dual_optimizer,
primal/dual_objective, status, Runtime, Params.Method, reset, resetParams, setParam), rows and
variables named exactly as controller.py:135-179.  `core(x0, lb, ub)` supplies the arithmetic
Gurobi would have done.  ``feedforward``, ``branch_and_bound``, ``_brancher``,
``SubproblemSolution.from_controller`` and ``construct_warm_start`` then run verbatim.
"""
import time
import numpy as np
from oracle.condense import Condensed


def unpack_solution(cond, model, x0, out):
    """condensed result -> the reference's named families (subproblem_solution.py:86-99, 137-166)."""
    c = cond
    T, nu, nuc, nub, nx = c.T, c.nu, c.nuc, c.nub, c.nx
    fam = {}
    y = out['y']
    mu = []
    for t in range(T):
        k = c.nh if t < T - 1 else c.nh1
        mu.append(np.maximum(y[c.row0[t]:c.row0[t] + k], 0.))
    yb = y[c.mc:].reshape(T, nub)
    nu_ub = np.maximum(yb, 0.); nu_lb = np.maximum(-yb, 0.)
    optimal = out['status'] == 2
    if optimal:
        z = out['z']
        x = c.states(x0, z)
        u = z.reshape(T, nu)
        for t in range(T + 1):
            fam['x_%d' % t] = x[t]
        for t in range(T):
            fam['uc_%d' % t] = u[t, :nuc]; fam['ub_%d' % t] = u[t, nuc:]
    # lam by the backward recursion of the dual constraints (test/cart_pole_with_wall.py:207-247)
    Q, Q_T, A = model['Q'], model['Q_T'], model['A']
    lam = [None] * (T + 1)
    lam[T] = -2. * Q_T.T.dot(Q_T.dot(x[T])) if optimal else np.zeros(nx)
    for t in range(T - 1, -1, -1):
        Ft = model['F'] if t < T - 1 else model['F_Tm1']
        lam[t] = A.T.dot(lam[t + 1]) - Ft.T.dot(mu[t])
        if optimal:
            lam[t] -= 2. * Q.T.dot(Q.dot(x[t]))
    for t in range(T + 1):
        fam['lam_%d' % t] = lam[t]
    for t in range(T):
        fam['mu_%d' % t] = mu[t]; fam['nu_lb_%d' % t] = nu_lb[t]; fam['nu_ub_%d' % t] = nu_ub[t]
    return fam


class _Params(object):
    Method = -1


class QPFacade(object):

    def __init__(self, model, core):
        self.model, self.core = model, core
        self.cond = Condensed(model)
        c = self.cond
        self.rhs = {'lam_0': np.zeros(c.nx)}
        for t in range(c.T):
            self.rhs['nu_lb_%d' % t] = np.zeros(c.nub)
            self.rhs['nu_ub_%d' % t] = np.ones(c.nub)
        self.status = 1
        self.Runtime = 0.
        self.Params = _Params()
        self.objVal = None
        self.log = []          # (x0, lb, ub, out) of every solve, for trace capture

    # -- bounded_qp.py:159-198
    def set_constraint_rhs(self, name, rhs):
        if len(self.rhs[name]) != len(rhs):
            raise ValueError('The rhs does not have the right dimension.')
        self.rhs[name] = np.array(rhs, dtype=float)

    def get_constraint_rhs(self, name):
        return self.rhs[name].copy()

    # -- bounded_qp.py:200-228
    def optimize(self):
        c = self.cond
        x0 = self.rhs['lam_0']
        lb = -np.concatenate([self.rhs['nu_lb_%d' % t] for t in range(c.T)])
        ub = np.concatenate([self.rhs['nu_ub_%d' % t] for t in range(c.T)])
        tic = time.perf_counter()
        out = self.core(x0, lb, ub)
        self.Runtime = time.perf_counter() - tic
        if out['status'] not in (2, 3):
            raise RuntimeError('QP core failed with status %r' % out['status'])
        self.status = out['status']
        self.out = out
        self.fam = unpack_solution(c, self.model, x0, out)
        if self.status == 2:
            self.objVal = out['cost']
        self.log.append((x0.copy(), lb, ub, out))

    def _raise_if_not_solved(self):
        if self.status == 1:
            raise RuntimeError('Problem not solved yet.')

    # -- bounded_qp.py:230-332
    def primal_optimizer(self, name):
        self._raise_if_not_solved()
        return self.fam[name].copy() if self.status == 2 else None

    def dual_optimizer(self, name):
        self._raise_if_not_solved()
        return self.fam[name].copy()

    def primal_objective(self):
        self._raise_if_not_solved()
        return self.objVal if self.status == 2 else np.inf

    def dual_objective(self):
        self._raise_if_not_solved()
        return self.objVal if self.status == 2 else self.out['farkas']

    def reset(self):
        self.status = 1

    def resetParams(self):
        pass

    def setParam(self, *a):
        pass

    def getConstrs(self):
        return []

    def getVars(self):
        return []


class _MLD(object):
    pass


def make_reference_controller(model, core):
    from oracle.refload import import_reference
    ctrl, bnb, sps, mlds = import_reference()
    mld = mlds.MLDSystem([model['A'], model['B']], [model['F'], model['G'], model['h']], int(model['nub']))

    class Controller(ctrl.HybridModelPredictiveController):
        def _build_mip(self_):
            return QPFacade(model, core)

        def _update_mu(self_):
            return model['M_mu']

    ts = [model['F_T'], model['h_T']]
    return Controller(mld, int(model['T']), [model['Q'], model['R'], model['Q_T']], ts)
