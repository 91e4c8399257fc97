"""``BoundedQP`` -- the reference's QP seam (bounded_qp.py:5-341) over the CUDA node solver (kernel K1).

The reference wraps a Gurobi model: named families of variables and rows, right-hand sides edited by
name per node, ``optimize()`` (with a Farkas proof when infeasible), read-back of primal values and
sign-normalised multipliers by name.  ``controller.py`` and ``subproblem_solution.py`` talk to the
solver ONLY through that surface, so an object with the same surface makes the reference's own
controller code (``_solve_subproblem``, ``feedforward``, ``SubproblemSolution.from_controller``) run
on the GPU solver unmodified (tests/test_gpu_facade.py).

The model behind this facade is the controller's fixed-structure relaxation (controller.py:119-184):

    variables   x_0 | x_{t+1}, uc_t, ub_t  (t = 0..T-1)           in the order _build_mip adds them
    rows        lam_0 | nu_lb_t, nu_ub_t, lam_{t+1}, mu_t         in the order _build_mip adds them

All matrices are shared by every node (they live in the device-resident shared operator); a node
only edits the right-hand sides ``lam_0`` (= x0), ``nu_lb_t`` (= -lb), ``nu_ub_t`` (= ub), which is
all ``_set_bound_binaries`` / ``_solve_subproblem`` ever do (controller.py:253-257, 273-298).

Hot start.  The reference hands a parent's simplex basis to its children through
``active_set = {'c': [CBasis...], 'v': [VBasis...]}`` when ``Params.Method == 1`` (controller.py:262-264,
426; subproblem_solution.py:38-43).  The dual active-set kernel starts from multipliers instead of a
basis, so here ``CBasis`` of a row IS its (sign-normalised, >= 0) multiplier and ``VBasis`` of the
first T*nu variables carries the proximal centre of the solve; the reference's code moves these lists
around without looking inside.  ``Params.Method`` defaults to 1 (the dual method is the only one).

There is no CPU fallback: ``optimize()`` needs the built library and a GPU.
"""
import numpy as np


class _Params(object):
    Method = 1


class _Constr(object):
    """Row proxy: the attributes of grb.Constr the reference touches (ConstrName, RHS, CBasis)."""
    __slots__ = ('_qp', '_i', 'ConstrName')

    def __init__(self, qp, i, name):
        self._qp, self._i, self.ConstrName = qp, i, name

    @property
    def RHS(self):
        return float(self._qp._rhs_flat[self._i])

    def getAttr(self, attr):
        if attr == 'CBasis':
            return float(self._qp._cbasis[self._i])
        if attr == 'RHS':
            return self.RHS
        raise AttributeError(attr)

    def setAttr(self, attr, value):
        if attr != 'CBasis':
            raise AttributeError(attr)
        self._qp._cbasis_in[self._i] = value
        self._qp._pushed = True


class _Var(object):
    """Variable proxy (VarName, x, VBasis)."""
    __slots__ = ('_qp', '_i', 'VarName')

    def __init__(self, qp, i, name):
        self._qp, self._i, self.VarName = qp, i, name

    @property
    def x(self):
        self._qp._raise_if_not_solved()
        return float(self._qp._x_flat[self._i])

    def getAttr(self, attr):
        if attr == 'VBasis':
            return float(self._qp._vbasis[self._i])
        raise AttributeError(attr)

    def setAttr(self, attr, value):
        if attr != 'VBasis':
            raise AttributeError(attr)
        self._qp._vbasis_in[self._i] = value
        self._qp._pushed = True


class BoundedQP(object):

    EDITABLE = 'only lam_0, nu_lb_t and nu_ub_t change between nodes; the other right-hand sides are part of the shared ' \
               'operator on the device (rebuild the controller to change them)'

    def __init__(self, problem, handle_factory):
        """problem: ProblemData; handle_factory(): capi.Handle (created lazily, needs a GPU)."""
        pd = self.pd = problem
        self._handle_factory = handle_factory
        T, nx, nuc, nub = pd.T, pd.nx, pd.nuc, pd.nub
        # families in the order controller.py:135-166 adds them
        self._var_fam, self._con_fam = {}, {}
        vo = co = 0

        def var(name, size):
            nonlocal vo
            self._var_fam[name] = (vo, size); vo += size

        def con(name, size):
            nonlocal co
            self._con_fam[name] = (co, size); co += size
        var('x_0', nx); con('lam_0', nx)
        for t in range(T):
            var('x_%d' % (t + 1), nx); var('uc_%d' % t, nuc); var('ub_%d' % t, nub)
            con('nu_lb_%d' % t, nub); con('nu_ub_%d' % t, nub); con('lam_%d' % (t + 1), nx)
            con('mu_%d' % t, pd.nh if t < T - 1 else pd.nh1)
        self.NumVars, self.NumConstrs = vo, co
        self._rhs_flat = np.zeros(co)
        for t in range(T):
            o, s = self._con_fam['nu_ub_%d' % t]; self._rhs_flat[o:o + s] = 1.
            o, s = self._con_fam['mu_%d' % t]; self._rhs_flat[o:o + s] = pd.h if t < T - 1 else pd.h_Tm1
        self._x_flat = np.zeros(vo)
        self._cbasis = np.zeros(co); self._vbasis = np.zeros(vo)            # of the last solve
        self._cbasis_in = np.zeros(co); self._vbasis_in = np.zeros(vo)      # pushed for the next solve
        self._pushed = False
        self._constrs = self._vars = None
        self.status = 1                  # Gurobi codes: 1 loaded, 2 optimal, 3 infeasible
        self.objVal = None
        self.Runtime = 0.
        self.IterCount = 0
        self.Params = _Params()
        self._user_params = {}
        self._primal = self._dual = None
        self._dobj = None

    # -- bounded_qp.py:19-157 ------------------------------------------------------------------------
    def add_variables(self, n, **kwargs):
        """bounded_qp.py:19-52.  Same argument check as the reference; the model itself is fixed."""
        if 'lb' in kwargs or 'ub' in kwargs:
            raise KeyError('Cannot set bounds with add_variables, use add_constraints instead.')
        raise NotImplementedError('the GPU-backed BoundedQP holds the fixed-structure relaxation of the controller; ' + self.EDITABLE)

    def add_constraints(self, x, op, y, **kwargs):
        """bounded_qp.py:87-125."""
        if len(x) != len(y):
            raise ValueError('Left- and right-hand side must have the same size.')
        raise NotImplementedError('the GPU-backed BoundedQP holds the fixed-structure relaxation of the controller; ' + self.EDITABLE)

    def get_variables(self, name):
        """bounded_qp.py:54-85: proxies of the family (empty if it does not exist)."""
        if name not in self._var_fam:
            return np.array([])
        o, s = self._var_fam[name]
        return np.array(self.getVars()[o:o + s])

    def get_constraints(self, name):
        """bounded_qp.py:127-157."""
        if name not in self._con_fam:
            return np.array([])
        o, s = self._con_fam[name]
        return np.array(self.getConstrs()[o:o + s])

    # -- bounded_qp.py:159-198 -----------------------------------------------------------------------
    def set_constraint_rhs(self, name, rhs):
        o, s = self._con_fam.get(name, (0, 0))
        if s != len(rhs):
            raise ValueError('The rhs does not have the right dimension.')
        if not (name == 'lam_0' or name.startswith('nu_lb_') or name.startswith('nu_ub_')):
            if np.array_equal(np.asarray(rhs, dtype=float), self._rhs_flat[o:o + s]):
                return
            raise NotImplementedError('rhs of %r: %s' % (name, self.EDITABLE))
        self._rhs_flat[o:o + s] = rhs
        self._pushed = False             # Gurobi discards a pushed basis when the model is modified (controller.py:259-261)

    def get_constraint_rhs(self, name):
        o, s = self._con_fam.get(name, (0, 0))
        return self._rhs_flat[o:o + s].copy()

    # -- hot start payload ---------------------------------------------------------------------------
    def _signed_multipliers(self, c):
        """[CBasis per row] -> signed multipliers of the kernel's rows (mu_0..mu_{T-1} | binaries): > 0 upper side."""
        pd = self.pd
        mu = np.concatenate([c[o:o + s] for o, s in (self._con_fam['mu_%d' % t] for t in range(pd.T))])
        lo = np.concatenate([c[o:o + s] for o, s in (self._con_fam['nu_lb_%d' % t] for t in range(pd.T))])
        up = np.concatenate([c[o:o + s] for o, s in (self._con_fam['nu_ub_%d' % t] for t in range(pd.T))])
        return np.concatenate((np.maximum(mu, 0.), np.maximum(up, 0.) - np.maximum(lo, 0.)))

    def active_set_from_dual(self, dual, centre=None):
        """The reference's dormant helper (bounded_qp.py:343-366): the `active_set` a node carrying `dual`
        (a DualSolution, e.g. a shifted one of a warm-start root) starts its solve from.  centre: proximal centre
        (T*nu) or None = 0."""
        c = np.zeros(self.NumConstrs)
        for fam in ('mu', 'nu_lb', 'nu_ub'):
            for t, v in enumerate(dual.variables[fam]):
                o, s = self._con_fam['%s_%d' % (fam, t)]
                c[o:o + s] = np.maximum(v, 0.)
        v = np.zeros(self.NumVars)
        if centre is not None:
            v[:self.pd.n] = centre
        return {'c': list(c), 'v': list(v)}

    def active_set_from_record(self, rec):
        """Same from a flat dual record of a device tree (wshmpc_layout, rec_stride doubles)."""
        pd, L = self.pd, self.pd.layout
        c = np.zeros(self.NumConstrs)
        mu = rec[L.off_mu:L.off_nu_lb]
        lo = rec[L.off_nu_lb:L.off_nu_ub].reshape(pd.T, pd.nub); up = rec[L.off_nu_ub:L.off_rho].reshape(pd.T, pd.nub)
        r = 0
        for t in range(pd.T):
            o, s = self._con_fam['mu_%d' % t]; c[o:o + s] = mu[r:r + s]; r += s
            o, s = self._con_fam['nu_lb_%d' % t]; c[o:o + s] = lo[t]
            o, s = self._con_fam['nu_ub_%d' % t]; c[o:o + s] = up[t]
        v = np.zeros(self.NumVars)
        v[:pd.n] = rec[L.dual:L.dual + pd.n]
        return {'c': list(c), 'v': list(v)}

    def set_active_set(self, active_set):
        """What controller.py:262-264 does row by row, in one call."""
        self._cbasis_in[:] = active_set['c']; self._vbasis_in[:] = active_set['v']
        self._pushed = True

    def get_active_set(self):
        """What subproblem_solution.py:38-43 collects row by row, in one call."""
        return {'c': list(self._cbasis), 'v': list(self._vbasis)}

    # -- bounded_qp.py:200-228 -----------------------------------------------------------------------
    def optimize(self):
        """One K1 launch (wshmpc_solve_nodes) on solver slot 0.  An infeasible node comes back with its Farkas
        proof from the same solve (the reference re-solves a zero-objective LP for it, bounded_qp.py:212-228)."""
        pd = self.pd
        x0 = self.get_constraint_rhs('lam_0') + 0.
        lb = -np.concatenate([self.get_constraint_rhs('nu_lb_%d' % t) for t in range(pd.T)]) + 0.     # + 0.: -0. -> 0.
        ub = np.concatenate([self.get_constraint_rhs('nu_ub_%d' % t) for t in range(pd.T)]) + 0.
        if np.any(lb > ub):
            # lb > ub cannot be written as a two-sided row of the kernel; no caller on the hot path produces it
            raise ValueError('lower bound above upper bound')
        y0 = yc0 = None
        if self._pushed and self.Params.Method == 1:
            y0 = self._signed_multipliers(self._cbasis_in)
            yc0 = self._vbasis_in[:pd.n].copy()
        self._pushed = False
        out = self._launch(x0, lb, ub, y0, yc0)
        status = int(out['status'])
        if status not in (2, 3):
            raise RuntimeError('QP kernel did not converge (status %d)' % status)
        self.status = status
        self.Runtime = float(out['runtime'])
        self.IterCount = int(out['iters'])
        self._primal = np.asarray(out['primal'], dtype=float)
        self._dual = np.asarray(out['dual'], dtype=float)
        self._dobj = float(out['dobj'])
        self.objVal = float(out['cost']) if status == 2 else None
        T, nx, nu, nuc = pd.T, pd.nx, pd.nu, pd.nuc
        if status == 2:
            X = self._primal[:(T + 1) * nx].reshape(T + 1, nx); U = self._primal[(T + 1) * nx:].reshape(T, nu)
            for t in range(T + 1):
                o, s = self._var_fam['x_%d' % t]; self._x_flat[o:o + s] = X[t]
            for t in range(T):
                o, s = self._var_fam['uc_%d' % t]; self._x_flat[o:o + s] = U[t, :nuc]
                o, s = self._var_fam['ub_%d' % t]; self._x_flat[o:o + s] = U[t, nuc:]
        # the payload the children start from: multipliers of this solve + its proximal centre
        a = self.active_set_from_record(np.concatenate((self._dual, np.asarray(out['yc'], dtype=float))))
        self._cbasis[:] = a['c']; self._vbasis[:] = a['v']

    def _launch(self, x0, lb, ub, y0, yc0):
        """The device call: one node through wshmpc_solve_nodes (y0 / yc0: start of the dual method or None = empty
        working set).  Returns host copies of the node's outputs."""
        import torch
        h = self._handle_factory()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        out = h.solve_nodes(x0[None], lb[None], ub[None], slot=np.zeros(1, np.int32),
                            hot=np.array([2 if y0 is not None else 0], np.int32),
                            y0=None if y0 is None else y0[None], yc0=None if yc0 is None else yc0[None])
        end.record(); end.synchronize()
        res = {k: out[k][0].cpu().numpy() for k in ('status', 'cost', 'dobj', 'iters', 'primal', 'dual', 'yc')}
        res['runtime'] = start.elapsed_time(end) * 1e-3
        return res

    def _raise_if_not_solved(self):
        if self.status == 1:
            raise RuntimeError('Problem not solved yet.')

    # -- bounded_qp.py:230-332 -----------------------------------------------------------------------
    def primal_optimizer(self, name):
        self._raise_if_not_solved()
        if self.status != 2:
            return None
        o, s = self._var_fam[name]
        return self._x_flat[o:o + s].copy()

    def dual_optimizer(self, name):
        """Multipliers of a row family, positive for `<=` rows (bounded_qp.py:260-290); the Farkas ray if infeasible."""
        self._raise_if_not_solved()
        pd, L = self.pd, self.pd.layout
        fam, t = name.rsplit('_', 1)
        t = int(t)
        d = self._dual
        if fam == 'lam':
            return d[L.off_lam + t * pd.nx:L.off_lam + (t + 1) * pd.nx].copy()
        if fam == 'mu':
            return d[L.off_mu + t * pd.nh:L.off_mu + t * pd.nh + (pd.nh if t < pd.T - 1 else pd.nh1)].copy()
        if fam == 'nu_lb':
            return d[L.off_nu_lb + t * pd.nub:L.off_nu_lb + (t + 1) * pd.nub].copy()
        if fam == 'nu_ub':
            return d[L.off_nu_ub + t * pd.nub:L.off_nu_ub + (t + 1) * pd.nub].copy()
        raise KeyError(name)

    def primal_objective(self):
        self._raise_if_not_solved()
        return self.objVal if self.status == 2 else np.inf

    def dual_objective(self):
        self._raise_if_not_solved()
        return self.objVal if self.status == 2 else self._dobj

    # -- the grb.Model attributes controller.py touches (:263-269, :362, :778-796) ----------------------
    def getConstrs(self):
        if self._constrs is None:
            self._constrs = [None] * self.NumConstrs
            for name, (o, s) in self._con_fam.items():
                for i in range(s):
                    self._constrs[o + i] = _Constr(self, o + i, '%s[%d]' % (name, i))
        return self._constrs

    def getVars(self):
        if self._vars is None:
            self._vars = [None] * self.NumVars
            for name, (o, s) in self._var_fam.items():
                for i in range(s):
                    self._vars[o + i] = _Var(self, o + i, '%s[%d]' % (name, i))
        return self._vars

    def getConstrByName(self, name):
        fam, i = name[:-1].split('[')
        o, s = self._con_fam.get(fam, (0, 0))
        return self.getConstrs()[o + int(i)] if int(i) < s else None

    def getVarByName(self, name):
        fam, i = name[:-1].split('[')
        o, s = self._var_fam.get(fam, (0, 0))
        return self.getVars()[o + int(i)] if int(i) < s else None

    def reset(self):
        self.status = 1
        self._pushed = False

    def resetParams(self):
        self.Params.Method = 1
        self._user_params = {}

    def setParam(self, name, value):
        if name == 'Method':
            self.Params.Method = int(value)
        else:
            self._user_params[name] = value      # Gurobi tuning knobs have no meaning for the CUDA solver

    def update(self):
        pass
