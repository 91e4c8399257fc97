"""TEST INFRASTRUCTURE (oracle side) -- CPU restatement of the whole hot path above the node QP.

/root/reference does not exist on the GPU box, so the parts of the reference that sit above the
QP solve are restated here in plain Python (each function cites the reference lines it follows) and
pinned, in this container, against the unmodified reference code running on the same QP core
(tests/test_oracle_cpu.py::test_bnb_restatement_equals_reference) and against tests/golden/:

* `branch_and_bound`      branch_and_bound.py:408-499 with best_first (:541-563)
* `OracleController`      controller.py:229-298 (_solve_subproblem), :329-393 (feedforward),
                          :395-429 (_brancher + branch_in_time :13-44), :431-564 (construct_warm_start),
                          :615-721 (_retain_leaf, _shift_dual_variables, _pi_sum)
* `closed_loop`           notebooks/cart_pole_with_walls/statistical_analysis.py:93-196 (no Gurobi legs)

The node QP itself is oracle/qp_core.c through oracle/qp_c.CoreC.  This module is what bench.py times as
the CPU baseline (`cpu_baseline.kind = "port"`, and the `--impl reference` arm on the GPU box).
Only tests/, __graft_entry__.smoke() and bench.py may import it.
"""
import time
import numpy as np

from oracle.condense import Condensed


class Node(object):
    """branch_and_bound.py:7-55."""

    def __init__(self, identifier, lb=-np.inf, dual=None):
        self.identifier = identifier
        self.lb = lb
        self.dual = dual            # dict(variables=..., objective=...) or None
        self.primal = None
        self.binary_feasible = None


def branch_in_time(identifier, nub):
    """controller.py:13-44."""
    t = max([k[0] for k in identifier.keys()] + [0])
    index = max([k[1] + 1 for k in identifier.keys() if k[0] == t] + [0])
    if index < nub:
        return [{(t, index): 0.}, {(t, index): 1.}]
    return [{(t + 1, 0): 0.}, {(t + 1, 0): 1.}]


def branch_and_bound(solve, branch, tol=0., warm_start=None, on_solve=None):
    """branch_and_bound.py:408-499 with candidate_selection = best_first (:541-563)."""
    ub = np.inf
    incumbent = None
    leaves = [Node({})] if warm_start is None else warm_start
    solves = 0
    while True:
        candidates = [l for l in leaves if l.lb < ub - tol]
        if not candidates:
            break
        node = candidates[int(np.argmin([l.lb for l in candidates]))]
        cutoff = ub - tol
        solve(node)
        solves += 1
        if on_solve is not None:
            on_solve(node)
        if node.lb >= cutoff:
            pass
        elif node.binary_feasible:
            incumbent = node
            ub = node.lb
        else:
            children = branch(node)
            leaves.remove(node)
            leaves.extend(children)
    return incumbent, leaves, solves


class OracleController(object):

    def __init__(self, model, core, hot_start=True, persistent=False):
        """model: dict of oracle/models.py; core: oracle.qp_c.CoreC (or any object with
        solve(x0, lb, ub, warm=None)).  `hot_start`: each node starts from the working set of the node
        solved before it; 'record' = each node starts from the multipliers of the dual solution IT carries (its
        parent's, or its own shifted one for a warm-start root) and from its parent's proximal centre -- the
        reference's dormant `active_set` hand-over (controller.py:262-264, 426) and what the CUDA path does;
        False = every node from scratch.
        `persistent`: the factorisation itself survives between the nodes of one feedforward call (exactly
        what a CUDA solver slot does) instead of being rebuilt from the inherited working set."""
        self.m = model
        self.core = core
        self.hot_start = hot_start
        self.c = Condensed(model)
        self.T, self.nub = int(model['T']), int(model['nub'])
        self.nx, self.nu = model['B'].shape
        self.nuc = self.nu - self.nub
        self.V = np.hstack((np.zeros((self.nub, self.nuc)), np.eye(self.nub)))
        self.qp_time = 0.
        self._warm = None
        self._state = core.new_state() if persistent else None
        self._reset = True

    # -- controller.py:300-327
    def bounds(self, identifier):
        lb = np.zeros((self.T, self.nub)); ub = np.ones((self.T, self.nub))
        for k, v in identifier.items():
            lb[k] = v; ub[k] = v
        return lb, ub

    # -- controller.py:229-271 + subproblem_solution.py:68-99, 119-168
    def solve_node(self, node, x0):
        m, c, T = self.m, self.c, self.T
        lb, ub = self.bounds(node.identifier)
        tic = time.perf_counter()
        if self._state is not None:
            out = self.core.solve(x0, lb.ravel(), ub.ravel(), state=self._state, reset=self._reset)
            self._reset = False
        elif self.hot_start == 'record':
            warm = None
            if node.dual is not None:
                var = node.dual['variables']
                rows, sides, lam = [], [], []
                for t in range(T):
                    for i in np.nonzero(var['mu'][t] > 0)[0]:
                        rows.append(c.row0[t] + i); sides.append(1); lam.append(var['mu'][t][i])
                for t in range(T):
                    for i in range(self.nub):
                        yy = max(var['nu_ub'][t][i], 0.) - max(var['nu_lb'][t][i], 0.)
                        if yy != 0.:
                            rows.append(c.mc + t * self.nub + i); sides.append(1 if yy > 0 else -1); lam.append(abs(yy))
                warm = dict(rows=rows, sides=sides, lam=lam, z=node.dual.get('yc'))
            out = self.core.solve(x0, lb.ravel(), ub.ravel(), warm=warm)
            if out['status'] not in (2, 3) and warm is not None:
                # a start from a stale, nearly dependent working set can degenerate: once more from the empty working
                # set (the CUDA solver does the same, qp_device.cuh qp_solve `restart`)
                out = self.core.solve(x0, lb.ravel(), ub.ravel(), warm=None)
        else:
            out = self.core.solve(x0, lb.ravel(), ub.ravel(), warm=self._warm if self.hot_start else None)
        self.qp_time += time.perf_counter() - tic
        if out['status'] not in (2, 3):
            raise RuntimeError('oracle QP core failed with status %r' % out['status'])
        self._warm = out['warm']
        y = out['y']
        mu = [np.maximum(y[c.row0[t]:c.row0[t] + (c.nh if t < T - 1 else c.nh1)], 0.) for t in range(T)]
        yb = y[c.mc:].reshape(T, self.nub)
        var = dict(mu=mu, nu_ub=list(np.maximum(yb, 0.)), nu_lb=list(np.maximum(-yb, 0.)))
        Q, R, Q_T, A = m['Q'], m['R'], m['Q_T'], m['A']
        if out['status'] == 2:
            x = c.states(x0, out['z']); u = out['z'].reshape(T, self.nu)
            node.primal = dict(x=x, u=u, objective=out['cost'])
            var['rho'] = [2. * Q.dot(x[t]) for t in range(T)] + [2. * Q_T.dot(x[T])]
            var['sigma'] = [2. * R.dot(u[t]) for t in range(T)]
            objective = out['cost']
            node.lb = out['cost']
        else:
            node.primal = None
            var['rho'] = [np.zeros(Q.shape[0]) for _ in range(T)] + [np.zeros(Q_T.shape[0])]
            var['sigma'] = [np.zeros(R.shape[0]) for _ in range(T)]
            objective = out['farkas']
            node.lb = np.inf
        lam = [None] * (T + 1)
        lam[T] = -Q_T.T.dot(var['rho'][T])
        for t in range(T - 1, -1, -1):
            Ft = m['F'] if t < T - 1 else m['F_Tm1']
            lam[t] = A.T.dot(lam[t + 1]) - Q.T.dot(var['rho'][t]) - Ft.T.dot(mu[t])
        var['lam'] = lam
        node.dual = dict(variables=var, objective=objective, yc=out['warm']['z'])      # yc: proximal centre (None if infeasible)
        node.binary_feasible = bool(np.array_equal(lb, ub))        # subproblem_solution.py:94-97

    # -- controller.py:395-429
    def branch(self, parent):
        children = []
        for br in branch_in_time(parent.identifier, self.nub):
            lb = parent.lb
            for k, v in br.items():
                lb += parent.dual['variables']['nu_lb' if v == 1 else 'nu_ub'][k[0]][k[1]]
            children.append(Node({**parent.identifier, **br}, lb, parent.dual))
        return children

    # -- controller.py:329-393
    def feedforward(self, x0, warm_start=None, tol=0., on_solve=None):
        self._warm = None
        self._reset = True
        inc, leaves, solves = branch_and_bound(lambda n: self.solve_node(n, x0), self.branch, tol, warm_start, on_solve)
        return inc, leaves, solves

    # -- controller.py:635-666
    def shift_dual(self, var):
        m = self.m
        sh = {}
        for k in ('lam', 'nu_lb', 'nu_ub', 'sigma'):
            sh[k] = list(var[k][1:]) + [np.zeros(var[k][-1].shape)]
        sh['mu'] = list(var['mu'][1:-1]) + [m['M_mu'].dot(var['mu'][-1]), np.zeros(var['mu'][-1].shape)]
        sh['rho'] = list(var['rho'][1:-1]) + [m['M_rho'].dot(var['rho'][-1]), np.zeros(var['rho'][-1].shape)]
        return sh

    # -- controller.py:668-721
    def pi_sum(self, identifier, var, sh, x0, u0):
        m, T = self.m, self.T
        sq = lambda v: v.dot(v)
        Qx0, Ru0 = m['Q'].dot(x0), m['R'].dot(u0)
        pi = -sq(Qx0) - sq(Ru0)
        pi += sq(.5 * var['rho'][0] - Qx0) + sq(.5 * var['sigma'][0] - Ru0)
        lb, ub = self.bounds(identifier)
        pi -= (m['F'].dot(x0) + m['G'].dot(u0) - m['h']).dot(var['mu'][0])
        pi -= (lb[0] - self.V.dot(u0)).dot(var['nu_lb'][0])
        pi -= (self.V.dot(u0) - ub[0]).dot(var['nu_ub'][0])
        pi += .25 * sq(var['rho'][T]) - .25 * sq(sh['rho'][T - 1])
        pi += m['h_Tm1'].dot(var['mu'][T - 1]) - m['h'].dot(sh['mu'][T - 2])
        return pi

    # -- controller.py:431-564, 615-633
    def construct_warm_start(self, leaves, x0, uc0, ub0, e0):
        u0 = np.concatenate((uc0, ub0))
        ws = []
        for leaf in leaves:
            if not all(v == ub0[k[1]] for k, v in leaf.identifier.items() if k[0] == 0):
                continue
            ident = {(k[0] - 1, k[1]): v for k, v in leaf.identifier.items() if k[0] > 0}
            var = leaf.dual['variables']
            sh = self.shift_dual(var)
            obj = leaf.dual['objective'] + self.pi_sum(leaf.identifier, var, sh, x0, u0)
            obj += -sh['lam'][0].dot(e0)
            obj = max(obj, 0)
            node = Node(ident, leaf.lb, dict(variables=sh, objective=obj))
            if not np.isinf(leaf.lb):
                node.lb = obj
            elif obj <= 0.:
                node.lb = 0.
                node.dual = None
            ws.append(node)
        return ws


def closed_loop(ctl, x0, n_steps, e=None, warm=True):
    """statistical_analysis.py:93-196 for one instance: returns per-step dicts."""
    x = np.array(x0, dtype=float)
    ws = None
    log = []
    for t in range(n_steps):
        inc, leaves, solves = ctl.feedforward(x, warm_start=ws if warm else None)
        if inc is None:
            log.append(dict(x=x.copy(), cost=np.inf, solves=solves))
            break
        u0 = inc.primal['u'][0]
        et = np.zeros(x.size) if e is None else e[t]
        if warm:
            ws = ctl.construct_warm_start(leaves, x, u0[:ctl.nuc], u0[ctl.nuc:], et)
        log.append(dict(x=x.copy(), cost=inc.primal['objective'], solves=solves, u0=u0.copy(),
                        ub=inc.primal['u'][:, ctl.nuc:].copy(), cover=len(ws) if warm else 0))
        x = inc.primal['x'][1] + et
    return log
