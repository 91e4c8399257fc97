"""Hybrid MPC controller with the reference's public surface (controller.py:13-843) on top of the
CUDA kernels.

Drop-in seams kept:
  * ``HybridModelPredictiveController(mld, T, objective, terminal_set)``
  * ``feedforward(x0, gurobi_params={}, search_rule=best_first, branch_rule=branch_in_time, tol=0.,
    warm_start=None, printing_period=3., draw_label=None) -> (PrimalSolution|None, leaves, n_qp, solver_time)``
  * ``construct_warm_start(leaves, x0, uc0, ub0, e0) -> (nodes, t_runtime, t_interstep)``
  * the per-node seam ``_solve_subproblem(identifier, x0) -> (SubproblemSolution, solve_time)``
Errors follow the reference: ValueError on inconsistent sizes; an infeasible node is not an error
(``primal.objective = inf``, Farkas multipliers in ``dual``).

Every QP relaxation is solved on the GPU (kernel K1 through the C ABI); host code here is the
reference's Python control flow.  The all-on-device batched search is ``feedforward_batch``.
"""
import gc
import numpy as np
from time import time

from .branch_and_bound import Node, branch_and_bound, best_first, depth_first, breadth_first  # noqa: F401
from .subproblem_solution import SubproblemSolution, PrimalSolution, DualSolution
from .problem import ProblemData


def _is_prefix(identifier, nub):
    """True iff the identifier pins exactly the first len(identifier) binaries in chronological order."""
    return all((q // nub, q % nub) in identifier for q in range(len(identifier)))


def branch_in_time(identifier, nub):
    """controller.py:13-44: fix the binaries in chronological order, children [value 0, value 1]."""
    t = max([k[0] for k in identifier.keys()] + [0])
    index = max([k[1] + 1 for k in identifier.keys() if k[0] == t] + [0])
    if index < nub:
        return [{(t, index): 0.}, {(t, index): 1.}]
    return [{(t + 1, 0): 0.}, {(t + 1, 0): 1.}]


class HybridModelPredictiveController(object):

    def __init__(self, mld, T, objective, terminal_set, device=0, qp_options=None):
        self.mld = mld
        self.T = T
        self.Q, self.R, self.Q_T = [np.asarray(M, dtype=float) for M in objective]
        # terminal constraint folded into the stage rows of time T-1 (controller.py:81-87)
        if terminal_set is None:
            terminal_set = [np.empty((0, self.mld.nx)), np.empty(0)]
        self.F_Tm1 = np.vstack((mld.F, terminal_set[0].dot(mld.A)))
        self.G_Tm1 = np.vstack((mld.G, terminal_set[0].dot(mld.B)))
        self.h_Tm1 = np.concatenate((mld.h, terminal_set[1]))
        self._check_input_sizes()
        # warm start construction (controller.py:94-97)
        self._update = {'mu': self._update_mu(), 'rho': np.linalg.pinv(self.Q.T).dot(self.Q_T.T)}
        # problem compiler (replaces _build_mip, controller.py:119-184)
        self.problem = ProblemData.from_controller_data(
            mld, T, [self.Q, self.R, self.Q_T], self.F_Tm1, self.G_Tm1, self.h_Tm1,
            self._update['mu'], self._update['rho'], **(qp_options or {}))
        self.device = device
        self._handle = None
        self._n_slots = 1
        self.device_search = True
        self.max_solves = 4096

    # -- construction ---------------------------------------------------------------------------
    def _check_input_sizes(self):
        """controller.py:99-117."""
        if self.Q.shape[1] != self.mld.nx:
            raise ValueError('Matrix Q has wrong number of columns.')
        if self.R.shape[1] != self.mld.nu:
            raise ValueError('Matrix R has wrong number of columns.')
        if self.Q_T.shape[1] != self.mld.nx:
            raise ValueError('Matrix Q_T has wrong number of columns.')
        if self.F_Tm1.shape[0] != self.h_Tm1.size:
            raise ValueError('Terminal-set matrices have wrong number of rows.')
        if self.G_Tm1.shape[0] != self.h_Tm1.size:
            raise ValueError('Terminal-set matrices have wrong number of rows.')

    def _update_mu(self):
        """controller.py:186-227: column i of M solves  min h.mu  s.t.  F'mu = F_Tm1[i], G'mu = G_Tm1[i],
        mu >= 0  (host precompute, once per controller; the reference routes these LPs through Gurobi,
        here HiGHS via scipy)."""
        from scipy.optimize import linprog
        mld = self.mld
        n = mld.h.size
        Aeq = np.vstack((mld.F.T, mld.G.T))
        cols = []
        for i in range(self.h_Tm1.size):
            if i < n and np.array_equal(self.F_Tm1[i], mld.F[i]) and np.array_equal(self.G_Tm1[i], mld.G[i]):
                # the first n rows of [F_Tm1 G_Tm1] ARE the rows of [F G]: unit vector unless the LP finds
                # a strictly cheaper certificate
                ei = np.zeros(n); ei[i] = 1.
                res = linprog(mld.h, A_eq=Aeq, b_eq=np.concatenate((self.F_Tm1[i], self.G_Tm1[i])),
                              bounds=[(0, None)] * n, method='highs')
                if res.status == 0 and res.fun < mld.h[i] - 1e-9 * max(1., abs(mld.h[i])):
                    cols.append(res.x)
                else:
                    cols.append(ei)
                continue
            res = linprog(mld.h, A_eq=Aeq, b_eq=np.concatenate((self.F_Tm1[i], self.G_Tm1[i])),
                          bounds=[(0, None)] * n, method='highs')
            if res.status != 0:
                raise ValueError('The conic hull of [F G] does not contain the one of [F_Tm1 G_Tm1].')
            cols.append(res.x)
        return np.vstack(cols).T if cols else np.zeros((n, 0))

    # -- device handle --------------------------------------------------------------------------
    @staticmethod
    def _require_gpu():
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('no CUDA device: the B&B hot path has no CPU fallback')

    def handle(self, n_slots=None):
        from .capi import Handle
        self._require_gpu()
        if n_slots is not None and (self._handle is None or n_slots > self._handle.n_slots):
            if self._handle is not None:
                self._handle.close()
            self._handle = None
            self._n_slots = n_slots
        if self._handle is None:
            self._handle = Handle(self.problem, device=self.device, n_slots=self._n_slots)
        return self._handle

    # -- per-node seam --------------------------------------------------------------------------
    def _get_bound_binaries(self, identifier):
        """controller.py:300-327."""
        ub_lb = np.zeros((self.T, self.mld.nub))
        ub_ub = np.ones((self.T, self.mld.nub))
        for k, v in identifier.items():
            ub_lb[k] = v
            ub_ub[k] = v
        return ub_lb, ub_ub

    def _start_from(self, extra):
        """Signed multipliers (and proximal centre) a node starts its dual active-set solve from: the dual
        solution it carries -- its parent's (controller.py:426) or its own shifted one (controller.py:487) --
        and the parent's `active_set` (controller.py:262-264).  None = empty working set."""
        if extra is None or extra.dual is None:
            return None, None
        v = extra.dual.variables
        y0 = np.concatenate([np.maximum(m_, 0.) for m_ in v['mu']]
                            + [np.maximum(np.concatenate(v['nu_ub']), 0.) - np.maximum(np.concatenate(v['nu_lb']), 0.)])
        return y0, extra.active_set

    def _solve_subproblem(self, identifier, x0, active_set=None, hot=True, extra=None):
        """controller.py:229-271: one node = one K1 launch on slot 0.  `extra` (the SubproblemSolution the node
        carries) gives the start of the solve, see _start_from; without it `hot` keeps the working set of the
        previously solved node (any multipliers >= 0 are dual feasible for every node)."""
        import torch
        h = self.handle()
        lb, ub = self._get_bound_binaries(identifier)
        y0, yc0 = self._start_from(extra)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        if extra is not None:
            mode = 2 if y0 is not None else 0
            out = h.solve_nodes(np.asarray(x0, dtype=float)[None], lb.reshape(1, -1), ub.reshape(1, -1),
                                slot=np.zeros(1, np.int32), hot=np.array([mode], np.int32),
                                y0=None if y0 is None else y0[None], yc0=None if yc0 is None else np.asarray(yc0)[None])
        else:
            out = h.solve_nodes(np.asarray(x0, dtype=float)[None], lb.reshape(1, -1), ub.reshape(1, -1),
                                slot=np.zeros(1, np.int32), hot=np.array([1 if hot else 0], np.int32))
        end.record()
        end.synchronize()
        solve_time = start.elapsed_time(end) * 1e-3
        status = int(out['status'][0])
        if status not in (2, 3):
            raise RuntimeError('QP kernel did not converge (status %d) for identifier %r' % (status, identifier))
        # subproblem_solution.py:94-97: binary feasible iff EVERY binary is pinned by the identifier
        binary_feasible = bool(np.array_equal(lb, ub))
        primal = PrimalSolution.from_record(self.problem, out['primal'][0].cpu().numpy(), float(out['cost'][0]),
                                            binary_feasible, status == 2)
        dual = DualSolution.from_record(self.problem, h.layout, out['dual'][0].cpu().numpy(), float(out['dobj'][0]))
        # active_set = proximal centre of this solve: what the children start from (controller.py:426)
        sol = SubproblemSolution(primal, dual, out['yc'][0].cpu().numpy() if status == 2 else None)
        sol.iters = int(out['iters'][0])
        return sol, solve_time

    # -- per-solve seam -------------------------------------------------------------------------
    def feedforward(self, x0, gurobi_params={}, search_rule=best_first, branch_rule=branch_in_time, **kwargs):
        """controller.py:329-393.  `gurobi_params` is accepted for signature compatibility and ignored
        (there is no Gurobi underneath).  With the reference's search rules (best_first, depth_first, breadth_first;
        branch_and_bound.py:501-563), branch_in_time and `device_search=True` the whole search runs in the device-side
        B&B kernel (K3); a user-defined rule, or `device_search=False`, runs the reference's host loop with one K1
        launch per node -- both visit the same nodes in the same order and return bit-identical results."""
        rules = {best_first: 0, depth_first: 1, breadth_first: 2}
        if self.device_search and search_rule in rules and branch_rule is branch_in_time:
            ws = kwargs.get('warm_start')
            if ws is None or all(_is_prefix(l.identifier, self.mld.nub) for l in ws):
                return self._feedforward_device(x0, kwargs.get('tol', 0.), ws, rules[search_rule])
        def solver(identifier, cutoff, extra):
            solution, solve_time = self._solve_subproblem(identifier, x0, None, extra=extra if extra is not None
                                                          else SubproblemSolution(None, None, None))
            return solution.primal.objective, solution.primal.binary_feasible, solve_time, solution

        def brancher(parent):
            return self._brancher(parent, branch_rule)

        incumbent, leaves, qp_solves, solver_time = branch_and_bound(solver, search_rule, brancher, **kwargs)
        if incumbent is None:
            return None, leaves, qp_solves, solver_time
        return incumbent.extra.primal, leaves, qp_solves, solver_time

    def _brancher(self, parent, branch_rule):
        """controller.py:395-429: child bound = parent bound + multiplier of the bound that moves."""
        branches = branch_rule(parent.identifier, self.mld.nub)
        children = []
        for branch in branches:
            lb = parent.lb
            for k, v in branch.items():
                nu = 'nu_lb' if v == 1 else 'nu_ub' if v == 0 else None
                lb += parent.extra.dual.variables[nu][k[0]][k[1]]
            identifier = {**parent.identifier, **branch}
            solution = SubproblemSolution(None, parent.extra.dual, parent.extra.active_set)
            children.append(Node(identifier, lb, solution))
        return children

    # -- warm start (host restatement; the batched device version is K4) ------------------------------
    def _construct_warm_start_interstep(self, leaves, x0, uc0, ub0):
        """controller.py:431-501."""
        u0 = np.concatenate((uc0, ub0))
        gc.disable()
        construction_time = time()
        warm_start = []
        for leaf in leaves:
            if self._retain_leaf(leaf.identifier, ub0):
                shifted_identifier = {(k[0] - 1, k[1]): v for k, v in leaf.identifier.items() if k[0] > 0}
                shifted_variables = self._shift_dual_variables(leaf.extra.dual.variables)
                pi_sum = self._pi_sum(leaf.identifier, leaf.extra.dual.variables, shifted_variables, x0, u0)
                shifted_dual = DualSolution(shifted_variables, leaf.extra.dual.objective + pi_sum)
                warm_start.append(Node(shifted_identifier, leaf.lb, SubproblemSolution(None, shifted_dual)))
        construction_time = time() - construction_time
        gc.enable()
        return warm_start, construction_time

    def construct_warm_start(self, leaves, x0, uc0, ub0, e0):
        """controller.py:503-564."""
        warm_start, interstep_time = self._construct_warm_start_interstep(leaves, x0, uc0, ub0)
        gc.disable()
        construction_time = time()
        for leaf in warm_start:
            pi3 = - leaf.extra.dual.variables['lam'][0].dot(e0)
            leaf.extra.dual.objective += pi3
            leaf.extra.dual.objective = max(leaf.extra.dual.objective, 0)
            if not np.isinf(leaf.lb):
                leaf.lb = leaf.extra.dual.objective
            else:
                if leaf.extra.dual.objective <= 0.:
                    leaf.lb = 0.
                    leaf.extra.dual = None
        construction_time = time() - construction_time
        gc.enable()
        return warm_start, construction_time, interstep_time

    @staticmethod
    def _retain_leaf(identifier, ub0):
        """controller.py:615-633."""
        return all(v == ub0[k[1]] for k, v in identifier.items() if k[0] == 0)

    def _shift_dual_variables(self, variables):
        """controller.py:635-666."""
        shifted = {}
        for k in ['lam', 'nu_lb', 'nu_ub', 'sigma']:
            shifted[k] = variables[k][1:]
            shifted[k].append(np.zeros(variables[k][-1].shape))
        for k in ['mu', 'rho']:
            shifted[k] = variables[k][1:-1]
            shifted[k].append(self._update[k].dot(variables[k][-1]))
            shifted[k].append(np.zeros(variables[k][-1].shape))
        return shifted

    def _pi_sum(self, identifier, variables, shifted_variables, x0, u0):
        """controller.py:668-721."""
        squared = lambda x: x.dot(x)
        Qx0 = self.Q.dot(x0)
        Ru0 = self.R.dot(u0)
        pi_sum = - squared(Qx0) - squared(Ru0)
        pi_sum += squared(.5 * variables['rho'][0] - Qx0) + squared(.5 * variables['sigma'][0] - Ru0)
        ub_lb, ub_ub = self._get_bound_binaries(identifier)
        residuals = {
            'mu': self.mld.F.dot(x0) + self.mld.G.dot(u0) - self.mld.h,
            'nu_lb': ub_lb[0] - self.mld.V.dot(u0),
            'nu_ub': self.mld.V.dot(u0) - ub_ub[0],
        }
        pi_sum -= sum(residual.dot(variables[k][0]) for k, residual in residuals.items())
        pi_sum += .25 * squared(variables['rho'][self.T]) - .25 * squared(shifted_variables['rho'][self.T - 1])
        pi_sum += self.h_Tm1.dot(variables['mu'][self.T - 1]) - self.mld.h.dot(shifted_variables['mu'][self.T - 2])
        return pi_sum

    # -- device-resident batch path (K3, K2 + K4) ----------------------------------------------------
    def _as_device(self, a, shape):
        import torch
        t = torch.as_tensor(np.asarray(a, dtype=float) if not torch.is_tensor(a) else a, dtype=torch.float64,
                            device=torch.device('cuda', self.device)).contiguous()
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))
        return t

    def new_tree(self, n_inst, n_roots=1, max_solves=None):
        """Device tree with room for `n_roots` initial leaves and `max_solves` QP solves per instance."""
        ms = self.max_solves if max_solves is None else max_solves
        return self.handle().new_tree(n_inst, n_roots + 2 * ms + 2, n_roots + ms + 1)

    def feedforward_batch(self, x0, warm_start=None, tol=0., max_solves=None, n_slots=None, active=None, trace=False,
                          out=None):
        """Batched feedforward (controller.py:329-393 for many independent initial states at once).
        x0 [N, nx] (numpy or CUDA tensor); warm_start: a device `Tree` (from construct_warm_start_batch)
        or None for the cold root.  Returns (result dict of CUDA tensors, tree); nothing is synchronised."""
        self._require_gpu()
        ms = self.max_solves if max_solves is None else max_solves
        N = x0.shape[0]
        h = self.handle(n_slots if n_slots is not None else max(self._n_slots, min(N, self.default_slots())))
        x0 = self._as_device(x0, (N, self.mld.nx))
        tree = warm_start
        if tree is None:
            tree = self.new_tree(N, 1, ms)
            h.tree_init_root(tree)
        res = h.bnb_solve(x0, tree, tol=tol, max_solves=ms, active=active, out=out, trace=trace)
        res['x0'] = x0
        return res, tree

    def construct_warm_start_batch(self, res, tree, e0=None, new_tree=None, active=None, x_next=None, u0=None,
                                   max_solves=None):
        """Batched construct_warm_start (controller.py:503-564) + plant update x <- x_1 + e0.
        Returns (new_tree, x_next [N, nx], u0 [N, nu])."""
        import torch
        h = self.handle()
        N = tree.n_inst
        dev = torch.device('cuda', self.device)
        if new_tree is None:
            ms = self.max_solves if max_solves is None else max_solves
            new_tree = h.new_tree(N, tree.cap_nodes + 2 * ms + 2, tree.cap_nodes + ms + 1)
        if e0 is not None:
            e0 = self._as_device(e0, (N, self.mld.nx))
        if x_next is None:
            x_next = torch.empty((N, self.mld.nx), dtype=torch.float64, device=dev)
        if u0 is None:
            u0 = torch.empty((N, self.mld.nu), dtype=torch.float64, device=dev)
        h.shift_tree(res['x0'], e0, tree, res['cost'], res['primal'], new_tree, active=active, x_next=x_next, u0=u0)
        return new_tree, x_next, u0

    @staticmethod
    def default_slots():
        """Solver states resident at once: SMs (148 on a B200) x solver CTAs per SM of the library build."""
        import torch
        from .capi import load_library
        return torch.cuda.get_device_properties(0).multi_processor_count * int(load_library().wshmpc_ctas_per_sm())

    # -- tree <-> reference Node lists ---------------------------------------------------------------
    def tree_to_leaves(self, tree, inst=0):
        """The reference's `leaves` list (branch_and_bound.py:497) of one instance of a device tree."""
        h = self.handle()
        T, nub = self.T, self.mld.nub
        nn = int(tree.n_nodes[inst])
        depth = tree.depth[inst, :nn].cpu().numpy(); alive = tree.alive[inst, :nn].cpu().numpy()
        rec = tree.rec[inst, :nn].cpu().numpy(); lb = tree.lb[inst, :nn].cpu().numpy()
        bits = tree.bits[inst, :nn].cpu().numpy().view(np.uint32)
        nr = int(tree.n_recs[inst])
        duals = tree.rec_dual[inst, :nr].cpu().numpy(); dobj = tree.rec_dobj[inst, :nr].cpu().numpy()
        cache = {}
        leaves = []
        for j in range(nn):
            if not alive[j]:
                continue
            ident = {(q // nub, q % nub): float((bits[j, q >> 5] >> (q & 31)) & 1) for q in range(depth[j])}
            r = int(rec[j])
            if r < 0:
                extra = SubproblemSolution(None, None) if depth[j] or np.isfinite(lb[j]) else None
            else:
                if r not in cache:          # children alias the parent's dual object (controller.py:426)
                    cache[r] = (DualSolution.from_record(self.problem, h.layout, duals[r], dobj[r]),
                                duals[r][h.layout.dual:h.layout.rec_stride].copy())
                extra = SubproblemSolution(None, cache[r][0], cache[r][1])
            node = Node(ident, float(lb[j]), extra)
            node.index = j
            leaves.append(node)
        return leaves

    def leaves_to_tree(self, leaves, max_solves=None):
        """Uploads a reference-style warm start (list of Node with prefix identifiers) as a device tree."""
        import torch
        h = self.handle()
        ms = self.max_solves if max_solves is None else max_solves
        n0 = len(leaves)
        tree = self.new_tree(1, n0, ms)
        nub = self.mld.nub
        depth = np.zeros(n0, np.int32); rec = np.full(n0, -1, np.int32); lb = np.zeros(n0)
        bits = np.zeros((n0, tree.words), np.uint32)
        recs, dobj, seen = [], [], {}
        for j, l in enumerate(leaves):
            depth[j] = len(l.identifier)
            for (t, i), v in l.identifier.items():
                if v:
                    q = t * nub + i
                    bits[j, q >> 5] |= np.uint32(1 << (q & 31))
            lb[j] = l.lb
            dual = None if l.extra is None else l.extra.dual
            if dual is not None:
                if id(dual) not in seen:
                    seen[id(dual)] = len(recs)
                    yc = np.zeros(h.layout.rec_stride - h.layout.dual) if l.extra.active_set is None else np.asarray(l.extra.active_set)
                    recs.append(np.concatenate((DualSolution.to_record(self.problem, h.layout, dual.variables), yc)))
                    dobj.append(dual.objective)
                rec[j] = seen[id(dual)]
        dev = tree.lb.device
        tree.n_nodes[0] = n0; tree.n_recs[0] = len(recs)
        tree.depth[0, :n0] = torch.as_tensor(depth, device=dev); tree.alive[0, :n0] = 1
        tree.rec[0, :n0] = torch.as_tensor(rec, device=dev); tree.lb[0, :n0] = torch.as_tensor(lb, device=dev)
        tree.bits[0, :n0] = torch.as_tensor(bits.view(np.int32), device=dev)
        if recs:
            tree.rec_dual[0, :len(recs)] = torch.as_tensor(np.vstack(recs), device=dev)
            tree.rec_dobj[0, :len(recs)] = torch.as_tensor(np.array(dobj), device=dev)
        return tree

    def _feedforward_device(self, x0, tol, warm_start, rule=0):
        import torch
        self._require_gpu()
        tree = None if warm_start is None else self.leaves_to_tree(warm_start)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        self.handle().set_search_rule(rule)
        try:
            res, tree = self.feedforward_batch(np.asarray(x0, dtype=float)[None], warm_start=tree, tol=tol, n_slots=1)
        finally:
            self.handle().set_search_rule(0)
        end.record()
        end.synchronize()
        solver_time = start.elapsed_time(end) * 1e-3
        status = int(res['status'][0])
        if status >= 2:
            raise RuntimeError('device branch and bound failed with status %d (2 = capacity, 3 = QP iteration limit)' % status)
        leaves = self.tree_to_leaves(tree, 0)
        if warm_start is not None:          # the reference mutates the list it is given (branch_and_bound.py:432)
            warm_start[:] = leaves
            leaves = warm_start
        self.last_batch = (res, tree)
        n_qp = int(res['n_solves'][0])
        if status == 1:
            return None, leaves, n_qp, solver_time
        primal = PrimalSolution.from_record(self.problem, res['primal'][0].cpu().numpy(), float(res['cost'][0]), True, True)
        return primal, leaves, n_qp, solver_time

    def shift_binary_solution(self, ub):
        """controller.py:811-812."""
        return np.vstack((ub[1:], np.zeros(self.mld.nub)))
