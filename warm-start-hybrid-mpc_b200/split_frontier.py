"""Split-frontier branch and bound: the frontier nodes of ONE MIQP solved on several GPUs, with the global incumbent
kept by a min-allreduce (north_star item 4: "NCCL over NVLink carries only the global incumbent min-allreduce and
periodic frontier load-balancing").

The reference's loop (branch_and_bound.py:408-499) solves one node per iteration because every decision depends on
the incumbent `ub` at that moment.  Here a ROUND takes the R = world x nodes_per_rank best candidates of the
frontier (best_first order: lower bound, ties by creation order), deals them to the ranks round robin, every rank
solves its share in one K1 launch (wshmpc_solve_nodes, each node started from the multipliers of the dual solution it
carries, controller.py:426), and then

  * ONE all-reduce(min) of the best binary-feasible cost of the round gives every rank the new incumbent value, and
  * ONE all-gather of the solved nodes' records (cost, status, dual solution, proximal centre: ~12 KB per node on the
    T = 20 cart-pole) keeps the frontier REPLICATED: every rank then applies the reference's prune / incumbent / branch
    rules to the round's nodes in best_first order, so all ranks hold the same leaves and the next round needs no
    further communication.  Child bounds are the reference's (parent bound + multiplier of the moved bound, :395-429).

The search explores a superset of the reference's nodes (the extra ones are the speculative part of a round: nodes whose
bound the round's own incumbent would have pruned); optimal cost, mode sequence and the leaves' validity as a cover are
unchanged, so the result can be time-shifted like any other (`construct_warm_start`).  How much it helps depends on the
frontier width: the cart-pole trees are caterpillars (2.5 candidates on average), so a round rarely has more
than two useful nodes -- DESIGN.md section 6 reports the measured latencies.

`solve` is pluggable (tests drive the same code on CPU with gloo and the oracle QP core).
"""
import numpy as np

from .branch_and_bound import Node
from .subproblem_solution import SubproblemSolution, PrimalSolution, DualSolution


class GpuBatchSolver(object):
    """solve(x0, nodes) -> records: one K1 launch for a list of nodes of the same MIQP."""

    def __init__(self, controller, n_slots=8):
        self.ctl = controller
        self.h = controller.handle(n_slots)

    def __call__(self, x0, nodes):
        import torch
        ctl, pd = self.ctl, self.ctl.problem
        L = pd.layout
        N = len(nodes)
        width = 3 + L.dual + pd.n + L.primal
        if N == 0:
            return torch.zeros((0, width), dtype=torch.float64, device=self.h.torch_device)
        lb = np.zeros((N, pd.nb)); ub = np.ones((N, pd.nb)); y0 = np.zeros((N, pd.m)); yc0 = np.zeros((N, pd.n))
        hot = np.zeros(N, np.int32)
        for k, node in enumerate(nodes):
            l, u = ctl._get_bound_binaries(node.identifier)
            lb[k], ub[k] = l.ravel(), u.ravel()
            start = None if node.extra is None else node.extra.active_set
            if start is not None:
                y0[k] = ctl.qp._signed_multipliers(np.asarray(start['c'], dtype=float))
                yc0[k] = np.asarray(start['v'], dtype=float)[:pd.n]
                hot[k] = 2
        out = self.h.solve_nodes(np.repeat(np.asarray(x0, dtype=float)[None], N, 0), lb, ub,
                                 slot=np.arange(N, dtype=np.int32) % self.h.n_slots, hot=hot, y0=y0, yc0=yc0)
        rec = torch.cat((out['status'].double()[:, None], out['cost'][:, None], out['dobj'][:, None], out['dual'], out['yc'],
                         out['primal']), dim=1)
        return rec


def split_frontier_bnb(controller, x0, solve, group=None, nodes_per_rank=1, tol=0., warm_start=None, max_rounds=100000):
    """Returns (PrimalSolution or None, leaves, qp_solves of the whole job, rounds).  All ranks return the same thing."""
    import torch
    import torch.distributed as dist
    ctl, pd = controller, controller.problem
    L = pd.layout
    dist_on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if dist_on else 0
    world = dist.get_world_size(group) if dist_on else 1
    R = world * nodes_per_rank
    width = 3 + L.dual + pd.n + L.primal
    nb = pd.nb
    leaves = [Node({})] if warm_start is None else list(warm_start)
    order = {id(l): i for i, l in enumerate(leaves)}          # creation order (ties of best_first)
    created = len(leaves)
    ub = np.inf
    incumbent = None
    solves = rounds = 0
    while rounds < max_rounds:
        cands = [l for l in leaves if l.lb < ub - tol]
        if not cands:
            break
        cands.sort(key=lambda l: (l.lb, order[id(l)]))
        batch = cands[:R]
        mine = batch[rank::world]
        rec = solve(x0, mine)                                  # [len(mine), width] on the solver's device
        dev = rec.device if hasattr(rec, 'device') else 'cpu'
        pad = torch.zeros((nodes_per_rank, width), dtype=torch.float64, device=dev)
        if len(mine):
            pad[:len(mine)] = rec
        # the global incumbent of the round: min over the ranks of the best binary-feasible cost solved locally
        best = torch.full((1,), float('inf'), dtype=torch.float64, device=dev)
        for k, node in enumerate(mine):
            if len(node.identifier) == nb and float(pad[k, 0]) == 2.:
                best = torch.minimum(best, pad[k, 1:2])
        if dist_on:
            dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
            parts = [torch.zeros_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad, group=group)
        else:
            parts = [pad]
        allrec = torch.stack(parts).cpu().numpy()               # [world, nodes_per_rank, width]
        round_best = float(best.cpu())
        rounds += 1
        # the reference's rules, node by node in best_first order (identical on every rank)
        for j, node in enumerate(batch):
            r = allrec[j % world, j // world]
            status, cost, dobj = int(r[0]), float(r[1]), float(r[2])
            if status not in (2, 3):
                raise RuntimeError('QP solver failed with status %d in the split-frontier search' % status)
            solves += 1
            dual_rec = r[3:3 + L.dual]; yc = r[3 + L.dual:3 + L.dual + pd.n]; prim = r[3 + L.dual + pd.n:]
            full = len(node.identifier) == nb
            dual = DualSolution.from_record(pd, L, dual_rec, cost if status == 2 else dobj)
            primal = PrimalSolution.from_record(pd, prim, cost, full, status == 2)
            start = ctl.qp.active_set_from_record(np.concatenate((dual_rec, yc)))
            node.lb = cost if status == 2 else np.inf
            node.binary_feasible = full
            node.extra = SubproblemSolution(primal, dual, start)
            cutoff = ub - tol
            if node.lb >= cutoff:
                continue                                        # pruned: stays a leaf with its new bound
            if full:
                incumbent, ub = node, node.lb
                continue
            children = ctl._brancher(node, lambda ident, nub: _branch_in_time(ident, nub))
            leaves.remove(node)
            for c in children:
                order[id(c)] = created; created += 1
            leaves.extend(children)
        if incumbent is not None and round_best < ub:
            raise RuntimeError('incumbent all-reduce and replicated frontier disagree')    # cannot happen: same records everywhere
    primal = None if incumbent is None else incumbent.extra.primal
    return primal, leaves, solves, rounds


def _branch_in_time(identifier, nub):
    from .controller import branch_in_time
    return branch_in_time(identifier, nub)
