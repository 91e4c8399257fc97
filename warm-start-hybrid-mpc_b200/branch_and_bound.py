"""Host-side branch and bound with the reference's callable seams (branch_and_bound.py:7-55,
408-563): ``Node``, ``branch_and_bound(solver, candidate_selection, brancher, tol, warm_start,
printing_period, draw_label)`` and the three selection rules.  This is the generic, problem-agnostic
drop-in (any solver / brancher callables).  The batched device-side search used for throughput lives in
csrc/ (K3) and is reached through ``HybridModelPredictiveController.feedforward_batch``.

The reference's ``Printer`` / ``Drawer`` (branch_and_bound.py:57-406) are logging / pygraphviz
visualisation and out of scope: ``printing_period`` and ``draw_label``
are accepted and ignored.
"""
import numpy as np


class Node(object):
    """branch_and_bound.py:7-55."""

    def __init__(self, identifier, lb=-np.inf, extra=None):
        self.identifier = identifier
        self.lb = lb
        self.extra = extra
        self.binary_feasible = None
        self.solve_time = None

    def solve(self, solver, cutoff=None):
        [self.lb, self.binary_feasible, self.solve_time, self.extra] = solver(self.identifier, cutoff, self.extra)


def branch_and_bound(solver, candidate_selection, brancher, tol=0., warm_start=None,
                     printing_period=3., draw_label=None, **kwargs):
    """branch_and_bound.py:408-499, same control flow and same return tuple
    (incumbent, leaves, solves, solver_time)."""
    ub = np.inf
    incumbent = None
    leaves = [Node({})] if warm_start is None else warm_start
    solves = 0
    solver_time = 0
    while True:
        candidate_nodes = [l for l in leaves if l.lb < ub - tol]
        if not candidate_nodes:
            break
        working_node = candidate_selection(candidate_nodes)
        cutoff = ub - tol
        working_node.solve(solver, cutoff)
        solves += 1
        solver_time += working_node.solve_time
        if working_node.lb >= cutoff:          # pruning: the node stays a leaf
            pass
        elif working_node.binary_feasible:     # new incumbent
            incumbent = working_node
            ub = working_node.lb
        else:                                  # branching
            children = brancher(working_node)
            leaves.remove(working_node)
            leaves.extend(children)
    return incumbent, leaves, solves, solver_time


def breadth_first(candidate_nodes):
    """branch_and_bound.py:501-518 (FIFO)."""
    return candidate_nodes[0]


def depth_first(candidate_nodes):
    """branch_and_bound.py:521-538 (LIFO)."""
    return candidate_nodes[-1]


def best_first(candidate_nodes):
    """branch_and_bound.py:541-563: lowest lower bound, first one wins on ties (np.argmin)."""
    return candidate_nodes[int(np.argmin([l.lb for l in candidate_nodes]))]
