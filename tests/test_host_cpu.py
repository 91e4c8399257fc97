"""CPU suite (-m "not gpu"), part 2: host logic of the drop-in surface and the C-ABI library.

No compute call is made (there is no GPU here): the library must load, export every symbol
include/wshmpc.h declares, and the product must FAIL LOUDLY (no CPU fallback) when asked to solve.
"""
import ctypes
import os
import re
import numpy as np
import pytest

import warm_start_hmpc_b200 as ws
from warm_start_hmpc_b200 import capi
from warm_start_hmpc_b200.closed_loop import shard
from warm_start_hmpc_b200.subproblem_solution import DualSolution
from oracle.models import load_model, GOLDEN
from oracle.refload import reference_available
from tests.util import make_controller, make_problem, leaves_from_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    hdr = open(os.path.join(ROOT, 'include', 'wshmpc.h')).read()
    declared = set(re.findall(r'\b(wshmpc_[a-z_]+)\s*\(', hdr))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    ctl = make_controller(load_model('cp1w40'))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ctl.feedforward(np.array([0., 0., 1., 0.]), printing_period=None)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ctl.feedforward_batch(np.zeros((2, 4)))


def test_mld_and_controller_size_errors():
    """test_mld_system.py:11-30, test_controller.py:12-38: ValueError on inconsistent sizes."""
    m = load_model('cp1w40')
    with pytest.raises(ValueError):
        ws.MLDSystem([m['A'][:3], m['B']], [m['F'], m['G'], m['h']], 2)
    with pytest.raises(ValueError):
        ws.MLDSystem([m['A'], m['B']], [m['F'][:, :3], m['G'], m['h']], 2)
    mld = ws.MLDSystem([m['A'], m['B']], [m['F'], m['G'], m['h']], 2)
    assert mld.V.shape == (2, mld.nu) and np.array_equal(mld.V[:, -2:], np.eye(2))
    with pytest.raises(ValueError):
        ws.HybridModelPredictiveController(mld, 5, [m['Q'][:, :3], m['R'], m['Q_T']], None)
    with pytest.raises(ValueError):
        ws.HybridModelPredictiveController(mld, 5, [m['Q'], m['R'], m['Q_T']], [m['F_T'], m['h_T'][:-1]])


def test_update_matrices():
    """test_controller.py:40-59: rho update exactly 1.1 I on the unit-test fixture; mu update = identity
    on the stage rows, >= 0 and [F G]' M = [F_Tm1 G_Tm1]' with a terminal set."""
    m = load_model('cp1w40')
    ctl = make_controller(m)
    assert np.array_equal(ctl._update['rho'], 1.1 * np.eye(4))
    M = ctl._update['mu']
    nh = m['h'].size
    assert np.array_equal(M[:, :nh], np.eye(nh))
    assert M.min() >= 0.
    assert np.allclose(np.vstack((m['F'].T, m['G'].T)).dot(M), np.vstack((ctl.F_Tm1.T, ctl.G_Tm1.T)), atol=1e-9)
    assert np.allclose(M, m['M_mu'], atol=1e-9)


def test_branch_in_time_and_bounds():
    """controller.py:13-44, 300-327."""
    assert ws.branch_in_time({}, 2) == [{(0, 0): 0.}, {(0, 0): 1.}]
    assert ws.branch_in_time({(0, 0): 1.}, 2) == [{(0, 1): 0.}, {(0, 1): 1.}]
    assert ws.branch_in_time({(0, 0): 1., (0, 1): 0.}, 2) == [{(1, 0): 0.}, {(1, 0): 1.}]
    ctl = make_controller(load_model('cp1w40'))
    lb, ub = ctl._get_bound_binaries({(0, 1): 1., (3, 0): 0.})
    assert lb[0, 1] == ub[0, 1] == 1. and lb[3, 0] == ub[3, 0] == 0. and lb.sum() == 1 and ub.sum() == ub.size - 1


def _knapsack_solver():
    """A toy problem for the generic B&B: minimise c.b over b in {0,1}^4 with a coupling penalty."""
    c = np.array([3., -2., 1.5, -1.])

    def solver(identifier, cutoff, extra):
        fixed = dict((k[1], v) for k, v in identifier.items())
        lb = sum(c[i] * v for i, v in fixed.items()) + sum(min(0., c[i]) for i in range(4) if i not in fixed)
        if fixed.get(1) == 1. and fixed.get(3) == 1.:
            lb = np.inf
        return lb, len(fixed) == 4, 0., None
    return solver


@pytest.mark.parametrize('rule', ['best_first', 'depth_first', 'breadth_first'])
def test_generic_branch_and_bound_equals_reference(rule):
    """branch_and_bound.py:408-563 with arbitrary callables: same incumbent, leaves and solve count as
    the reference implementation (when /root/reference is there), known optimum otherwise."""
    def brancher(node):
        i = len(node.identifier)
        return [ws.Node({**node.identifier, (0, i): v}, node.lb) for v in (0., 1.)]
    inc, leaves, solves, _ = ws.branch_and_bound(_knapsack_solver(), getattr(ws, rule), brancher, printing_period=None)
    assert inc.lb == -2. and inc.identifier == {(0, 0): 0., (0, 1): 1., (0, 2): 0., (0, 3): 0.}
    if reference_available():
        from oracle.refload import import_reference
        _, rb, _, _ = import_reference()
        rinc, rleaves, rsolves, _ = rb.branch_and_bound(
            _knapsack_solver(), getattr(rb, rule),
            lambda node: [rb.Node({**node.identifier, (0, len(node.identifier)): v}, node.lb) for v in (0., 1.)],
            printing_period=None)
        assert rsolves == solves and rinc.identifier == inc.identifier
        assert [l.identifier for l in rleaves] == [l.identifier for l in leaves]
        assert [l.lb for l in rleaves] == [l.lb for l in leaves]


def test_record_round_trip_and_layout():
    model = load_model('cp20')
    pd = make_problem(model)
    L = pd.layout
    assert L.primal == 21 * 4 + 20 * 7 and L.dual == 84 + pd.mc + 160 + 84 + 20
    rec = np.random.default_rng(0).standard_normal(L.dual)
    d = DualSolution.from_record(pd, L, rec, 1.5)
    assert len(d.variables['mu']) == 20 and d.variables['mu'][-1].size == pd.nh1
    assert np.array_equal(DualSolution.to_record(pd, L, d.variables), rec)


def test_host_construct_warm_start_equals_reference_golden():
    """The host restatement of construct_warm_start in the product (used by the drop-in API with custom
    rules) on the reference's golden leaves: bounds equal to the reference code's output."""
    model = load_model('cp20')
    ctl = make_controller(model)
    g = np.load(os.path.join(GOLDEN, 'cp20_warmstart.npz'))
    for tag, e0 in (('zero', np.zeros(4)), ('rand', g['e_rand'])):
        leaves = [ws.Node(i, lb, ws.SubproblemSolution(None, d)) for i, lb, d in leaves_from_golden(ctl.problem, g)]
        nodes, _, _ = ctl.construct_warm_start(leaves, g['x0'], g['uc0'], g['ub0'], e0)
        assert len(nodes) == 77
        assert np.array_equal(np.array([n.lb for n in nodes]), g['ws_%s_lb' % tag])
        assert [n.extra.dual is None for n in nodes] == list(g['ws_%s_none' % tag])


def test_shard_partition():
    for n, w in ((4096, 8), (10, 4), (3, 8)):
        blocks = [shard(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[r][1] == blocks[r + 1][0] for r in range(w - 1))
        assert max(b - a for a, b in blocks) - min(b - a for a, b in blocks) <= 1
