"""-m gpu: the round-2 execution model and the multi-GPU / comparator code paths on a real device.

* one-lane (255 registers) and two-lane (128 registers) builds of the kernels, and any number of solver slots, give
  bit-identical results (the factor's summation order does not depend on where a column lives);
* capacity overflow of the warm start switches an instance off with status 2 instead of truncating its cover;
* split-frontier search (single rank: no collective) and the MIQP comparator against the device-side search.
"""
import os
import numpy as np
import pytest
import torch

from oracle.models import load_model, MODELS
from tests.util import make_controller

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cp20():
    model = load_model('cp20')
    return model, make_controller(model)


def test_one_lane_and_two_lane_builds_agree_bit_for_bit(cp20):
    model, ctl = cp20
    from warm_start_hmpc_b200.closed_loop import ClosedLoop
    N, S = 300, 3
    x0 = np.load(os.path.join(MODELS, 'cp20_instances.npy'))[200:200 + N]
    e = torch.as_tensor(0.003 * np.random.default_rng(3).standard_normal((S, N, 4)) * model['x_max'], device='cuda')
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    runs = {}
    for slots in (8, sms, ctl.default_slots()):       # <= SMs: one lane per CTA (255 registers); more: two lanes (128 registers)
        L = ClosedLoop(ctl, N, warm=True, max_solves=1024, max_roots=512, n_slots=slots)
        L.reset(x0)
        logs = L.run(S, e=e)
        torch.cuda.synchronize()
        runs[slots] = (logs['cost'].clone(), logs['n_solves'].clone(), L.x.clone(), L.active.clone(), int(L.totals[1]))
        del L
    ref = runs[8]
    for slots, r in runs.items():
        assert torch.equal(r[0], ref[0]) and torch.equal(r[1], ref[1]) and torch.equal(r[2], ref[2]) and torch.equal(r[3], ref[3])
        assert r[4] == ref[4]                           # the same number of active-set iterations, too
    assert ctl.default_slots() >= 2 * sms               # CP20 runs two lanes per SM


def test_warm_start_overflow_switches_the_instance_off(cp20):
    """a cover that does not fit the new tree must not be truncated (a truncated cover can report a
    suboptimal or 'infeasible' MIQP as solved): the instance gets status 2 and leaves the loop."""
    model, ctl = cp20
    from warm_start_hmpc_b200.closed_loop import ClosedLoop
    L = ClosedLoop(ctl, 2, warm=True, max_solves=256, max_roots=8, n_slots=2)      # cap_recs = 265 >= 160 solves, cap for roots tiny
    # new trees with room for fewer records than the 77 leaves the shift retains
    small = [ctl.handle().new_tree(2, 600, 40) for _ in range(2)]
    L.reset(np.repeat(model['x0_nominal'][None], 2, 0))
    out = L.step()                                       # cold solve in the regular tree, shift into ... the regular other tree
    assert int(out['status'][0]) == 0 and int(L.trees[L.cur].n_nodes[0]) == 77
    # shift the same leaves into the small tree through the C ABI: overflow
    h = L.h
    L2 = ClosedLoop(ctl, 2, warm=True, max_solves=256, max_roots=8, n_slots=2)
    L2.reset(np.repeat(model['x0_nominal'][None], 2, 0))
    tree = L2.trees[0]
    h.tree_init_root(tree)
    h.bnb_solve(L2.x, tree, max_solves=256, active=L2.active, out=L2.out)
    act = L2.active.clone()
    h.shift_tree(L2.x, None, tree, L2.out['cost'], L2.out['primal'], small[0], active=act)
    torch.cuda.synchronize()
    assert act.tolist() == [0, 0] and small[0].n_nodes.tolist() == [0, 0]
    # ... and in the fused loop the status says why
    L3 = ClosedLoop(ctl, 2, warm=True, max_solves=256, max_roots=8, n_slots=2)
    L3.trees = [ctl.handle().new_tree(2, 600, 200), ctl.handle().new_tree(2, 600, 40)]
    L3.reset(np.repeat(model['x0_nominal'][None], 2, 0))
    logs = L3.run(2)
    torch.cuda.synchronize()
    assert logs['status'][0].tolist() == [2, 2] and L3.active.tolist() == [0, 0]
    assert np.isfinite(logs['cost'][0].cpu().numpy()).all()          # the step itself was solved; only its warm start did not fit


def test_split_frontier_single_rank_equals_device_search(cp20):
    model, ctl = cp20
    from warm_start_hmpc_b200.split_frontier import split_frontier_bnb, GpuBatchSolver
    x0 = model['x0_nominal']
    sol_d, leaves_d, n_d, _ = ctl.feedforward(x0, printing_period=None)
    solver = GpuBatchSolver(ctl, n_slots=4)
    for npr in (1, 3):
        sol, leaves, n, rounds = split_frontier_bnb(ctl, x0, solver, nodes_per_rank=npr)
        assert abs(sol.objective - sol_d.objective) <= 1e-12 * sol_d.objective
        assert np.array_equal(np.array(sol.variables['ub']), np.array(sol_d.variables['ub']))
        assert n >= n_d and rounds <= n
        if npr == 1:
            # one node per round on one rank IS the reference's loop: same nodes, same leaves, same bounds
            assert n == n_d and rounds == n
            assert [sorted(l.identifier.items()) for l in leaves] == [sorted(l.identifier.items()) for l in leaves_d]
            assert np.array_equal(np.array([l.lb for l in leaves]), np.array([l.lb for l in leaves_d]))
    # the leaves of a multi-node round are a cover the device shifts
    ws, _, _ = ctl.construct_warm_start(leaves, x0, sol.variables['uc'][0], sol.variables['ub'][0], np.zeros(4))
    assert len(ws) >= 70


def test_miqp_comparator_on_the_device(cp20):
    """the conventional MIQP branch and bound through the QP seam returns the structured search's optimum."""
    model, ctl = cp20
    x0 = model['x0_nominal']
    sol, _, n_qp, _ = ctl.feedforward(x0, printing_period=None)
    variables, objective, nodes, seconds = ctl.feedforward_gurobi(x0, {'OutputFlag': 0, 'MIPGap': 0, 'Threads': 1})
    assert abs(objective - sol.objective) <= 1e-9 * sol.objective
    assert np.array_equal(variables['ub'], np.array(sol.variables['ub']))
    assert nodes > n_qp and seconds > 0.


def test_closed_loop_against_the_nonlinear_plant(cp20):
    """warm-started hybrid MPC driving the TRUE plant (cart-pole between soft walls, Euler sub-steps of
    1 ms): the model error fed to the warm start is the measured one; the cart that starts moving towards the right wall
    (x0 = [0, 0, 1, 0]) is stopped and brought back, every step stays feasible, warm-started steps stay cheap."""
    model, ctl = cp20
    from warm_start_hmpc_b200.closed_loop import ClosedLoop
    from warm_start_hmpc_b200.plants import CartPoleWithWalls
    plant = CartPoleWithWalls()
    h = 0.05
    sim = lambda x, u0: plant.simulate(x, h, u0[:, 0])
    N = 3
    x0 = np.repeat(model['x0_nominal'][None], N, 0) * np.array([1., 0.9, 0.8])[:, None]
    L = ClosedLoop(ctl, N, warm=True, max_solves=2048, max_roots=1024, n_slots=N)
    L.reset(x0)
    costs, solves, errs = [], [], []
    for t in range(40):
        out, u0, e = L.step_with_plant(sim)
        assert out['status'].tolist() == [0] * N, (t, out['status'].tolist())
        costs.append(out['cost'].cpu().numpy().copy()); solves.append(out['n_solves'].cpu().numpy().copy()); errs.append(np.abs(e).max())
        assert np.all(np.abs(u0[:, 0]) <= 1. + 1e-9)
    costs, solves = np.array(costs), np.array(solves)
    x_end = L.x.cpu().numpy()
    assert np.all(np.abs(x_end[:, 2]) < 0.5 * np.abs(x0[:, 2]))              # the cart has been slowed down
    assert np.all(costs[-1] < costs[0])
    assert solves[1:].mean() * 3 < solves[0].mean()                            # warm start pays off against the true plant too
    # linear model vs plant: a one-step error of 1e-3 in free flight, up to ~0.3 (velocity) on the steps the stiff wall acts
    assert errs[0] < 0.01 and max(errs) < 1.


@pytest.mark.parametrize('order_name', ['branch_in_reverse_time', 'branch_by_index'])
def test_static_branch_orders_on_the_device_equal_the_host_loop(cp20, order_name):
    """a user branch_rule of the static-order family runs in the device-side search (identifiers are
    (assigned mask, values) pairs, wshmpc_set_branch_order): same explored nodes, leaves, bounds and optimum as the
    reference's host loop driven by the same rule through the QP seam; the non-prefix leaves are time-shifted by K2+K4
    like the host formulation does, and the warm-started next step finds the optimum the default rule finds."""
    import warm_start_hmpc_b200 as ws
    model, ctl = cp20
    rule = getattr(ws, order_name)(ctl.T, ctl.mld.nub)
    x0 = model['x0_nominal'] * 0.5                       # a state whose tree stays small under any branching order
    ctl.device_search = False
    try:
        sol_h, leaves_h, n_h, _ = ctl.feedforward(x0, branch_rule=rule, printing_period=None)
        ws_h, _, _ = ctl.construct_warm_start(leaves_h, x0, sol_h.variables['uc'][0], sol_h.variables['ub'][0], np.zeros(4))
    finally:
        ctl.device_search = True
    sol_d, leaves_d, n_d, _ = ctl.feedforward(x0, branch_rule=rule, printing_period=None)
    sol_0, _, n_0, _ = ctl.feedforward(x0, printing_period=None)                      # branch_in_time
    assert n_d == n_h and sol_d.objective == sol_h.objective
    assert abs(sol_d.objective - sol_0.objective) <= 1e-9 * sol_0.objective
    key = lambda ls: [sorted(l.identifier.items()) for l in ls]
    assert key(leaves_d) == key(leaves_h)
    assert np.array_equal(np.array([l.lb for l in leaves_d]), np.array([l.lb for l in leaves_h]))
    assert any(not all((q // ctl.mld.nub, q % ctl.mld.nub) in l.identifier for q in range(len(l.identifier))) for l in leaves_d)   # non-prefix identifiers
    # warm start of the non-prefix leaves: K2+K4 against the host formulation
    e0 = 0.003 * np.random.default_rng(0).standard_normal(4) * model['x_max']
    ws_d, _, _ = ctl.construct_warm_start(leaves_d, x0, sol_d.variables['uc'][0], sol_d.variables['ub'][0], e0)
    ctl.device_search = False
    try:
        ws_h, _, _ = ctl.construct_warm_start(leaves_h, x0, sol_h.variables['uc'][0], sol_h.variables['ub'][0], e0)
    finally:
        ctl.device_search = True
    assert key(ws_d) == key(ws_h)
    a, b = np.array([l.lb for l in ws_d]), np.array([l.lb for l in ws_h])
    assert np.array_equal(np.isinf(a), np.isinf(b)) and np.all(np.abs(a[np.isfinite(a)] - b[np.isfinite(b)]) <= 1e-11)
    assert [l.extra.dual is None for l in ws_d] == [l.extra.dual is None for l in ws_h]
    # next step, warm-started with the same rule on the device, against a cold default-rule solve
    x1 = sol_d.variables['x'][1] + e0
    sol_w, _, n_w, _ = ctl.feedforward(x1, branch_rule=rule, warm_start=ws_d, printing_period=None)
    sol_c, _, n_c, _ = ctl.feedforward(x1, printing_period=None)
    assert abs(sol_w.objective - sol_c.objective) <= 1e-9 * sol_c.objective
    assert np.array_equal(np.array(sol_w.variables['ub']), np.array(sol_c.variables['ub']))


def test_nonlinear_plant_through_the_mailbox_equals_lock_step(cp20):
    """The TRUE plant in the loop every step of every instance WITHOUT lock step (ClosedLoop.run_mailbox, one persistent
    launch, plant called per ready instance) gives bit for bit what ClosedLoop.step_with_plant gives step by step: the
    plant is a deterministic function of (measured state, applied input), and the kernel's answers do not depend on
    which lane solves what when."""
    model, ctl = cp20
    from warm_start_hmpc_b200.closed_loop import ClosedLoop
    from warm_start_hmpc_b200.plants import CartPoleWithWalls
    plant = CartPoleWithWalls()
    h = 0.05
    N, S = 6, 12
    x0 = np.repeat(model['x0_nominal'][None], N, 0) * np.linspace(1., 0.8, N)[:, None]
    lock = ClosedLoop(ctl, N, warm=True, max_solves=2048, max_roots=1024, n_slots=N)
    lock.reset(x0)
    costs, inputs = [], []
    for t in range(S):
        out, u0, e = lock.step_with_plant(lambda x, u: plant.simulate(x, h, u[:, 0]))
        costs.append(out['cost'].cpu().numpy().copy()); inputs.append(u0.copy())
    mbx = ClosedLoop(ctl, N, warm=True, max_solves=2048, max_roots=1024, n_slots=N)
    mbx.reset(x0)
    x_meas = np.array(x0)                                      # the host keeps the measured state of every plant

    def host_plant(idx, step, u0, x_pred):
        x_meas[idx] = plant.simulate(x_meas[idx], h, u0[:, 0])
        return x_meas[idx]
    logs = mbx.run_mailbox(S, host_plant, timeout_s=60.)
    torch.cuda.synchronize()
    costs = np.array(costs)
    assert np.isfinite(costs).all()
    assert np.array_equal(logs['cost'].cpu().numpy(), costs)
    assert np.array_equal(logs['host']['u0'], np.array(inputs))
    assert torch.equal(mbx.x, lock.x)
