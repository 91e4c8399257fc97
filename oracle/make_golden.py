"""TEST INFRASTRUCTURE (oracle side) -- generates the golden fixtures under tests/golden/.

Runs in the AUTHORING container only (needs /root/reference): the reference's own, unmodified
``HybridModelPredictiveController.feedforward`` / ``branch_and_bound`` / ``construct_warm_start``
(warm_start_hmpc/controller.py:329-564, branch_and_bound.py:408-499) are executed with the oracle's
C core (oracle/qp_core.c) standing in for Gurobi, and everything the parity tests need is frozen:

* cp20_closed_loop.npz  nominal (e = 0) and noisy closed loop of the notebook two-wall cart-pole, T = 20:
                        per step state, optimal cost, first input, optimal mode sequence, cold / warm QP
                        counts, warm-start cover size;  plus the PUBLISHED node counts of
                        notebooks/cart_pole_with_walls/data/nodes_{cs,ws,len_ws}_sd_0.000.npy (Gurobi).
* cp20_nodes.npz        the 160 node QPs of the step-0 cold solve, in the order the reference B&B
                        solved them: bounds, status, cost / Farkas cost.
* cp20_warmstart.npz    the 81 leaves of that solve (identifier, lb, dual record), the inputs of
                        construct_warm_start and the shifted identifiers / bounds / dual objectives the
                        reference code returns for e0 = 0 and for a random e0.

    python -m oracle.make_golden
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle.models import load_model, GOLDEN                      # noqa: E402
from oracle.qp_c import CoreC                                     # noqa: E402
from oracle.ref_facade import make_reference_controller           # noqa: E402

DATA = '/root/reference/notebooks/cart_pole_with_walls/data/'


def ident_arrays(identifiers, nub, nb):
    n = len(identifiers)
    depth = np.zeros(n, np.int32); bits = np.zeros((n, (nb + 31) // 32), np.uint32)
    for j, ident in enumerate(identifiers):
        depth[j] = len(ident)
        assert all((q // nub, q % nub) in ident for q in range(len(ident))), 'identifier is not a prefix'
        for (t, i), v in ident.items():
            if v:
                q = t * nub + i
                bits[j, q >> 5] |= np.uint32(1 << (q & 31))
    return depth, bits


def records_of(leaves, pd):
    """dual records (product layout) of a list of reference Nodes; children alias the parent's record."""
    from warm_start_hmpc_b200.subproblem_solution import DualSolution
    rec = np.full(len(leaves), -1, np.int32); recs = []; dobj = []; seen = {}
    for j, l in enumerate(leaves):
        dual = None if l.extra is None else l.extra.dual
        if dual is None:
            continue
        if id(dual) not in seen:
            seen[id(dual)] = len(recs)
            recs.append(DualSolution.to_record(pd, pd.layout, dual.variables)); dobj.append(dual.objective)
        rec[j] = seen[id(dual)]
    return rec, np.array(recs), np.array(dobj)


def closed_loop(ctl, model, n_steps, sigma, seed):
    """statistical_analysis.py:73-194 without the Gurobi legs."""
    T, nub = int(model['T']), int(model['nub'])
    np.random.seed(seed)
    x = model['x0_nominal'].copy()
    ws = None
    out = dict(x=[], cost=[], u0=[], ub=[], n_cold=[], n_warm=[], cover=[], e=[], cost_warm=[])
    for t in range(n_steps):
        sol_c, leaves_c, n_c, _ = ctl.feedforward(x, printing_period=None)
        sol_w, leaves_w, n_w, _ = ctl.feedforward(x, warm_start=ws, printing_period=None)
        assert np.isclose(sol_c.objective, sol_w.objective)
        e = sigma * np.multiply(np.random.randn(x.size), model['x_max'])
        ws, _, _ = ctl.construct_warm_start(leaves_w, sol_w.variables['x'][0], sol_c.variables['uc'][0],
                                            sol_c.variables['ub'][0], e)
        out['x'].append(x.copy()); out['cost'].append(sol_c.objective); out['cost_warm'].append(sol_w.objective)
        out['u0'].append(np.concatenate((sol_c.variables['uc'][0], sol_c.variables['ub'][0])))
        out['ub'].append(np.array(sol_c.variables['ub']).reshape(T, nub))
        out['n_cold'].append(n_c); out['n_warm'].append(n_w); out['cover'].append(len(ws)); out['e'].append(e)
        x = sol_w.variables['x'][1] + e
        print('sigma %.3f step %d: cold %d warm %d cover %d cost %.10f' % (sigma, t, n_c, n_w, len(ws), sol_c.objective), flush=True)
    return {k: np.array(v) for k, v in out.items()}


def main():
    from tests.util import make_problem
    model = load_model('cp20')
    pd = make_problem(model)
    core = CoreC(model)
    T, nub = int(model['T']), int(model['nub'])
    nb = T * nub

    # ---- 1. closed loops
    ctl = make_reference_controller(model, lambda x0, lb, ub: core.solve(x0, lb, ub))
    nom = closed_loop(ctl, model, 8, 0., 0)
    noisy = closed_loop(ctl, model, 6, 0.003, 0)
    pub = {k: np.load(DATA + 'nodes_%s_sd_0.000.npy' % k).ravel() for k in ('cs', 'ws', 'len_ws')}
    d = {'nom_' + k: v for k, v in nom.items()}
    d.update({'noisy_' + k: v for k, v in noisy.items()})
    d.update({'published_nodes_' + k: v for k, v in pub.items()})
    d['noisy_sigma'] = 0.003
    np.savez_compressed(os.path.join(GOLDEN, 'cp20_closed_loop.npz'), **d)

    # ---- 2. the node QPs of the step-0 cold solve, in B&B order
    ctl = make_reference_controller(model, lambda x0, lb, ub: core.solve(x0, lb, ub))
    x0 = model['x0_nominal'].copy()
    sol, leaves, n_qp, _ = ctl.feedforward(x0, printing_period=None)
    log = ctl.qp.log
    assert len(log) == n_qp
    np.savez_compressed(os.path.join(GOLDEN, 'cp20_nodes.npz'), x0=x0,
                        lb=np.array([l[1] for l in log]), ub=np.array([l[2] for l in log]),
                        status=np.array([l[3]['status'] for l in log], np.int32),
                        cost=np.array([l[3]['cost'] for l in log]),
                        farkas=np.array([l[3].get('farkas', 0.) for l in log]),
                        opt_cost=sol.objective, opt_ub=np.array(sol.variables['ub']).reshape(T, nub),
                        opt_u0=np.concatenate((sol.variables['uc'][0], sol.variables['ub'][0])))

    # ---- 3. warm start of step 1 by the reference's construct_warm_start
    depth, bits = ident_arrays([l.identifier for l in leaves], nub, nb)
    lbs = np.array([l.lb for l in leaves])
    rec, recs, dobj = records_of(leaves, pd)
    uc0, ub0 = sol.variables['uc'][0], sol.variables['ub'][0]
    x1 = sol.variables['x'][1]
    rng = np.random.default_rng(5)
    e_rand = 0.01 * rng.standard_normal(x0.size) * model['x_max']
    d = dict(x0=x0, uc0=uc0, ub0=ub0, x1=x1, e_rand=e_rand, depth=depth, bits=bits, lb=lbs, rec=rec, recs=recs, dobj=dobj)
    for tag, e0 in (('zero', np.zeros(x0.size)), ('rand', e_rand)):
        import copy
        ws, _, _ = ctl.construct_warm_start(copy.deepcopy(leaves), x0, uc0, ub0, e0)
        wd, wb = ident_arrays([l.identifier for l in ws], nub, nb)
        d['ws_%s_depth' % tag] = wd; d['ws_%s_bits' % tag] = wb
        d['ws_%s_lb' % tag] = np.array([l.lb for l in ws])
        d['ws_%s_none' % tag] = np.array([l.extra.dual is None for l in ws])
        d['ws_%s_dobj' % tag] = np.array([0. if l.extra.dual is None else l.extra.dual.objective for l in ws])
        if tag == 'rand':
            # one full shifted record, to pin _shift_dual_variables entry by entry
            j = int(np.argmax([0 if l.extra.dual is None else len(l.identifier) for l in ws]))
            from warm_start_hmpc_b200.subproblem_solution import DualSolution
            d['ws_rand_sample'] = j
            d['ws_rand_sample_rec'] = DualSolution.to_record(pd, pd.layout, ws[j].extra.dual.variables)
    np.savez_compressed(os.path.join(GOLDEN, 'cp20_warmstart.npz'), **d)
    print('cover', len(ws), 'leaves', len(leaves), 'records', recs.shape)


if __name__ == '__main__':
    main()
