"""The UNMODIFIED reference controller (imported from /root/reference) consuming the CUDA solver's QP answers.

/root/reference exists only in the authoring container (no GPU); the GPU box has no reference.  So the GPU's side
is frozen once as a transcript (tools/record_transcript.py, run on a B200: every node QP kernel K1 answered, what the
device-side search K3 explored, what the device-side warm start K2+K4 produced) and replayed HERE through the
reference's own ``HybridModelPredictiveController`` with only ``_build_mip`` overridden to return the product's
``BoundedQP`` front end (its device call answered from the transcript).  ``feedforward``, ``branch_and_bound``,
``best_first``, ``_brancher``, ``_solve_subproblem``, ``SubproblemSolution.from_controller`` and
``construct_warm_start`` run verbatim.  Asserted:

* the reference asks EXACTLY the QPs the device search solved, in the same order (every request, including the
  multipliers the node starts from, is found in the transcript bit for bit);
* explored node sequence, leaves (identifiers, order) and their lower bounds are bit-identical to K3;
* optimal cost bit-identical;
* the reference's warm start equals K2+K4: same cover, identifiers, order, dual = None pattern; bounds to 1e-11
  (the two sum the same terms in a different order).
"""
import os
import numpy as np
import pytest

from oracle.models import load_model, GOLDEN
from oracle.refload import reference_available, import_reference
from tests.util import make_problem

TRANSCRIPT = os.path.join(GOLDEN, 'cp20_transcript.npz')
pytestmark = pytest.mark.skipif(not (reference_available() and os.path.exists(TRANSCRIPT)),
                                reason='needs /root/reference (authoring container) and the recorded transcript')


def _key(x0, lb, ub, y0, yc0):
    return b''.join(np.ascontiguousarray(a + 0., dtype=np.float64).tobytes() for a in (x0, lb, ub, y0, yc0))


def _idents(depth, bits, nub):
    return [{(q // nub, q % nub): float(bits[j, q]) for q in range(depth[j])} for j in range(len(depth))]


def _replay_qp(pd, z):
    from warm_start_hmpc_b200.bounded_qp import BoundedQP
    table = {}
    for i in range(len(z['qp_status'])):
        table[_key(z['qp_x0'][i], z['qp_lb'][i], z['qp_ub'][i], z['qp_y0'][i], z['qp_yc0'][i])] = i

    class ReplayQP(BoundedQP):
        asked = []

        def _launch(self, x0, lb, ub, y0, yc0):
            k = _key(x0, lb, ub, np.zeros(pd.m) if y0 is None else y0, np.zeros(pd.n) if yc0 is None else yc0)
            assert k in table, 'the reference asked a QP (or a start) the device search never solved'
            i = table[k]
            assert bool(z['qp_hot'][i]) == (y0 is not None)
            self.asked.append(i)
            return dict(status=z['qp_status'][i], cost=z['qp_cost'][i], dobj=z['qp_dobj'][i], iters=z['qp_iters'][i],
                        primal=z['qp_primal'][i], dual=z['qp_dual'][i], yc=z['qp_yc'][i], runtime=0.)
    return ReplayQP(pd, lambda: None)


def test_unmodified_reference_controller_on_gpu_answers_equals_device_search():
    model = load_model('cp20')
    pd = make_problem(model)
    z = np.load(TRANSCRIPT)
    ctrl, bnb, sps, mlds = import_reference()
    qp = _replay_qp(pd, z)
    mld = mlds.MLDSystem([model['A'], model['B']], [model['F'], model['G'], model['h']], int(model['nub']))

    class OnGpuAnswers(ctrl.HybridModelPredictiveController):
        def _build_mip(self_):
            return qp

        def _update_mu(self_):
            return model['M_mu']
    ref = OnGpuAnswers(mld, int(model['T']), [model['Q'], model['R'], model['Q_T']], [model['F_T'], model['h_T']])
    nub = int(model['nub'])
    ws = None
    x = model['x0_nominal'].copy()
    for t in range(int(z['n_steps'])):
        assert np.array_equal(x, z['k3_%d_x' % t])
        qp.asked.clear()
        sol, leaves, n_qp, _ = ref.feedforward(x, warm_start=ws, printing_period=None)        # default Params.Method = 1
        c0, c1 = z['k3_%d_calls' % t]
        # same QPs, same order as the device-side search
        assert n_qp == c1 - c0
        assert qp.asked == list(range(c0, c1))
        order = _idents(z['k3_%d_order_depth' % t], z['k3_%d_order_bits' % t], nub)
        for i, ident in zip(qp.asked, order):
            lb, ub = ref._get_bound_binaries(ident)
            assert np.array_equal(lb.ravel(), z['qp_lb'][i]) and np.array_equal(ub.ravel(), z['qp_ub'][i])
        # same leaves, same order, bit-identical bounds; bit-identical optimum
        dev = _idents(z['k3_%d_leaf_depth' % t], z['k3_%d_leaf_bits' % t], nub)
        assert [l.identifier for l in leaves] == dev
        assert np.array_equal(np.array([l.lb for l in leaves]), z['k3_%d_leaf_lb' % t])
        assert sol.objective == float(z['k3_%d_cost' % t])
        T, nx, nu = ref.T, mld.nx, mld.nu
        U = z['k3_%d_primal' % t][(T + 1) * nx:].reshape(T, nu)
        assert np.array_equal(np.array(sol.variables['ub']), U[:, nu - nub:])
        # warm start by the reference's own construct_warm_start vs K2 + K4
        e = z['k4_%d_e' % t]
        ws, _, _ = ref.construct_warm_start(leaves, x, sol.variables['uc'][0], sol.variables['ub'][0], e)
        k4 = _idents(z['k4_%d_depth' % t], z['k4_%d_bits' % t], nub)
        assert [l.identifier for l in ws] == k4                                             # cover, identifiers, order
        lb_ref, lb_dev = np.array([l.lb for l in ws]), z['k4_%d_lb' % t]
        assert np.array_equal(np.isinf(lb_ref), np.isinf(lb_dev))
        fin = np.isfinite(lb_dev)
        scale = max(1., np.abs(z['k4_%d_dobj' % t]).max())
        assert np.all(np.abs(lb_ref[fin] - lb_dev[fin]) <= 1e-11 * scale)
        none = np.array([l.extra.dual is None for l in ws])
        assert np.array_equal(none, z['k4_%d_none' % t])
        dobj = np.array([0. if l.extra.dual is None else l.extra.dual.objective for l in ws])
        assert np.all(np.abs(dobj[~none] - z['k4_%d_dobj' % t][~none]) <= 1e-11 * scale)
        # every root starts its solve from its shifted multipliers (the reference's dormant active_set hand-over,
        # controller.py:262-264; bounded_qp.py:343-366): take them from the device records so that the next step's
        # requests can be looked up bit for bit, after checking they ARE the reference's shifted multipliers
        for j, l in enumerate(ws):
            a = dict(c=list(z['k4_%d_start_c' % t][j]), v=list(z['k4_%d_start_v' % t][j]))
            if l.extra.dual is not None:
                mine = qp.active_set_from_dual(l.extra.dual)
                assert np.allclose(mine['c'], a['c'], rtol=1e-12, atol=1e-13)
            l.extra.active_set = a
        x = sol.variables['x'][1] + e
