"""-m gpu: batched small LPs on the device (wshmpc_lp_batch) against HiGHS (scipy):
the `_update_mu` family (controller.py:186-227), random standard-form LPs incl. infeasible and unbounded ones, and an
inequality-form LP of the mcais kind (mcais.py:147-184) through its dual."""
import numpy as np
import pytest
from scipy.optimize import linprog

from oracle.models import load_model
from tests.util import make_controller

pytestmark = pytest.mark.gpu


def test_update_mu_on_the_device_equals_highs():
    for name in ('cp20', 'cp1w40'):
        model = load_model(name)
        ctl = make_controller(model)
        M_host = ctl._update['mu']
        M_dev = ctl._update_mu_device()
        assert M_dev.shape == M_host.shape and M_dev.min() >= -1e-12
        mld = ctl.mld
        lhs = np.vstack((mld.F.T, mld.G.T)).dot(M_dev)
        rhs = np.hstack((ctl.F_Tm1, ctl.G_Tm1)).T
        assert np.abs(lhs - rhs).max() <= 1e-9 * max(1., np.abs(rhs).max())               # [F G]' M = [F_Tm1 G_Tm1]'
        # same optimal values column by column (never above the unit-vector certificate the host version may keep)
        v_dev, v_host = mld.h.dot(M_dev), mld.h.dot(M_host)
        assert np.all(v_dev <= v_host + 1e-9 * np.maximum(1., np.abs(v_host)))
        n = mld.h.size
        Aeq = np.vstack((mld.F.T, mld.G.T))
        for i in (0, n - 1, n, M_host.shape[1] - 1):
            res = linprog(mld.h, A_eq=Aeq, b_eq=rhs[:, i], bounds=[(0, None)] * n, method='highs')
            assert res.status == 0 and abs(res.fun - v_dev[i]) <= 1e-8 * max(1., abs(res.fun)), (name, i, res.fun, v_dev[i])


def test_random_standard_form_lps_match_highs():
    from warm_start_hmpc_b200.capi import lp_batch
    rng = np.random.default_rng(0)
    m, n, K = 7, 30, 64
    E = rng.standard_normal((K, m, n)); c = rng.standard_normal((K, n)); r = np.zeros((K, m))
    kind = []
    for k in range(K):
        if k % 4 == 0:                                   # feasible and bounded: c = E'pi + s with s >= 0, r = E y0 with y0 >= 0
            y0 = np.maximum(rng.standard_normal(n), 0.); r[k] = E[k].dot(y0)
            c[k] = E[k].T.dot(rng.standard_normal(m)) + np.abs(rng.standard_normal(n)); kind.append(2)
        elif k % 4 == 1:                                 # infeasible: a row with non-negative coefficients and negative rhs
            E[k, 0] = np.abs(E[k, 0]); r[k] = E[k].dot(np.abs(rng.standard_normal(n))); r[k, 0] = -1.; kind.append(3)
        elif k % 4 == 2:                                 # unbounded: a ray d >= 0 with E d = 0 and c.d < 0
            d = np.abs(rng.standard_normal(n)); N = np.linalg.svd(E[k])[2][m:]                 # null space
            d = np.abs(N.T.dot(rng.standard_normal(n - m)))
            E[k] = E[k] - np.outer(E[k].dot(d), d) / d.dot(d)                                  # now E d = 0
            r[k] = E[k].dot(np.abs(rng.standard_normal(n))); c[k] = -np.abs(rng.standard_normal(n)); kind.append(5)
        else:                                            # whatever comes out
            r[k] = E[k].dot(np.abs(rng.standard_normal(n))); kind.append(0)
    out = lp_batch(E, c, r)
    st = out['status'].cpu().numpy(); obj = out['obj'].cpu().numpy(); y = out['y'].cpu().numpy(); pi = out['dual'].cpu().numpy()
    for k in range(K):
        res = linprog(c[k], A_eq=E[k], b_eq=r[k], bounds=[(0, None)] * n, method='highs')
        want = {0: 2, 2: 3, 3: 5}[res.status]
        assert st[k] == want, (k, st[k], res.status)
        if kind[k]:
            assert st[k] == kind[k]
        if want == 2:
            assert abs(obj[k] - res.fun) <= 1e-8 * max(1., abs(res.fun))
            assert np.abs(E[k].dot(y[k]) - r[k]).max() <= 1e-8 and y[k].min() >= -1e-10          # primal feasible
            assert (c[k] - E[k].T.dot(pi[k])).min() >= -1e-7 and abs(r[k].dot(pi[k]) - obj[k]) <= 1e-7 * max(1., abs(obj[k]))   # dual feasible, no gap


def test_inequality_form_lp_through_its_dual():
    """max c.x s.t. D x <= e, x free (the LPs of mcais.py:147-184) = -min e.y s.t. D'y = c, y >= 0; x = multipliers."""
    from warm_start_hmpc_b200.capi import lp_batch
    rng = np.random.default_rng(1)
    nx, nf = 4, 60
    D = rng.standard_normal((nf, nx)); e = 1. + np.abs(rng.standard_normal(nf))            # a bounded polytope around 0
    cs = rng.standard_normal((16, nx))
    out = lp_batch(D.T, e, cs)
    assert np.all(out['status'].cpu().numpy() == 2)
    obj = out['obj'].cpu().numpy(); x = out['dual'].cpu().numpy()
    for k in range(16):
        res = linprog(-cs[k], A_ub=D, b_ub=e, bounds=[(None, None)] * nx, method='highs')
        assert res.status == 0 and abs(-res.fun - obj[k]) <= 1e-8 * max(1., abs(res.fun))
        assert np.all(D.dot(x[k]) <= e + 1e-8) and abs(cs[k].dot(x[k]) - obj[k]) <= 1e-8 * max(1., abs(obj[k]))
