"""TEST INFRASTRUCTURE (oracle side) -- solver-independent certificates for a node-QP result.

Restates the plug-in checkers of the reference fixture
(warm_start_hmpc/test/cart_pole_with_wall.py:171-205 primal, :207-247 dual, :249-268 dual objective)
for an arbitrary model.  A result that passes with residuals r proves its cost to O(r) regardless of
who computed it: primal feasibility gives cost >= optimum, dual feasibility gives
dual objective <= optimum.
"""
import numpy as np


def families(model, cond, x0, status, z, y):
    """(z, y) of the condensed problem -> reference families x, u, lam, mu, nu_lb, nu_ub, rho, sigma
    (subproblem_solution.py:137-166)."""
    T, nu, nub, nx = cond.T, cond.nu, cond.nub, cond.nx
    A, Q, R, Q_T = model['A'], model['Q'], model['R'], model['Q_T']
    mu = []
    for t in range(T):
        k = cond.nh if t < T - 1 else cond.nh1
        mu.append(np.maximum(y[cond.row0[t]:cond.row0[t] + k], 0.))
    yb = y[cond.mc:].reshape(T, nub)
    out = dict(mu=mu, nu_ub=list(np.maximum(yb, 0.)), nu_lb=list(np.maximum(-yb, 0.)))
    if status == 2:
        x = cond.states(x0, z); u = z.reshape(T, nu)
        out['x'], out['u'] = x, u
        out['rho'] = [2. * Q.dot(x[t]) for t in range(T)] + [2. * Q_T.dot(x[T])]
        out['sigma'] = [2. * R.dot(u[t]) for t in range(T)]
    else:
        out['rho'] = [np.zeros(Q.shape[0]) for t in range(T)] + [np.zeros(Q_T.shape[0])]
        out['sigma'] = [np.zeros(R.shape[0]) for t in range(T)]
    lam = [None] * (T + 1)
    lam[T] = -Q_T.T.dot(out['rho'][T])
    for t in range(T - 1, -1, -1):
        Ft = model['F'] if t < T - 1 else model['F_Tm1']
        lam[t] = A.T.dot(lam[t + 1]) - Q.T.dot(out['rho'][t]) - Ft.T.dot(mu[t])
    out['lam'] = lam
    return out


def primal_residuals(model, cond, x0, lb, ub, fam):
    """cart_pole_with_wall.py:171-205 -> (max |equality|, max inequality violation)."""
    T, nub = cond.T, cond.nub
    x, u = fam['x'], fam['u']
    eq = [x0 - x[0]]
    viol = []
    for t in range(T):
        eq.append(model['A'].dot(x[t]) + model['B'].dot(u[t]) - x[t + 1])
        Ft, Gt, ht = (model['F'], model['G'], model['h']) if t < T - 1 else (model['F_Tm1'], model['G_Tm1'], model['h_Tm1'])
        viol.append(Ft.dot(x[t]) + Gt.dot(u[t]) - ht)
        ubt = u[t, -nub:]
        viol.append(lb[t * nub:(t + 1) * nub] - ubt)
        viol.append(ubt - ub[t * nub:(t + 1) * nub])
    return np.max(np.abs(np.concatenate(eq))), max(0., np.max(np.concatenate(viol)))


def primal_violation_scaled(model, cond, x0, lb, ub, fam):
    """Largest inequality violation with every row divided by max(1, |row of [F G]|): the scaling in which the CUDA
    solver's primal tolerance is stated (qp_device.cuh pricing: tol_p, escalated to at most 100 tol_p on rows that
    entered the working set degenerately).  Bound rows have unit norm."""
    T, nub = cond.T, cond.nub
    x, u = fam['x'], fam['u']
    worst = 0.
    for t in range(T):
        Ft, Gt, ht = (model['F'], model['G'], model['h']) if t < T - 1 else (model['F_Tm1'], model['G_Tm1'], model['h_Tm1'])
        sc = np.maximum(1., np.sqrt((Ft ** 2).sum(1) + (Gt ** 2).sum(1)))
        worst = max(worst, np.max((Ft.dot(x[t]) + Gt.dot(u[t]) - ht) / sc))
        ubt = u[t, -nub:]
        worst = max(worst, np.max(lb[t * nub:(t + 1) * nub] - ubt), np.max(ubt - ub[t * nub:(t + 1) * nub]))
    return max(0., worst)


def dual_residuals(model, cond, fam):
    """cart_pole_with_wall.py:207-247 -> (max |stationarity|, most negative multiplier)."""
    T, nuc = cond.T, cond.nuc
    A, B, Q, R, Q_T = model['A'], model['B'], model['Q'], model['R'], model['Q_T']
    V = np.hstack((np.zeros((cond.nub, nuc)), np.eye(cond.nub)))
    z = [Q_T.T.dot(fam['rho'][T]) + fam['lam'][T]]
    for t in range(T):
        Ft, Gt = (model['F'], model['G']) if t < T - 1 else (model['F_Tm1'], model['G_Tm1'])
        z.append(Q.T.dot(fam['rho'][t]) + fam['lam'][t] - A.T.dot(fam['lam'][t + 1]) + Ft.T.dot(fam['mu'][t]))
        z.append(R.T.dot(fam['sigma'][t]) - B.T.dot(fam['lam'][t + 1]) + Gt.T.dot(fam['mu'][t])
                 + V.T.dot(fam['nu_ub'][t] - fam['nu_lb'][t]))
    neg = min(np.min(np.concatenate(fam['mu'])), np.min(np.concatenate(fam['nu_lb'])), np.min(np.concatenate(fam['nu_ub'])))
    return np.max(np.abs(np.concatenate(z))), neg


def dual_objective(model, cond, x0, lb, ub, fam):
    """cart_pole_with_wall.py:249-268 (with x0 in place of the fixture's x1)."""
    T, nub = cond.T, cond.nub
    obj = 0.
    for k in ('rho', 'sigma'):
        obj -= sum(v.dot(v) for v in fam[k]) / 4.
    obj -= fam['lam'][0].dot(x0)
    obj += sum(lb[t * nub:(t + 1) * nub].dot(fam['nu_lb'][t]) for t in range(T))
    obj -= sum(ub[t * nub:(t + 1) * nub].dot(fam['nu_ub'][t]) for t in range(T))
    obj -= sum(model['h'].dot(v) for v in fam['mu'][:-1])
    obj -= model['h_Tm1'].dot(fam['mu'][-1])
    return obj


def certify(model, cond, x0, lb, ub, out):
    """Returns dict of residuals.  Optimal: primal eq/ineq violation, dual stationarity, relative
    duality gap.  Infeasible: stationarity of the ray (rho = sigma = 0) and its (positive) cost."""
    fam = families(model, cond, x0, out['status'], out.get('z'), out['y'])
    ds, neg = dual_residuals(model, cond, fam)
    dobj = dual_objective(model, cond, x0, lb, ub, fam)
    r = dict(dual_stat=ds, dual_neg=neg, dual_obj=dobj)
    if out['status'] == 2:
        pe, pv = primal_residuals(model, cond, x0, lb, ub, fam)
        x, u = fam['x'], fam['u']
        cost = sum(np.sum(model['Q'].dot(x[t]) ** 2) + np.sum(model['R'].dot(u[t]) ** 2) for t in range(cond.T))
        cost += np.sum(model['Q_T'].dot(x[cond.T]) ** 2)
        r.update(prim_eq=pe, prim_viol=pv, cost=cost, gap=(cost - dobj) / max(1e-12, abs(cost)))
    else:
        scale = max(1e-300, np.max(np.abs(out['y'])))
        r.update(ray_cost=dobj, ray_cost_rel=dobj / scale, dual_stat_rel=ds / scale)
    return r
