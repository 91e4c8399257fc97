"""TEST INFRASTRUCTURE (oracle side) -- reference-grade CPU solver for the node QP.

Stands in for the Gurobi call of bounded_qp.py:200-228 (gurobipy, version unpinned in
setup.py:16-20, is absent here -- SURVEY.md section 8c).  Algorithm: dual active-set on the
least-distance form of the proximal-regularised condensed QP (Goldfarb-Idnani family /
Arnstrom's DAQP), with an outer proximal-point loop because the condensed Hessian is only
positive SEMI-definite (rank 40 of 140 on CP20).  Linear algebra here is
deliberately naive and robust (dense re-factorisation by QR / least squares at every
iteration) -- this file is the slow, independent referee of oracle/qp_core.c and of the CUDA
kernel, which use incremental LDL' updates.

Outputs per node follow bounded_qp.py:230-332 semantics:
  status 2 (optimal): z, multipliers >= 0 of every row, cost
  status 3 (infeasible): Farkas ray y >= 0 with  Aall' y_signed = 0  and  -rhs.y > 0
"""
import numpy as np

OPTIMAL, INFEASIBLE, ITER_LIMIT = 2, 3, 9


class LDP(object):
    """min 1/2|v|^2  s.t.  bl <= M v <= bu  by a dual active-set method."""

    def __init__(self, M, row_scale=None, tol_p=1e-9, tol_d=1e-12, tol_sing=1e-7, max_iter=5000):
        # rows are normalised to unit length (multipliers are scaled back on exit);
        # violations are measured in the units of the ORIGINAL rows a_r'z <= b_r
        self.norm = np.linalg.norm(M, axis=1)
        self.norm[self.norm == 0.] = 1.       # constant rows (stage-0 state rows: no z dependence)
        self.M = M / self.norm[:, None]
        self.tol_p, self.tol_d, self.tol_sing, self.max_iter = tol_p, tol_d, tol_sing, max_iter
        self.trace = None
        self.tol_ray = 1e-9
        row_scale = np.ones(M.shape[0]) if row_scale is None else row_scale
        self.row_scale = row_scale / self.norm

    def solve(self, bl, bu, W=None, lam=None):
        """W: list of (row, side) ; lam: nonnegative multipliers in the sign-normalised form."""
        M = self.M
        bl = bl / self.norm; bu = bu / self.norm
        W = [] if W is None else list(W)
        lam = np.zeros(0) if lam is None else np.array(lam, dtype=float) * self.norm[[w[0] for w in W]]
        unscale = lambda W_, l_: l_ / self.norm[[w[0] for w in W_]] if len(W_) else l_
        it = 0
        singular = False
        ignored = set()
        while True:
            it += 1
            if it > self.max_iter:
                return ITER_LIMIT, None, W, unscale(W, lam), it
            rows = np.array([w[0] for w in W], dtype=int)
            sgn = np.array([w[1] for w in W], dtype=float)
            Mw = M[rows] * sgn[:, None] if W else np.zeros((0, M.shape[1]))
            dw = np.array([bu[r] if s > 0 else -bl[r] for r, s in W])
            if not singular:
                if W:
                    # lam* = argmin 1/2|Mw' lam|^2 + dw.lam  <=>  Mw Mw' lam = -dw
                    Qf, Rf = np.linalg.qr(Mw.T)
                    lam_star = -np.linalg.solve(Rf, np.linalg.solve(Rf.T, dw))
                    # one step of refinement on the normal equations
                    res = -dw - Mw.dot(Mw.T.dot(lam_star))
                    lam_star += np.linalg.solve(Rf, np.linalg.solve(Rf.T, res))
                else:
                    lam_star = np.zeros(0)
                neg = lam_star < -self.tol_d
                if not np.any(neg):
                    lam = np.maximum(lam_star, 0.)
                    v = -Mw.T.dot(lam) if W else np.zeros(M.shape[1])
                    Mv = M.dot(v)
                    vu = (Mv - bu) / self.row_scale
                    vl = (bl - Mv) / self.row_scale
                    vu[rows[sgn > 0]] = -np.inf
                    vl[rows[sgn < 0]] = -np.inf
                    for (r, s_) in ignored:
                        (vu if s_ > 0 else vl)[r] = -np.inf
                    ju, jl = np.argmax(vu), np.argmax(vl)
                    if max(vu[ju], vl[jl]) <= self.tol_p:
                        return OPTIMAL, v, W, unscale(W, lam), it
                    j, s = (ju, 1) if vu[ju] >= vl[jl] else (jl, -1)
                    # a row that is active on the other side cannot be added again: swap sides is
                    # impossible for bl < bu; for bl == bu the violated side replaces the present one
                    mj = M[j] * s
                    if W:
                        proj = Qf.dot(Qf.T.dot(mj))
                        dep = np.linalg.norm(mj - proj) <= self.tol_sing
                    else:
                        dep = np.linalg.norm(mj) <= self.tol_sing
                    W.append((j, s)); lam = np.append(lam, 0.)
                    if self.trace is not None: self.trace.append(('add', int(j), s, float(max(vu[ju], vl[jl])), bool(dep)))
                    singular = dep
                else:
                    p = lam_star - lam
                    cand = np.where(neg)[0]
                    ratios = lam[cand] / (lam[cand] - lam_star[cand])
                    k = cand[np.argmin(ratios)]
                    alpha = np.min(ratios)
                    lam = lam + alpha * p
                    if self.trace is not None: self.trace.append(('drop', W[k], float(alpha)))
                    lam = np.delete(lam, k); W.pop(k)
            else:
                # W = W' + [j] with mj in span(Mw'): dual ray p, p_j = 1, Mw' p = 0
                Mp = Mw[:-1]
                sol = np.linalg.lstsq(Mp.T, Mw[-1], rcond=None)[0]
                p = np.append(-sol, 1.)
                blocking = np.where(p < -self.tol_ray * max(1., np.max(np.abs(p))))[0]
                if blocking.size == 0:
                    # Farkas ray p >= 0, M~'p = 0, cost -d~.p = violation of the entering row.
                    # The rows are infeasible by cost/|p|_1 (weighted max-violation); below the
                    # primal tolerance the entering row is merely redundant-and-tight: skip it.
                    p = np.maximum(p, 0.)
                    cost = -dw.dot(p)
                    if cost > self.tol_p * np.sum(p * self.row_scale[rows]):
                        return INFEASIBLE, unscale(W, p), W, unscale(W, lam), it
                    ignored.add(W[-1]); W.pop(); lam = lam[:-1]
                    singular = False
                    continue
                ratios = lam[blocking] / (-p[blocking])
                k = blocking[np.argmin(ratios)]
                lam = lam + np.min(ratios) * p
                if self.trace is not None: self.trace.append(('raydrop', W[k], float(np.min(ratios))))
                lam = np.delete(lam, k); W.pop(k)
                singular = False


class NodeQP(object):
    """Shared data for all nodes: R'R = H + eps I,  M = Aall R^-1."""

    def __init__(self, cond, eps=1e-8, prox_tol=1e-11, max_prox=50):
        self.c = cond
        self.eps, self.prox_tol, self.max_prox = eps, prox_tol, max_prox
        He = cond.H + eps * np.eye(cond.n)
        L = np.linalg.cholesky(He)              # He = L L'
        self.Rinv = np.linalg.inv(L.T)          # z = Rinv (v - Rinv' f)
        self.M = cond.Aall.dot(self.Rinv)
        self.ldp = LDP(self.M, np.maximum(1., np.linalg.norm(cond.Aall, axis=1)))

    def solve(self, x0, lb, ub, W=None, lam=None, z0=None):
        c = self.c
        bl, bu = c.bounds(x0, lb, ub)
        f = c.Fx.dot(x0)
        z = np.zeros(c.n) if z0 is None else z0.copy()
        iters = 0
        for k in range(self.max_prox):
            fk = f - self.eps * z
            w = self.Rinv.T.dot(fk)             # v = R z + w
            g = self.M.dot(w)
            st, v, W, lam, it = self.ldp.solve(bl + g, bu + g, W, lam)
            iters += it
            if st != OPTIMAL:
                break
            z_new = self.Rinv.dot(v - w)
            dz = np.max(np.abs(z_new - z))
            z = z_new
            if self.eps * dz <= self.prox_tol:
                break
        out = dict(status=st, iters=iters, prox=k + 1, W=W, lam_w=lam)
        y = np.zeros(c.m)                        # signed multipliers: >0 upper side, <0 lower side
        if st == OPTIMAL:
            for (r, s), l in zip(W, lam):
                y[r] += s * l
            pinned = lb == ub                    # rows lb <= z_i <= ub with lb == ub hold exactly
            z[c.bin_idx[pinned]] = lb[pinned]
            out.update(z=z, y=y, cost=c.cost(x0, z))
        elif st == INFEASIBLE:
            for (r, s), l in zip(W, v):
                y[r] += s * l
            rhs = np.where(y > 0, bu, bl)
            rhs[y == 0] = 0.
            out.update(z=None, y=y, cost=np.inf, farkas=-(rhs.dot(y)))
        return out
