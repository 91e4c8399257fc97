"""TEST INFRASTRUCTURE (oracle side) -- condensed form of the node QP of controller.py:119-184.

The reference hands Gurobi the sparse stage-wise QP (variables x_t, uc_t, ub_t; rows lam_t, nu_lb_t,
nu_ub_t, mu_t).  Eliminating the states with the dynamics rows (x = Phi x0 + Gam z, z = (u_0..u_{T-1}))
gives the *shared* data of north_star: every node of every time step has the same

    cost(z) = 1/2 z'H z + (Fx x0)'z + x0'Cx x0           H = 2 (Gam'Qb'Qb Gam + Rb'Rb)
    rows    : C z <= hbar - E x0          (mu_0 .. mu_{T-1}, in the reference's row order)
    bounds  : lb <= z[bin] <= ub          (nu_lb_t / nu_ub_t rows; the only node-dependent data)

This module is numpy-only and is used by the oracle; the product builds the same matrices in
warm-start-hybrid-mpc_b200/problem.py (host precompute, north_star "MLD condensing stays host-side").
"""
import numpy as np


class Condensed(object):

    def __init__(self, model):
        m = model
        A, B, F, G, h = m['A'], m['B'], m['F'], m['G'], m['h']
        F1, G1, h1 = m['F_Tm1'], m['G_Tm1'], m['h_Tm1']
        Q, R, Q_T = m['Q'], m['R'], m['Q_T']
        T, nub = int(m['T']), int(m['nub'])
        nx, nu = B.shape
        nuc = nu - nub
        self.T, self.nx, self.nu, self.nub, self.nuc = T, nx, nu, nub, nuc
        self.n = n = T * nu
        self.nh, self.nh1 = h.size, h1.size
        self.mc = mc = (T - 1) * h.size + h1.size
        # x_t = Phi[t] x0 + Gam[t] z
        Phi = np.zeros((T + 1, nx, nx))
        Gam = np.zeros((T + 1, nx, n))
        Phi[0] = np.eye(nx)
        for t in range(T):
            Phi[t + 1] = A.dot(Phi[t])
            Gam[t + 1] = A.dot(Gam[t])
            Gam[t + 1][:, t * nu:(t + 1) * nu] += B
        self.Phi, self.Gam = Phi, Gam
        # cost
        H = np.zeros((n, n)); Fx = np.zeros((n, nx)); Cx = np.zeros((nx, nx))
        for t in range(T + 1):
            W = Q_T if t == T else Q
            WG = W.dot(Gam[t]); WP = W.dot(Phi[t])
            H += 2. * WG.T.dot(WG); Fx += 2. * WG.T.dot(WP); Cx += WP.T.dot(WP)
        RtR = R.T.dot(R)
        for t in range(T):
            H[t * nu:(t + 1) * nu, t * nu:(t + 1) * nu] += 2. * RtR
        self.H, self.Fx, self.Cx = .5 * (H + H.T), Fx, Cx
        # general rows
        C = np.zeros((mc, n)); E = np.zeros((mc, nx)); hb = np.zeros(mc)
        self.row0 = []
        r = 0
        for t in range(T):
            Ft, Gt, ht = (F, G, h) if t < T - 1 else (F1, G1, h1)
            k = ht.size
            self.row0.append(r)
            C[r:r + k] = Ft.dot(Gam[t]); C[r:r + k, t * nu:(t + 1) * nu] += Gt
            E[r:r + k] = Ft.dot(Phi[t]); hb[r:r + k] = ht
            r += k
        self.C, self.E, self.hbar = C, E, hb
        # binaries: flat index (t, i) -> z index
        self.bin_idx = np.array([t * nu + nuc + i for t in range(T) for i in range(nub)])
        self.nb = self.bin_idx.size
        self.m = mc + self.nb
        # all rows (general rows, then one two-sided row per binary)
        Ab = np.zeros((self.nb, n)); Ab[np.arange(self.nb), self.bin_idx] = 1.
        self.Aall = np.vstack((C, Ab))

    def bounds(self, x0, lb, ub):
        """two-sided row bounds bl <= Aall z <= bu for a node (lb, ub flat over (t, i))."""
        bl = np.concatenate((np.full(self.mc, -np.inf), lb))
        bu = np.concatenate((self.hbar - self.E.dot(x0), ub))
        return bl, bu

    def states(self, x0, z):
        return np.einsum('tij,j->ti', self.Phi, x0) + np.einsum('tij,j->ti', self.Gam, z)

    def cost(self, x0, z):
        return .5 * z.dot(self.H.dot(z)) + z.dot(self.Fx.dot(x0)) + x0.dot(self.Cx.dot(x0))


def identifier_to_bounds(identifier, T, nub):
    """controller.py:273-298 / 300-327."""
    lb = np.zeros(T * nub); ub = np.ones(T * nub)
    for (t, i), v in identifier.items():
        lb[t * nub + i] = v; ub[t * nub + i] = v
    return lb, ub


class OrthoForm(object):
    """The same node QP in ORTHONORMAL null-space coordinates of the dynamics rows.

    Plain condensing (class Condensed) parametrises the feasible affine subspace of the dynamics rows
    by xi = (z, X) = P0 x0 + G z with G = [I; Gam]; for open-loop-unstable A the columns of G are huge
    and nearly collinear (cond(H) ~ 1e7 .. 1e9 on CP20/CP40).  With the thin QR
    G = N T the coordinates y = T z have  xi = P0 x0 + N y,  N'N = I, and

        cost  = 1/2 y'Hy y + (Fy x0)'y + const(x0)      eig(Hy) in [0, max eig(2 Q'Q, 2 R'R, 2 Q_T'Q_T)]
        rows  : bl <= Ay y <= bu,   general rows bu = hbar - Ey x0, binary rows [lb, ub]
        z     = Zmap y

    i.e. the regularised Hessian Hy + eps I has condition <= max-eig / eps instead of 1e9 / eps, which
    is what makes the dependent / independent row decision of the active-set method clean.
    """

    def __init__(self, cond, model):
        c = cond
        T, nx, nu, n = c.T, c.nx, c.nu, c.n
        Q, R, Q_T = model['Q'], model['R'], model['Q_T']
        Gbar = np.vstack([c.Gam[t] for t in range(1, T + 1)])
        Pbar = np.vstack([c.Phi[t] for t in range(1, T + 1)])
        G = np.vstack((np.eye(n), Gbar))
        N, Tm = np.linalg.qr(G)                       # G = N Tm
        P0 = np.vstack((np.zeros((n, nx)), Pbar))
        ns = n + T * nx
        Hs = np.zeros((ns, ns))
        RtR, QtQ, QTtQT = R.T.dot(R), Q.T.dot(Q), Q_T.T.dot(Q_T)
        for t in range(T):
            Hs[t * nu:(t + 1) * nu, t * nu:(t + 1) * nu] = 2. * RtR
            blk = QtQ if t < T - 1 else QTtQT
            o = n + t * nx
            Hs[o:o + nx, o:o + nx] = 2. * blk
        As = np.zeros((c.mc, ns)); E0 = np.zeros((c.mc, nx))
        for t in range(T):
            Ft, Gt = (model['F'], model['G']) if t < T - 1 else (model['F_Tm1'], model['G_Tm1'])
            r = c.row0[t]; k = Ft.shape[0]
            As[r:r + k, t * nu:(t + 1) * nu] = Gt
            if t == 0:
                E0[r:r + k] = Ft
            else:
                o = n + (t - 1) * nx
                As[r:r + k, o:o + nx] = Ft
        Sel = np.zeros((c.nb, ns)); Sel[np.arange(c.nb), c.bin_idx] = 1.
        HsN = Hs.dot(N)
        self.n, self.m, self.mc, self.nb, self.nx = n, c.m, c.mc, c.nb, nx
        Hy = N.T.dot(HsN)
        self.Hy = .5 * (Hy + Hy.T)
        self.Fy = HsN.T.dot(P0)
        self.Ay = np.vstack((As.dot(N), Sel.dot(N)))
        self.Ey = np.vstack((E0 + As.dot(P0), np.zeros((c.nb, nx))))
        self.hbar = c.hbar
        self.Zmap = N[:n]                               # z = Zmap y  (u-part of P0 is zero)
        self.Tm = Tm
        # row scale for violations: norm of the ORIGINAL sparse row (stage row / unit vector)
        self.arow = np.concatenate((np.linalg.norm(np.hstack((As, E0 * 0.)), axis=1), np.ones(c.nb)))
