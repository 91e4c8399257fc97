"""MLD system container -- host-side modelling object consumed by the controller.

Mirrors the constructor contract of the reference's ``MLDSystem`` (mld_system.py:9-66): same
attribute names (A, B, F, G, h, nx, nu, nub, nuc, V) and the same ValueErrors on inconsistent sizes.
The builders `from_symbolic`, `from_pwa`, `from_symbolic_pwa` (mld_system.py:68-214) are host-side
modelling helpers: they produce the matrices the problem compiler (problem.py) consumes.
"""
import numpy as np


def sym2mat(x, expr):
    """utils.py:4-31: (J, b) with expr(x) = J x + b for an affine sympy expression (column matrix) in the symbols x."""
    jac = np.array(expr.jacobian(x)).astype(np.float64)          # raises TypeError if a symbol survives: not affine
    off = np.array(expr.subs({xi: 0 for xi in x})).astype(np.float64).ravel()
    return jac, off


def unpack_bmat(A, indices, direction):
    """utils.py:33-72: cuts A into consecutive blocks of the given widths ('h') or heights ('v')."""
    if direction not in ('h', 'v'):
        raise ValueError('Unknown unpacking direction %s.' % direction)
    edges = np.concatenate(([0], np.cumsum(indices)))
    return [A[:, a:b] if direction == 'h' else A[a:b, :] for a, b in zip(edges[:-1], edges[1:])]


class MLDSystem(object):
    """x(t+1) = A x(t) + B u(t),  F x(t) + G u(t) <= h,  last `nub` inputs binary."""

    def __init__(self, dynamics, constraints, nub):
        [self.A, self.B] = [np.asarray(M, dtype=float) for M in dynamics]
        [self.F, self.G, self.h] = [np.asarray(M, dtype=float) for M in constraints]
        self.nx = self.A.shape[1]
        self.nu = self.B.shape[1]
        self.nub = nub
        self.nuc = self.nu - nub
        self.V = np.hstack((np.zeros((nub, self.nuc)), np.eye(nub)))
        self._check_input_sizes()

    def _check_input_sizes(self):
        if self.A.shape[0] != self.A.shape[1]:
            raise ValueError('Nonsquare A matrix.')
        if self.B.shape[0] != self.nx:
            raise ValueError('A and B matrices have incompatible size.')
        if self.F.shape != (self.h.size, self.nx):
            raise ValueError('Matrix F has incompatible size.')
        if self.G.shape != (self.h.size, self.nu):
            raise ValueError('Matrix G has incompatible size.')

    # -- builders (mld_system.py:68-214) ---------------------------------------------------------------
    @staticmethod
    def from_symbolic(dynamics, constraints, x, u, nub):
        """mld_system.py:68-108: `dynamics` = next state (LINEAR in (x, u)), `constraints` <= 0 (affine), as sympy column
        matrices in the symbols x, u; the last `nub` inputs are binary."""
        import sympy as sp
        v = sp.Matrix([x, u])
        widths = [x.shape[0], u.shape[0]]
        J, off = sym2mat(v, dynamics)
        if not np.allclose(off, 0.):
            raise ValueError('The dynamics seems to be affine and not linear.')
        A, B = unpack_bmat(J, widths, 'h')
        J, off = sym2mat(v, constraints)
        F, G = unpack_bmat(J, widths, 'h')
        return MLDSystem([A, B], [F, G, -off], nub)

    @staticmethod
    def from_pwa(dynamics, domains):
        """mld_system.py:110-184: piecewise-affine system  x+ = A_i x + B_i u + c_i  on  F_i x + G_i u <= h_i  (mode i)
        as an MLD system by the convex-hull formulation.  MLD input = (u, x_1..x_I, u_1..u_I, mu_1..mu_I) with the mode
        indicators mu binary: x+ = sum_i A_i x_i + B_i u_i + c_i mu_i, F_i x_i + G_i u_i <= h_i mu_i, x = sum x_i,
        u = sum u_i, sum mu_i = 1.  Rows: mode domains, then x = sum x_i as two inequalities, u = sum u_i, sum mu = 1."""
        I = len(dynamics)
        nx, nu = dynamics[0][0].shape[0], dynamics[0][1].shape[1]
        rows = [np.asarray(d[0]).shape[0] for d in domains]
        nc = sum(rows)
        # columns of the MLD input
        cu = slice(0, nu)
        cx = [slice(nu + i * nx, nu + (i + 1) * nx) for i in range(I)]
        cv = [slice(nu + I * nx + i * nu, nu + I * nx + (i + 1) * nu) for i in range(I)]
        cm = [nu + I * (nx + nu) + i for i in range(I)]
        width = nu + I * (nx + nu) + I
        B = np.zeros((nx, width))
        for i, (Ai, Bi, ci) in enumerate(dynamics):
            B[:, cx[i]] = Ai; B[:, cv[i]] = Bi; B[:, cm[i]] = np.asarray(ci, dtype=float).ravel()
        F = np.zeros((nc + 2 * nx + 2 * nu + 2, nx)); G = np.zeros((nc + 2 * nx + 2 * nu + 2, width)); h = np.zeros(F.shape[0])
        r = 0
        for i, (Fi, Gi, hi) in enumerate(domains):                       # F_i x_i + G_i u_i - h_i mu_i <= 0
            G[r:r + rows[i], cx[i]] = Fi; G[r:r + rows[i], cv[i]] = Gi; G[r:r + rows[i], cm[i]] = -np.asarray(hi, dtype=float).ravel()
            r += rows[i]
        for sgn in (1., -1.):                                            # x - sum x_i  (<= 0 and >= 0)
            F[r:r + nx] = sgn * np.eye(nx)
            for i in range(I):
                G[r:r + nx, cx[i]] = -sgn * np.eye(nx)
            r += nx
        for sgn in (1., -1.):                                            # u - sum u_i
            G[r:r + nu, cu] = sgn * np.eye(nu)
            for i in range(I):
                G[r:r + nu, cv[i]] = -sgn * np.eye(nu)
            r += nu
        for sgn in (1., -1.):                                            # sum mu_i = 1
            G[r, cm] = sgn; h[r] = sgn
            r += 1
        return MLDSystem([np.zeros((nx, nx)), B], [F, G, h], I)

    @staticmethod
    def from_symbolic_pwa(dynamics_sym, domains_sym, x, u):
        """mld_system.py:186-214: every mode given by sympy expressions (next state affine in (x, u), domain <= 0)."""
        import sympy as sp
        v = sp.Matrix([x, u])
        widths = [x.shape[0], u.shape[0]]
        dynamics, domains = [], []
        for d in dynamics_sym:
            J, c = sym2mat(v, d)
            dynamics.append(unpack_bmat(J, widths, 'h') + [c])
        for d in domains_sym:
            J, off = sym2mat(v, d)
            domains.append(unpack_bmat(J, widths, 'h') + [-off])
        return MLDSystem.from_pwa(dynamics, domains)
