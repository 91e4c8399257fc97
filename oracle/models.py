"""TEST INFRASTRUCTURE (oracle side) -- benchmark instance builders.

Builds the problem instances of BASELINE.json / SURVEY.md section 8(d) and freezes them to
``warm-start-hybrid-mpc_b200/data/model_*.npz`` (product data; golden VECTORS stay in tests/golden/) so that CPU oracle and GPU path consume identical bits.

* CP20 / CP40: the notebook two-wall cart-pole.  The MLD matrices come from the reference's
  own ``notebooks/cart_pole_with_walls/mld_dynamics.py`` (imported unmodified, sympy);
  weights / LQR terminal cost as in ``notebooks/cart_pole_with_walls/controller.py:8-21``;
  MCAIS terminal set by `mcais` below (restatement of ``warm_start_hmpc/mcais.py:44-184``
  over scipy/HiGHS LPs because the reference routes these LPs through gurobipy).
* CP1W40: the unit-test fixture ``warm_start_hmpc/test/cart_pole_with_wall.py:1-117`` (one wall, T=40).
* SYN: synthetic random MLD (the reference has no generator).

The update matrices of the warm start (``controller.py:94-97, 186-227``) are host-side
precompute and are stored with the model: ``M_mu`` (LPs) and ``M_rho = pinv(Q')Q_T'``.
"""
import os
import numpy as np
from scipy.optimize import linprog
from scipy.linalg import solve_discrete_are

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(_ROOT, 'tests', 'golden')
MODELS = os.path.join(_ROOT, 'warm-start-hybrid-mpc_b200', 'data')


# ---------------------------------------------------------------------------------------------
# restatements of host-side precompute (mcais.py, controller.py:186-227)
# ---------------------------------------------------------------------------------------------
def solve_dare(A, B, Q, R):
    """mcais.py:10-42."""
    P = solve_discrete_are(A, B, Q, R)
    K = - np.linalg.inv(B.T.dot(P).dot(B) + R).dot(B.T).dot(P).dot(A)
    return P, K


def _lp_max(c, D, e):
    """max c.x s.t. D x <= e  (free x)."""
    res = linprog(-c, A_ub=D, b_ub=e, bounds=[(None, None)] * D.shape[1], method='highs')
    if res.status != 0:
        raise RuntimeError('LP failed: ' + res.message)
    return -res.fun


def remove_redundant_inequalities(E, f, tol=1.e-7):
    """mcais.py:147-184: facet i is kept iff max E_i x s.t. E x <= f + e_i exceeds f_i by >= tol.
    (As in the reference, facets found redundant are NOT removed from the LP, only from the list.)"""
    nc = E.shape[0]
    minimal = []
    for i in range(nc):
        f_add = np.zeros(nc)
        f_add[i] = 1.
        if not (_lp_max(E[i], E, f + f_add) - f[i] < tol):
            minimal.append(i)
    return E[minimal], f[minimal]


def mcais(A, D, e):
    """mcais.py:44-145 (Gilbert & Tan algorithm 3.2)."""
    if np.max(np.absolute(np.linalg.eig(A)[0])) > 1.:
        raise ValueError('Unstable system, cannot derive maximal constraint-admissible set.')
    if np.min(e) < 0.:
        raise ValueError('The origin is not in the constraint set, cannot derive maximal constraint-admissible set.')
    D_inf, e_inf = D.copy(), e.copy()
    t = 1
    while True:
        J = D.dot(np.linalg.matrix_power(A, t))
        residuals = [_lp_max(J[i], D_inf, e_inf) - e[i] for i in range(D.shape[0])]
        new = [i for i, r in enumerate(residuals) if r > 0.]
        if not new:
            break
        D_inf = np.vstack((D_inf, J[new]))
        e_inf = np.concatenate((e_inf, e[new]))
        t += 1
    return remove_redundant_inequalities(D_inf, e_inf)


def update_mu(F, G, h, F_Tm1, G_Tm1):
    """controller.py:186-227:  M[:, i] = argmin h.mu s.t. F'mu = F_Tm1[i], G'mu = G_Tm1[i], mu >= 0."""
    n = h.size
    Aeq = np.vstack((F.T, G.T))
    cols = []
    for i in range(F_Tm1.shape[0]):
        beq = np.concatenate((F_Tm1[i], G_Tm1[i]))
        res = linprog(h, A_eq=Aeq, b_eq=beq, bounds=[(0, None)] * n, method='highs')
        if res.status != 0:
            raise ValueError('The conic hull of [F G] does not contain the one of [F_Tm1 G_Tm1].')
        cols.append(res.x)
    M = np.vstack(cols).T
    # rows of [F_Tm1 G_Tm1] that are literally rows of [F G] get the exact unit vector
    # (the LP returns it up to ~1e-17 noise; test_controller.py:52 expects the identity exactly)
    for i in range(min(n, F_Tm1.shape[0])):
        if np.array_equal(F_Tm1[i], F[i]) and np.array_equal(G_Tm1[i], G[i]):
            ei = np.zeros(n); ei[i] = 1.
            if abs(h.dot(M[:, i]) - h[i]) <= 1e-9 * max(1., abs(h[i])):
                M[:, i] = ei
    return M


def pack_model(name, A, B, F, G, h, nub, T, Q, R, Q_T, F_T, h_T, meta=None):
    """Controller-level data exactly as HybridModelPredictiveController.__init__ derives it
    (controller.py:76-97)."""
    F_Tm1 = np.vstack((F, F_T.dot(A)))
    G_Tm1 = np.vstack((G, F_T.dot(B)))
    h_Tm1 = np.concatenate((h, h_T))
    M_mu = update_mu(F, G, h, F_Tm1, G_Tm1)
    M_rho = np.linalg.pinv(Q.T).dot(Q_T.T)
    d = dict(name=name, A=A, B=B, F=F, G=G, h=h, nub=nub, T=T, Q=Q, R=R, Q_T=Q_T,
             F_T=F_T, h_T=h_T, F_Tm1=F_Tm1, G_Tm1=G_Tm1, h_Tm1=h_Tm1, M_mu=M_mu, M_rho=M_rho)
    if meta:
        d.update(meta)
    return d


def save_model(d):
    os.makedirs(MODELS, exist_ok=True)
    np.savez_compressed(os.path.join(MODELS, 'model_%s.npz' % d['name']), **d)


def load_model(name):
    z = np.load(os.path.join(MODELS, 'model_%s.npz' % name), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d['name'] = str(d['name'])
    d['nub'] = int(d['nub'])
    d['T'] = int(d['T'])
    return d


# ---------------------------------------------------------------------------------------------
# instances
# ---------------------------------------------------------------------------------------------
def build_cart_pole_two_walls(T, name):
    """Needs /root/reference (authoring container only)."""
    import sys
    from oracle.refload import install_stubs, REFERENCE_ROOT, NOTEBOOK_DIR
    install_stubs()
    for p in (REFERENCE_ROOT, NOTEBOOK_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    import mld_dynamics as md            # reference file, unmodified
    mld, hstep = md.mld, md.h
    # notebooks/cart_pole_with_walls/controller.py:11-27
    Q = np.eye(mld.nx) * hstep
    R = np.vstack([1.] + [0.] * (mld.nu - 1)).T * hstep
    Bu = mld.B[:, :1]
    Ru = R[:, :1]
    P, K = solve_dare(mld.A, Bu, Q.dot(Q), Ru.dot(Ru))
    Q_T = np.linalg.cholesky(P).T
    A_cl = mld.A + Bu.dot(K)
    lhs = mld.F + mld.G[:, :1].dot(K)
    F_T, h_T = mcais(A_cl, lhs, mld.h)
    return pack_model(name, mld.A, mld.B, mld.F, mld.G, mld.h, mld.nub, T, Q, R, Q_T, F_T, h_T,
                      meta=dict(x_max=md.x_max, x0_nominal=np.array([0., 0., 1., 0.])))


def build_cart_pole_one_wall_test_fixture(name='cp1w40'):
    """warm_start_hmpc/test/cart_pole_with_wall.py:1-117, built through the reference's own
    MLDSystem.from_symbolic (the fixture module itself cannot be imported: it solves at import)."""
    import sys
    import sympy as sp
    from oracle.refload import import_reference
    _, _, _, mlds = import_reference()
    mc = 1.; mp = 1.; l = 1.; d = .5; k = 100.; nu = 30.; g = 10.; h = .05
    x = sp.Matrix(sp.symbols('q t qd td'))
    u = sp.Matrix([sp.symbols('u')])
    f = sp.Matrix([sp.symbols('f')])
    b = sp.Matrix(sp.symbols('el dam'))
    inputs = sp.Matrix([u, f, b])
    x2_dot = x[1] * g * mp / mc + u[0] / mc
    x3_dot = x[1] * g * (mc + mp) / (l * mc) + u[0] / (l * mc) + f[0] / (l * mp)
    dynamics = sp.Matrix([x[0] + h * x[2], x[1] + h * x[3], x[2] + h * x2_dot, x[3] + h * x3_dot])
    x_max = np.array([d, np.pi / 8., 2., 1.]); x_min = -x_max
    u_max = np.array([2.]); u_min = -u_max
    p = x[0] - l * x[1] - d; p_dot = x[2] - l * x[3]
    p_min = x_min[0] - l * x_max[1] - d; p_max = x_max[0] - l * x_min[1] - d
    p_dot_min = x_min[2] - l * x_max[3]; p_dot_max = x_max[2] - l * x_min[3]
    f_min = k * p_min + nu * p_dot_min; f_max = k * p_max + nu * p_dot_max
    contacts = [
        sp.Matrix([p_min * (1. - b[0]) - p]), sp.Matrix([p - p_max * b[0]]),
        sp.Matrix([f_min * (1. - b[1]) - k * p - nu * p_dot]), sp.Matrix([k * p + nu * p_dot - f_max * b[1]]),
        sp.Matrix([- f[0]]), sp.Matrix([f[0] - f_max * b[0]]), sp.Matrix([f[0] - f_max * b[1]]),
        sp.Matrix([k * p + nu * p_dot + nu * p_dot_max * (b[0] - 1.) - f[0]]),
        sp.Matrix([f[0] - k * p - nu * p_dot - f_min * (b[1] - 1.)])]
    constraints = sp.Matrix([
        x - x_max.reshape(4, 1), x_min.reshape(4, 1) - x,
        u - u_max.reshape(1, 1), u_min.reshape(1, 1) - u, sp.Matrix(contacts)])
    mld = mlds.MLDSystem.from_symbolic(dynamics, constraints, x, inputs, b.shape[0])
    T = 40
    Q = np.eye(mld.nx)
    R = np.vstack([1.] + [0.] * (mld.nu - 1)).T
    Q_T = Q * 1.1
    F_T = np.vstack((np.eye(mld.nx), -np.eye(mld.nx)))
    h_T = np.concatenate((x_max, x_max)) / 1.1
    return pack_model(name, mld.A, mld.B, mld.F, mld.G, mld.h, mld.nub, T, Q, R, Q_T, F_T, h_T,
                      meta=dict(x_max=x_max, x0_nominal=np.array([0., 0., 1., 0.])))


def build_synthetic(name='syn30', nx=20, nuc=4, nub=8, T=30, seed=0):
    """SURVEY.md section 8(d)5 (the reference ships no generator)."""
    rng = np.random.default_rng(seed)
    hstep = .05
    A = np.eye(nx) + 0.05 * rng.standard_normal((nx, nx)) / np.sqrt(nx)
    A *= 1.02 / np.max(np.abs(np.linalg.eigvals(A)))
    Bc = 0.05 * rng.standard_normal((nx, nuc))
    nu = nuc + nub
    B = np.hstack((Bc, np.zeros((nx, nub))))
    rows_F, rows_G, rows_h = [], [], []
    I = np.eye(nx)
    for s in (1., -1.):
        rows_F.append(s * I); rows_G.append(np.zeros((nx, nu))); rows_h.append(np.ones(nx))
    Iu = np.hstack((np.eye(nuc), np.zeros((nuc, nub))))
    for s in (1., -1.):
        rows_F.append(np.zeros((nuc, nx))); rows_G.append(s * Iu); rows_h.append(np.ones(nuc))
    for j in range(nub):
        a = rng.standard_normal(nx); a /= np.linalg.norm(a, 1)      # |a'x| <= 1 on the state box
        c = 0.3 * rng.uniform(-1, 1)                                  # threshold: b_j = 1 iff a'x >= c
        pmin, pmax = -1. - c, 1. - c
        ej = np.zeros(nu); ej[nuc + j] = 1.
        k = j % nuc
        ek = np.zeros(nu); ek[k] = 1.
        # p_min (1-b) - p <= 0 ; p - p_max b <= 0   with p = a'x - c   (mld_dynamics.py:104-106 pattern)
        rows_F.append(-a[None]); rows_G.append((-pmin * ej)[None]); rows_h.append(np.array([-pmin - c]))
        rows_F.append(a[None]); rows_G.append((-pmax * ej)[None]); rows_h.append(np.array([c]))
        # mode-dependent input limits: uc_k <= 1 - 0.7 b ; -uc_k <= 1 - 0.7 (1-b)
        rows_F.append(np.zeros((1, nx))); rows_G.append((ek + .7 * ej)[None]); rows_h.append(np.array([1.]))
        rows_F.append(np.zeros((1, nx))); rows_G.append((-ek - .7 * ej)[None]); rows_h.append(np.array([.3]))
    F = np.vstack(rows_F); G = np.vstack(rows_G); h = np.concatenate(rows_h)
    Q = np.eye(nx) * hstep
    R = np.hstack((np.eye(nuc), np.zeros((nuc, nub)))) * hstep
    P, _ = solve_dare(A, Bc, Q.dot(Q), (R[:, :nuc]).T.dot(R[:, :nuc]))
    Q_T = np.linalg.cholesky(P).T
    F_T = np.empty((0, nx)); h_T = np.empty(0)
    return pack_model(name, A, B, F, G, h, nub, T, Q, R, Q_T, F_T, h_T,
                      meta=dict(x_max=np.ones(nx), x0_nominal=0.5 * rng.uniform(-1, 1, nx)))


if __name__ == '__main__':
    import time
    t = time.time()
    cp20 = build_cart_pole_two_walls(20, 'cp20'); save_model(cp20)
    print('cp20: terminal facets', cp20['h_T'].size, 'h_Tm1', cp20['h_Tm1'].size, '%.1fs' % (time.time() - t))
    cp40 = dict(cp20); cp40['name'] = 'cp40'; cp40['T'] = 40; save_model(cp40)
    fx = build_cart_pole_one_wall_test_fixture(); save_model(fx)
    print('cp1w40:', fx['F'].shape, fx['h_Tm1'].size)
    syn = build_synthetic(); save_model(syn)
    print('syn30:', syn['F'].shape, '%.1fs' % (time.time() - t))
