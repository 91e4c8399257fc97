"""TEST INFRASTRUCTURE (oracle side) -- ctypes binding of oracle/qp_core.c + shared-data packing.

`build()` compiles the C restatement with gcc into oracle/_build/ (git-ignored).  `CoreC(model)`
packs the shared least-distance data exactly as documented in qp_core.c and exposes
`solve(x0, lb, ub, warm=None)` returning the same dict as oracle/qp_numpy.NodeQP.solve.
"""
import ctypes as C
import os
import subprocess
import numpy as np
from oracle.condense import Condensed, OrthoForm

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, '_build', 'liboracle_qp.so')
OPTIMAL, INFEASIBLE, ITER_LIMIT = 2, 3, 9


def build(force=False):
    src = os.path.join(HERE, 'qp_core.c')
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-std=c11', '-o', SO, src, '-lm'])
    return SO


class _Shared(C.Structure):
    _fields_ = [('n', C.c_int), ('m', C.c_int), ('mc', C.c_int), ('nb', C.c_int), ('nx', C.c_int),
                ('Mh', C.c_void_p), ('nrm', C.c_void_p), ('vscale', C.c_void_p), ('Eh', C.c_void_p),
                ('hh', C.c_void_p), ('Rinv', C.c_void_p), ('Kx', C.c_void_p), ('Zmap', C.c_void_p), ('bin_idx', C.c_void_p),
                ('eps', C.c_double), ('tol_p', C.c_double), ('tol_d', C.c_double), ('tol_sing', C.c_double),
                ('tol_ray', C.c_double), ('prox_tol', C.c_double), ('max_iter', C.c_int), ('max_prox', C.c_int),
                ('variant', C.c_int)]


def pack_shared(of, eps):
    """Shared least-distance data from the orthonormal form (oracle/condense.OrthoForm):
    Rinv'(Hy + eps I) Rinv = I ;  M = Ay Rinv (rows normalised) ;  y = Rinv (v - w), w = Kx x0 - eps Rinv' y_c."""
    n = of.n
    # symmetric square root by eigen-decomposition, NOT Cholesky: U separates the curved and the
    # (regularised) null directions exactly, so every column of M = Ay U Lam^-1/2 is accurate to
    # 1e-16 relative to its own scale (a triangular factor mixes the 1/sqrt(eps) and O(10) scales)
    lamb, U = np.linalg.eigh(of.Hy)
    Rinv = U / np.sqrt(np.maximum(lamb, 0.) + eps)[None, :]
    M = of.Ay.dot(Rinv)
    nrm = np.linalg.norm(M, axis=1)
    nrm[nrm == 0.] = 1.
    d = dict(
        Mh=np.ascontiguousarray(M / nrm[:, None]), nrm=nrm,
        vscale=nrm / np.maximum(1., of.arow),
        Eh=np.ascontiguousarray(of.Ey[:of.mc] / nrm[:of.mc, None]), hh=of.hbar / nrm[:of.mc],
        Rinv=np.ascontiguousarray(Rinv), Kx=np.ascontiguousarray(Rinv.T.dot(of.Fy)),
        Zmap=np.ascontiguousarray(of.Zmap))
    return d


def default_eps(Hy):
    """Proximal weight: 1% of the smallest NON-ZERO curvature of the cost, so that the proximal-point
    outer loop contracts by >= 100x per iteration while the 1/sqrt(eps) scale separation between
    the flat and the curved directions stays as small as the problem allows."""
    ev = np.linalg.eigvalsh(Hy)
    pos = ev[ev > 1e-9 * ev.max()]
    return 1e-2 * float(pos.min())


class _State(object):
    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        try:
            self.lib.qp_state_free(C.c_void_p(self.ptr))
        except Exception:
            pass


class CoreC(object):

    def __init__(self, model, eps=None, tol_p=1e-7, tol_d=1e-12, tol_sing=1e-7, tol_ray=1e-9,
                 prox_tol=1e-11, max_iter=5000, max_prox=50, variant=0):
        self.lib = C.CDLL(build())
        self.lib.qp_solve.restype = C.c_int
        self.cond = c = Condensed(model)
        self.of = OrthoForm(c, model)
        if eps is None:
            eps = default_eps(self.of.Hy)
        self.eps = eps
        self.arr = pack_shared(self.of, eps)
        self.arr['bin_idx'] = c.bin_idx.astype(np.int32)
        S = _Shared()
        S.n, S.m, S.mc, S.nb, S.nx = c.n, c.m, c.mc, c.nb, c.nx
        for k, v in self.arr.items():
            setattr(S, k, v.ctypes.data)
        S.eps, S.tol_p, S.tol_d, S.tol_sing, S.tol_ray, S.prox_tol = eps, tol_p, tol_d, tol_sing, tol_ray, prox_tol
        S.max_iter, S.max_prox = max_iter, max_prox
        S.variant = variant
        self.S = S

    def new_state(self):
        """A persistent solver state (one CUDA slot): pass it as `state=` to solve()."""
        self.lib.qp_state_new.restype = C.c_void_p
        return _State(self.lib, self.lib.qp_state_new(self.cond.n, self.cond.m))

    def solve(self, x0, lb, ub, warm=None, state=None, reset=False):
        c = self.cond
        n, m = c.n, c.m
        x0 = np.ascontiguousarray(x0, dtype=float); lb = np.ascontiguousarray(lb, dtype=float)
        ub = np.ascontiguousarray(ub, dtype=float)
        z = np.zeros(n); yc = np.zeros(n); y = np.zeros(m); fark = C.c_double(0.)
        nW = C.c_int(0); Wr = np.zeros(n + 1, np.int32); Ws = np.zeros(n + 1, np.int32); Wl = np.zeros(n + 1)
        it = C.c_int(0); px = C.c_int(0)
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        if warm is None:
            nW0, a0, a1, a2, a3 = 0, None, None, None, None
        else:
            r0 = np.ascontiguousarray(warm['rows'], np.int32); s0 = np.ascontiguousarray(warm['sides'], np.int32)
            l0 = np.ascontiguousarray(warm['lam'], float)
            z0 = None if warm.get('z') is None else np.ascontiguousarray(warm['z'], float)
            nW0, a0, a1, a2, a3 = r0.size, P(r0), P(s0), P(l0), (None if z0 is None else P(z0))
        if state is not None:
            st = self.lib.qp_solve_state(C.byref(self.S), C.c_void_p(state.ptr), int(bool(reset)), P(x0), P(lb), P(ub),
                                         P(z), P(yc), P(y), C.byref(fark), C.byref(nW), P(Wr), P(Ws), P(Wl),
                                         C.byref(it), C.byref(px))
        else:
            st = self.lib.qp_solve(C.byref(self.S), P(x0), P(lb), P(ub), nW0, a0, a1, a2, a3,
                                   P(z), P(yc), P(y), C.byref(fark), C.byref(nW), P(Wr), P(Ws), P(Wl),
                                   C.byref(it), C.byref(px))
        k = nW.value
        out = dict(status=st, iters=it.value, prox=px.value,
                   warm=dict(rows=Wr[:k].copy(), sides=Ws[:k].copy(), lam=Wl[:k].copy(), z=yc if st == 2 else None))
        if st == OPTIMAL:
            out.update(z=z, y=y, cost=c.cost(x0, z))
        elif st == INFEASIBLE:
            out.update(z=None, y=y, cost=np.inf, farkas=fark.value)
        return out
