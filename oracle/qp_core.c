/* TEST INFRASTRUCTURE (oracle side): sequential C restatement of the node-QP solve.
 *
 * Stands in for the Gurobi call the reference makes once per branch-and-bound node
 * (warm_start_hmpc/bounded_qp.py:200-228, reached from controller.py:229-271).  gurobipy is a
 * third-party, closed-source dependency, unpinned (setup.py:16-20) and absent from /root/reference;
 * the algorithm restated here is the published dual active-set method for the least-distance form
 * of a strictly convex QP (Goldfarb & Idnani 1983; Arnstrom, Bemporad & Axehill 2022 "DAQP"), with a
 * proximal-point outer loop because the condensed Hessian of this problem is only positive
 * SEMI-definite.  Parity of this file is pinned in tests/ against
 *   (i)  the known-answer tests of warm_start_hmpc/test/test_bounded_qp.py:104-189,
 *   (ii) KKT / Farkas certificates (test/cart_pole_with_wall.py:171-268) <= 1e-8,
 *   (iii) the golden node counts 160 / 77 of notebooks/cart_pole_with_walls/data/ when it drives the
 *        reference's own branch_and_bound (oracle/make_golden.py),
 *   (iv) the slow numpy referee oracle/qp_numpy.py (re-factorises by QR at every iteration).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this file.  The product (warm-start-hybrid-mpc_b200/csrc) never links it.
 *
 * Problem (all nodes share everything but x0, lb, ub):
 *     min_v 1/2 |v|^2   s.t.   bl_r <= mh_r . v <= bu_r ,  r = 0..m-1        (rows normalised)
 * with y = Rinv (v - w),  w = Kx x0 - eps Rinv' y_c  (y_c = proximal centre), where y are the
 * ORTHONORMAL null-space coordinates of the dynamics rows (oracle/condense.py OrthoForm; this is
 * what keeps cond(H + eps I) at max-eig(Q'Q, R'R)/eps instead of 1e9/eps), and z = Zmap y.
 * The warm-start argument z0 / the proximal centre are in y coordinates.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define QP_OPTIMAL 2
#define QP_INFEASIBLE 3
#define QP_ITER_LIMIT 9

typedef struct {
    int n, m, mc, nb, nx;
    const double *Mh;      /* m x n, unit rows (zero rows stay zero) */
    const double *nrm;     /* m : |a_r Rinv| (1 for zero rows) */
    const double *vscale;  /* m : nrm_r / max(1, |a_r|) -> violation in units of the original row */
    const double *Eh;      /* mc x nx : E / nrm */
    const double *hh;      /* mc : hbar / nrm */
    const double *Rinv;    /* n x n dense, row major :  y = Rinv (v - w),  Rinv'(Hy + eps I) Rinv = I */
    const double *Kx;      /* n x nx : Rinv' Fy */
    const double *Zmap;    /* n x n : z = Zmap y  (inputs from orthonormal coordinates) */
    const int *bin_idx;    /* nb : position of binary (t,i) in z */
    double eps, tol_p, tol_d, tol_sing, tol_ray, prox_tol;
    int max_iter, max_prox;
    int variant;           /* 0: full Q (Householder) + R ; 1: thin Q1 (Gram-Schmidt, re-orthogonalised) + R^-1 only */
} qp_shared;

typedef struct {
    int nW, n;
    int *row, *side;       /* n+1 */
    double *lam;           /* n+1 : multipliers of the sign-normalised rows, >= 0 */
    double *Q;             /* n x n orthogonal, COLUMN major: column j at Q + j*n */
    double *R;             /* (n+1) x (n+1) upper triangular, row major, ld = n+1 :  Mw' = Q[:, :nW] R */
    int ld;
} qp_ws;

static double dot(const double *a, const double *b, int n) {
    double s = 0.; for (int i = 0; i < n; ++i) s += a[i] * b[i]; return s;
}

/* t = R^-1 c  (leading k x k block) */
static void r_backsolve(const qp_ws *w, int k, const double *c, double *t) {
    const int ld = w->ld;
    for (int i = k - 1; i >= 0; --i) {
        double s = c[i];
        for (int j = i + 1; j < k; ++j) s -= w->R[i * ld + j] * t[j];
        t[i] = s / w->R[i * ld + i];
    }
}

/* u = R^-T c */
static void rt_forwardsolve(const qp_ws *w, int k, const double *c, double *u) {
    const int ld = w->ld;
    for (int i = 0; i < k; ++i) {
        double s = c[i];
        for (int j = 0; j < i; ++j) s -= w->R[j * ld + i] * u[j];
        u[i] = s / w->R[i * ld + i];
    }
}

/* try to append sign-normalised row (r, s).  c = Q' mj ; the part of mj outside span(Mw') has norm
 * |c[k:]|.  Returns 1 if appended; 0 if the row is (numerically) dependent, with t = R^-1 c[:k]
 * (mj = Mw' t) for the dual ray. */
static int qr_append(const qp_shared *S, qp_ws *w, int r, int s, double *t, double *c, double *hv) {
    const int n = S->n, k = w->nW, ld = w->ld;
    const double *mj = S->Mh + (size_t)r * n;
    for (int j = 0; j < n; ++j) c[j] = (double)s * dot(w->Q + (size_t)j * n, mj, n);
    double rho2 = 0.;
    for (int j = k; j < n; ++j) rho2 += c[j] * c[j];
    if (k >= n || rho2 <= S->tol_sing * S->tol_sing) { r_backsolve(w, k, c, t); return 0; }
    const double rho = sqrt(rho2);
    /* Householder H = I - beta hv hv' on the trailing n-k coordinates: H c[k:] = -sign(c_k) rho e_1 */
    const double sg = c[k] >= 0. ? 1. : -1.;
    for (int j = k; j < n; ++j) hv[j] = c[j];
    hv[k] += sg * rho;
    double hh = 0.; for (int j = k; j < n; ++j) hh += hv[j] * hv[j];
    const double beta = 2. / hh;
    /* Q2 <- Q2 H = Q2 - beta (Q2 hv) hv' */
    for (int i = 0; i < n; ++i) {
        double a = 0.;
        for (int j = k; j < n; ++j) a += w->Q[(size_t)j * n + i] * hv[j];
        a *= beta;
        for (int j = k; j < n; ++j) w->Q[(size_t)j * n + i] -= a * hv[j];
    }
    for (int i = 0; i < k; ++i) w->R[i * ld + k] = c[i];
    w->R[k * ld + k] = -sg * rho;
    w->row[k] = r; w->side[k] = s; w->lam[k] = 0.;
    w->nW = k + 1;
    return 1;
}

/* remove position k: delete column k of R, restore triangularity by Givens rotations (applied to
 * the columns of Q as well) */
static void qr_remove(qp_ws *w, int k) {
    const int n = w->n, nW = w->nW, ld = w->ld;
    for (int j = k + 1; j < nW; ++j) {
        for (int i = 0; i <= j; ++i) w->R[i * ld + (j - 1)] = w->R[i * ld + j];
        w->row[j - 1] = w->row[j]; w->side[j - 1] = w->side[j]; w->lam[j - 1] = w->lam[j];
    }
    for (int i = k; i < nW - 1; ++i) {
        /* zero R[i+1][i] with a rotation of rows i, i+1 */
        const double a = w->R[i * ld + i], b = w->R[(i + 1) * ld + i];
        const double h = hypot(a, b);
        if (h == 0.) continue;
        const double cs = a / h, sn = b / h;
        for (int j = i; j < nW - 1; ++j) {
            const double x = w->R[i * ld + j], y = w->R[(i + 1) * ld + j];
            w->R[i * ld + j] = cs * x + sn * y;
            w->R[(i + 1) * ld + j] = -sn * x + cs * y;
        }
        double *qi = w->Q + (size_t)i * n, *qj = w->Q + (size_t)(i + 1) * n;
        for (int c = 0; c < n; ++c) {
            const double x = qi[c], y = qj[c];
            qi[c] = cs * x + sn * y;
            qj[c] = -sn * x + cs * y;
        }
    }
    w->nW = nW - 1;
}

/* remove position k, then keep removing rows whose diagonal of R collapsed: a subset of an
 * independent set can be NEARLY dependent, and a tiny |R_ii| would amplify rounding by 1/R_ii^2.
 * Dropped rows lose their multiplier (lam >= 0 stays dual feasible); if still needed they re-enter
 * through the dependent-row (ray) branch. */
static void ws_remove(const qp_shared *S, qp_ws *w, int k, int *inW) {
    inW[w->row[k]] = 0;
    qr_remove(w, k);
    for (;;) {
        int bad = -1;
        for (int i = k; i < w->nW; ++i) if (fabs(w->R[i * w->ld + i]) <= S->tol_sing) { bad = i; break; }
        if (bad < 0) break;
        inW[w->row[bad]] = 0;
        qr_remove(w, bad);
        k = bad;
    }
}


/* ------------------------------------------------------------------------------------------------
 * variant 1: THIN factorisation  Mw' = Q1 R  kept as  Q1 (n x k, the first k columns of w->Q) and
 * Ri = R^-1 (packed by columns: column j at Ri + j(j+1)/2), no R and no null-space basis.  This is the
 * layout the CUDA kernel uses (csrc/qp_device.cuh): appends are classical Gram-Schmidt with one
 * re-orthogonalisation pass (Daniel, Gragg, Kaufman & Stewart 1976), removals rotate the columns of
 * Q1 and Ri with the Givens sequence that carries row kp of Ri into its last entry -- every rotation
 * is known up front from prefix sums of squares of that row, so nothing is a sequential chain of
 * square roots.  Same pivoting rules as variant 0; the two variants must agree (tests/).
 * ---------------------------------------------------------------------------------------------- */
#define TRI(j) ((size_t)(j) * ((size_t)(j) + 1) / 2)

/* t = Ri c  (leading k x k) */
static void thin_ri_mul(const double *Ri, int k, const double *c, double *t) {
    for (int i = 0; i < k; ++i) {
        double s = 0.;
        for (int j = i; j < k; ++j) s += Ri[TRI(j) + i] * c[j];
        t[i] = s;
    }
}

/* u = Ri' c */
static void thin_rit_mul(const double *Ri, int k, const double *c, double *u) {
    for (int j = 0; j < k; ++j) {
        double s = 0.;
        for (int i = 0; i <= j; ++i) s += Ri[TRI(j) + i] * c[i];
        u[j] = s;
    }
}

static int thin_append(const qp_shared *S, qp_ws *w, double *Ri, int r, int s, double *t, double *c1, double *z, double *c2) {
    const int n = S->n, k = w->nW;
    const double *mj = S->Mh + (size_t)r * n;
    for (int i = 0; i < n; ++i) z[i] = (double)s * mj[i];
    for (int j = 0; j < k; ++j) c1[j] = dot(w->Q + (size_t)j * n, z, n);
    for (int j = 0; j < k; ++j) { const double cj = c1[j]; const double *q = w->Q + (size_t)j * n; for (int i = 0; i < n; ++i) z[i] -= cj * q[i]; }
    double rho2 = dot(z, z, n);
    if (rho2 < 1e-2) {          /* lost more than one digit: project once more ("twice is enough") */
        for (int j = 0; j < k; ++j) c2[j] = dot(w->Q + (size_t)j * n, z, n);
        for (int j = 0; j < k; ++j) { const double cj = c2[j]; const double *q = w->Q + (size_t)j * n; for (int i = 0; i < n; ++i) z[i] -= cj * q[i]; c1[j] += cj; }
        rho2 = dot(z, z, n);
    }
    thin_ri_mul(Ri, k, c1, t);
    if (k >= n || rho2 <= S->tol_sing * S->tol_sing) return 0;
    const double rho = sqrt(rho2), ir = 1. / rho;
    double *qk = w->Q + (size_t)k * n, *rk = Ri + TRI(k);
    for (int i = 0; i < n; ++i) qk[i] = z[i] * ir;
    for (int i = 0; i < k; ++i) rk[i] = -t[i] * ir;
    rk[k] = ir;
    w->row[k] = r; w->side[k] = s; w->lam[k] = 0.;
    w->nW = k + 1;
    return 1;
}

/* remove position kp; returns the smallest position >= kp whose diagonal of R collapsed (|Ri_ii| >= 1/tol), or -1 */
static int thin_remove(const qp_shared *S, qp_ws *w, double *Ri, int kp, double *gc, double *gs, double *rowbuf) {
    const int n = w->n, k = w->nW;
    /* rotations i = kp .. k-2 on the column pairs (i, i+1): they carry row kp of Ri into its last entry */
    double run = Ri[TRI(kp) + kp] * Ri[TRI(kp) + kp];
    double tau = Ri[TRI(kp) + kp];
    for (int i = kp; i < k - 1; ++i) {
        const double wj = Ri[TRI(i + 1) + kp];
        run += wj * wj;
        const double sj = sqrt(run);
        gc[i] = wj / sj; gs[i] = -tau / sj;
        tau = sj;
    }
    for (int r = 0; r < n; ++r) {
        double carry = w->Q[(size_t)kp * n + r];
        for (int i = kp; i < k - 1; ++i) {
            const double b = w->Q[(size_t)(i + 1) * n + r];
            w->Q[(size_t)i * n + r] = gc[i] * carry + gs[i] * b;
            carry = -gs[i] * carry + gc[i] * b;
        }
    }
    int bad = -1;
    /* Ri: delete row kp, rotate the columns, drop the last column.  Rows are processed top-down so that
     * new row rr (= old row rr or rr+1) can be written in place once old row ro has been buffered. */
    for (int rr = 0; rr < k - 1; ++rr) {
        const int ro = rr < kp ? rr : rr + 1;
        for (int j = kp; j < k; ++j) rowbuf[j] = ro <= j ? Ri[TRI(j) + ro] : 0.;
        double carry = rowbuf[kp];
        for (int i = kp; i < k - 1; ++i) {
            const double b = rowbuf[i + 1];
            const double out = gc[i] * carry + gs[i] * b;
            if (rr <= i) Ri[TRI(i) + rr] = out;
            carry = -gs[i] * carry + gc[i] * b;
            if (rr == i && bad < 0 && fabs(out) * S->tol_sing >= 1.) bad = rr;
        }
    }
    for (int j = kp + 1; j < k; ++j) { w->row[j - 1] = w->row[j]; w->side[j - 1] = w->side[j]; w->lam[j - 1] = w->lam[j]; }
    w->nW = k - 1;
    return bad;
}

static void thin_ws_remove(const qp_shared *S, qp_ws *w, double *Ri, int kp, int *inW, double *gc, double *gs, double *rowbuf) {
    inW[w->row[kp]] = 0;
    int bad = thin_remove(S, w, Ri, kp, gc, gs, rowbuf);
    while (bad >= 0) {
        inW[w->row[bad]] = 0;
        bad = thin_remove(S, w, Ri, bad, gc, gs, rowbuf);
    }
}

size_t qp_work_doubles(int n, int m) {
    return (size_t)(n + 1) * (n + 1) + (size_t)n * n + 20 * (size_t)(n + 1) + 4 * (size_t)m + 64;
}

/* returns status.  Outputs: z (n), ycoord (n, the same point in orthonormal coordinates), y (m, signed multipliers of the ORIGINAL rows: > 0 upper side,
 * < 0 lower side; Farkas ray if infeasible), farkas (cost of the ray), final working set.
 * buf / ibuf: workspace (qp_work_doubles / qp_work_ints); keep != 0: the workspace holds the state left by the
 * previous solve (factor, working set, multipliers, proximal centre) and the solve continues from it -- this is
 * what a CUDA solver slot does between the nodes of one instance. */
size_t qp_work_ints(int n, int m) { return 2 * (size_t)(n + 1) + 3 * (size_t)m + 4; }

static int qp_solve_impl(const qp_shared *S, double *buf, int *ibuf, int keep, const double *x0, const double *lb, const double *ub,
             int nW0, const int *W0row, const int *W0side, const double *lam0, const double *z0,
             double *z, double *ycoord, double *y, double *farkas,
             int *nWout, int *Wrow, int *Wside, double *Wlam, int *iters, int *prox_iters)
{
    const int n = S->n, m = S->m, mc = S->mc, nb = S->nb, nx = S->nx;
    qp_ws w;
    w.ld = n + 1;
    static const double tenpow[3] = {1., 10., 100.};
    double *p = buf;
    w.R = p; p += (size_t)(n + 1) * (n + 1);
    w.Q = p; p += (size_t)n * n;
    w.n = n;
    if (!keep) {
        memset(w.Q, 0, sizeof(double) * (size_t)n * n);
        for (int i = 0; i < n; ++i) w.Q[(size_t)i * n + i] = 1.;
    }
    double *hv = p; p += n + 1;
    w.lam = p; p += n + 1;
    double *lstar = p; p += n + 1;
    double *t = p; p += n + 1;
    double *tmp = p; p += n + 1;
    double *dW = p; p += n + 1;
    double *v = p; p += n + 1;
    double *wv = p; p += n + 1;
    double *res = p; p += n + 1;
    double *zc = p; p += n + 1;
    double *bl = p; p += m;
    double *bu = p; p += m;
    double *g = p; p += m;
    /* variant 1 (thin factor): Ri aliases the storage of R */
    double *Ri = w.R;
    double *gc = p; p += n + 1;
    double *gs = p; p += n + 1;
    double *rowbuf = p; p += n + 1;
    double *zt = p; p += n + 1;
    double *c2 = p; p += n + 1;
    const int thin = S->variant == 1;
#define APPEND(r_, s_) (thin ? thin_append(S, &w, Ri, (r_), (s_), t, tmp, zt, c2) : qr_append(S, &w, (r_), (s_), t, tmp, hv))
#define REMOVE(k_) do { if (thin) thin_ws_remove(S, &w, Ri, (k_), inW, gc, gs, rowbuf); else ws_remove(S, &w, (k_), inW); } while (0)
    w.row = ibuf; w.side = ibuf + (n + 1);
    int *inW = ibuf + 2 * (n + 1);        /* +1 upper, -1 lower, 0 not in W */
    int *ign = inW + m;                   /* bit0: upper side ignored, bit1: lower */
    int *nadd = ign + m;                  /* times a row left the working set on a zero step */
    int *nWkeep = nadd + m;
    memset(ign, 0, sizeof(int) * m); memset(nadd, 0, sizeof(int) * m);
    int it = 0, status = QP_ITER_LIMIT, pk = 0;
    if (keep) {
        w.nW = *nWkeep; nW0 = 0;
    } else {
        memset(inW, 0, sizeof(int) * m);
        w.nW = 0;
        for (int c = 0; c < n; ++c) zc[c] = z0 ? z0[c] : 0.;
    }

    /* warm start: rebuild the factor of the inherited working set, dropping dependent rows */
    for (int i = 0; i < nW0; ++i) {
        const int r = W0row[i], s = W0side[i];
        if (inW[r]) continue;
        if (APPEND(r, s)) { w.lam[w.nW - 1] = lam0[i] * S->nrm[r]; inW[r] = s; }
    }

    int pending = -1, pside = 0;          /* entering row that is dependent on W (singular case) */
    double plam = 0.;
    int just_added = -1;
    for (pk = 0; pk < S->max_prox; ++pk) {
        /* w = Kx x0 - eps Rinv' zc ;  g = Mh w ; bounds of this proximal sub-problem */
        for (int c = 0; c < n; ++c) {
            double s = dot(S->Kx + (size_t)c * nx, x0, nx), a = 0.;
            for (int rr = 0; rr < n; ++rr) a += S->Rinv[(size_t)rr * n + c] * zc[rr];
            wv[c] = s - S->eps * a;
        }
        for (int r = 0; r < m; ++r) g[r] = dot(S->Mh + (size_t)r * n, wv, n);
        for (int r = 0; r < mc; ++r) { bl[r] = -INFINITY; bu[r] = S->hh[r] - dot(S->Eh + (size_t)r * nx, x0, nx) + g[r]; }
        for (int i = 0; i < nb; ++i) { const int r = mc + i; bl[r] = lb[i] / S->nrm[r] + g[r]; bu[r] = ub[i] / S->nrm[r] + g[r]; }

        status = QP_ITER_LIMIT;
        while (it < S->max_iter) {
            ++it;
            const int k = w.nW;
            if (pending < 0) {
                for (int i = 0; i < k; ++i) dW[i] = -(w.side[i] > 0 ? bu[w.row[i]] : -bl[w.row[i]]);
                if (thin) { thin_rit_mul(Ri, k, dW, res); thin_ri_mul(Ri, k, res, lstar); }
                else {
                    rt_forwardsolve(&w, k, dW, res);      /* u = R^-T (-d) ; lam* = R^-1 u ; v = -Q1 u */
                    r_backsolve(&w, k, res, lstar);
                }
                int kmin = -1; double amin = INFINITY;
                for (int i = 0; i < k; ++i) if (lstar[i] < -S->tol_d) {
                    const double a = w.lam[i] / (w.lam[i] - lstar[i]);
                    if (a < amin) { amin = a; kmin = i; }
                }
                if (kmin >= 0) {           /* partial step towards lam*, drop the blocking row */
                    if (w.row[kmin] == just_added && w.lam[kmin] == 0.)   /* add/drop cycle on noise */
                        ign[just_added] |= (w.side[kmin] > 0 ? 1 : 2);
                    just_added = -1;
                    for (int i = 0; i < k; ++i) w.lam[i] += amin * (lstar[i] - w.lam[i]);
                    if (amin <= 1e-9) ++nadd[w.row[kmin]];
                    REMOVE(kmin);
                    continue;
                }
                for (int i = 0; i < k; ++i) w.lam[i] = lstar[i] > 0. ? lstar[i] : 0.;
                /* primal iterate and most violated row */
                for (int c = 0; c < n; ++c) v[c] = 0.;
                for (int i = 0; i < k; ++i) {
                    const double f = -w.lam[i] * (double)w.side[i];
                    const double *mi = S->Mh + (size_t)w.row[i] * n;
                    for (int c = 0; c < n; ++c) v[c] += f * mi[c];
                }
                /* anti-cycling: a row that has left the working set c times on a zero-length step
                 * must be violated by more than tol_p 10^min(c,2) to re-enter (degenerate add/drop
                 * cycles live at the tolerance level) */
                /* rounding floor: v = -sum lam_i m_i (unit rows) carries ~ eps_mach sum(lam_i) of noise,
                 * i.e. vnoise * vscale_r in the units of row r; capped so that a blown-up iterate
                 * can never be accepted */
                double lsum = 0.; for (int i = 0; i < k; ++i) lsum += w.lam[i];
                const double vnoise = 1e-14 * lsum, vcap = 100. * S->tol_p;
                int jbest = -1, sbest = 0; double vbest = 0.;
                for (int r = 0; r < m; ++r) {
                    const double sv = dot(S->Mh + (size_t)r * n, v, n);
                    if (inW[r]) continue;   /* active rows are tight; bl < bu or bl == bu both fine */
                    double tolr = S->tol_p * tenpow[nadd[r] < 2 ? nadd[r] : 2];
                    const double fl = vnoise * S->vscale[r] < vcap ? vnoise * S->vscale[r] : vcap;
                    if (fl > tolr) tolr = fl;
                    if (!(ign[r] & 1)) {
                        const double vu = (sv - bu[r]) * S->vscale[r];
                        if (vu > tolr && vu > vbest) { vbest = vu; jbest = r; sbest = 1; }
                    }
                    if (!(ign[r] & 2) && bl[r] > -INFINITY) {
                        const double vl = (bl[r] - sv) * S->vscale[r];
                        if (vl > tolr && vl > vbest) { vbest = vl; jbest = r; sbest = -1; }
                    }
                }
                if (jbest < 0) { status = QP_OPTIMAL; break; }
                if (APPEND(jbest, sbest)) { inW[jbest] = sbest; just_added = jbest; }
                else { pending = jbest; pside = sbest; plam = 0.; }
            } else {
                /* singular case: dual ray (p_W, 1) with p_W = -t  (t from the failed append) */
                double pmax = 1.;
                for (int i = 0; i < k; ++i) if (fabs(t[i]) > pmax) pmax = fabs(t[i]);
                int kmin = -1; double amin = INFINITY;
                for (int i = 0; i < k; ++i) if (-t[i] < -S->tol_ray * pmax) {
                    const double a = w.lam[i] / t[i];
                    if (a < amin) { amin = a; kmin = i; }
                }
                if (kmin < 0) {
                    /* Farkas ray: cost = violation of the entering row; rows are infeasible by
                     * cost / sum(p_i * scale_i).  Below tol_p the row is only redundant-and-tight. */
                    double cost = -(pside > 0 ? bu[pending] : -bl[pending]), wsum = 1. / S->vscale[pending];
                    for (int i = 0; i < k; ++i) {
                        const double pi = -t[i] > 0. ? -t[i] : 0.;
                        cost -= pi * (w.side[i] > 0 ? bu[w.row[i]] : -bl[w.row[i]]);
                        wsum += pi / S->vscale[w.row[i]];
                    }
                    if (cost > S->tol_p * wsum) {
                        for (int r = 0; r < m; ++r) y[r] = 0.;
                        for (int i = 0; i < k; ++i) {
                            const double pi = -t[i] > 0. ? -t[i] : 0.;
                            y[w.row[i]] += (double)w.side[i] * pi / S->nrm[w.row[i]];
                        }
                        y[pending] += (double)pside / S->nrm[pending];
                        status = QP_INFEASIBLE;
                        break;
                    }
                    ign[pending] |= (pside > 0 ? 1 : 2);
                    pending = -1;
                    continue;
                }
                for (int i = 0; i < k; ++i) w.lam[i] -= amin * t[i];
                plam += amin;
                if (amin <= 1e-9 * (1. + plam)) ++nadd[w.row[kmin]];
                REMOVE(kmin);
                if (APPEND(pending, pside)) {
                    w.lam[w.nW - 1] = plam; inW[pending] = pside; pending = -1;
                }
            }
        }
        if (status != QP_OPTIMAL) break;
        /* z = Rinv (v - w) ; proximal convergence */
        double dz = 0.;
        for (int r = 0; r < n; ++r) {
            double s = 0.;
            for (int c = 0; c < n; ++c) s += S->Rinv[(size_t)r * n + c] * (v[c] - wv[c]);
            const double d = fabs(s - zc[r]); if (d > dz) dz = d;
            zc[r] = s;
        }
        if (S->eps * dz <= S->prox_tol) { ++pk; break; }
    }

    *iters = it; *prox_iters = pk;
    *farkas = 0.;
    if (status == QP_OPTIMAL) {
        for (int c = 0; c < n; ++c) z[c] = dot(S->Zmap + (size_t)c * n, zc, n);
        for (int c = 0; c < n; ++c) ycoord[c] = zc[c];
        for (int i = 0; i < nb; ++i) if (lb[i] == ub[i]) z[S->bin_idx[i]] = lb[i];
        for (int r = 0; r < m; ++r) y[r] = 0.;
        for (int i = 0; i < w.nW; ++i) y[w.row[i]] += (double)w.side[i] * w.lam[i] / S->nrm[w.row[i]];
    } else if (status == QP_INFEASIBLE) {
        /* cost of the ray in the ORIGINAL bounds:  -(sum rhs_r y_r)  (bounded_qp.py:328-332) */
        double c = 0.;
        for (int r = 0; r < mc; ++r) if (y[r] != 0.) c += y[r] * (S->hh[r] - dot(S->Eh + (size_t)r * nx, x0, nx)) * S->nrm[r];
        for (int i = 0; i < nb; ++i) { const double yy = y[mc + i]; if (yy > 0.) c += yy * ub[i]; else if (yy < 0.) c += yy * lb[i]; }
        *farkas = -c;
    }
    *nWout = w.nW;
    for (int i = 0; i < w.nW; ++i) { Wrow[i] = w.row[i]; Wside[i] = w.side[i]; Wlam[i] = w.lam[i] / S->nrm[w.row[i]]; }
    *nWkeep = w.nW;
    return status;
}

int qp_solve(const qp_shared *S, const double *x0, const double *lb, const double *ub,
             int nW0, const int *W0row, const int *W0side, const double *lam0, const double *z0,
             double *z, double *ycoord, double *y, double *farkas,
             int *nWout, int *Wrow, int *Wside, double *Wlam, int *iters, int *prox_iters)
{
    double *buf = (double *)malloc(sizeof(double) * qp_work_doubles(S->n, S->m));
    int *ibuf = (int *)malloc(sizeof(int) * qp_work_ints(S->n, S->m));
    const int st = qp_solve_impl(S, buf, ibuf, 0, x0, lb, ub, nW0, W0row, W0side, lam0, z0, z, ycoord, y, farkas,
                                 nWout, Wrow, Wside, Wlam, iters, prox_iters);
    free(buf); free(ibuf);
    return st;
}

/* persistent solver state = one CUDA "slot": the factor survives between solves */
typedef struct { double *buf; int *ibuf; int used; } qp_state;

void *qp_state_new(int n, int m) {
    qp_state *st = (qp_state *)malloc(sizeof(qp_state));
    st->buf = (double *)malloc(sizeof(double) * qp_work_doubles(n, m));
    st->ibuf = (int *)malloc(sizeof(int) * qp_work_ints(n, m));
    st->used = 0;
    return st;
}

void qp_state_free(void *p) { qp_state *st = (qp_state *)p; if (st) { free(st->buf); free(st->ibuf); free(st); } }

/* reset != 0: start from the empty working set (first node of an instance) */
int qp_solve_state(const qp_shared *S, void *state, int reset, const double *x0, const double *lb, const double *ub,
                   double *z, double *ycoord, double *y, double *farkas,
                   int *nWout, int *Wrow, int *Wside, double *Wlam, int *iters, int *prox_iters)
{
    qp_state *st = (qp_state *)state;
    const int keep = st->used && !reset;
    st->used = 1;
    return qp_solve_impl(S, st->buf, st->ibuf, keep, x0, lb, ub, 0, NULL, NULL, NULL, NULL, z, ycoord, y, farkas,
                         nWout, Wrow, Wside, Wlam, iters, prox_iters);
}


/* diagnostics of a persistent state: out[0] = max |Mw' - Q1 R| (variant 0) or max |Mw' Ri - Q1| (variant 1),
 * out[1] = max |Q1'Q1 - I|, out[2] = nW, out[3] = max |diag| of Ri resp. 1/min |diag R| */
void qp_state_diag(const qp_shared *S, void *state, double *out)
{
    qp_state *st = (qp_state *)state;
    const int n = S->n, m = S->m;
    double *R = st->buf, *Q = st->buf + (size_t)(n + 1) * (n + 1);
    int *row = st->ibuf, *side = st->ibuf + (n + 1);
    const int k = *(st->ibuf + 2 * (n + 1) + 3 * m);
    double e0 = 0., e1 = 0., dmax = 0.;
    for (int j = 0; j < k; ++j) {
        for (int i = 0; i < n; ++i) {
            double a = 0.;
            if (S->variant == 1) {
                /* (Mw' Ri)[i][j] = sum_{l<=j} s_l Mh[row_l][i] Ri[l][j] */
                for (int l = 0; l <= j; ++l) a += (double)side[l] * S->Mh[(size_t)row[l] * n + i] * R[TRI(j) + l];
                a -= Q[(size_t)j * n + i];
            } else {
                for (int l = 0; l <= j; ++l) a += Q[(size_t)l * n + i] * R[l * (n + 1) + j];
                a -= (double)side[j] * S->Mh[(size_t)row[j] * n + i];
            }
            if (fabs(a) > e0) e0 = fabs(a);
        }
        const double d = S->variant == 1 ? fabs(R[TRI(j) + j]) : 1. / fabs(R[j * (n + 1) + j]);
        if (d > dmax) dmax = d;
        for (int l = 0; l <= j; ++l) {
            double a = dot(Q + (size_t)j * n, Q + (size_t)l * n, n) - (l == j ? 1. : 0.);
            if (fabs(a) > e1) e1 = fabs(a);
        }
    }
    out[0] = e0; out[1] = e1; out[2] = k; out[3] = dmax;
}
