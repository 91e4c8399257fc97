// qp_device.cuh -- CTA-cooperative dual active-set solver for one branch-and-bound node QP (K1).
//
// Replaces the Gurobi call of bounded_qp.py:200-228 (reached from controller.py:229-271) for the
// relaxation of one node.  All nodes of all instances share the least-distance operator
//     min_v 1/2 |v|^2   s.t.   bl_r <= mh_r . v <= bu_r        (unit rows mh_r, r < m)
// and differ only in the bounds (x0 and the bounds on the relaxed binaries), see DESIGN.md.
// One CTA (NT threads) owns one solver state ("slot"):
//     working set W (rows, sides, multipliers lam >= 0 of the sign-normalised rows),
//     Mw' = Q[:, :k] R   (Q n x n orthogonal, column major, global/L2;  R upper triangular, packed by
//     columns, in SHARED memory when it fits),  Ri = R^-1 (packed, global) so that every solve with R
//     is a parallel mat-vec instead of a sequential substitution,  yc = proximal centre.
// The state survives between nodes: any lam >= 0 is dual feasible for every node, so each node is
// hot-started from whatever node the slot solved last.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define WS_NT 256
#define WS_NW (WS_NT / 32)
#define WS_OPTIMAL 2
#define WS_INFEASIBLE 3
#define WS_ITER_LIMIT 9

struct DevProblem {
    int nx, nu, nub, nuc, T, nh, nh1, nq, nqT, nr, n, m, mc, nb;
    const double *A, *B, *F, *G, *h, *F1, *G1, *h1, *Q, *R, *QT, *Mmu, *Mrho;
    const double *Mh, *MhT, *nrm, *vscale, *Eh, *hh, *Rinv, *RinvT, *Kx, *ZmapT;
    const int *bin_idx;
    double eps, tol_p, tol_d, tol_sing, tol_ray, prox_tol;
    int max_iter, max_prox;
    int r_in_smem;           // 1: R lives in shared memory while a CTA works on the slot
    int tri;                 // (n+1)(n+2)/2 packed triangle size
    // record layout
    int n_primal, n_dual, off_lam, off_mu, off_nulb, off_nuub, off_rho, off_sigma;
};

// per-slot persistent state in global memory
struct SlotPtrs {
    double *Q;      // n*n
    double *Ri;     // tri
    double *Rg;     // tri (home of R)
    double *tmp;    // tri (scratch for the Ri down-date)
    double *lam;    // n+1
    double *yc;     // n
    int *row;       // n+1
    int *side;      // n+1
    int *nW;        // 1
};

__host__ __device__ inline size_t slot_doubles(int n) {
    size_t tri = (size_t)(n + 1) * (n + 2) / 2;
    return (size_t)n * n + 3 * tri + (n + 1) + n;
}
__host__ __device__ inline size_t slot_ints(int n) { return 2 * (size_t)(n + 1) + 4; }

__device__ inline SlotPtrs slot_ptrs(double *dbase, int *ibase, int slot, int n) {
    SlotPtrs s;
    const size_t tri = (size_t)(n + 1) * (n + 2) / 2;
    double *d = dbase + (size_t)slot * slot_doubles(n);
    int *i = ibase + (size_t)slot * slot_ints(n);
    s.Q = d; d += (size_t)n * n;
    s.Ri = d; d += tri;
    s.Rg = d; d += tri;
    s.tmp = d; d += tri;
    s.lam = d; d += n + 1;
    s.yc = d;
    s.row = i; i += n + 1;
    s.side = i; i += n + 1;
    s.nW = i;
    return s;
}

// shared-memory working vectors of one CTA
struct Smem {
    double *v, *wv, *c, *hv, *t, *ls, *u, *yc, *lam, *bu, *blb, *gc, *gs, *red;
    double *R;          // packed triangle (smem or global)
    int *row, *side, *ired;
    signed char *inW;
    unsigned char *ign, *nadd;
};

__host__ __device__ inline size_t smem_bytes(int n, int m, int nb, int r_in_smem) {
    size_t d = 9 * (size_t)(n + 1) + m + nb + 2 * (size_t)(n + 1) + 4 * WS_NW + 8;
    if (r_in_smem) d += (size_t)(n + 1) * (n + 2) / 2;
    size_t b = d * 8 + (2 * (size_t)(n + 1) + 2 * WS_NW + 8) * 4 + 3 * (size_t)m + 16;
    return (b + 15) & ~(size_t)15;
}

__device__ inline Smem carve_smem(unsigned char *base, int n, int m, int nb, int r_in_smem, double *Rglobal) {
    Smem s;
    double *d = reinterpret_cast<double *>(base);
    s.v = d; d += n + 1;  s.wv = d; d += n + 1;  s.c = d; d += n + 1;  s.hv = d; d += n + 1;
    s.t = d; d += n + 1;  s.ls = d; d += n + 1;  s.u = d; d += n + 1;  s.yc = d; d += n + 1;
    s.lam = d; d += n + 1;
    s.bu = d; d += m;  s.blb = d; d += nb;
    s.gc = d; d += n + 1;  s.gs = d; d += n + 1;
    s.red = d; d += 4 * WS_NW + 8;
    if (r_in_smem) { s.R = d; d += (size_t)(n + 1) * (n + 2) / 2; } else s.R = Rglobal;
    int *i = reinterpret_cast<int *>(d);
    s.row = i; i += n + 1;  s.side = i; i += n + 1;  s.ired = i; i += 2 * WS_NW + 8;
    signed char *b = reinterpret_cast<signed char *>(i);
    s.inW = b; b += m;
    s.ign = reinterpret_cast<unsigned char *>(b); b += m;
    s.nadd = reinterpret_cast<unsigned char *>(b);
    return s;
}

__device__ __forceinline__ int tri_off(int j) { return j * (j + 1) / 2; }

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// block-wide sum, result to all threads
__device__ inline double block_sum(double x, double *red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    x = warp_sum(x);
    __syncthreads();
    if (lane == 0) red[w] = x;
    __syncthreads();
    double s = 0.;
#pragma unroll
    for (int i = 0; i < WS_NW; ++i) s += red[i];
    return s;
}

// block-wide arg-max of (val, idx): larger val wins, ties -> smaller idx.  idx < 0 = no candidate.
__device__ inline void block_argmax(double &val, int &idx, double *red, int *ired) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, val, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (oi >= 0 && (idx < 0 || ov > val || (ov == val && oi < idx))) { val = ov; idx = oi; }
    }
    __syncthreads();
    if (lane == 0) { red[w] = val; ired[w] = idx; }
    __syncthreads();
    val = red[0]; idx = ired[0];
#pragma unroll
    for (int i = 1; i < WS_NW; ++i) {
        const double ov = red[i]; const int oi = ired[i];
        if (oi >= 0 && (idx < 0 || ov > val || (ov == val && oi < idx))) { val = ov; idx = oi; }
    }
}

// block-wide arg-min (ratio tests): smaller val wins, ties -> smaller idx
__device__ inline void block_argmin(double &val, int &idx, double *red, int *ired) {
    double nv = -val;
    block_argmax(nv, idx, red, ired);
    val = -nv;
}

// ---------------------------------------------------------------------------------------------
// factor updates
// ---------------------------------------------------------------------------------------------

// t = Ri * c[:k]   (thread per row; Ri packed by columns -> coalesced over rows)
__device__ inline void ri_matvec(const double *Ri, int k, const double *c, double *t) {
    for (int i = threadIdx.x; i < k; i += WS_NT) {
        double s = 0.;
        for (int j = i; j < k; ++j) s += Ri[tri_off(j) + i] * c[j];
        t[i] = s;
    }
}

// u = Ri' * d[:k]  (thread per column: u_j = sum_{i<=j} Ri[i][j] d_i)
__device__ inline void rit_matvec(const double *Ri, int k, const double *d, double *u) {
    for (int j = threadIdx.x; j < k; j += WS_NT) {
        const double *col = Ri + tri_off(j);
        double s = 0.;
        for (int i = 0; i <= j; ++i) s += col[i] * d[i];
        u[j] = s;
    }
}

// Try to append the sign-normalised row (r, sgn).  Returns 1 if appended (lam = 0), 0 if the row is
// numerically in the span of the working rows; in that case sm.t = R^-1 c[:k]  (mj = Mw' t).
__device__ inline int qr_append(const DevProblem &P, const SlotPtrs &sp, Smem &sm, int &k, int r, int sgn) {
    const int n = P.n, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const double *mj = P.Mh + (size_t)r * n;
    // c = sgn * Q' mj   (warp per column)
    for (int col = w; col < n; col += WS_NW) {
        const double *q = sp.Q + (size_t)col * n;
        double s = 0.;
        for (int i = lane; i < n; i += 32) s += q[i] * mj[i];
        s = warp_sum(s);
        if (lane == 0) sm.c[col] = (double)sgn * s;
    }
    __syncthreads();
    double part = 0.;
    for (int j = k + threadIdx.x; j < n; j += WS_NT) part += sm.c[j] * sm.c[j];
    const double rho2 = block_sum(part, sm.red);
    if (k >= n || rho2 <= P.tol_sing * P.tol_sing) {
        ri_matvec(sp.Ri, k, sm.c, sm.t);
        __syncthreads();
        return 0;
    }
    const double rho = sqrt(rho2);
    const double ck = sm.c[k];
    const double sg = ck >= 0. ? 1. : -1.;
    const double hk = ck + sg * rho;
    const double hh = rho2 - ck * ck + hk * hk;
    const double beta = 2. / hh;
    __syncthreads();
    for (int j = k + threadIdx.x; j < n; j += WS_NT) sm.hv[j] = (j == k) ? hk : sm.c[j];
    // t = Ri c1 for the new column of Ri (before R / Ri are touched)
    ri_matvec(sp.Ri, k, sm.c, sm.t);
    __syncthreads();
    // Q2 <- Q2 - beta (Q2 hv) hv'   (thread per row, coalesced in the column-major Q)
    for (int i = threadIdx.x; i < n; i += WS_NT) {
        double a = 0.;
        for (int j = k; j < n; ++j) a += sp.Q[(size_t)j * n + i] * sm.hv[j];
        a *= beta;
        for (int j = k; j < n; ++j) sp.Q[(size_t)j * n + i] -= a * sm.hv[j];
    }
    const double rkk = -sg * rho, irkk = 1. / rkk;
    double *Rc = sm.R + tri_off(k), *Ric = sp.Ri + tri_off(k);
    for (int i = threadIdx.x; i < k; i += WS_NT) { Rc[i] = sm.c[i]; Ric[i] = -sm.t[i] * irkk; }
    if (threadIdx.x == 0) {
        Rc[k] = rkk; Ric[k] = irkk;
        sm.row[k] = r; sm.side[k] = sgn; sm.lam[k] = 0.;
    }
    k += 1;
    __syncthreads();
    return 1;
}

// Remove position kp from the working set: delete column kp of R, restore triangularity by Givens
// rotations of rows (i, i+1), i = kp..k-2; the same rotations act on the columns of Q and of R^-1.
__device__ inline void qr_remove(const DevProblem &P, const SlotPtrs &sp, Smem &sm, int &k, int kp) {
    const int n = P.n, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nrot = k - 1 - kp;          // rotations i = kp .. k-2
    double *R = sm.R;
    // ---- 1. Givens chain on the Hessenberg part of R, one warp, columns distributed over lanes.
    // Old column j (> kp) keeps its storage while it is rotated; it becomes new column j-1.
    if (w == 0 && nrot > 0) {
        for (int i = kp; i < k - 1; ++i) {
            const int jown = i + 1;                       // old column that defines rotation i
            double cs = 1., sn = 0.;
            {
                // every lane reads the two defining entries (broadcast read, already up to date
                // because the owner lane finished rotation i-1 on this column in the last step)
                const double a = R[tri_off(jown) + i], b = R[tri_off(jown) + i + 1];
                const double hyp = sqrt(a * a + b * b);
                if (hyp > 0.) { const double ih = 1. / hyp; cs = a * ih; sn = b * ih; }
            }
            if (lane == 0) { sm.gc[i] = cs; sm.gs[i] = sn; }
            for (int j = jown + lane; j < k; j += 32) {
                double *col = R + tri_off(j);
                const double x = col[i], y = col[i + 1];
                col[i] = cs * x + sn * y;
                col[i + 1] = -sn * x + cs * y;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- 2. compact R: new column j-1 <- old column j (rows 0..j-1), ascending chunks (src > dst)
    {
        const int dst0 = tri_off(kp), dst1 = tri_off(k - 1);       // [dst0, dst1) in new indexing
        for (int base = dst0; base < dst1; base += WS_NT) {
            const int a = base + threadIdx.x;
            double val = 0.;
            if (a < dst1) {
                // column jn of address a: largest jn with tri_off(jn) <= a
                int jn = (int)((sqrt(8. * (double)a + 1.) - 1.) * .5);
                while (tri_off(jn + 1) <= a) ++jn;
                while (tri_off(jn) > a) --jn;
                const int i = a - tri_off(jn);
                val = R[tri_off(jn + 1) + i];
            }
            __syncthreads();
            if (a < dst1) R[a] = val;
            __syncthreads();
        }
    }
    // ---- 3. Q columns: q_i <- c q_i + s q_{i+1} ; carry the other combination (thread per row)
    for (int r = threadIdx.x; r < n; r += WS_NT) {
        double carry = sp.Q[(size_t)kp * n + r];
        for (int i = kp; i < k - 1; ++i) {
            const double b = sp.Q[(size_t)(i + 1) * n + r];
            const double cs = sm.gc[i], sn = sm.gs[i];
            sp.Q[(size_t)i * n + r] = cs * carry + sn * b;
            carry = -sn * carry + cs * b;
        }
        sp.Q[(size_t)(k - 1) * n + r] = carry;
    }
    // ---- 4. Ri: delete row kp, rotate columns, drop the last column.  new row rr <- old row
    // (rr < kp ? rr : rr+1).  Written to tmp then copied back (threads own rows, columns interleave).
    for (int rr = threadIdx.x; rr < k - 1; rr += WS_NT) {
        const int ro = rr < kp ? rr : rr + 1;
        // X[rr][j] = old Ri[ro][j] if ro <= j else 0
        double carry = (ro <= kp) ? sp.Ri[tri_off(kp) + ro] : 0.;
        for (int i = kp; i < k - 1; ++i) {
            const double b = (ro <= i + 1) ? sp.Ri[tri_off(i + 1) + ro] : 0.;
            const double cs = sm.gc[i], sn = sm.gs[i];
            if (rr <= i) sp.tmp[tri_off(i) + rr] = cs * carry + sn * b;
            carry = -sn * carry + cs * b;
        }
    }
    __syncthreads();
    {
        const int a0 = tri_off(kp), a1 = tri_off(k - 1);
        for (int a = a0 + threadIdx.x; a < a1; a += WS_NT) sp.Ri[a] = sp.tmp[a];
    }
    // ---- 5. shift the bookkeeping (chunked: k may exceed the block size)
    __syncthreads();
    for (int base = kp + 1; base < k; base += WS_NT) {
        const int tsrc = base + threadIdx.x;
        int rw = 0, sd = 0; double lm = 0.;
        if (tsrc < k) { rw = sm.row[tsrc]; sd = sm.side[tsrc]; lm = sm.lam[tsrc]; }
        __syncthreads();
        if (tsrc < k) { sm.row[tsrc - 1] = rw; sm.side[tsrc - 1] = sd; sm.lam[tsrc - 1] = lm; }
        __syncthreads();
    }
    k -= 1;
    __syncthreads();
}

// remove position kp, then every row whose diagonal of R collapsed (see oracle/qp_core.c ws_remove)
__device__ inline void ws_remove(const DevProblem &P, const SlotPtrs &sp, Smem &sm, int &k, int kp) {
    if (threadIdx.x == 0) sm.inW[sm.row[kp]] = 0;
    __syncthreads();
    qr_remove(P, sp, sm, k, kp);
    for (;;) {
        double val = 0.; int bad = -1;
        for (int i = kp + threadIdx.x; i < k; i += WS_NT)
            if (fabs(sm.R[tri_off(i) + i]) <= P.tol_sing && (bad < 0 || i < bad)) bad = i;
        // smallest index wins: encode as arg-max of -index
        val = bad >= 0 ? -(double)bad : 0.;
        block_argmax(val, bad, sm.red, sm.ired);
        if (bad < 0) break;
        if (threadIdx.x == 0) sm.inW[sm.row[bad]] = 0;
        __syncthreads();
        qr_remove(P, sp, sm, k, bad);
        kp = bad;
    }
}

// sv_r = mh_r . x for all rows, thread per row through the transposed operator (coalesced);
// calls f(r, sv) for every row.
template <class Fn>
__device__ inline void price_rows(const DevProblem &P, const double *x, Fn f) {
    const int n = P.n, m = P.m;
    for (int r0 = threadIdx.x; r0 < m; r0 += 2 * WS_NT) {
        const int r1 = r0 + WS_NT;
        double s0 = 0., s1 = 0.;
        const double *p0 = P.MhT + r0;
        if (r1 < m) {
            for (int c = 0; c < n; ++c) { const double xc = x[c]; s0 += p0[(size_t)c * m] * xc; s1 += p0[(size_t)c * m + WS_NT] * xc; }
            f(r0, s0); f(r1, s1);
        } else {
            for (int c = 0; c < n; ++c) s0 += p0[(size_t)c * m] * x[c];
            f(r0, s0);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the solver.  Inputs: x0 (global), lb/ub (global, nb).  Slot state must be loaded (load_slot).
// Outputs: status; sm.yc = solution in orthonormal coordinates (if optimal); y_out (global, m):
// signed multipliers of the ORIGINAL rows (>0 upper side, <0 lower side; Farkas ray if infeasible).
// ---------------------------------------------------------------------------------------------
__device__ inline void load_slot(const DevProblem &P, const SlotPtrs &sp, Smem &sm, int &k, bool reset) {
    const int n = P.n;
    if (reset) {
        for (size_t a = threadIdx.x; a < (size_t)n * n; a += WS_NT) sp.Q[a] = 0.;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += WS_NT) { sp.Q[(size_t)i * n + i] = 1.; sm.yc[i] = 0.; }
        k = 0;
    } else {
        k = *sp.nW;
        for (int i = threadIdx.x; i < n; i += WS_NT) sm.yc[i] = sp.yc[i];
        for (int i = threadIdx.x; i < k; i += WS_NT) { sm.row[i] = sp.row[i]; sm.side[i] = sp.side[i]; sm.lam[i] = sp.lam[i]; }
        if (P.r_in_smem) for (int a = threadIdx.x; a < tri_off(k); a += WS_NT) sm.R[a] = sp.Rg[a];
    }
    for (int r = threadIdx.x; r < P.m; r += WS_NT) { sm.inW[r] = 0; sm.ign[r] = 0; sm.nadd[r] = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += WS_NT) sm.inW[sm.row[i]] = (signed char)sm.side[i];
    __syncthreads();
}

__device__ inline void store_slot(const DevProblem &P, const SlotPtrs &sp, Smem &sm, int k) {
    const int n = P.n;
    for (int i = threadIdx.x; i < n; i += WS_NT) sp.yc[i] = sm.yc[i];
    for (int i = threadIdx.x; i < k; i += WS_NT) { sp.row[i] = sm.row[i]; sp.side[i] = sm.side[i]; sp.lam[i] = sm.lam[i]; }
    if (P.r_in_smem) for (int a = threadIdx.x; a < tri_off(k); a += WS_NT) sp.Rg[a] = sm.R[a];
    if (threadIdx.x == 0) *sp.nW = k;
    __syncthreads();
}

// per-node reset of the anti-cycling bookkeeping (the working set itself is kept)
__device__ inline void begin_node(const DevProblem &P, Smem &sm) {
    for (int r = threadIdx.x; r < P.m; r += WS_NT) { sm.ign[r] = 0; sm.nadd[r] = 0; }
    __syncthreads();
}

__device__ inline int qp_solve(const DevProblem &P, const SlotPtrs &sp, Smem &sm, int &k,
                               const double *x0, const double *lb, const double *ub,
                               double *y_out, int *iters_out)
{
    const int n = P.n, m = P.m, mc = P.mc, nb = P.nb, nx = P.nx;
    int it = 0, status = WS_ITER_LIMIT;
    int pending = -1, pside = 0, just_added = -1;
    double plam = 0.;

    for (int pk = 0; pk < P.max_prox; ++pk) {
        // wv = Kx x0 - eps Rinv' yc
        for (int c = threadIdx.x; c < n; c += WS_NT) {
            double s = 0.;
            for (int j = 0; j < nx; ++j) s += P.Kx[(size_t)c * nx + j] * x0[j];
            double a = 0.;
            const double *rt = P.RinvT + (size_t)c * n;          // row c of Rinv' = column c of Rinv
            for (int rr = 0; rr < n; ++rr) a += rt[rr] * sm.yc[rr];
            sm.wv[c] = s - P.eps * a;
        }
        __syncthreads();
        // bounds of this proximal sub-problem: g = Mh wv
        price_rows(P, sm.wv, [&](int r, double g) {
            if (r < mc) {
                double e = 0.;
                for (int j = 0; j < nx; ++j) e += P.Eh[(size_t)r * nx + j] * x0[j];
                sm.bu[r] = P.hh[r] - e + g;
            } else {
                const int i = r - mc;
                const double inr = 1. / P.nrm[r];
                sm.bu[r] = ub[i] * inr + g;
                sm.blb[i] = lb[i] * inr + g;
            }
        });
        __syncthreads();

        status = WS_ITER_LIMIT;
        while (it < P.max_iter) {
            ++it;
            if (pending < 0) {
                // u = R^-T (-d_W) ; lam* = R^-1 u ; v = -Q1 u
                for (int i = threadIdx.x; i < k; i += WS_NT) {
                    const int r = sm.row[i];
                    sm.c[i] = -(sm.side[i] > 0 ? sm.bu[r] : -sm.blb[r - mc]);
                }
                __syncthreads();
                rit_matvec(sp.Ri, k, sm.c, sm.u);
                __syncthreads();
                ri_matvec(sp.Ri, k, sm.u, sm.ls);
                __syncthreads();
                double amin = INFINITY; int kmin = -1;
                for (int i = threadIdx.x; i < k; i += WS_NT) if (sm.ls[i] < -P.tol_d) {
                    const double a = sm.lam[i] / (sm.lam[i] - sm.ls[i]);
                    if (a < amin || (a == amin && i < kmin)) { amin = a; kmin = i; }
                }
                block_argmin(amin, kmin, sm.red, sm.ired);
                if (kmin >= 0) {
                    for (int i = threadIdx.x; i < k; i += WS_NT) sm.lam[i] += amin * (sm.ls[i] - sm.lam[i]);
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        const int rr = sm.row[kmin];
                        if (rr == just_added && sm.lam[kmin] == 0.) sm.ign[rr] |= (sm.side[kmin] > 0 ? 1 : 2);
                        if (amin <= 1e-9 && sm.nadd[rr] < 255) ++sm.nadd[rr];
                    }
                    just_added = -1;
                    ws_remove(P, sp, sm, k, kmin);
                    continue;
                }
                for (int i = threadIdx.x; i < k; i += WS_NT) sm.lam[i] = sm.ls[i] > 0. ? sm.ls[i] : 0.;
                // v = -Q1 u   (thread per row)
                for (int i = threadIdx.x; i < n; i += WS_NT) {
                    double s = 0.;
                    for (int j = 0; j < k; ++j) s += sp.Q[(size_t)j * n + i] * sm.u[j];
                    sm.v[i] = -s;
                }
                double lpart = 0.;
                for (int i = threadIdx.x; i < k; i += WS_NT) lpart += sm.ls[i] > 0. ? sm.ls[i] : 0.;
                const double lsum = block_sum(lpart, sm.red);      // (also the barrier after v)
                const double vnoise = 1e-14 * lsum, vcap = 100. * P.tol_p;
                double vbest = 0.; int ibest = -1;                 // ibest = 2 r + (lower side)
                price_rows(P, sm.v, [&](int r, double sv) {
                    if (sm.inW[r]) return;
                    const int na = sm.nadd[r];
                    double tolr = P.tol_p * (na == 0 ? 1. : (na == 1 ? 10. : 100.));
                    const double vs = P.vscale[r];
                    const double fl = vnoise * vs < vcap ? vnoise * vs : vcap;
                    if (fl > tolr) tolr = fl;
                    if (!(sm.ign[r] & 1)) {
                        const double vu = (sv - sm.bu[r]) * vs;
                        if (vu > tolr && (vu > vbest || (vu == vbest && 2 * r < ibest))) { vbest = vu; ibest = 2 * r; }
                    }
                    if (r >= mc && !(sm.ign[r] & 2)) {
                        const double vl = (sm.blb[r - mc] - sv) * vs;
                        if (vl > tolr && (vl > vbest || (vl == vbest && 2 * r + 1 < ibest))) { vbest = vl; ibest = 2 * r + 1; }
                    }
                });
                block_argmax(vbest, ibest, sm.red, sm.ired);
                if (ibest < 0) { status = WS_OPTIMAL; break; }
                const int jb = ibest >> 1, sb = (ibest & 1) ? -1 : 1;
                if (qr_append(P, sp, sm, k, jb, sb)) {
                    if (threadIdx.x == 0) sm.inW[jb] = (signed char)sb;
                    just_added = jb;
                    __syncthreads();
                } else { pending = jb; pside = sb; plam = 0.; }
            } else {
                // dependent entering row: dual ray (p_W, 1), p_W = -t
                double pm = 1.;
                for (int i = threadIdx.x; i < k; i += WS_NT) pm = fmax(pm, fabs(sm.t[i]));
                { int dummy = 0; block_argmax(pm, dummy, sm.red, sm.ired); }
                double amin = INFINITY; int kmin = -1;
                for (int i = threadIdx.x; i < k; i += WS_NT) if (sm.t[i] > P.tol_ray * pm) {
                    const double a = sm.lam[i] / sm.t[i];
                    if (a < amin || (a == amin && i < kmin)) { amin = a; kmin = i; }
                }
                block_argmin(amin, kmin, sm.red, sm.ired);
                if (kmin < 0) {
                    double cpart = 0., wpart = 0.;
                    for (int i = threadIdx.x; i < k; i += WS_NT) {
                        const double pi = sm.t[i] < 0. ? -sm.t[i] : 0.;
                        const int r = sm.row[i];
                        cpart -= pi * (sm.side[i] > 0 ? sm.bu[r] : -sm.blb[r - mc]);
                        wpart += pi / P.vscale[r];
                    }
                    double cost = block_sum(cpart, sm.red);
                    double wsum = block_sum(wpart, sm.red);
                    cost -= (pside > 0 ? sm.bu[pending] : -sm.blb[pending - mc]);
                    wsum += 1. / P.vscale[pending];
                    if (cost > P.tol_p * wsum) {
                        for (int r = threadIdx.x; r < m; r += WS_NT) y_out[r] = 0.;
                        __syncthreads();
                        for (int i = threadIdx.x; i < k; i += WS_NT) {
                            const double pi = sm.t[i] < 0. ? -sm.t[i] : 0.;
                            const int r = sm.row[i];
                            y_out[r] = (double)sm.side[i] * pi / P.nrm[r];
                        }
                        if (threadIdx.x == 0) y_out[pending] = (double)pside / P.nrm[pending];
                        __syncthreads();
                        status = WS_INFEASIBLE;
                        break;
                    }
                    if (threadIdx.x == 0) sm.ign[pending] |= (pside > 0 ? 1 : 2);
                    pending = -1;
                    __syncthreads();
                    continue;
                }
                for (int i = threadIdx.x; i < k; i += WS_NT) sm.lam[i] -= amin * sm.t[i];
                plam += amin;
                __syncthreads();
                if (threadIdx.x == 0 && amin <= 1e-9 * (1. + plam) && sm.nadd[sm.row[kmin]] < 255) ++sm.nadd[sm.row[kmin]];
                ws_remove(P, sp, sm, k, kmin);
                if (qr_append(P, sp, sm, k, pending, pside)) {
                    if (threadIdx.x == 0) { sm.lam[k - 1] = plam; sm.inW[pending] = (signed char)pside; }
                    pending = -1;
                    __syncthreads();
                }
            }
        }
        if (status != WS_OPTIMAL) break;
        // yc <- Rinv (v - wv) ; proximal convergence
        double dz = 0.;
        __syncthreads();
        for (int r = threadIdx.x; r < n; r += WS_NT) sm.c[r] = sm.v[r] - sm.wv[r];
        __syncthreads();
        for (int r = threadIdx.x; r < n; r += WS_NT) {
            double s = 0.;
            const double *ri = P.Rinv + (size_t)r * n;
            for (int c = 0; c < n; ++c) s += ri[c] * sm.c[c];
            dz = fmax(dz, fabs(s - sm.yc[r]));
            sm.hv[r] = s;
        }
        { int dummy = 0; block_argmax(dz, dummy, sm.red, sm.ired); }
        for (int r = threadIdx.x; r < n; r += WS_NT) sm.yc[r] = sm.hv[r];
        __syncthreads();
        if (P.eps * dz <= P.prox_tol) break;
    }
    if (status == WS_OPTIMAL) {
        for (int r = threadIdx.x; r < m; r += WS_NT) y_out[r] = 0.;
        __syncthreads();
        for (int i = threadIdx.x; i < k; i += WS_NT) {
            const int r = sm.row[i];
            y_out[r] = (double)sm.side[i] * sm.lam[i] / P.nrm[r];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *iters_out = it;
    return status;
}
