// qp_device.cuh -- CTA-cooperative dual active-set solver for one branch-and-bound node QP (K1).
//
// Replaces the Gurobi call of bounded_qp.py:200-228 (reached from controller.py:229-271) for the
// relaxation of one node.  All nodes of all instances share the least-distance operator
//     min_v 1/2 |v|^2   s.t.   bl_r <= mh_r . v <= bu_r        (unit rows mh_r, r < m)
// and differ only in the bounds (x0 and the bounds on the relaxed binaries), see DESIGN.md.
// One CTA (WS_NT threads, one CTA per SM) owns one solver state ("slot"):
//     working set W (rows, sides, multipliers lam >= 0 of the sign-normalised rows),
//     Mw' = Q[:, :k] R   Q n x n orthogonal, column major, in SHARED memory when it fits (157 KB for the
//                        T=20 cart-pole), else in global memory / L2;
//                        R upper triangular, packed by columns, global memory / L2;
//     Ri = R^-1 (packed) so that every solve with R is a parallel mat-vec instead of a substitution,
//     yc = proximal centre.
// Rows are priced in FACTORED form: mh_r . x = a_r . (Wf x) / nrm_r with Wf = N Rinv (ns x n, shared by
// every CTA, L2 resident) and a_r the sparse stage row [F_t G_t] of the MLD system, instead of streaming
// the dense m x n operator (6x less L2 traffic on the cart-pole).
// The state survives between nodes: any lam >= 0 is dual feasible for every node, so each node is
// hot-started from whatever node the slot solved last.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define WS_NT 512
#define WS_NW (WS_NT / 32)
#define WS_OPTIMAL 2
#define WS_INFEASIBLE 3
#define WS_ITER_LIMIT 9

struct DevProblem {
    int nx, nu, nub, nuc, T, nh, nh1, nq, nqT, nr, n, m, mc, nb, ns;
    const double *A, *B, *F, *G, *h, *F1, *G1, *h1, *Q, *R, *QT, *Mmu, *Mrho;
    const double *Mh, *Wf, *nrm, *inr, *vscale, *Eh, *hh, *Rinv, *RinvT, *Kx, *ZmapT;
    const int *bin_idx;
    double eps, tol_p, tol_d, tol_sing, tol_ray, prox_tol;
    int max_iter, max_prox;
    int q_in_smem;           // 1: Q lives in shared memory while a CTA works on the slot
    int tri;                 // (n+1)(n+2)/2 packed triangle size
    // record layout
    int n_primal, n_dual, off_lam, off_mu, off_nulb, off_nuub, off_rho, off_sigma;
};

// per-slot persistent state in global memory
struct SlotPtrs {
    double *Q;      // n*n   (home of Q; working copy when it does not fit in shared memory)
    double *R;      // tri
    double *Ri;     // tri
    double *tmp;    // tri (scratch: R down-date)
    double *tmp2;   // tri (scratch: Ri down-date)
    double *lam;    // n+1
    double *yc;     // n
    int *row;       // n+1
    int *side;      // n+1
    int *nW;        // 1
};

__host__ __device__ inline size_t slot_doubles(int n) {
    size_t tri = (size_t)(n + 1) * (n + 2) / 2;
    return (size_t)n * n + 4 * tri + (n + 1) + n;
}
__host__ __device__ inline size_t slot_ints(int n) { return 2 * (size_t)(n + 1) + 4; }

__device__ inline SlotPtrs slot_ptrs(double *dbase, int *ibase, int slot, int n) {
    SlotPtrs s;
    const size_t tri = (size_t)(n + 1) * (n + 2) / 2;
    double *d = dbase + (size_t)slot * slot_doubles(n);
    int *i = ibase + (size_t)slot * slot_ints(n);
    s.Q = d; d += (size_t)n * n;
    s.R = d; d += tri;
    s.Ri = d; d += tri;
    s.tmp = d; d += tri;
    s.tmp2 = d; d += tri;
    s.lam = d; d += n + 1;
    s.yc = d;
    s.row = i; i += n + 1;
    s.side = i; i += n + 1;
    s.nW = i;
    return s;
}

// shared-memory working vectors of one CTA
struct Smem {
    double *v, *wv, *c, *hv, *t, *ls, *u, *yc, *lam, *bu, *blb, *gc, *gs, *red, *xi, *mj, *part, *stage;
    double *Q;          // n x n column major (smem or global)
    int *row, *side, *ired;
    signed char *inW;
    unsigned char *ign, *nadd;
};

__host__ __device__ inline size_t smem_doubles(int n, int m, int nb, int ns) {
    const size_t part = (size_t)(n > WS_NT ? n : WS_NT);
    return 12 * (size_t)(n + 1) + m + nb + ns + part + 32 * 33 + 4 * WS_NW + 8;
}
__host__ __device__ inline size_t smem_bytes(int n, int m, int nb, int ns, int q_in_smem) {
    size_t d = smem_doubles(n, m, nb, ns);
    if (q_in_smem) d += (size_t)n * n;
    size_t b = d * 8 + (2 * (size_t)(n + 1) + 2 * WS_NW + 8) * 4 + 3 * (size_t)m + 16;
    return (b + 15) & ~(size_t)15;
}

__device__ inline Smem carve_smem(unsigned char *base, int n, int m, int nb, int ns, int q_in_smem, double *Qglobal) {
    Smem s;
    double *d = reinterpret_cast<double *>(base);
    if (q_in_smem) { s.Q = d; d += (size_t)n * n; } else s.Q = Qglobal;
    s.v = d; d += n + 1;  s.wv = d; d += n + 1;  s.c = d; d += n + 1;  s.hv = d; d += n + 1;
    s.t = d; d += n + 1;  s.ls = d; d += n + 1;  s.u = d; d += n + 1;  s.yc = d; d += n + 1;
    s.lam = d; d += n + 1;  s.mj = d; d += n + 1;
    s.gc = d; d += n + 1;  s.gs = d; d += n + 1;
    s.bu = d; d += m;  s.blb = d; d += nb;  s.xi = d; d += ns;
    s.part = d; d += (n > WS_NT ? n : WS_NT);
    s.stage = d; d += 32 * 33;
    s.red = d; d += 4 * WS_NW + 8;
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
// grouped mat-vec:  out(i, sum_{j0 <= j < j1} M[j * ld + i] * x[j])  for i < rows.
// Adjacent threads take adjacent i (coalesced / conflict free); when rows <= WS_NT the threads form
// G = WS_NT / rows groups that split the j range and combine through `part` (>= max(WS_NT, rows) doubles).
// Ends with a barrier-protected combine; `out` is called once per row.  Contains __syncthreads.
// ---------------------------------------------------------------------------------------------
template <class Out>
__device__ inline void grouped_matvec(const double *__restrict__ M, int ld, int rows, int j0, int j1,
                                      const double *x, double *part, Out out) {
    if (rows <= WS_NT) {
        const int G = WS_NT / rows;
        const int g = threadIdx.x / rows, i = threadIdx.x - g * rows;
        if (g < G) {
            double s0 = 0., s1 = 0.;
            int j = j0 + g;
            for (; j + G < j1; j += 2 * G) {
                s0 += M[(size_t)j * ld + i] * x[j];
                s1 += M[(size_t)(j + G) * ld + i] * x[j + G];
            }
            if (j < j1) s0 += M[(size_t)j * ld + i] * x[j];
            part[g * rows + i] = s0 + s1;
        }
        __syncthreads();
        if (threadIdx.x < rows) {
            double s = part[threadIdx.x];
            for (int q = 1; q < G; ++q) s += part[q * rows + threadIdx.x];
            out(threadIdx.x, s);
        }
        __syncthreads();
    } else {
        for (int i = threadIdx.x; i < rows; i += WS_NT) {
            double s0 = 0., s1 = 0.;
            int j = j0;
            for (; j + 1 < j1; j += 2) { s0 += M[(size_t)j * ld + i] * x[j]; s1 += M[(size_t)(j + 1) * ld + i] * x[j + 1]; }
            if (j < j1) s0 += M[(size_t)j * ld + i] * x[j];
            out(i, s0 + s1);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// factor updates
// ---------------------------------------------------------------------------------------------

// t = Ri * c[:k]   (t_i = sum_{j >= i} Ri[tri_off(j) + i] c_j ; adjacent threads -> adjacent i, coalesced)
__device__ inline void ri_matvec(const double *__restrict__ Ri, int k, const double *c, double *t, double *part) {
    if (k <= 0) { __syncthreads(); return; }
    if (k <= WS_NT) {
        const int G = WS_NT / k;
        const int g = threadIdx.x / k, i = threadIdx.x - g * k;
        if (g < G) {
            double s = 0.;
            for (int j = i + g; j < k; j += G) s += Ri[tri_off(j) + i] * c[j];
            part[g * k + i] = s;
        }
        __syncthreads();
        if (threadIdx.x < k) {
            double s = part[threadIdx.x];
            for (int q = 1; q < G; ++q) s += part[q * k + threadIdx.x];
            t[threadIdx.x] = s;
        }
    } else {
        for (int i = threadIdx.x; i < k; i += WS_NT) {
            double s = 0.;
            for (int j = i; j < k; ++j) s += Ri[tri_off(j) + i] * c[j];
            t[i] = s;
        }
    }
    __syncthreads();
}

// u = Ri' * d[:k]  (u_j = sum_{i <= j} Ri[tri_off(j) + i] d_i ; warp per column, lanes over rows)
__device__ inline void rit_matvec(const double *__restrict__ Ri, int k, const double *d, double *u) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int j = w; j < k; j += WS_NW) {
        const double *col = Ri + tri_off(j);
        double s = 0.;
        for (int i = lane; i <= j; i += 32) s += col[i] * d[i];
        s = warp_sum(s);
        if (lane == 0) u[j] = s;
    }
    __syncthreads();
}

// Try to append the sign-normalised row (r, sgn).  Returns 1 if appended (lam = 0), 0 if the row is
// numerically in the span of the working rows; in that case sm.t = R^-1 c[:k]  (mj = Mw' t).
__device__ inline int qr_append(const DevProblem &P, const SlotPtrs &sp, Smem &sm, int &k, int r, int sgn) {
    const int n = P.n, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const double *Q = sm.Q;
    for (int i = threadIdx.x; i < n; i += WS_NT) sm.mj[i] = P.Mh[(size_t)r * n + i];
    __syncthreads();
    // c = sgn * Q' mj   (warp per column, two columns in flight)
    for (int col = w; col < n; col += 2 * WS_NW) {
        const int col2 = col + WS_NW;
        const double *q0 = Q + (size_t)col * n, *q1 = Q + (size_t)(col2 < n ? col2 : col) * n;
        double s0 = 0., s1 = 0.;
        for (int i = lane; i < n; i += 32) { const double x = sm.mj[i]; s0 += q0[i] * x; s1 += q1[i] * x; }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        if (lane == 0) { sm.c[col] = (double)sgn * s0; if (col2 < n) sm.c[col2] = (double)sgn * s1; }
    }
    __syncthreads();
    double part = 0.;
    for (int j = k + threadIdx.x; j < n; j += WS_NT) part += sm.c[j] * sm.c[j];
    const double rho2 = block_sum(part, sm.red);
    // t = Ri c1: the dual ray of a dependent row / the new column of Ri (before R, Ri are touched)
    ri_matvec(sp.Ri, k, sm.c, sm.t, sm.part);
    if (k >= n || rho2 <= P.tol_sing * P.tol_sing) return 0;
    const double rho = sqrt(rho2);
    const double ck = sm.c[k];
    const double sg = ck >= 0. ? 1. : -1.;
    const double hk = ck + sg * rho;
    const double hh = rho2 - ck * ck + hk * hk;
    const double beta = 2. / hh;
    __syncthreads();
    for (int j = k + threadIdx.x; j < n; j += WS_NT) sm.hv[j] = (j == k) ? hk : sm.c[j];
    __syncthreads();
    // Q2 <- Q2 - beta (Q2 hv) hv'   (adjacent threads -> adjacent rows of the column-major Q)
    {
        double *Qw = sm.Q;
        if (n <= WS_NT) {
            const int G = WS_NT / n;
            const int g = threadIdx.x / n, i = threadIdx.x - g * n;
            if (g < G) {
                double a = 0.;
                for (int j = k + g; j < n; j += G) a += Qw[(size_t)j * n + i] * sm.hv[j];
                sm.part[g * n + i] = a;
            }
            __syncthreads();
            if (g < G) {
                double a = sm.part[i];
                for (int q = 1; q < G; ++q) a += sm.part[q * n + i];
                a *= beta;
                for (int j = k + g; j < n; j += G) Qw[(size_t)j * n + i] -= a * sm.hv[j];
            }
        } else {
            for (int i = threadIdx.x; i < n; i += WS_NT) {
                double a = 0.;
                for (int j = k; j < n; ++j) a += Qw[(size_t)j * n + i] * sm.hv[j];
                a *= beta;
                for (int j = k; j < n; ++j) Qw[(size_t)j * n + i] -= a * sm.hv[j];
            }
        }
    }
    const double rkk = -sg * rho, irkk = 1. / rkk;
    double *Rc = sp.R + tri_off(k), *Ric = sp.Ri + tri_off(k);
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
    const double *R = sp.R;
    double *Rn = sp.tmp;                  // new R (columns kp .. k-2), built out of place
    // ---- 1. warp 0: Givens chain.  Old column j (> kp) becomes new column j-1.  A lane owns one
    // column of a block of 32 and carries its running entry in a register: entries of R are read once
    // (no read-after-write through L2).  The other warps copy the untouched rows < kp meanwhile.
    if (w == 0) {
        for (int j0 = kp + 1; j0 < k; j0 += 32) {
            const int j = j0 + lane;
            const bool has = j < k;
            const double *col = R + tri_off(has ? j : kp);
            double *out = Rn + tri_off(has ? j - 1 : kp);
            double carry = has ? col[kp] : 0.;
            // rotations defined by earlier blocks: i = kp .. j0-2
#pragma unroll 4
            for (int i = kp; i < j0 - 1; ++i) {
                const double b = has ? col[i + 1] : 0.;
                const double cs = sm.gc[i], sn = sm.gs[i];
                if (has) out[i] = cs * carry + sn * b;
                carry = -sn * carry + cs * b;
            }
            // stage rows j0 .. j of the own column (the triangular part of the block)
            for (int q = 0; q < 32; ++q) sm.stage[q * 33 + lane] = (has && j0 + q <= j) ? col[j0 + q] : 0.;
            __syncwarp();
            const int iend = (j0 + 32 < k ? j0 + 32 : k) - 1;          // rotations i = j0-1 .. iend-1
            for (int i = j0 - 1; i < iend; ++i) {
                const int q = i + 1 - j0;                               // defining column = lane q, its row i+1
                const double b = sm.stage[q * 33 + lane];
                const double a_d = __shfl_sync(0xffffffffu, carry, q), b_d = __shfl_sync(0xffffffffu, b, q);
                double cs = 1., sn = 0.;
                const double hyp = sqrt(a_d * a_d + b_d * b_d);
                if (hyp > 0.) { const double ih = 1. / hyp; cs = a_d * ih; sn = b_d * ih; }
                if (lane == 0) { sm.gc[i] = cs; sm.gs[i] = sn; }
                if (has && j >= i + 1) {
                    out[i] = cs * carry + sn * b;
                    carry = -sn * carry + cs * b;
                }
            }
            __syncwarp();
        }
    } else {
        // rows < kp of the shifted columns are unchanged
        for (int j = kp + 1 + (w - 1); j < k; j += WS_NW - 1) {
            const double *col = R + tri_off(j);
            double *out = Rn + tri_off(j - 1);
            for (int i = lane; i < kp; i += 32) out[i] = col[i];
        }
    }
    __syncthreads();
    // ---- 2. copy the new columns back
    {
        const int a0 = tri_off(kp), a1 = tri_off(k - 1);
        for (int a = a0 + threadIdx.x; a < a1; a += WS_NT) sp.R[a] = Rn[a];
    }
    // ---- 3. Q columns: q_i <- c q_i + s q_{i+1} ; carry the other combination (thread per row)
    {
        double *Qw = sm.Q;
        for (int r = threadIdx.x; r < n; r += WS_NT) {
            double carry = Qw[(size_t)kp * n + r];
#pragma unroll 4
            for (int i = kp; i < k - 1; ++i) {
                const double b = Qw[(size_t)(i + 1) * n + r];
                const double cs = sm.gc[i], sn = sm.gs[i];
                Qw[(size_t)i * n + r] = cs * carry + sn * b;
                carry = -sn * carry + cs * b;
            }
            Qw[(size_t)(k - 1) * n + r] = carry;
        }
    }
    // ---- 4. Ri: delete row kp, rotate columns, drop the last column.  new row rr <- old row
    // (rr < kp ? rr : rr+1).  Written to tmp2 then copied back (threads own rows, columns interleave).
    // Threads WS_NT/2.. take the rows from the top so that both halves of the CTA have work.
    for (int rr = threadIdx.x; rr < k - 1; rr += WS_NT) {
        const int ro = rr < kp ? rr : rr + 1;
        // X[rr][j] = old Ri[ro][j] if ro <= j else 0
        double carry = (ro <= kp) ? sp.Ri[tri_off(kp) + ro] : 0.;
#pragma unroll 4
        for (int i = kp; i < k - 1; ++i) {
            const double b = (ro <= i + 1) ? sp.Ri[tri_off(i + 1) + ro] : 0.;
            const double cs = sm.gc[i], sn = sm.gs[i];
            if (rr <= i) sp.tmp2[tri_off(i) + rr] = cs * carry + sn * b;
            carry = -sn * carry + cs * b;
        }
    }
    __syncthreads();
    {
        const int a0 = tri_off(kp), a1 = tri_off(k - 1);
        for (int a = a0 + threadIdx.x; a < a1; a += WS_NT) sp.Ri[a] = sp.tmp2[a];
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
            if (fabs(sp.R[tri_off(i) + i]) <= P.tol_sing && (bad < 0 || i < bad)) bad = i;
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

// sv_r = mh_r . x for all rows in factored form; calls f(r, sv) once per row.
//   xi = Wf x  (ns = n + T nx entries: inputs zeta_t, then states xi_1 .. xi_T of the homogeneous dynamics)
//   row (t, i):  sv = (F_t[i] . xi_t + G_t[i] . zeta_t) / nrm_r   (xi_0 = 0: the x0 part is in the bounds)
//   binary (t, i): sv = zeta_t[nuc + i] / nrm_r
template <class Fn>
__device__ inline void price_rows(const DevProblem &P, Smem &sm, const double *x, Fn f) {
    const int n = P.n, m = P.m, mc = P.mc, nx = P.nx, nu = P.nu, ns = P.ns;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // xi = Wf x : warp per row of Wf (row major, coalesced over lanes), two rows in flight
    for (int r = w; r < ns; r += 2 * WS_NW) {
        const int r2 = r + WS_NW;
        const double *w0 = P.Wf + (size_t)r * n, *w1 = P.Wf + (size_t)(r2 < ns ? r2 : r) * n;
        double s0 = 0., s1 = 0.;
        for (int c = lane; c < n; c += 32) { const double xc = x[c]; s0 += w0[c] * xc; s1 += w1[c] * xc; }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        if (lane == 0) { sm.xi[r] = s0; if (r2 < ns) sm.xi[r2] = s1; }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < m; r += WS_NT) {
        double s;
        if (r < mc) {
            int t = r / P.nh; if (t > P.T - 1) t = P.T - 1;
            const int i = r - t * P.nh;
            const double *Fr = (t < P.T - 1 ? P.F : P.F1) + (size_t)i * nx;
            const double *Gr = (t < P.T - 1 ? P.G : P.G1) + (size_t)i * nu;
            const double *zt = sm.xi + (size_t)t * nu;
            s = 0.;
            for (int c = 0; c < nu; ++c) s += Gr[c] * zt[c];
            if (t > 0) {
                const double *xt = sm.xi + n + (size_t)(t - 1) * nx;
                for (int c = 0; c < nx; ++c) s += Fr[c] * xt[c];
            }
        } else {
            s = sm.xi[P.bin_idx[r - mc]];
        }
        f(r, s * P.inr[r]);
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
        for (size_t a = threadIdx.x; a < (size_t)n * n; a += WS_NT) sm.Q[a] = 0.;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += WS_NT) { sm.Q[(size_t)i * n + i] = 1.; sm.yc[i] = 0.; }
        k = 0;
    } else {
        k = *sp.nW;
        for (int i = threadIdx.x; i < n; i += WS_NT) sm.yc[i] = sp.yc[i];
        for (int i = threadIdx.x; i < k; i += WS_NT) { sm.row[i] = sp.row[i]; sm.side[i] = sp.side[i]; sm.lam[i] = sp.lam[i]; }
        if (P.q_in_smem) for (size_t a = threadIdx.x; a < (size_t)n * n; a += WS_NT) sm.Q[a] = sp.Q[a];
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
    if (P.q_in_smem) for (size_t a = threadIdx.x; a < (size_t)n * n; a += WS_NT) sp.Q[a] = sm.Q[a];
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
    const int n = P.n, m = P.m, mc = P.mc, nx = P.nx;
    int it = 0, status = WS_ITER_LIMIT;
    int pending = -1, pside = 0, just_added = -1;
    double plam = 0.;

    for (int pk = 0; pk < P.max_prox; ++pk) {
        // wv = Kx x0 - eps Rinv' yc     ((Rinv' yc)_c = sum_r Rinv[r][c] yc[r], coalesced over c)
        grouped_matvec(P.Rinv, n, n, 0, n, sm.yc, sm.part, [&](int c, double a) {
            double s = 0.;
            for (int j = 0; j < nx; ++j) s += P.Kx[(size_t)c * nx + j] * x0[j];
            sm.wv[c] = s - P.eps * a;
        });
        // bounds of this proximal sub-problem: g = Mh wv
        price_rows(P, sm, sm.wv, [&](int r, double g) {
            if (r < mc) {
                double e = 0.;
                for (int j = 0; j < nx; ++j) e += P.Eh[(size_t)r * nx + j] * x0[j];
                sm.bu[r] = P.hh[r] - e + g;
            } else {
                const int i = r - mc;
                const double inr = P.inr[r];
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
                ri_matvec(sp.Ri, k, sm.u, sm.ls, sm.part);
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
                // v = -Q1 u
                grouped_matvec(sm.Q, n, n, 0, k, sm.u, sm.part, [&](int i, double s) { sm.v[i] = -s; });
                double lpart = 0.;
                for (int i = threadIdx.x; i < k; i += WS_NT) lpart += sm.ls[i] > 0. ? sm.ls[i] : 0.;
                const double lsum = block_sum(lpart, sm.red);
                const double vnoise = 1e-14 * lsum, vcap = 100. * P.tol_p;
                double vbest = 0.; int ibest = -1;                 // ibest = 2 r + (lower side)
                price_rows(P, sm, sm.v, [&](int r, double sv) {
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
                            y_out[r] = (double)sm.side[i] * pi * P.inr[r];
                        }
                        if (threadIdx.x == 0) y_out[pending] = (double)pside * P.inr[pending];
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
        __syncthreads();
        for (int r = threadIdx.x; r < n; r += WS_NT) sm.c[r] = sm.v[r] - sm.wv[r];
        __syncthreads();
        double dz = 0.;
        grouped_matvec(P.RinvT, n, n, 0, n, sm.c, sm.part, [&](int r, double s) {
            dz = fmax(dz, fabs(s - sm.yc[r]));
            sm.hv[r] = s;
        });
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
            y_out[r] = (double)sm.side[i] * sm.lam[i] * P.inr[r];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *iters_out = it;
    return status;
}
