// lp_batch.cuh -- batched small LPs on the device.
//
// The reference solves two families of tiny LPs on the host through BoundedQP / Gurobi while a controller is built:
//   * _update_mu (controller.py:186-227): h_Tm1.size LPs  min h.mu  s.t.  [F G]' mu = [F_Tm1 G_Tm1]_i',  mu >= 0
//     (130 LPs of 28 variables and 11 equality rows on the T = 20 cart-pole) -- the LPs share matrix and cost and
//     differ in the right-hand side;
//   * mcais (mcais.py:44-184): max c.x s.t. D x <= e with free x, whose dual  min e.y s.t. D'y = c, y >= 0  has the
//     same standard form (few equality rows, many columns).
// One CTA solves one standard-form LP   min c'y  s.t.  E y = r,  y >= 0   (m <= LP_MAX_M rows, n columns) by the
// two-phase tableau simplex in shared memory: threads own tableau columns, the ratio test is a block arg-min, entering
// column by Dantzig's rule with a switch to Bland's rule (no cycling) after 2 (n + m) pivots.  E and c may be shared by
// the whole batch (stride 0).  fp64 throughout.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define LP_NT 128
#define LP_MAX_M 64
#define LP_OPTIMAL 2
#define LP_INFEASIBLE 3
#define LP_UNBOUNDED 5
#define LP_ITER_LIMIT 9

// tableau: (m + 1) rows x (n + m + 1) columns, row major in shared memory: structural | artificial | rhs
__global__ void __launch_bounds__(LP_NT)
lp_std_kernel(int n_lp, int m, int n, const double *__restrict__ E, long long strideE, const double *__restrict__ c, long long stridec,
              const double *__restrict__ r, double tol, int max_iter,
              int *status, double *obj, double *y, double *dual, int *iters)
{
    extern __shared__ __align__(16) double T[];
    __shared__ int basis[LP_MAX_M];
    __shared__ double sgn[LP_MAX_M];
    __shared__ double redv[LP_NT / 32];
    __shared__ int redi[LP_NT / 32];
    const int lp = blockIdx.x;
    if (lp >= n_lp) return;
    const int nc = n + m + 1, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const double *El = E + (size_t)lp * strideE, *cl = c + (size_t)lp * stridec, *rl = r + (size_t)lp * m;
    double *orow = T + (size_t)m * nc;                       // objective row

    // block arg-min over (value, index): smaller value wins, ties -> smaller index; idx < 0 = no candidate
    auto argmin = [&](double v, int i, int &iout) -> double {
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, o); const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (oi >= 0 && (i < 0 || ov < v || (ov == v && oi < i))) { v = ov; i = oi; }
        }
        __syncthreads();
        if (lane == 0) { redv[w] = v; redi[w] = i; }
        __syncthreads();
        v = redv[0]; i = redi[0];
        for (int q = 1; q < LP_NT / 32; ++q) { const double ov = redv[q]; const int oi = redi[q]; if (oi >= 0 && (i < 0 || ov < v || (ov == v && oi < i))) { v = ov; i = oi; } }
        iout = i;
        return v;
    };
    auto pivot = [&](int pr, int pc) {
        // row pr /= T[pr][pc]; every other row (objective row included) -= T[i][pc] * row pr
        const double piv = T[(size_t)pr * nc + pc];
        __syncthreads();
        for (int j = tid; j < nc; j += LP_NT) T[(size_t)pr * nc + j] /= piv;
        __syncthreads();
        for (int i = 0; i <= m; ++i) {
            if (i == pr) continue;
            const double f = T[(size_t)i * nc + pc];
            __syncthreads();
            if (f != 0.) for (int j = tid; j < nc; j += LP_NT) T[(size_t)i * nc + j] -= f * T[(size_t)pr * nc + j];
        }
        __syncthreads();
        if (tid == 0) basis[pr] = pc;
        __syncthreads();
    };

    // ---- tableau with artificial basis, right-hand sides made non-negative
    for (int i = tid; i < m; i += LP_NT) { sgn[i] = rl[i] < 0. ? -1. : 1.; basis[i] = n + i; }
    __syncthreads();
    for (int e = tid; e < m * nc; e += LP_NT) {
        const int i = e / nc, j = e - i * nc;
        double v;
        if (j < n) v = sgn[i] * El[(size_t)i * n + j];
        else if (j < n + m) v = (j - n == i) ? 1. : 0.;
        else v = sgn[i] * rl[i];
        T[e] = v;
    }
    __syncthreads();
    // phase-1 objective: minimise the sum of the artificials -> reduced cost of column j = -sum_i T[i][j] (0 on artificials)
    for (int j = tid; j < nc; j += LP_NT) {
        double s = 0.;
        if (j < n || j == nc - 1) for (int i = 0; i < m; ++i) s -= T[(size_t)i * nc + j];
        orow[j] = s;
    }
    __syncthreads();
    int it = 0, st = LP_ITER_LIMIT;
    for (int phase = 1; phase <= 2; ++phase) {
        const int n_enter = n;                                // artificials never (re-)enter
        st = LP_ITER_LIMIT;
        while (it < max_iter) {
            // entering column
            const bool bland = it > 2 * (n + m);
            double bv = 0.; int bj = -1;
            for (int j = tid; j < n_enter; j += LP_NT) {
                const double rc = orow[j];
                if (rc < -tol) {
                    const double key = bland ? (double)j : rc;
                    if (bj < 0 || key < bv) { bv = key; bj = j; }
                }
            }
            int pc; argmin(bv, bj, pc);
            if (pc < 0) { st = LP_OPTIMAL; break; }
            // ratio test; ties -> smallest BASIS index (Bland): the arg-min runs on the key basis * LP_MAX_M + row
            double rv = 0.; int ri = -1;
            for (int i = tid; i < m; i += LP_NT) {
                const double a = T[(size_t)i * nc + pc];
                if (a > tol) {
                    const double q = T[(size_t)i * nc + nc - 1] / a;
                    const int key = basis[i] * LP_MAX_M + i;
                    if (ri < 0 || q < rv || (q == rv && key < ri)) { rv = q; ri = key; }
                }
            }
            int pk; argmin(rv, ri, pk);
            if (pk < 0) { st = LP_UNBOUNDED; break; }
            const int pr = pk % LP_MAX_M;
            pivot(pr, pc);
            ++it;
        }
        if (st != LP_OPTIMAL) break;
        if (phase == 1) {
            // feasible iff the artificials sum to zero
            double scale = 1.;
            for (int i = 0; i < m; ++i) scale = fmax(scale, fabs(rl[i]));
            if (-orow[nc - 1] > 1e3 * tol * scale) { st = LP_INFEASIBLE; break; }
            // artificials still basic (at level zero): pivot them out on any structural entry, or the row is redundant
            for (int i = 0; i < m; ++i) {
                if (basis[i] < n) continue;
                double bv = 0.; int bj = -1;
                for (int j = tid; j < n; j += LP_NT) { const double a = fabs(T[(size_t)i * nc + j]); if (a > 1e-9 && (bj < 0 || -a < bv)) { bv = -a; bj = j; } }
                int pc; argmin(bv, bj, pc);
                if (pc >= 0) pivot(i, pc);
            }
            // phase-2 objective row: reduced costs c_j - c_B' B^-1 a_j, value -c_B' B^-1 r
            __syncthreads();
            for (int j = tid; j < nc; j += LP_NT) {
                double s = (j < n) ? cl[j] : 0.;
                for (int i = 0; i < m; ++i) { const int b = basis[i]; if (b < n) s -= cl[b] * T[(size_t)i * nc + j]; }
                orow[j] = s;
            }
            __syncthreads();
        }
    }
    // ---- outputs
    if (tid == 0) { status[lp] = st; iters[lp] = it; obj[lp] = st == LP_OPTIMAL ? -orow[nc - 1] : (st == LP_INFEASIBLE ? INFINITY : (st == LP_UNBOUNDED ? -INFINITY : NAN)); }
    for (int j = tid; j < n; j += LP_NT) y[(size_t)lp * n + j] = 0.;
    __syncthreads();
    if (st == LP_OPTIMAL) {
        for (int i = tid; i < m; i += LP_NT) if (basis[i] < n) y[(size_t)lp * n + basis[i]] = T[(size_t)i * nc + nc - 1];
        // multipliers of the equality rows: pi = B^-T c_B; the artificial columns hold B^-1 of the sign-normalised rows
        for (int i = tid; i < m; i += LP_NT) {
            double s = 0.;
            for (int k = 0; k < m; ++k) { const int b = basis[k]; if (b < n) s += cl[b] * T[(size_t)k * nc + n + i]; }
            dual[(size_t)lp * m + i] = s * sgn[i];
        }
    } else {
        for (int i = tid; i < m; i += LP_NT) dual[(size_t)lp * m + i] = 0.;
    }
}
