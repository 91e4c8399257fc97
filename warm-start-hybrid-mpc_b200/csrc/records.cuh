// records.cuh -- per-node outputs in the reference's layout (K1 epilogue).
//
// Replaces PrimalSolution.from_controller / DualSolution.from_controller
// (subproblem_solution.py:68-99, 119-168) and BoundedQP.primal_objective / dual_objective
// (bounded_qp.py:292-332):  primal record  x_0..x_T | u_0..u_{T-1},
// dual record  lam_0..lam_T | mu_0..mu_{T-1} | nu_lb | nu_ub | rho_0..rho_T | sigma_0..sigma_{T-1}.
#pragma once
#include "qp_device.cuh"

// yc: solution in orthonormal coordinates (shared memory), y: signed row multipliers (global, m).
// scratch: >= max(WS_NT, n, 2 (T + 1) nx) doubles of shared memory (Smem::part).  Returns cost (inf if infeasible) and dual objective.
__device__ __forceinline__ void build_records(const DevProblem &P, int status, const double *yc, const double *y,
                                     const double *x0, const double *lb, const double *ub,
                                     double *primal, double *dual, double *cost_out, double *dobj_out,
                                     double *scratch, double *red)
{
    const int n = P.n, nx = P.nx, nu = P.nu, nub = P.nub, nuc = P.nuc, T = P.T, mc = P.mc, nb = P.nb;
    double *X = primal, *U = primal + (size_t)(T + 1) * nx;
    double *lam = dual + P.off_lam, *mu = dual + P.off_mu, *nulb = dual + P.off_nulb,
           *nuub = dual + P.off_nuub, *rho = dual + P.off_rho, *sigma = dual + P.off_sigma;
    const bool opt = status == WS_OPTIMAL;
    double cost = INFINITY;
    if (opt) {
        // z = Zmap yc ; pinned binaries hold exactly (rows lb <= z_i <= ub with lb == ub)
        grouped_matvec(P.ZmapT, n, n, 0, n, yc, scratch, [&](int c, double s) { U[c] = s; });
        for (int i = WS_TID; i < nb; i += WS_NT) if (lb[i] == ub[i]) U[P.bin_idx[i]] = lb[i];
        for (int j = WS_TID; j < nx; j += WS_NT) X[j] = x0[j];
        WS_SYNC();
        if (nx <= 32) {
            // x_{t+1} = A x_t + B u_t: the input terms of all stages in parallel, then ONE warp runs the recursion with
            // the state in registers (lane j holds x_t[j]) -- no barrier per stage
            // (the input terms wait in SHARED memory: read back from the global record inside the recursion, every stage
            // would expose an L2 round trip -- the stores of the loop keep the compiler from hoisting the loads)
            double *bx = scratch + (size_t)(T + 1) * nx;
            for (int e = WS_TID; e < T * nx; e += WS_NT) {
                const int t = e / nx, j = e - t * nx;
                double s = 0.;
                for (int c = 0; c < nu; ++c) s += P.B[j * nu + c] * U[(size_t)t * nu + c];
                bx[e] = s;
            }
            WS_SYNC();
            if (WS_TID < 32) {
                const int j = WS_TID < nx ? WS_TID : 0;
                double xj = X[j];
                for (int t = 0; t < T; ++t) {
                    double s = bx[(size_t)t * nx + j];
                    for (int c = 0; c < nx; ++c) s += P.A[j * nx + c] * __shfl_sync(0xffffffffu, xj, c);
                    if (WS_TID < nx) X[(size_t)(t + 1) * nx + j] = s;
                    xj = s;
                }
            }
            WS_SYNC();
        } else {
            for (int t = 0; t < T; ++t) {
                for (int j = WS_TID; j < nx; j += WS_NT) {
                    double s = 0.;
                    for (int c = 0; c < nx; ++c) s += P.A[j * nx + c] * X[(size_t)t * nx + c];
                    for (int c = 0; c < nu; ++c) s += P.B[j * nu + c] * U[(size_t)t * nu + c];
                    X[(size_t)(t + 1) * nx + j] = s;
                }
                WS_SYNC();
            }
        }
        // rho_t = 2 Q x_t, rho_T = 2 Q_T x_T, sigma_t = 2 R u_t ; cost = 1/4 (|rho|^2 + |sigma|^2)
        double part = 0.;
        for (int e = WS_TID; e < T * P.nq; e += WS_NT) {
            const int t = e / P.nq, i = e % P.nq;
            double s = 0.;
            for (int c = 0; c < nx; ++c) s += P.Q[i * nx + c] * X[(size_t)t * nx + c];
            rho[e] = 2. * s; part += s * s;
        }
        for (int i = WS_TID; i < P.nqT; i += WS_NT) {
            double s = 0.;
            for (int c = 0; c < nx; ++c) s += P.QT[i * nx + c] * X[(size_t)T * nx + c];
            rho[T * P.nq + i] = 2. * s; part += s * s;
        }
        for (int e = WS_TID; e < T * P.nr; e += WS_NT) {
            const int t = e / P.nr, i = e % P.nr;
            double s = 0.;
            for (int c = 0; c < nu; ++c) s += P.R[i * nu + c] * U[(size_t)t * nu + c];
            sigma[e] = 2. * s; part += s * s;
        }
        cost = block_sum(part, red);
    } else {
        for (int e = WS_TID; e < T * P.nq + P.nqT; e += WS_NT) rho[e] = 0.;
        for (int e = WS_TID; e < T * P.nr; e += WS_NT) sigma[e] = 0.;
    }
    // multipliers of the inequality rows
    for (int r = WS_TID; r < mc; r += WS_NT) mu[r] = y[r] > 0. ? y[r] : 0.;
    for (int i = WS_TID; i < nb; i += WS_NT) {
        const double yy = y[mc + i];
        nuub[i] = yy > 0. ? yy : 0.;
        nulb[i] = yy < 0. ? -yy : 0.;
    }
    WS_SYNC();
    // lam_T = -Q_T' rho_T ; lam_t = A' lam_{t+1} - Q' rho_t - F_t' mu_t
    double *g = scratch;                       // (T + 1) nx: lam_T and the g_t in shared memory for the backward recursion
    for (int j = WS_TID; j < nx; j += WS_NT) {
        double s = 0.;
        for (int i = 0; i < P.nqT; ++i) s += P.QT[i * nx + j] * rho[T * P.nq + i];
        lam[(size_t)T * nx + j] = -s;
        g[(size_t)T * nx + j] = -s;
    }
    WS_SYNC();
    if (nx <= 32) {
        // lam_t = A' lam_{t+1} - g_t,  g_t = Q' rho_t + F_t' mu_t: one warp per stage forms g_t (lanes split the rows,
        // one shuffle reduction per state), then ONE warp runs the backward recursion in registers
        const int lane = WS_TID & 31;
        for (int t = WS_TID >> 5; t < T; t += WS_NW) {
            const double *Ft = t < T - 1 ? P.F : P.F1;
            const int k = t < T - 1 ? P.nh : P.nh1;
            const double *mut = mu + (size_t)t * P.nh;
            for (int j = 0; j < nx; ++j) {
                double s = 0.;
                for (int i = lane; i < k; i += 32) s += Ft[i * nx + j] * mut[i];
                for (int i = lane; i < P.nq; i += 32) s += P.Q[i * nx + j] * rho[(size_t)t * P.nq + i];
                s = warp_sum(s);
                if (lane == 0) g[(size_t)t * nx + j] = -s;
            }
        }
        WS_SYNC();
        if (WS_TID < 32) {
            const int j = WS_TID < nx ? WS_TID : 0;
            double lj = g[(size_t)T * nx + j];
            for (int t = T - 1; t >= 0; --t) {
                double s = g[(size_t)t * nx + j];
                for (int c = 0; c < nx; ++c) s += P.A[c * nx + j] * __shfl_sync(0xffffffffu, lj, c);
                if (WS_TID < nx) lam[(size_t)t * nx + j] = s;
                lj = s;
            }
        }
        WS_SYNC();
    } else {
        for (int t = T - 1; t >= 0; --t) {
            const double *Ft = t < T - 1 ? P.F : P.F1;
            const int k = t < T - 1 ? P.nh : P.nh1;
            const double *mut = mu + (size_t)t * P.nh;
            for (int j = WS_TID; j < nx; j += WS_NT) {
                double s = 0.;
                for (int c = 0; c < nx; ++c) s += P.A[c * nx + j] * lam[(size_t)(t + 1) * nx + c];
                for (int i = 0; i < P.nq; ++i) s -= P.Q[i * nx + j] * rho[(size_t)t * P.nq + i];
                for (int i = 0; i < k; ++i) s -= Ft[i * nx + j] * mut[i];
                lam[(size_t)t * nx + j] = s;
            }
            WS_SYNC();
        }
    }
    double dobj = cost;
    if (!opt) {
        // cost of the Farkas proof: -(sum rhs_r y_r)  (bounded_qp.py:328-332)
        double part = 0.;
        for (int j = WS_TID; j < nx; j += WS_NT) part -= lam[j] * x0[j];
        for (int r = WS_TID; r < mc; r += WS_NT) {
            const int t = r / P.nh < T - 1 ? r / P.nh : T - 1;
            const double hr = t < T - 1 ? P.h[r - t * P.nh] : P.h1[r - (T - 1) * P.nh];
            part -= hr * mu[r];
        }
        for (int i = WS_TID; i < nb; i += WS_NT) part += lb[i] * nulb[i] - ub[i] * nuub[i];
        dobj = block_sum(part, red);
    }
    if (WS_TID == 0) { *cost_out = cost; *dobj_out = dobj; }
    WS_SYNC();
    (void)nub; (void)nuc; (void)scratch;
}
