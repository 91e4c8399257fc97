// bnb.cuh -- device-side branch and bound (K3) and warm-start construction (K2 + K4).
//
// K3 replaces branch_and_bound.py:408-499 driven by the closures of controller.py:365-380 and the
// child-bound rule of controller.py:395-429.  K2/K4 replace controller.py:431-564, 615-721.
// One CTA owns one MPC instance at a time; instances are handed out by an atomic work counter, so
// thousands of independent closed-loop instances keep every SM busy with no host round trip per node.
#pragma once
#include "qp_device.cuh"
#include "records.cuh"

struct TreeView {
    int cap_nodes, cap_recs, words;
    int *n_nodes, *n_recs, *depth, *alive, *rec;
    unsigned int *bits;
    double *lb, *rec_dobj, *rec_dual;
};

#define BNB_OK 0
#define BNB_INFEASIBLE 1
#define BNB_CAPACITY 2
#define BNB_QP_LIMIT 3

// per-slot scratch of the B&B kernel (doubles): lb | ub | primal record | cost | dobj
__host__ __device__ inline size_t bnb_scratch_doubles(int nb, int n_primal) { return 2 * (size_t)nb + n_primal + 4; }

__global__ void init_root_kernel(int n_inst, TreeView tr)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_inst) return;
    const size_t o = (size_t)k * tr.cap_nodes;
    tr.n_nodes[k] = 1; tr.n_recs[k] = 0;
    tr.depth[o] = 0; tr.alive[o] = 1; tr.rec[o] = -1; tr.lb[o] = -INFINITY;
    for (int w = 0; w < tr.words; ++w) tr.bits[o * tr.words + w] = 0u;
}

// ---------------------------------------------------------------------------------------------
// K3
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WS_NT, 1)
bnb_kernel(DevProblem P, double *slot_d, int *slot_i, double *ybuf, double *scratch, int *work_counter,
           int n_inst, const double *__restrict__ x0, const int *__restrict__ active, TreeView tr,
           double tol, int max_solves,
           double *inc_cost, int *inc_node, double *inc_primal, int *n_solves, int *status_out, int *trace,
           unsigned long long *totals)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_inst;
    const int slot = blockIdx.x;
    const int nb = P.nb;
    SlotPtrs sp = slot_ptrs(slot_d, slot_i, slot, P.n, P.ld);
    const Ctx cx = make_ctx(P, smem_raw, sp);
    init_shared_tables(P, cx);
    double *y = ybuf + (size_t)slot * P.m;
    double *sc = scratch + (size_t)slot * bnb_scratch_doubles(nb, P.n_primal);
    double *lbv = sc, *ubv = sc + nb, *prim = sc + 2 * nb, *cost_s = prim + P.n_primal, *dobj_s = cost_s + 1;
    int *iters_s = slot_i + (size_t)slot * slot_ints(P.n) + 2 * (P.n + 1) + 1;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_inst = atomicAdd(work_counter, 1);
        __syncthreads();
        const int inst = s_inst;
        if (inst >= n_inst) break;
        if (active && !active[inst]) {
            if (threadIdx.x == 0) { inc_cost[inst] = INFINITY; inc_node[inst] = -1; n_solves[inst] = 0; status_out[inst] = BNB_INFEASIBLE; }
            continue;
        }

        const size_t no = (size_t)inst * tr.cap_nodes;
        int *depth = tr.depth + no, *alive = tr.alive + no, *rec = tr.rec + no;
        unsigned int *bits = tr.bits + no * tr.words;
        double *lb = tr.lb + no;
        double *rdobj = tr.rec_dobj + (size_t)inst * tr.cap_recs;
        double *rdual = tr.rec_dual + (size_t)inst * tr.cap_recs * P.n_dual;
        const double *xi = x0 + (size_t)inst * P.nx;
        int *tr_i = trace ? trace + (size_t)inst * 2 * max_solves : nullptr;

        int nn = tr.n_nodes[inst], nr = tr.n_recs[inst];
        double ub = INFINITY;
        int inc = -1, solves = 0, st = -1, k = 0;
        long long iters = 0;
        bool first = true;

        while (st < 0) {
            // ---- select: candidates = alive leaves with lb < ub - tol ; best_first = first minimum
            const double cutoff = ub - tol;
            double best = INFINITY; int bi = -1;
            for (int j = threadIdx.x; j < nn; j += WS_NT)
                if (alive[j]) {
                    const double l = lb[j];
                    if (l < cutoff && (bi < 0 || l < best)) { best = l; bi = j; }
                }
            block_argmin(best, bi, SMV(red), SMI(ired));
            if (bi < 0) { st = inc >= 0 ? BNB_OK : BNB_INFEASIBLE; break; }
            if (solves >= max_solves || nn + 2 > tr.cap_nodes || nr + 1 > tr.cap_recs) { st = BNB_CAPACITY; break; }
            // ---- bounds of the node (controller.py:273-298)
            const int d = depth[bi];
            const unsigned int *bw = bits + (size_t)bi * tr.words;
            for (int j = threadIdx.x; j < nb; j += WS_NT) {
                const double v = (double)((bw[j >> 5] >> (j & 31)) & 1u);
                lbv[j] = j < d ? v : 0.;
                ubv[j] = j < d ? v : 1.;
            }
            __syncthreads();
            // ---- solve (K1), hot-started from the node solved before it
            if (first) { load_slot(P, cx, sp, k, true); first = false; }
            const int qs = qp_solve(P, cx, k, xi, lbv, ubv, y, iters_s);
            if (qs == WS_ITER_LIMIT) { st = BNB_QP_LIMIT; break; }
            double *dual = rdual + (size_t)nr * P.n_dual;
            build_records(P, qs, SMV(yc), y, xi, lbv, ubv, prim, dual, cost_s, dobj_s, SMV(part), SMV(red));
            const double cost = *cost_s;
            if (threadIdx.x == 0) {
                lb[bi] = cost; rec[bi] = nr; rdobj[nr] = *dobj_s;
                if (tr_i) { tr_i[2 * solves] = bi; tr_i[2 * solves + 1] = *iters_s; }
            }
            const int myrec = nr;
            iters += *iters_s;
            ++nr; ++solves;
            // ---- prune / incumbent / branch (branch_and_bound.py:476-489)
            if (cost >= cutoff) {
                // pruned: the node stays a leaf with its new bound
            } else if (d == nb) {
                inc = bi; ub = cost;
                double *ip = inc_primal + (size_t)inst * P.n_primal;
                for (int j = threadIdx.x; j < P.n_primal; j += WS_NT) ip[j] = prim[j];
            } else {
                // children [value 0, value 1] of binary d = (t, i); bound += multiplier of the bound that moves
                const double l0 = cost + dual[P.off_nuub + d], l1 = cost + dual[P.off_nulb + d];
                unsigned int *c0 = bits + (size_t)nn * tr.words, *c1 = c0 + tr.words;
                for (int w = threadIdx.x; w < tr.words; w += WS_NT) {
                    const unsigned int b = bw[w];
                    c0[w] = b & ~((w == (d >> 5)) ? (1u << (d & 31)) : 0u);
                    c1[w] = b | ((w == (d >> 5)) ? (1u << (d & 31)) : 0u);
                }
                if (threadIdx.x == 0) {
                    alive[bi] = 0;
                    depth[nn] = d + 1; alive[nn] = 1; rec[nn] = myrec; lb[nn] = l0;
                    depth[nn + 1] = d + 1; alive[nn + 1] = 1; rec[nn + 1] = myrec; lb[nn + 1] = l1;
                }
                nn += 2;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            tr.n_nodes[inst] = nn; tr.n_recs[inst] = nr;
            inc_cost[inst] = ub; inc_node[inst] = inc; n_solves[inst] = solves; status_out[inst] = st;
            if (totals) { atomicAdd(totals, (unsigned long long)solves); atomicAdd(totals + 1, (unsigned long long)iters); }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2 + K4: retain / shift identifiers / shift duals / re-evaluate the bounds / plant update
// ---------------------------------------------------------------------------------------------
#define SH_NT 256
#define SH_NW (SH_NT / 32)

__global__ void __launch_bounds__(SH_NT)
shift_tree_kernel(DevProblem P, int n_inst, const double *__restrict__ x0, const double *__restrict__ e0,
                  TreeView ot, const double *__restrict__ inc_cost, const double *__restrict__ inc_primal,
                  int *active, TreeView nt, double *x_next, double *u0_out)
{
    extern __shared__ __align__(16) double shm[];
    __shared__ int s_wsum[SH_NW];
    __shared__ int s_base;
    const int nx = P.nx, nu = P.nu, nub = P.nub, nuc = P.nuc, T = P.T, nh = P.nh, nh1 = P.nh1;
    const int nq = P.nq, nqT = P.nqT, nr_ = P.nr;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // shared: x0 | u0 | e0 | Qx0 | Ru0 | res_mu | per-warp scratch (nh1 + nqT each)
    double *xs = shm, *us = xs + nx, *es = us + nu, *Qx = es + nx, *Ru = Qx + nq, *rmu = Ru + nr_;
    double *wscr = rmu + nh + (size_t)w * (nh1 + nqT + nh + nq);

    for (int inst = blockIdx.x; inst < n_inst; inst += gridDim.x) {
        __syncthreads();
        const double *ip = inc_primal + (size_t)inst * P.n_primal;
        const bool on = (!active || active[inst]) && inc_cost[inst] < INFINITY;
        if (!on) {
            if (threadIdx.x == 0) {
                if (active) active[inst] = 0;
                nt.n_nodes[inst] = 0; nt.n_recs[inst] = 0;
            }
            for (int j = threadIdx.x; j < nx; j += SH_NT) if (x_next) x_next[(size_t)inst * nx + j] = x0[(size_t)inst * nx + j];
            for (int j = threadIdx.x; j < nu; j += SH_NT) if (u0_out) u0_out[(size_t)inst * nu + j] = nan("");
            continue;
        }
        for (int j = threadIdx.x; j < nx; j += SH_NT) { xs[j] = x0[(size_t)inst * nx + j]; es[j] = e0 ? e0[(size_t)inst * nx + j] : 0.; }
        for (int j = threadIdx.x; j < nu; j += SH_NT) us[j] = ip[(size_t)(T + 1) * nx + j];
        if (threadIdx.x == 0) s_base = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < nq; i += SH_NT) { double s = 0.; for (int c = 0; c < nx; ++c) s += P.Q[i * nx + c] * xs[c]; Qx[i] = s; }
        for (int i = threadIdx.x; i < nr_; i += SH_NT) { double s = 0.; for (int c = 0; c < nu; ++c) s += P.R[i * nu + c] * us[c]; Ru[i] = s; }
        for (int i = threadIdx.x; i < nh; i += SH_NT) {
            double s = -P.h[i];
            for (int c = 0; c < nx; ++c) s += P.F[i * nx + c] * xs[c];
            for (int c = 0; c < nu; ++c) s += P.G[i * nu + c] * us[c];
            rmu[i] = s;
        }
        // plant update and applied input
        for (int j = threadIdx.x; j < nx; j += SH_NT) if (x_next) x_next[(size_t)inst * nx + j] = ip[nx + j] + es[j];
        for (int j = threadIdx.x; j < nu; j += SH_NT) if (u0_out) u0_out[(size_t)inst * nu + j] = us[j];
        __syncthreads();

        const size_t oo = (size_t)inst * ot.cap_nodes, on_ = (size_t)inst * nt.cap_nodes;
        const int nn = ot.n_nodes[inst];
        // ---- pass 1: _retain_leaf (controller.py:615-633) + ordered compaction + identifier shift (:476)
        for (int base = 0; base < nn; base += SH_NT) {
            const int j = base + threadIdx.x;
            int keep = 0;
            if (j < nn && ot.alive[oo + j]) {
                keep = 1;
                const int d = ot.depth[oo + j];
                const unsigned int b0 = ot.bits[(oo + j) * ot.words];
                const int lim = d < nub ? d : nub;
                for (int i = 0; i < lim; ++i)
                    if ((double)((b0 >> i) & 1u) != us[nuc + i]) keep = 0;
            }
            const unsigned int bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) s_wsum[w] = __popc(bal);
            __syncthreads();
            int pre = s_base;
            for (int q = 0; q < w; ++q) pre += s_wsum[q];
            const int idx = pre + __popc(bal & ((1u << lane) - 1u));
            if (keep && idx < nt.cap_nodes) {
                const int d = ot.depth[oo + j];
                const int dn = d > nub ? d - nub : 0;
                nt.depth[on_ + idx] = dn; nt.alive[on_ + idx] = 1; nt.rec[on_ + idx] = j;   // rec = source node (pass 2 rewrites it)
                const unsigned int *src = ot.bits + (oo + j) * ot.words;
                unsigned int *dst = nt.bits + (on_ + idx) * nt.words;
                for (int q = 0; q < nt.words; ++q) {
                    // shift the bit string right by nub bits
                    const int sb = q * 32 + nub;
                    const int wq = sb >> 5, sh = sb & 31;
                    unsigned int lo = wq < ot.words ? src[wq] : 0u, hi = wq + 1 < ot.words ? src[wq + 1] : 0u;
                    unsigned int v = sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
                    const int valid = dn - q * 32;      // keep only bits < dn
                    if (valid <= 0) v = 0u; else if (valid < 32) v &= (1u << valid) - 1u;
                    dst[q] = v;
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) { int s = s_base; for (int q = 0; q < SH_NW; ++q) s += s_wsum[q]; s_base = s; }
            __syncthreads();
        }
        const int nnew = s_base < nt.cap_nodes ? s_base : nt.cap_nodes;
        // ---- pass 2: one warp per retained leaf
        for (int idx = w; idx < nnew; idx += SH_NW) {
            const int j = nt.rec[on_ + idx];
            const int ro = ot.rec[oo + j];
            const double lbo = ot.lb[oo + j];
            double *E = nt.rec_dual + ((size_t)inst * nt.cap_recs + idx) * P.n_dual;
            if (ro < 0) {
                // dual = None (controller.py:556-558 on the previous step, never solved since): trivial bound
                for (int e = lane; e < P.n_dual; e += 32) E[e] = 0.;
                if (lane == 0) { nt.lb[on_ + idx] = 0.; nt.rec[on_ + idx] = -1; nt.rec_dobj[(size_t)inst * nt.cap_recs + idx] = 0.; }
                continue;
            }
            const double *D = ot.rec_dual + ((size_t)inst * ot.cap_recs + ro) * P.n_dual;
            const int d_old = ot.depth[oo + j];
            const unsigned int b0 = ot.bits[(oo + j) * ot.words];
            double acc = 0.;                     // pi_sum + pi3, lane-partial
            // lam: drop t = 0, append zero ; pi3 = -lam'_0 . e0 (controller.py:544)
            {
                const double *s = D + P.off_lam + nx; double *t = E + P.off_lam;
                for (int e = lane; e < T * nx; e += 32) { const double v = s[e]; t[e] = v; if (e < nx) acc -= v * es[e]; }
                for (int e = lane; e < nx; e += 32) t[T * nx + e] = 0.;
            }
            // nu_lb, nu_ub: complementarity terms with the OLD identifier's bounds at t = 0 (controller.py:703-709)
            {
                const double *sl = D + P.off_nulb, *su = D + P.off_nuub;
                double *tl = E + P.off_nulb, *tu = E + P.off_nuub;
                for (int e = lane; e < (T - 1) * nub; e += 32) { tl[e] = sl[nub + e]; tu[e] = su[nub + e]; }
                for (int e = lane; e < nub; e += 32) {
                    tl[(T - 1) * nub + e] = 0.; tu[(T - 1) * nub + e] = 0.;
                    const double bit = (double)((b0 >> e) & 1u);
                    const double l0 = e < d_old ? bit : 0., u0b = e < d_old ? bit : 1.;
                    const double vu = us[nuc + e];
                    acc -= (l0 - vu) * sl[e] + (vu - u0b) * su[e];
                }
            }
            // sigma: suboptimality term |sigma_0 / 2 - R u0|^2 - |R u0|^2
            {
                const double *s = D + P.off_sigma; double *t = E + P.off_sigma;
                for (int e = lane; e < (T - 1) * nr_; e += 32) t[e] = s[nr_ + e];
                for (int e = lane; e < nr_; e += 32) { t[(T - 1) * nr_ + e] = 0.; const double a = .5 * s[e] - Ru[e]; acc += a * a - Ru[e] * Ru[e]; }
            }
            // rho: rho'_{T-1} = M_rho rho_T (controller.py:96, 662-664)
            {
                const double *s = D + P.off_rho; double *t = E + P.off_rho;
                for (int e = lane; e < (T - 1) * nq; e += 32) t[e] = s[nq + e];
                for (int e = lane; e < nq; e += 32) { const double a = .5 * s[e] - Qx[e]; acc += a * a - Qx[e] * Qx[e]; }
                double *rT = wscr;                              // rho_T staged for the small mat-vec
                for (int e = lane; e < nqT; e += 32) { const double v = s[T * nq + e]; rT[e] = v; acc += .25 * v * v; t[T * nq + e] = 0.; }
                __syncwarp();
                for (int i = lane; i < nq; i += 32) {
                    double v = 0.;
                    for (int c = 0; c < nqT; ++c) v += P.Mrho[i * nqT + c] * rT[c];
                    t[(T - 1) * nq + i] = v; acc -= .25 * v * v;
                }
            }
            // mu: mu'_{T-2} = M_mu mu_{T-1} (controller.py:186-227, 662-664)
            {
                const double *s = D + P.off_mu; double *t = E + P.off_mu;
                for (int e = lane; e < (T - 2) * nh; e += 32) t[e] = s[nh + e];
                for (int e = lane; e < nh; e += 32) acc -= rmu[e] * s[e];
                double *mT = wscr + nqT;
                for (int e = lane; e < nh1; e += 32) { const double v = s[(T - 1) * nh + e]; mT[e] = v; acc += P.h1[e] * v; t[(T - 1) * nh + e] = 0.; }
                __syncwarp();
                for (int i = lane; i < nh; i += 32) {
                    double v = 0.;
                    for (int c = 0; c < nh1; ++c) v += P.Mmu[(size_t)i * nh1 + c] * mT[c];
                    t[(T - 2) * nh + i] = v; acc -= P.h[i] * v;
                }
                __syncwarp();
            }
            acc = warp_sum(acc);
            if (lane == 0) {
                double obj = ot.rec_dobj[(size_t)inst * ot.cap_recs + ro] + acc;
                obj = obj > 0. ? obj : 0.;                     // controller.py:546
                double lbn; int rn = idx;
                if (!isinf(lbo)) lbn = obj;                    // :550-551
                else if (obj <= 0.) { lbn = 0.; rn = -1; }     // :555-558
                else lbn = INFINITY;
                nt.lb[on_ + idx] = lbn; nt.rec[on_ + idx] = rn; nt.rec_dobj[(size_t)inst * nt.cap_recs + idx] = obj;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) { nt.n_nodes[inst] = nnew; nt.n_recs[inst] = nnew; }
    }
}

__host__ inline size_t shift_smem_bytes(const DevProblem &P) {
    return sizeof(double) * ((size_t)2 * P.nx + P.nu + P.nq + P.nr + P.nh + (size_t)SH_NW * (P.nh1 + P.nqT + P.nh + P.nq) + 8);
}
