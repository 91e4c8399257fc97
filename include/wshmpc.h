/* wshmpc.h -- C ABI of the B200-native hybrid-MPC branch-and-bound hot path.
 *
 * Drop-in boundary for the per-node / per-solve seams of TobiaMarcucci/warm-start-hybrid-mpc
 * (SURVEY.md section 8b).  The reference has no FFI layer: its seams are Python callables, so every
 * entry point below cites the reference callable it replaces.  All matrices are fp64, row-major,
 * caller-owned; "d_" arguments are DEVICE pointers (torch CUDA tensors' data_ptr()), everything else
 * is host memory.  Every call returns 0 on success, < 0 on error (message in wshmpc_last_error).
 * One handle per (GPU, stream); not thread-safe; never throws; never falls back to the CPU.
 */
#ifndef WSHMPC_H
#define WSHMPC_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct wshmpc_handle wshmpc_handle;

/* Problem description: the MLD system + controller data of
 * HybridModelPredictiveController.__init__ (controller.py:58-97) and the shared least-distance
 * operator the QP kernel runs on (host-side precompute, north_star "condensing stays host-side";
 * built by warm-start-hybrid-mpc_b200/problem.py). */
typedef struct {
    /* sizes */
    int nx, nu, nub, T;          /* mld_system.py:37-41, controller.py:78 */
    int nh, nh1;                 /* rows of F (stage) and of F_Tm1 (last stage, controller.py:85-87) */
    int nq, nqT, nr;             /* rows of Q, Q_T, R (controller.py:79) */
    int n, m, mc, nb;            /* n = T*nu condensed inputs, m = mc + nb rows, nb = T*nub */
    /* stage data (host pointers, row-major) */
    const double *A, *B, *F, *G, *h, *F_Tm1, *G_Tm1, *h_Tm1, *Q, *R, *Q_T;
    const double *M_mu;          /* nh x nh1, controller.py:186-227 */
    const double *M_rho;         /* nq x nqT, controller.py:96 */
    /* shared least-distance operator (host pointers) */
    const double *Mh;            /* m x n, unit rows */
    const double *nrm;           /* m */
    const double *vscale;        /* m */
    const double *Eh;            /* mc x nx */
    const double *hh;            /* mc */
    const double *Rinv;          /* n x n */
    const double *Kx;            /* n x nx */
    const double *Zmap;          /* n x n */
    const int *bin_idx;          /* nb */
    /* solver parameters */
    double eps, tol_p, tol_d, tol_sing, tol_ray, prox_tol;
    int max_iter, max_prox;
} wshmpc_problem;

/* sizes of the per-node records, in doubles */
typedef struct {
    int primal;                  /* (T+1)*nx + T*nu : x_0..x_T, u_0..u_{T-1}  (subproblem_solution.py:86-91) */
    int dual;                    /* lam | mu | nu_lb | nu_ub | rho | sigma     (subproblem_solution.py:137-166) */
    int off_lam, off_mu, off_nu_lb, off_nu_ub, off_rho, off_sigma;
} wshmpc_layout;

const char *wshmpc_last_error(void);

/* create / destroy.  `n_slots` = number of independent solver states (one per concurrently solved
 * MPC instance); `stream` is a cudaStream_t passed as void* (0 = default stream). */
int wshmpc_create(const wshmpc_problem *p, int device, int n_slots, void *stream, wshmpc_handle **out);
int wshmpc_destroy(wshmpc_handle *h);
int wshmpc_get_layout(const wshmpc_handle *h, wshmpc_layout *out);

/* K1 -- batched node QP relaxation.
 * Replaces, per node: controller._solve_subproblem (controller.py:229-271) = _set_bound_binaries
 * (:273-298) + BoundedQP.optimize (bounded_qp.py:200-228) + SubproblemSolution.from_controller
 * (subproblem_solution.py:18-45).
 *   d_x0   [n_nodes][nx]     initial state of each node's problem          (rhs of 'lam_0')
 *   d_lb   [n_nodes][nb]     lower bounds on the relaxed binaries, (t,i) -> t*nub+i   (-rhs of 'nu_lb_t')
 *   d_ub   [n_nodes][nb]     upper bounds                                  (rhs of 'nu_ub_t')
 *   d_slot [n_nodes]         solver state used by the node; nodes sharing a slot are solved in index
 *                            order by one CTA, each hot-started from the previous one (any dual
 *                            feasible point of one node is dual feasible for every other node)
 *   d_hot  [n_nodes]         0: reset the slot to the empty working set before the solve, 1: keep it
 * outputs
 *   d_status [n_nodes]       2 optimal, 3 infeasible (Gurobi status codes, bounded_qp.py:212), 9 iteration limit
 *   d_cost   [n_nodes]       primal objective, +inf if infeasible          (bounded_qp.py:292-311)
 *   d_dobj   [n_nodes]       dual objective / cost of the Farkas proof     (bounded_qp.py:313-332)
 *   d_iters  [n_nodes]       active-set iterations
 *   d_primal [n_nodes][layout.primal]   (undefined if infeasible)
 *   d_dual   [n_nodes][layout.dual]
 */
int wshmpc_solve_nodes(wshmpc_handle *h, int n_nodes, const double *d_x0, const double *d_lb,
                       const double *d_ub, const int *d_slot, const int *d_hot,
                       int *d_status, double *d_cost, double *d_dobj, int *d_iters,
                       double *d_primal, double *d_dual);

#ifdef __cplusplus
}
#endif
#endif
