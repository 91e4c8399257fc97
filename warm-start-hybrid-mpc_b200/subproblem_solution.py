"""Per-node solution containers, same attribute names as the reference
(subproblem_solution.py:4-168): ``SubproblemSolution(primal, dual, active_set)``,
``PrimalSolution(variables, objective, binary_feasible)``, ``DualSolution(variables, objective)``.
Here they are views built from the flat records the CUDA kernels write (include/wshmpc.h layout)."""
import numpy as np


class SubproblemSolution(object):

    def __init__(self, primal, dual, active_set=None):
        self.primal = primal
        self.dual = dual
        self.active_set = active_set

    @staticmethod
    def from_controller(controller):
        """subproblem_solution.py:18-45: primal + dual solution of the QP the controller has just solved, and (with
        Params.Method == 1) the active set its children start from."""
        primal = PrimalSolution.from_controller(controller)
        dual = DualSolution.from_controller(controller)
        active_set = controller.qp.get_active_set() if controller.qp.Params.Method == 1 else None
        return SubproblemSolution(primal, dual, active_set)


class PrimalSolution(object):
    """variables = {'x': [T+1 arrays], 'uc': [T arrays], 'ub': [T arrays]} (None if infeasible)."""

    def __init__(self, variables, objective, binary_feasible):
        self.variables = variables
        self.objective = objective
        self.binary_feasible = binary_feasible

    @staticmethod
    def from_controller(controller):
        """subproblem_solution.py:68-99.  binary_feasible iff EVERY binary is pinned by the node (:94-97)."""
        qp = controller.qp
        qp._raise_if_not_solved()
        T = controller.T
        lb = np.concatenate([qp.get_constraint_rhs('nu_lb_%d' % t) for t in range(T)])
        ub = np.concatenate([qp.get_constraint_rhs('nu_ub_%d' % t) for t in range(T)])
        return PrimalSolution.from_record(controller.problem, qp._primal, qp.primal_objective(),
                                          bool(np.array_equal(lb, -ub)), qp.status == 2)

    @staticmethod
    def from_record(pd, rec, objective, binary_feasible, feasible):
        """subproblem_solution.py:86-99: every family is None when the node is infeasible."""
        T, nx, nu, nuc = pd.T, pd.nx, pd.nu, pd.nuc
        if not feasible:
            variables = {'x': [None] * (T + 1), 'uc': [None] * T, 'ub': [None] * T}
            return PrimalSolution(variables, np.inf, binary_feasible)
        X = rec[:(T + 1) * nx].reshape(T + 1, nx)
        U = rec[(T + 1) * nx:].reshape(T, nu)
        variables = {'x': [X[t].copy() for t in range(T + 1)],
                     'uc': [U[t, :nuc].copy() for t in range(T)],
                     'ub': [U[t, nuc:].copy() for t in range(T)]}
        return PrimalSolution(variables, float(objective), binary_feasible)


class DualSolution(object):
    """variables = {'lam','mu','nu_lb','nu_ub','rho','sigma'}: lists over time of numpy arrays;
    objective = dual objective, or cost of the Farkas proof for an infeasible node."""

    def __init__(self, variables, objective):
        self.variables = variables
        self.objective = objective

    @staticmethod
    def from_controller(controller):
        """subproblem_solution.py:119-168.  rho_t = 2 Q x_t, sigma_t = 2 R u_t (zero for a Farkas proof) are part of the
        record the kernel writes (csrc/records.cuh), so they are the very numbers the device tree holds."""
        qp = controller.qp
        qp._raise_if_not_solved()
        return DualSolution.from_record(controller.problem, controller.problem.layout, qp._dual, qp.dual_objective())

    @staticmethod
    def from_record(pd, layout, rec, objective):
        T, nx, nub, nh, nh1, nq, nqT, nr = pd.T, pd.nx, pd.nub, pd.nh, pd.nh1, pd.nq, pd.nqT, pd.nr
        lam = rec[layout.off_lam:layout.off_mu].reshape(T + 1, nx)
        mu = rec[layout.off_mu:layout.off_nu_lb]
        nulb = rec[layout.off_nu_lb:layout.off_nu_ub].reshape(T, nub)
        nuub = rec[layout.off_nu_ub:layout.off_rho].reshape(T, nub)
        rho = rec[layout.off_rho:layout.off_sigma]
        sigma = rec[layout.off_sigma:layout.dual].reshape(T, nr)
        variables = {
            'lam': [lam[t].copy() for t in range(T + 1)],
            'mu': [mu[t * nh:(t + 1) * nh].copy() for t in range(T - 1)] + [mu[(T - 1) * nh:(T - 1) * nh + nh1].copy()],
            'nu_lb': [nulb[t].copy() for t in range(T)],
            'nu_ub': [nuub[t].copy() for t in range(T)],
            'rho': [rho[t * nq:(t + 1) * nq].copy() for t in range(T)] + [rho[T * nq:T * nq + nqT].copy()],
            'sigma': [sigma[t].copy() for t in range(T)],
        }
        return DualSolution(variables, float(objective))

    @staticmethod
    def to_record(pd, layout, variables):
        """inverse of from_record (used to hand warm-start duals back to the device kernels)."""
        rec = np.zeros(layout.dual)
        rec[layout.off_lam:layout.off_mu] = np.concatenate(variables['lam'])
        rec[layout.off_mu:layout.off_nu_lb] = np.concatenate(variables['mu'])
        rec[layout.off_nu_lb:layout.off_nu_ub] = np.concatenate(variables['nu_lb'])
        rec[layout.off_nu_ub:layout.off_rho] = np.concatenate(variables['nu_ub'])
        rec[layout.off_rho:layout.off_sigma] = np.concatenate(variables['rho'])
        rec[layout.off_sigma:layout.dual] = np.concatenate(variables['sigma'])
        return rec
