"""B200-native branch-and-bound hot path of warm-start-hybrid-mpc.

Same public names as the reference package ``warm_start_hmpc`` for the hot path:
``MLDSystem``, ``BoundedQP``, ``HybridModelPredictiveController`` (``feedforward`` / ``construct_warm_start``),
``branch_and_bound``, ``Node``, ``best_first`` / ``depth_first`` / ``breadth_first``,
``branch_in_time``, ``SubproblemSolution`` / ``PrimalSolution`` / ``DualSolution``.
The QP relaxations run in hand-written sm_100a CUDA kernels behind a C ABI (include/wshmpc.h,
csrc/); there is no CPU fallback: importing is fine anywhere, solving needs the built library and a GPU.
"""
from .mld_system import MLDSystem                                               # noqa: F401
from .subproblem_solution import SubproblemSolution, PrimalSolution, DualSolution  # noqa: F401
from .branch_and_bound import Node, branch_and_bound, best_first, depth_first, breadth_first  # noqa: F401
from .controller import (HybridModelPredictiveController, branch_in_time, StaticBranchOrder,   # noqa: F401
                         branch_in_reverse_time, branch_by_index)
from .bounded_qp import BoundedQP                                                # noqa: F401
