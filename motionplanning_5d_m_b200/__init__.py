"""motionplanning_5d_m_b200 -- B200-native (sm_100a) Convex-Feasible-Set hot path behind the reference's interface.

Compute lives in libcfs_b200.so (csrc/, C ABI in include/cfs_b200.h); this package is the host-side mirror of the
reference's MATLAB interface for that path (CFS_FANUC, PSGCFS_FANUC, EVAL, robotproperty2, cost set-up).
"""
from ._lib import (CfsError, Context, FLAG_TOUCH, GRAD_DERIVEST, GRAD_NUMJAC, SOLVER_CFS, SOLVER_PSGCFS,  # noqa: F401
                   STATUS_CONVERGED, STATUS_INFEASIBLE, STATUS_MAX_ITER, STATUS_NO_ROUTE, STATUS_NUMERICAL)
from .cfs import CFS_FANUC, PSGCFS_FANUC, CHOMP_FANUC, BatchCFS, EVAL  # noqa: F401
from .problem import (build_cost_matrices, build_linear_term, make_sys_info, straight_line_reference)  # noqa: F401
from .robot import robotproperty2  # noqa: F401
from .rrt import RRT_FANUC, s_Parallel_rrt, rrtstar_cfs  # noqa: F401
