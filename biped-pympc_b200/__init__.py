"""biped-pympc_b200 -- B200-native batched SRBD-MPC solve (QP assembly + PDIPM).

Drop-in for the hot path of rl-augmented-mpc/Biped-PyMPC: the two CusADi functions
``qp_former`` and ``sparse_pdipm_multiple_iterations`` and the glue in
``biped_pympc/convex_mpc/mpc_controller_cusadi.py``.  See DESIGN.md / INTEGRATION.md.

The directory name contains a hyphen; import it through the ``biped_pympc_b200`` shim at
the repository root (``import biped_pympc_b200 as bp``)."""
from . import _native, workload  # noqa: F401
from ._native import (COLD_START, STAGE_INVARIANT, Context, SrbdError, dfma_peak, is_stage_invariant, marshal, pattern,  # noqa: F401
                      pdipm, qp_form, t1_inverse_kinematics, t1_kinematics, wrench)
from .cusadi_function import B200Function, PDIPM_FUNCTION, QP_FORMER_FUNCTION  # noqa: F401
from .biped_controller import BipedControllerB200  # noqa: F401
from .gait_generator import GaitGeneratorB200  # noqa: F401
from .mpc_controller_b200 import MPCControllerB200, solve_tick  # noqa: F401
from .sharding import shard_bounds, ShardedSolver  # noqa: F401
from .t1 import T1B200  # noqa: F401

__all__ = ["qp_form", "pdipm", "wrench", "Context", "B200Function", "MPCControllerB200", "GaitGeneratorB200", "BipedControllerB200", "solve_tick",
           "shard_bounds", "ShardedSolver", "workload", "dfma_peak", "pattern", "T1B200", "t1_kinematics", "t1_inverse_kinematics"]
