"""scpn_fusion_core_b200 - B200-native Grad-Shafranov equilibrium hot path.

A drop-in for the equilibrium entry points of anulum/scpn-fusion-core
(``FusionKernel``, ``multigrid_solve``, the ``gs_rb_sor_smooth`` tier and the
``libscpn_solver.so`` C ABI), built from scratch as hand-written sm_100a CUDA kernels
behind a C ABI (``include/gsb200.h``).  See DESIGN.md / INTEGRATION.md.
"""
from . import _lib  # noqa: F401
from . import dataset, eqdsk, free_boundary  # noqa: F401
from .fusion_kernel import (  # noqa: F401
    BatchedFusionKernel, CoilSet, FusionKernel, shard_range, solve_sharded, validate_config,
)
from .multigrid_solve import (  # noqa: F401
    mg_residual, mg_smooth, multigrid_solve, multigrid_vcycle, prolongate_bilinear, residual_linf,
    restrict_full_weight, validate_sor_omega,
)

__version__ = "0.1.0"
