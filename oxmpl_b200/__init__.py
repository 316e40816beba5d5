"""oxmpl_b200 — B200 (sm_100a) nearest-neighbour / validity backend for oxmpl's planners.

The product is ``liboxg.so`` (hand-written CUDA behind the C ABI in ``include/oxg.h``);
this package is the thin host-side binding.  Importing it requires the built library —
there is no CPU fallback.
"""
from .api import (  # noqa: F401
    INVALID_INDEX,
    So3Cones,
    Space,
    StateStore,
    World,
    device_count,
    fp32_peak,
    fp64_peak,
    kernel_launches,
    l2_peak,
    motion_resolution,
    prm_build_knn,
    prm_build_radius,
    prm_solve,
    prm_solve_knn,
    real_vector_max_extent,
    rrt_connect_solve,
    rrt_solve,
    rrt_star_solve,
    so3_rrt_connect_solve,
    so3_rrt_solve,
    tuning_set,
)
from ._lib import OxgError, PlanParams, PlanResult, PrmQueryResult, PrmStats, SpaceDesc  # noqa: F401

__version__ = "0.1.0"
