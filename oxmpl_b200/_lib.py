"""ctypes binding of liboxg.so — the C ABI declared in include/oxg.h.

There is no CPU fallback: if the shared library has not been built (``python -c
"import __graft_entry__ as g; g.build()"`` or ``make -C oxmpl_b200/csrc``) importing this
module raises, and every compute entry point fails with OXG_E_CUDA when no CUDA device is
present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OXG_LIB_PATH") or os.path.join(_HERE, "liboxg.so")  # override: A/B builds

OXG_OK = 0
OXG_E_NULL = -1
OXG_E_DIM = -2
OXG_E_CAPACITY = -3
OXG_E_CUDA = -4
OXG_E_RESOLUTION = -5
OXG_E_ARG = -6
OXG_E_EMPTY = -7
OXG_E_ALLOC = -8
OXG_E_INTERNAL = -9
OXG_E_NCCL = -10
OXG_INVALID_INDEX = 0xFFFFFFFF

ERROR_NAMES = {
    OXG_E_NULL: "OXG_E_NULL",
    OXG_E_DIM: "OXG_E_DIM",
    OXG_E_CAPACITY: "OXG_E_CAPACITY",
    OXG_E_CUDA: "OXG_E_CUDA",
    OXG_E_RESOLUTION: "OXG_E_RESOLUTION",
    OXG_E_ARG: "OXG_E_ARG",
    OXG_E_EMPTY: "OXG_E_EMPTY",
    OXG_E_ALLOC: "OXG_E_ALLOC",
    OXG_E_INTERNAL: "OXG_E_INTERNAL",
    OXG_E_NCCL: "OXG_E_NCCL",
}


class SpaceDesc(C.Structure):
    """struct oxg_space_desc."""

    _fields_ = [
        ("n_real", C.c_uint32),
        ("n_so2", C.c_uint32),
        ("w_real", C.c_double),
        ("w_so2", C.c_double),
    ]


class PrmStats(C.Structure):
    """struct oxg_prm_stats."""

    _fields_ = [
        ("n_rows", C.c_uint64),
        ("n_candidates", C.c_uint64),
        ("n_edges_checked", C.c_uint64),
        ("n_edges_valid", C.c_uint64),
        ("knn_ms", C.c_double),
        ("edge_ms", C.c_double),
    ]


class PrmQueryResult(C.Structure):
    """struct oxg_prm_query_result."""

    _fields_ = [
        ("status", C.c_int32),
        ("path_len", C.c_uint32),
        ("goal_node", C.c_uint32),
        ("n_start_connections", C.c_uint32),
        ("n_goal_nodes", C.c_uint32),
        ("bfs_levels", C.c_uint32),
        ("n_visited", C.c_uint32),
        ("reserved", C.c_uint32),
        ("n_adjacency", C.c_uint64),
        ("device_ms", C.c_double),
    ]


class PlanParams(C.Structure):
    """struct oxg_plan_params."""

    _fields_ = [
        ("max_distance", C.c_double),
        ("search_radius", C.c_double),
        ("resolution", C.c_double),
        ("goal_radius", C.c_double),
        ("timeout_s", C.c_double),
        ("max_nodes", C.c_uint64),
        ("stop_at_goal", C.c_int32),
        ("reserved", C.c_int32),
    ]


class PlanResult(C.Structure):
    """struct oxg_plan_result."""

    _fields_ = [
        ("n_nodes", C.c_uint64 * 2),
        ("iters", C.c_uint64),
        ("motion_checks", C.c_uint64),
        ("motion_checks_executed", C.c_uint64),
        ("goal_node", C.c_int64 * 2),
        ("solved", C.c_int32),
        ("status", C.c_int32),
        ("device_ms", C.c_double),
        ("ctas", C.c_int32),
        ("reserved", C.c_int32),
        ("phase_ms", C.c_double * 8),
    ]


_vp = C.c_void_p
_u64 = C.c_uint64
_u32 = C.c_uint32
_dbl = C.c_double
_pd = C.c_void_p  # double*   (passed as raw addresses from numpy / torch)
_pu32 = C.c_void_p
_pu64 = C.c_void_p
_pu8 = C.c_void_p

# name -> (restype, argtypes); every symbol include/oxg.h and include/oxg_so3.h declare.
SIGNATURES = {
    "oxg_version": (C.c_int, []),
    "oxg_last_error": (C.c_char_p, []),
    "oxg_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "oxg_set_device": (C.c_int, [C.c_int]),
    "oxg_store_create": (C.c_int, [C.POINTER(SpaceDesc), _u64, C.POINTER(_vp)]),
    "oxg_store_destroy": (C.c_int, [_vp]),
    "oxg_store_clear": (C.c_int, [_vp]),
    "oxg_store_reserve": (C.c_int, [_vp, _u64]),
    "oxg_store_append": (C.c_int, [_vp, _pd, _u64, C.POINTER(_u64)]),
    "oxg_store_size": (C.c_int, [_vp, C.POINTER(_u64)]),
    "oxg_store_read": (C.c_int, [_vp, _u64, _u64, _pd]),
    "oxg_store_sync": (C.c_int, [_vp]),
    "oxg_store_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "oxg_nn1": (C.c_int, [_vp, _pd, _u64, _pu32, _pd]),
    "oxg_knn": (C.c_int, [_vp, _pd, _u64, _u32, _pu32, _pu32, _pd]),
    "oxg_radius": (C.c_int, [_vp, _pd, _u64, _dbl, _pu32, _pu64, _pu32, _u64, _pd]),
    "oxg_world_create": (C.c_int, [_u32, _pd, _u64, _pd, _u64, _pd, _u64, C.POINTER(_vp)]),
    "oxg_world_add_so2_interval": (C.c_int, [_vp, _u32, _dbl, _dbl]),
    "oxg_world_destroy": (C.c_int, [_vp]),
    "oxg_world_sync": (C.c_int, [_vp]),
    "oxg_is_valid": (C.c_int, [_vp, C.POINTER(SpaceDesc), _pd, _u64, _pu8]),
    "oxg_check_motion": (C.c_int, [_vp, C.POINTER(SpaceDesc), _pd, _pd, _u64, _dbl, _pu8, _pu32]),
    "oxg_check_edges": (C.c_int, [_vp, _vp, _pu32, _pu32, _u64, _dbl, _pu8]),
    "oxg_prm_build_radius": (
        C.c_int,
        [_vp, _vp, _dbl, _dbl, _u64, _u64, _pu64, _pu32, _u64, C.POINTER(PrmStats)],
    ),
    "oxg_prm_build_knn": (
        C.c_int,
        [_vp, _vp, _u32, _dbl, _u64, _u64, _pu32, _pd, _pu8, C.POINTER(PrmStats)],
    ),
    "oxg_prm_solve": (
        C.c_int,
        [_vp, _vp, _pu64, _pu32, _dbl, _dbl, _pd, _pd, _dbl, _pu32, _u32, C.POINTER(PrmQueryResult)],
    ),
    "oxg_prm_solve_knn": (
        C.c_int,
        [_vp, _vp, _pu32, _pu8, _u32, _dbl, _dbl, _pd, _pd, _dbl, _pu32, _u32, C.POINTER(PrmQueryResult)],
    ),
    "oxg_rrt_solve": (
        C.c_int,
        [_vp, _vp, C.POINTER(PlanParams), _pd, _pd, _u64, _vp, C.POINTER(PlanResult)],
    ),
    "oxg_rrt_connect_solve": (
        C.c_int,
        [_vp, _vp, _vp, C.POINTER(PlanParams), _pd, _pd, _u64, _vp, _vp, C.POINTER(PlanResult)],
    ),
    "oxg_rrt_star_solve": (
        C.c_int,
        [_vp, _vp, C.POINTER(PlanParams), _pd, _pd, _u64, _vp, _vp, C.POINTER(PlanResult)],
    ),
    "oxg_knn_rows_dev": (C.c_int, [_vp, _u64, _u64, _u32, _pu32, _pd]),
    "oxg_prm_edges_dev": (C.c_int, [_vp, _vp, _pu32, _u32, _dbl, _u64, _u64, _pu8, _pu64]),
    "oxg_prm_resolve_dev": (C.c_int, [_vp, _pu32, _u32, _pu8]),
    "oxg_dist_unique_id": (C.c_int, [_vp]),
    "oxg_dist_create": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "oxg_dist_destroy": (C.c_int, [_vp]),
    "oxg_dist_row_range": (C.c_int, [_u64, C.c_int, C.c_int, C.c_int, C.POINTER(_u64), C.POINTER(_u64)]),
    "oxg_store_append_dist": (C.c_int, [_vp, _vp, _pd, _u64, C.POINTER(_u64)]),
    "oxg_prm_build_knn_dist": (C.c_int, [_vp, _vp, _vp, _u32, _dbl, _pu32, _pd, _pu8, C.POINTER(PrmStats)]),
    "oxg_prm_build_knn_dist_dev": (C.c_int, [_vp, _vp, _vp, _u32, _dbl, _pu32, _pd, _pu8, _pu64]),
    "oxg_prm_build_radius_dist": (
        C.c_int, [_vp, _vp, _vp, _dbl, _dbl, _pu64, _pu32, _u64, C.POINTER(PrmStats)]),
    "oxg_store_last_scan_ms": (C.c_int, [_vp, C.POINTER(_dbl)]),
    "oxg_fp64_peak": (C.c_int, [C.POINTER(_dbl), C.POINTER(_dbl)]),
    "oxg_fp32_peak": (C.c_int, [C.POINTER(_dbl), C.POINTER(_dbl)]),
    "oxg_l2_peak": (C.c_int, [C.POINTER(_dbl), C.POINTER(_dbl)]),
    "oxg_kernel_launches": (_u64, [C.c_int]),
    "oxg_tuning_set": (C.c_int, [C.c_char_p, C.c_int64]),
    "oxg_store_tree_stats": (C.c_int, [_vp, C.POINTER(_dbl)]),
    # include/oxg_so3.h
    "oxg_so3_nn1": (C.c_int, [_vp, _pd, _u64, _pu32, _pd]),
    "oxg_so3_radius": (C.c_int, [_vp, _pd, _u64, _dbl, _pu64, _pu32, _u64, _pd]),
    "oxg_so3_knn": (C.c_int, [_vp, _pd, _u64, C.c_uint32, _pu32, _pu32, _pd]),
    "oxg_so3_rrt_solve": (
        C.c_int,
        [_vp, _vp, C.POINTER(PlanParams), _pd, _pd, _u64, _vp, C.POINTER(PlanResult)],
    ),
    "oxg_so3_rrt_connect_solve": (
        C.c_int,
        [_vp, _vp, _vp, C.POINTER(PlanParams), _pd, _pd, _u64, _vp, _vp, C.POINTER(PlanResult)],
    ),
    "oxg_so3_cones_create": (C.c_int, [_pd, _u64, C.POINTER(_vp)]),
    "oxg_so3_cones_destroy": (C.c_int, [_vp]),
    "oxg_so3_is_valid": (C.c_int, [_vp, _pd, _u64, _pu8]),
    "oxg_so3_check_motion": (C.c_int, [_vp, _pd, _pd, _u64, _dbl, _pu8, _pu32]),
}


class OxgError(RuntimeError):
    """A liboxg entry point returned a negative OXG_E_* code."""

    def __init__(self, code: int, message: str):
        self.code = code
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()'). "
            "oxmpl_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(code: int) -> None:
    if code != OXG_OK:
        raise OxgError(code, lib.oxg_last_error().decode("utf-8", "replace"))
