//! Raw `extern "C"` declarations of include/oxg.h (one-to-one; keep in sync with the header —
//! tests/test_abi.py checks the header against the built library).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct oxg_space_desc {
    pub n_real: u32,
    pub n_so2: u32,
    pub w_real: f64,
    pub w_so2: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct oxg_prm_stats {
    pub n_rows: u64,
    pub n_candidates: u64,
    pub n_edges_checked: u64,
    pub n_edges_valid: u64,
    pub knn_ms: f64,
    pub edge_ms: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct oxg_plan_params {
    pub max_distance: f64,
    pub search_radius: f64,
    pub resolution: f64,
    pub goal_radius: f64,
    pub timeout_s: f64,
    pub max_nodes: u64,
    pub stop_at_goal: i32,
    pub reserved: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct oxg_plan_result {
    pub n_nodes: [u64; 2],
    pub iters: u64,
    pub motion_checks: u64,
    pub motion_checks_executed: u64,
    pub goal_node: [i64; 2],
    pub solved: i32,
    pub status: i32,
    pub device_ms: f64,
    pub ctas: i32,
    pub reserved: i32,
    pub phase_ms: [f64; 8],
}

pub type oxg_store = *mut c_void;
pub type oxg_world = *mut c_void;

pub const OXG_OK: c_int = 0;
pub const OXG_E_CAPACITY: c_int = -3;
pub const OXG_E_NCCL: c_int = -10;

extern "C" {
    pub fn oxg_last_error() -> *const c_char;
    pub fn oxg_device_count(out_count: *mut c_int) -> c_int;
    pub fn oxg_set_device(device: c_int) -> c_int;

    pub fn oxg_store_create(space: *const oxg_space_desc, capacity: u64, out: *mut oxg_store) -> c_int;
    pub fn oxg_store_destroy(s: oxg_store) -> c_int;
    pub fn oxg_store_clear(s: oxg_store) -> c_int;
    pub fn oxg_store_append(s: oxg_store, rows: *const f64, n: u64, first_index: *mut u64) -> c_int;
    pub fn oxg_store_size(s: oxg_store, out_n: *mut u64) -> c_int;
    pub fn oxg_store_read(s: oxg_store, first: u64, n: u64, rows_out: *mut f64) -> c_int;

    pub fn oxg_nn1(s: oxg_store, queries: *const f64, q: u64, idx: *mut u32, dist: *mut f64) -> c_int;
    pub fn oxg_knn(s: oxg_store, queries: *const f64, q: u64, k: u32, exclude_idx: *const u32,
                   idx: *mut u32, dist: *mut f64) -> c_int;
    pub fn oxg_radius(s: oxg_store, queries: *const f64, q: u64, r: f64, idx_limit: *const u32,
                      offsets: *mut u64, idx: *mut u32, cap: u64, dist: *mut f64) -> c_int;

    pub fn oxg_world_create(dim: u32, spheres: *const f64, n_spheres: u64, aabbs: *const f64,
                            n_aabbs: u64, capsules: *const f64, n_capsules: u64,
                            out: *mut oxg_world) -> c_int;
    pub fn oxg_world_add_so2_interval(w: oxg_world, so2_component: u32, lo: f64, hi: f64) -> c_int;
    pub fn oxg_world_destroy(w: oxg_world) -> c_int;
    pub fn oxg_is_valid(w: oxg_world, space: *const oxg_space_desc, states: *const f64, q: u64,
                        out: *mut u8) -> c_int;
    pub fn oxg_check_motion(w: oxg_world, space: *const oxg_space_desc, from: *const f64,
                            to: *const f64, e: u64, resolution: f64, valid: *mut u8,
                            first_invalid_step: *mut u32) -> c_int;
    pub fn oxg_check_edges(w: oxg_world, s: oxg_store, from_idx: *const u32, to_idx: *const u32,
                           e: u64, resolution: f64, valid: *mut u8) -> c_int;

    pub fn oxg_prm_build_radius(w: oxg_world, s: oxg_store, connection_radius: f64, resolution: f64,
                                row_begin: u64, row_end: u64, offsets: *mut u64, idx: *mut u32,
                                cap: u64, stats: *mut oxg_prm_stats) -> c_int;
    pub fn oxg_prm_build_knn(w: oxg_world, s: oxg_store, k: u32, resolution: f64, row_begin: u64,
                             row_end: u64, idx: *mut u32, dist: *mut f64, valid: *mut u8,
                             stats: *mut oxg_prm_stats) -> c_int;

    pub fn oxg_rrt_solve(w: oxg_world, tree: oxg_store, params: *const oxg_plan_params,
                         goal_target: *const f64, samples: *const f64, n_samples: u64,
                         parents: *mut i64, out: *mut oxg_plan_result) -> c_int;
    pub fn oxg_rrt_connect_solve(w: oxg_world, start_tree: oxg_store, goal_tree: oxg_store,
                                 params: *const oxg_plan_params, goal_target: *const f64,
                                 samples: *const f64, n_samples: u64, start_parents: *mut i64,
                                 goal_parents: *mut i64, out: *mut oxg_plan_result) -> c_int;
    pub fn oxg_rrt_star_solve(w: oxg_world, tree: oxg_store, params: *const oxg_plan_params,
                              goal_target: *const f64, samples: *const f64, n_samples: u64,
                              parents: *mut i64, costs: *mut f64, out: *mut oxg_plan_result) -> c_int;

    // housekeeping, device-pointer pipeline of the multi-GPU PRM shard, measurement helpers
    pub fn oxg_version() -> c_int;
    pub fn oxg_store_reserve(s: oxg_store, capacity: u64) -> c_int;
    pub fn oxg_store_sync(s: oxg_store) -> c_int;
    pub fn oxg_store_stream(s: oxg_store, out_stream: *mut *mut c_void) -> c_int;
    pub fn oxg_store_last_scan_ms(s: oxg_store, out_ms: *mut f64) -> c_int;
    pub fn oxg_world_sync(w: oxg_world) -> c_int;
    pub fn oxg_knn_rows_dev(s: oxg_store, row_begin: u64, row_end: u64, k: u32, d_idx: *mut u32,
                            d_dist: *mut f64) -> c_int;
    pub fn oxg_prm_edges_dev(w: oxg_world, s: oxg_store, d_idx_all: *const u32, k: u32,
                             resolution: f64, row_begin: u64, row_end: u64, d_valid_all: *mut u8,
                             d_counters: *mut u64) -> c_int;
    pub fn oxg_prm_resolve_dev(s: oxg_store, d_idx_all: *const u32, k: u32,
                               d_valid_all: *mut u8) -> c_int;
    // PRM query phase (PRM::solve, prm.rs:227-307)
    pub fn oxg_prm_solve(w: oxg_world, s: oxg_store, lower_offsets: *const u64, lower_idx: *const u32,
                         connection_radius: f64, resolution: f64, start: *const f64,
                         goal_target: *const f64, goal_radius: f64, path: *mut u32, path_cap: u32,
                         out: *mut oxg_prm_query_result) -> c_int;
    pub fn oxg_prm_solve_knn(w: oxg_world, s: oxg_store, idx: *const u32, valid: *const u8, k: u32,
                             connection_radius: f64, resolution: f64, start: *const f64,
                             goal_target: *const f64, goal_radius: f64, path: *mut u32,
                             path_cap: u32, out: *mut oxg_prm_query_result) -> c_int;
    // multi-GPU PRM build (NCCL inside the library; the id travels by the host's rendezvous)
    pub fn oxg_dist_unique_id(out_id_128: *mut c_void) -> c_int;
    pub fn oxg_dist_create(unique_id_128: *const c_void, rank: c_int, world_size: c_int,
                           out: *mut oxg_dist) -> c_int;
    pub fn oxg_dist_destroy(d: oxg_dist) -> c_int;
    pub fn oxg_dist_row_range(n: u64, rank: c_int, world_size: c_int, radius_mode: c_int,
                              row_begin: *mut u64, row_end: *mut u64) -> c_int;
    pub fn oxg_store_append_dist(d: oxg_dist, s: oxg_store, rows_local: *const f64, n_local: u64,
                                 first_index: *mut u64) -> c_int;
    pub fn oxg_prm_build_knn_dist(d: oxg_dist, w: oxg_world, s: oxg_store, k: u32, resolution: f64,
                                  idx: *mut u32, dist: *mut f64, valid: *mut u8,
                                  stats: *mut oxg_prm_stats) -> c_int;
    pub fn oxg_prm_build_knn_dist_dev(d: oxg_dist, w: oxg_world, s: oxg_store, k: u32,
                                      resolution: f64, d_idx_all: *mut u32, d_dist_all: *mut f64,
                                      d_valid_all: *mut u8, d_counters: *mut u64) -> c_int;
    pub fn oxg_prm_build_radius_dist(d: oxg_dist, w: oxg_world, s: oxg_store,
                                     connection_radius: f64, resolution: f64, offsets: *mut u64,
                                     idx: *mut u32, cap: u64, stats: *mut oxg_prm_stats) -> c_int;
    pub fn oxg_fp64_peak(out_ops_per_s: *mut f64, out_ms: *mut f64) -> c_int;
    pub fn oxg_fp32_peak(out_ops_per_s: *mut f64, out_ms: *mut f64) -> c_int;
    pub fn oxg_l2_peak(out_bytes_per_s: *mut f64, out_ms: *mut f64) -> c_int;
    pub fn oxg_kernel_launches(reset: c_int) -> u64;
    pub fn oxg_tuning_set(key: *const std::os::raw::c_char, value: i64) -> c_int;
    pub fn oxg_store_tree_stats(s: oxg_store, out8: *mut f64) -> c_int;

    // include/oxg_so3.h — SO(3) scans, cone validity, check_motion
    pub fn oxg_so3_nn1(s: oxg_store, queries: *const f64, q: u64, idx: *mut u32, dist: *mut f64) -> c_int;
    pub fn oxg_so3_radius(s: oxg_store, queries: *const f64, q: u64, r: f64, offsets: *mut u64,
                          idx: *mut u32, cap: u64, dist: *mut f64) -> c_int;
    pub fn oxg_so3_knn(s: oxg_store, queries: *const f64, q: u64, k: u32, exclude_idx: *const u32,
                       idx: *mut u32, dist: *mut f64) -> c_int;
    pub fn oxg_so3_rrt_solve(tree: oxg_store, cones: oxg_so3_cones, params: *const oxg_plan_params,
                             goal_target: *const f64, samples: *const f64, n_samples: u64,
                             parents: *mut i64, out: *mut oxg_plan_result) -> c_int;
    pub fn oxg_so3_rrt_connect_solve(start_tree: oxg_store, goal_tree: oxg_store, cones: oxg_so3_cones,
                                     params: *const oxg_plan_params, goal_target: *const f64,
                                     samples: *const f64, n_samples: u64, start_parents: *mut i64,
                                     goal_parents: *mut i64, out: *mut oxg_plan_result) -> c_int;
    pub fn oxg_so3_cones_create(cones: *const f64, n_cones: u64, out: *mut oxg_so3_cones) -> c_int;
    pub fn oxg_so3_cones_destroy(cones: oxg_so3_cones) -> c_int;
    pub fn oxg_so3_is_valid(cones: oxg_so3_cones, states: *const f64, q: u64, out: *mut u8) -> c_int;
    pub fn oxg_so3_check_motion(cones: oxg_so3_cones, from: *const f64, to: *const f64, e: u64,
                                resolution: f64, valid: *mut u8, first_invalid: *mut u32) -> c_int;
}

#[allow(non_camel_case_types)]
pub type oxg_so3_cones = *mut std::os::raw::c_void;
#[allow(non_camel_case_types)]
pub type oxg_dist = *mut std::os::raw::c_void;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
#[allow(non_camel_case_types)]
pub struct oxg_prm_query_result {
    pub status: i32,
    pub path_len: u32,
    pub goal_node: u32,
    pub n_start_connections: u32,
    pub n_goal_nodes: u32,
    pub bfs_levels: u32,
    pub n_visited: u32,
    pub reserved: u32,
    pub n_adjacency: u64,
    pub device_ms: f64,
}
