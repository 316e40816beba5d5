/*
 * oxg.h — C ABI of liboxg: the B200 (sm_100a) nearest-neighbour / validity backend
 * for oxmpl's planners.
 *
 * This header is the single seam between oxmpl (Rust, via an `oxmpl-cuda` crate's
 * `extern "C"` block — see INTEGRATION.md) and the CUDA kernels.  Every entry point
 * cites the reference code it replaces (paths relative to the oxmpl v0.4.1 tree).
 *
 * Conventions
 *  - All functions return 0 on success or a negative OXG_E_* code; the message for
 *    the calling thread's last failure is oxg_last_error().  Nothing unwinds across
 *    this boundary.
 *  - A *state* is `dim = n_real + n_so2` f64 values: the RealVector block first,
 *    then one angle (radians) per SO2 component.  Host state arrays are row-major
 *    (`rows[i*dim + d]`), exactly the `Vec<f64>` of RealVectorState / the `value`
 *    of SO2State.  On the device the store is structure-of-arrays.
 *  - Indices are append order, identical to `tree.len()` / `roadmap.len()` at push
 *    time in the reference.
 *  - Arithmetic is IEEE-754 binary64, round-to-nearest, never fused, in the
 *    reference's operation order; indices and booleans are bit-exact with the
 *    reference's CPU loops, distances are bit-identical.
 *  - Host-pointer entry points are synchronous: outputs are valid on return.
 *    `_dev` entry points take device pointers, enqueue on the handle's stream and
 *    return without synchronising (use oxg_store_sync / oxg_world_sync).
 *  - A handle may be used from one host thread at a time.
 *  - There is no CPU fallback: without a CUDA device every compute entry point
 *    fails with OXG_E_CUDA.
 */
#ifndef OXG_H
#define OXG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OXG_VERSION 100 /* 0.1.0 */

/* ---- error codes ------------------------------------------------------- */
#define OXG_OK 0
#define OXG_E_NULL (-1)       /* required pointer was NULL                      */
#define OXG_E_DIM (-2)        /* bad dimension / space descriptor               */
#define OXG_E_CAPACITY (-3)   /* caller-provided output capacity too small      */
#define OXG_E_CUDA (-4)       /* CUDA runtime error (incl. no device)           */
#define OXG_E_RESOLUTION (-5) /* motion resolution <= 0 or NaN                  */
#define OXG_E_ARG (-6)        /* other invalid argument                         */
#define OXG_E_EMPTY (-7)      /* query against an empty store                   */
#define OXG_E_ALLOC (-8)      /* host/device allocation failed                  */
#define OXG_E_INTERNAL (-9)   /* unexpected C++ exception caught at the boundary */
#define OXG_E_NCCL (-10)      /* NCCL unavailable in this process, or a NCCL call failed */

#define OXG_INVALID_INDEX 0xFFFFFFFFu

/* ---- space descriptor ---------------------------------------------------
 * Replaces the `SP: StateSpace` type parameter the planners are generic over
 * (oxmpl/src/base/space.rs:78-145).  distance():
 *   n_so2 == 0             -> RealVectorStateSpace::distance
 *                             (oxmpl/src/base/spaces/real_vector_state_space.rs:137-155)
 *   n_real == 0, n_so2 == 1 -> SO2StateSpace::distance
 *                             (oxmpl/src/base/spaces/so2_state_space.rs:97-101)
 *   otherwise (compound; not in the reference — SURVEY F2): left-to-right
 *        acc = w_real * realvector_distance ; acc = acc + w_so2 * so2_distance_c  (c ascending)
 *     with every product and sum separately rounded.
 * interpolate(): RealVector block per real_vector_state_space.rs:161-186, each SO2
 * component per so2_state_space.rs:107-122.
 */
typedef struct oxg_space_desc {
  uint32_t n_real;
  uint32_t n_so2;
  double w_real; /* ignored unless compound */
  double w_so2;  /* ignored unless compound */
} oxg_space_desc;

typedef struct oxg_store_s* oxg_store; /* GPU-resident SoA f64 state store */
typedef struct oxg_world_s* oxg_world; /* declarative obstacle set          */

/* ---- library / device ---------------------------------------------------- */
int oxg_version(void);
const char* oxg_last_error(void);
int oxg_device_count(int* out_count);
int oxg_set_device(int device);

/* ---- state store ---------------------------------------------------------
 * Replaces the planners' `Vec<Node<S>>` state storage: rrt.rs:24-27,61,
 * rrt_star.rs:24-28, prm.rs:24-29.  append == `tree.push` (rrt.rs:213-217,
 * prm.rs:140-141); clear == `Planner::setup` (rrt.rs:147, prm.rs:224).
 * Parents / costs / adjacency stay with the caller.
 */
int oxg_store_create(const oxg_space_desc* space, uint64_t capacity, oxg_store* out);
int oxg_store_destroy(oxg_store s);
int oxg_store_clear(oxg_store s);
int oxg_store_reserve(oxg_store s, uint64_t capacity);
/* Appends n row-major states; *first_index receives the index of the first one. */
int oxg_store_append(oxg_store s, const double* rows, uint64_t n, uint64_t* first_index);
int oxg_store_size(oxg_store s, uint64_t* out_n);
int oxg_store_read(oxg_store s, uint64_t first, uint64_t n, double* rows_out);
int oxg_store_sync(oxg_store s);

/* ---- nearest neighbour ----------------------------------------------------
 * oxg_nn1: the 1-NN linear scan of rrt.rs:187-196, rrt_connect.rs:128-136,
 * rrt_star.rs:199-208 for Q queries at once.  distance(stored, query); values
 * compared after sqrt with strict `<`, so the lowest index wins ties.  `dist[q]`
 * is the reference's `min_dist` (bit-identical; it feeds the steering division).
 */
int oxg_nn1(oxg_store s, const double* queries, uint64_t Q, uint32_t* idx, double* dist);

/* oxg_knn: k nearest by (distance(query, stored), index) ascending.  Not in the
 * reference — semantics pinned by oracle/ (k smallest by the reference's
 * distance value, ties to the lowest index).  exclude_idx (nullable, one per query)
 * removes that stored index from the candidates (self-exclusion by index).  Rows are
 * padded with OXG_INVALID_INDEX / +inf when fewer than k candidates exist.
 */
int oxg_knn(oxg_store s, const double* queries, uint64_t Q, uint32_t k,
            const uint32_t* exclude_idx, uint32_t* idx, double* dist);

/* oxg_radius: the radius scans of rrt_star.rs:121-131 (find_neighbours),
 * prm.rs:131-138 and prm.rs:249-256: all i with distance(query, stored_i) < r
 * (strict), ascending i.  idx_limit (nullable, one per query) restricts query q
 * to stored indices < idx_limit[q] (PRM's "node sees only earlier nodes").
 * CSR output: offsets[Q+1]; idx (and dist, nullable) hold offsets[Q] entries.
 * If offsets[Q] > cap the call fails with OXG_E_CAPACITY after filling offsets
 * (so the caller can size the buffers and call again).
 */
int oxg_radius(oxg_store s, const double* queries, uint64_t Q, double r,
               const uint32_t* idx_limit, uint64_t* offsets, uint32_t* idx,
               uint64_t cap, double* dist);

/* ---- declarative validity checker -----------------------------------------
 * Implements `StateValidityChecker::is_valid` (oxmpl/src/base/validity.rs:39-48)
 * for obstacle sets over the first `dim` (RealVector) coordinates of a state:
 *   sphere  (dim+1 f64: centre, r):   valid iff sqrt(sum (x_i-c_i)^2) > r
 *                                      (README.md:147-150; boundary is invalid)
 *   aabb    (2*dim f64: lo, hi):       invalid iff for all i lo_i <= x_i <= hi_i
 *                                      (oxmpl/tests/rrt_rvss_tests.rs:24-36; inclusive)
 *   capsule (2*dim+1 f64: a, b, r):    valid iff dist(x, segment ab) > r, operation
 *                                      order fixed by oracle/ (no reference form)
 *   so2 interval (comp, lo, hi):       invalid iff lo <= normalise(theta_comp) <= hi
 *                                      (oxmpl/tests/rrt_so2ss_tests.rs:17-27)
 * A state is valid iff no obstacle marks it invalid.
 */
int oxg_world_create(uint32_t dim, const double* spheres, uint64_t n_spheres,
                     const double* aabbs, uint64_t n_aabbs, const double* capsules,
                     uint64_t n_capsules, oxg_world* out);
int oxg_world_add_so2_interval(oxg_world w, uint32_t so2_component, double lo, double hi);
int oxg_world_destroy(oxg_world w);
int oxg_world_sync(oxg_world w);

/* is_valid for Q row-major states (validity.rs:47). out[q] = 1 valid, 0 invalid. */
int oxg_is_valid(oxg_world w, const oxg_space_desc* space, const double* states,
                 uint64_t Q, uint8_t* out);

/* ---- motion validity --------------------------------------------------------
 * The planners' private `check_motion(from, to)` — rrt.rs:90-116,
 * rrt_connect.rs:166-189, rrt_star.rs:76-104, prm.rs:161-187 — for E edges.
 * `resolution` is the caller-computed `space.get_longest_valid_segment_length() * 0.1`
 * (rrt.rs:97).  first_invalid_step (nullable): 0 if the motion is valid, else the
 * 1-based step at which the sequential reference loop returns false (1 when
 * num_steps <= 1 and `to` is invalid).
 */
int oxg_check_motion(oxg_world w, const oxg_space_desc* space, const double* from,
                     const double* to, uint64_t E, double resolution, uint8_t* valid,
                     uint32_t* first_invalid_step);

/* Same, endpoints given as store indices (PRM: check_motion(q_rand, roadmap[i]),
 * prm.rs:134; RRT* rewire rrt_star.rs:244,275). */
int oxg_check_edges(oxg_world w, oxg_store s, const uint32_t* from_idx,
                    const uint32_t* to_idx, uint64_t E, double resolution, uint8_t* valid);

/* ---- batched PRM roadmap construction -----------------------------------------
 * Replaces PRM::construct_roadmap's body (prm.rs:117-146) for a store pre-filled
 * with the valid samples, in sample order.
 *
 * oxg_prm_build_radius — reference-faithful: node i connects to every j < i with
 *   distance(q_i, q_j) < connection_radius && check_motion(q_i, q_j) (prm.rs:131-138).
 *   Output is the lower-index adjacency in CSR form (ascending j per i), i.e. exactly
 *   `new_node.edges` at push time; symmetrise to obtain `roadmap[i].edges`
 *   (prm.rs:143-145).  Rows [row_begin,row_end) only (shard unit for multi-GPU);
 *   offsets has (row_end-row_begin)+1 entries.
 *
 * oxg_prm_build_knn — count-bounded k-NN variant: for each
 *   node i in [row_begin,row_end) its k nearest other nodes (oxg_knn semantics,
 *   self excluded by index); every candidate edge {i,j} is motion-checked once as
 *   check_motion(from = node max(i,j), to = node min(i,j)).  idx/dist/valid are
 *   [rows][k]; valid is 1/0, and 0 for padded slots.
 */
typedef struct oxg_prm_stats {
  uint64_t n_rows;          /* rows processed                                   */
  uint64_t n_candidates;    /* directed candidate slots (non-padded)            */
  uint64_t n_edges_checked; /* distinct motion checks executed                  */
  uint64_t n_edges_valid;   /* of those, valid                                  */
  double knn_ms;            /* device time of the neighbour phase               */
  double edge_ms;           /* device time of the motion-check phase            */
} oxg_prm_stats;

int oxg_prm_build_radius(oxg_world w, oxg_store s, double connection_radius,
                         double resolution, uint64_t row_begin, uint64_t row_end,
                         uint64_t* offsets, uint32_t* idx, uint64_t cap,
                         oxg_prm_stats* stats);

int oxg_prm_build_knn(oxg_world w, oxg_store s, uint32_t k, double resolution,
                      uint64_t row_begin, uint64_t row_end, uint32_t* idx, double* dist,
                      uint8_t* valid, oxg_prm_stats* stats);

/* ---- PRM query phase -----------------------------------------------------------------
 * PRM::solve (prm.rs:227-307) on a roadmap the library built, all on the device: is_valid(start)
 * (:244), start connection = radius scan + check_motion(start, roadmap[i]) (:249-256), goal scan
 * over all nodes (:259-264; the goal is the ball distance(state, goal_target) <= goal_radius of
 * the reference's tests), and the breadth-first search over roadmap[i].edges (:271-301) in the
 * reference's queue order — the same parents, hence the same path, not merely a shortest one.
 * The adjacency is symmetrised on the device (prm.rs:143-145), rows ascending, which is the
 * order the reference's edge lists grow in.
 *  oxg_prm_solve     : roadmap = the lower-index CSR oxg_prm_build_radius(_dist) returns for all
 *                      N rows (offsets[N+1], idx).
 *  oxg_prm_solve_knn : roadmap = idx / valid [N][k] of oxg_prm_build_knn(_dist); an edge exists
 *                      where valid = 1 (no reference form: the reference has no k-NN PRM).
 * path receives the roadmap indices from the start connection to the goal node (the reference's
 * Path is the start state followed by these nodes' states, prm.rs:189-208).
 * out->status: 0 = Ok(path), 1 = NoSolutionFound, 2 = InvalidStartState, 3 = UnsampledStateSpace.
 */
typedef struct oxg_prm_query_result {
  int32_t status;
  uint32_t path_len;            /* roadmap nodes on the path                        */
  uint32_t goal_node;           /* OXG_INVALID_INDEX when no path                   */
  uint32_t n_start_connections; /* start_connections.len()                          */
  uint32_t n_goal_nodes;        /* goal_indices.len()                               */
  uint32_t bfs_levels;
  uint32_t n_visited;           /* nodes that entered the queue                     */
  uint32_t reserved;
  uint64_t n_adjacency;         /* directed adjacency entries after symmetrising    */
  double device_ms;
} oxg_prm_query_result;

int oxg_prm_solve(oxg_world w, oxg_store s, const uint64_t* lower_offsets, const uint32_t* lower_idx,
                  double connection_radius, double resolution, const double* start,
                  const double* goal_target, double goal_radius, uint32_t* path, uint32_t path_cap,
                  oxg_prm_query_result* out);
int oxg_prm_solve_knn(oxg_world w, oxg_store s, const uint32_t* idx, const uint8_t* valid, uint32_t k,
                      double connection_radius, double resolution, const double* start,
                      const double* goal_target, double goal_radius, uint32_t* path,
                      uint32_t path_cap, oxg_prm_query_result* out);

/* ---- fused sequential planners ---------------------------------------------------------
 * The bodies of RRT::solve (rrt.rs:158-227), RRTConnect::solve (rrt_connect.rs:227-309 with
 * extend :121-159) and RRTStar::solve (rrt_star.rs:171-290) as ONE persistent kernel each:
 * 1-NN scan, steering (rrt.rs:200-208), check_motion, append, goal test and — for RRT* —
 * find_neighbours, choose-parent and rewire run on the device for a whole sample stream, so a
 * planner pays one launch per solve instead of several launches and a D2H sync per iteration.
 *
 *  - `tree` must hold the start state at index 0 (what Planner::setup pushes, rrt.rs:147-155);
 *    RRT-Connect's goal tree must hold the goal state (rrt_connect.rs:201-216).  On return the
 *    stores hold the grown trees (oxg_store_read); parents[i] is Node::parent_index (-1 = None),
 *    costs[i] is Node::cost (RRT*).  Calling again on the same stores continues the same trees.
 *  - `samples` is the pre-drawn stream of q_rand (row-major, n_samples x dim): the reference
 *    draws `random_bool(goal_bias) ? goal.sample_goal : space.sample_uniform` once per iteration
 *    independently of the tree (rrt.rs:177-184), so the caller draws the same sequence up front.
 *  - The goal is the ball `distance(state, goal_target) <= goal_radius`
 *    (CircularGoalRegion::is_satisfied, oxmpl/tests/rrt_rvss_tests.rs:46-50); goal_target NULL or
 *    goal_radius < 0 disables the goal test.
 *  - The reference loops until a wall-clock timeout; here a solve ends when it is solved, the
 *    stream is consumed, a tree holds max_nodes states, or timeout_s (if > 0) elapses.
 * status: 0 solved, 1 stream exhausted, 2 tree full, 3 timeout (PlanningError::Timeout).
 *
 * Limit (RRT* only): one iteration may find at most OXG_RRT_STAR_MAX_NEIGHBOURS tree nodes inside
 * search_radius (find_neighbours, rrt_star.rs:121-131, is unbounded; the kernel keeps the
 * neighbour list of an iteration in shared memory).  A denser neighbourhood ends the solve with
 * OXG_E_CAPACITY; the tree, parents, costs and *out describe the iterations completed before it,
 * and the solve can be continued with a smaller search_radius.  The host mirrors report it as
 * PlanningError::NoSolutionFound with the library's message.
 */
#define OXG_RRT_STAR_MAX_NEIGHBOURS 2048
typedef struct oxg_plan_params {
  double max_distance;  /* RRT::new(max_distance, ..) rrt.rs:75, rrt_connect.rs:86, rrt_star.rs:65 */
  double search_radius; /* RRTStar::new(.., .., search_radius); ignored by RRT / RRT-Connect        */
  double resolution;    /* space.get_longest_valid_segment_length() * 0.1 (rrt.rs:97)                */
  double goal_radius;
  double timeout_s;
  uint64_t max_nodes;   /* capacity of each tree                                                     */
  int32_t stop_at_goal; /* RRT*: 0 = keep optimising after the goal is first reached                 */
  int32_t reserved;
} oxg_plan_params;

typedef struct oxg_plan_result {
  uint64_t n_nodes[2];             /* tree sizes ([1]: RRT-Connect's goal tree)                      */
  uint64_t iters;                  /* samples consumed                                               */
  uint64_t motion_checks;          /* check_motion calls the sequential reference makes              */
  uint64_t motion_checks_executed; /* evaluated here (RRT* checks a superset in parallel)            */
  int64_t goal_node[2];            /* RRT / RRT*: node that satisfied the goal; RRT-Connect: the
                                      joined nodes (start tree, goal tree), [1] = -1 when the start
                                      tree reached the goal itself (rrt_connect.rs:271-274)          */
  int32_t solved;
  int32_t status;
  double device_ms;
  int32_t ctas;                    /* CTAs of the persistent grid                                    */
  int32_t reserved;
  double phase_ms[8];              /* where the kernel's time went (CTA 0's clock): [0] 1-NN scans,
                                      [1] barrier after the scan, [2] steer + motion check (+ RRT*
                                      radius scan), [3] barrier, [4] RRT* choose-parent batch,
                                      [5] RRT* barriers, [6] RRT* replay + append + rewire batch,
                                      [7] everything else (loop control, append, goal test)         */
} oxg_plan_result;

int oxg_rrt_solve(oxg_world w, oxg_store tree, const oxg_plan_params* params,
                  const double* goal_target, const double* samples, uint64_t n_samples,
                  int64_t* parents /* nullable, max_nodes entries */, oxg_plan_result* out);
int oxg_rrt_connect_solve(oxg_world w, oxg_store start_tree, oxg_store goal_tree,
                          const oxg_plan_params* params, const double* goal_target,
                          const double* samples, uint64_t n_samples, int64_t* start_parents,
                          int64_t* goal_parents, oxg_plan_result* out);
int oxg_rrt_star_solve(oxg_world w, oxg_store tree, const oxg_plan_params* params,
                       const double* goal_target, const double* samples, uint64_t n_samples,
                       int64_t* parents, double* costs, oxg_plan_result* out);

/* ---- device-pointer pipeline (resident data; multi-GPU sharding) ----------------
 * The same phases with device buffers, enqueued on the store's stream.
 *  oxg_knn_rows_dev   : k-NN of stored rows [row_begin,row_end) against the whole
 *                       store, self excluded; d_idx/d_dist are [rows][k] device.
 *  oxg_prm_edges_dev  : given the complete d_idx_all [N][k] (after the all-gather),
 *                       motion-checks the candidate edges owned by rows
 *                       [row_begin,row_end) and writes d_valid_all[row][slot] for
 *                       those rows: 1/0, or 2 = "owned by the mirror slot".
 *  oxg_prm_resolve_dev: after d_valid_all is complete on this device, replaces every
 *                       2 by the mirror slot's value.
 */
int oxg_knn_rows_dev(oxg_store s, uint64_t row_begin, uint64_t row_end, uint32_t k,
                     uint32_t* d_idx, double* d_dist);
int oxg_prm_edges_dev(oxg_world w, oxg_store s, const uint32_t* d_idx_all, uint32_t k,
                      double resolution, uint64_t row_begin, uint64_t row_end,
                      uint8_t* d_valid_all, uint64_t* d_counters /* 2 x u64: checked, valid */);
int oxg_prm_resolve_dev(oxg_store s, const uint32_t* d_idx_all, uint32_t k,
                        uint8_t* d_valid_all);
/* The CUDA stream of a store (cudaStream_t as void*), so callers can order their
 * own work (NCCL collectives, torch ops) after the library's. */
int oxg_store_stream(oxg_store s, void** out_stream);

/* ---- multi-GPU PRM build (one process per GPU, NCCL over NVLink) -----------------
 * The sharded form of PRM::construct_roadmap's body (prm.rs:117-146): every rank holds
 * the same store and world (each appended the same samples), owns a slice of the query
 * rows and of the candidate edges, and the per-node rows are exchanged with NCCL
 * all-gathers inside the library, on the store's stream.  Every rank returns the
 * complete roadmap, identical to what the single-GPU entry point returns.
 *
 *  oxg_dist_unique_id : ncclGetUniqueId (128 bytes).  Rank 0 calls it and hands the bytes
 *                       to the other ranks by whatever rendezvous the host has (MPI,
 *                       torch.distributed, a file).
 *  oxg_dist_create    : ncclCommInitRank on the calling thread's current CUDA device;
 *                       collective over all ranks.  world_size 1 needs no id and no NCCL.
 *  oxg_dist_row_range : the query rows rank owns.  k-NN mode: [rank n / world,
 *                       (rank+1) n / world) (uniform work per row).  Radius mode:
 *                       [n sqrt(rank/world), n sqrt((rank+1)/world)) — node i scans the i
 *                       earlier nodes (prm.rs:131), so equal areas of the triangle balance it.
 *  oxg_store_append_dist : collective append: every rank passes ITS slice of a common row set
 *                       (any sizes); afterwards every rank's store holds all the rows in rank
 *                       order.  Each slice crosses PCIe once and reaches the other ranks over
 *                       NVLink (one broadcast per rank in a group call).
 *  oxg_prm_build_knn_dist      : oxg_prm_build_knn over all rows; idx/dist/valid are
 *                       [N][k] host arrays (each nullable: a rank that does not want the
 *                       roadmap on the host passes NULL).  Exchanges: one grouped
 *                       all-gather of the idx and dist rows, one of the validity rows.
 *                       stats->knn_ms is the device time of the whole sharded build on
 *                       this rank; n_edges_checked / n_edges_valid count this rank's shard.
 *  oxg_prm_build_knn_dist_dev  : the same with device outputs ([N][k]; d_counters: 2 x u64
 *                       = motion checks executed / valid on this rank), enqueued on the
 *                       store's stream without synchronising.
 *  oxg_prm_build_radius_dist   : oxg_prm_build_radius over all rows (reference-faithful
 *                       mode); the ragged lower-index adjacency lists are exchanged with one
 *                       broadcast per rank inside a group call.  offsets has N+1 entries;
 *                       fails with OXG_E_CAPACITY (offsets filled) on every rank alike.
 * NCCL is looked up at run time (the libnccl.so.2 already loaded in the process, else the
 * system one); without it, world sizes above 1 fail with OXG_E_NCCL.
 */
typedef struct oxg_dist_s* oxg_dist;
int oxg_dist_unique_id(void* out_id_128);
int oxg_dist_create(const void* unique_id_128, int rank, int world_size, oxg_dist* out);
int oxg_dist_destroy(oxg_dist d);
int oxg_dist_row_range(uint64_t n, int rank, int world_size, int radius_mode,
                       uint64_t* row_begin, uint64_t* row_end);
int oxg_store_append_dist(oxg_dist d, oxg_store s, const double* rows_local, uint64_t n_local,
                          uint64_t* first_index);
int oxg_prm_build_knn_dist(oxg_dist d, oxg_world w, oxg_store s, uint32_t k, double resolution,
                           uint32_t* idx, double* dist, uint8_t* valid, oxg_prm_stats* stats);
int oxg_prm_build_knn_dist_dev(oxg_dist d, oxg_world w, oxg_store s, uint32_t k,
                               double resolution, uint32_t* d_idx_all, double* d_dist_all,
                               uint8_t* d_valid_all, uint64_t* d_counters);
int oxg_prm_build_radius_dist(oxg_dist d, oxg_world w, oxg_store s, double connection_radius,
                              double resolution, uint64_t* offsets, uint32_t* idx, uint64_t cap,
                              oxg_prm_stats* stats);

/* ---- measurement helpers ----------------------------------------------------------
 * oxg_fp64_peak: measures this GPU's unfused FP64 issue rate (independent DADD/DMUL
 * chains, no FMA) — the roofline denominator for the batched k-NN / validity kernels
 * (SURVEY §8d).  Returns operations per second.
 * oxg_kernel_launches: number of liboxg kernel launches since the counter was reset.
 */
/* Device time (CUDA events on the store's stream) of the kernels of the most recent
 * oxg_nn1 / oxg_knn call on this store, excluding the host<->device copies. */
int oxg_store_last_scan_ms(oxg_store s, double* out_ms);
int oxg_fp64_peak(double* out_ops_per_s, double* out_ms);
/* oxg_fp32_peak: FP32 pipe rate with the certified pre-filter's instruction mix (packed
 * subtract + packed FMA, FADD2/FFMA2), in lane-instructions per second (an FMA counts once;
 * the mix's flop rate is 1.5x this figure). */
int oxg_fp32_peak(double* out_ops_per_s, double* out_ms);
/* oxg_l2_peak: bytes per second all SMs read from an L2-resident buffer with coalesced 16-byte
 * loads that bypass L1 — the memory roofline of kernels whose working set (store + hierarchy) is
 * L2-resident by design. */
int oxg_l2_peak(double* out_bytes_per_s, double* out_ms);
uint64_t oxg_kernel_launches(int reset);
/* A/B switches for measurements; every setting runs hand-written sm_100a kernels and gives
 * bit-identical results.  "knn_tree": 1 = always answer oxg_knn / oxg_prm_build_knn through the
 * store's bounding-box hierarchy when the space allows it, 0 = always scan every pair,
 * -1 = choose by batch and store size (default).  "tree_cache": 0 = rebuild the hierarchy for
 * every batch instead of reusing it while the store is unchanged (benchmarks that must pay the
 * build inside the timed step), 1 = reuse (default).  "tree_kind": 0 = order the states by a space-filling key (cheapest build),
 * 1 = balanced partition order (a few ms per 1e6 states, about half the leaf visits per query),
 * -1 = by batch size (default).  "tree_stats": 1 = time the hierarchy build and
 * the query kernel for oxg_store_tree_stats, 2 = also count the visits (the counters cost the
 * query kernel ~15 %), 0 = neither (default).  "row_cull": 0 = the k-NN PRM edge phase tests every
 * obstacle against every edge's own ball only, 1 = it first rules obstacles out once per row
 * (default). */
int oxg_tuning_set(const char* key, int64_t value);
/* Counters of the most recent k-NN answered through the hierarchy with "tree_stats" on:
 * out8 = { build ms, query-kernel ms, leaves visited, 32-child box groups tested, candidates
 * admitted, list merges, leaves whose states reached the exact f64 evaluation (the FP32 leaf
 * filter discarded the others whole), queries }. */
int oxg_store_tree_stats(oxg_store s, double* out8);

#ifdef __cplusplus
}
#endif
#endif /* OXG_H */
