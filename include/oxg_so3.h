/* oxg_so3.h — SO(3) scans and validity for liboxg.
 *
 * The reference's SO3 metric (oxmpl/src/base/spaces/so3_state_space.rs:102-112) is
 *
 *     a = |x1*x2 + y1*y2 + z1*z2 + w1*w2|        (left-to-right, plain IEEE f64)
 *     distance = a > 1 - 1e-9 ? 0 : acos(a)      (acos = the host's libm)
 *
 * `a` is reproducible bit for bit on the GPU; libm's `acos` is not (no device routine matches
 * glibc to the last bit).  So the device never evaluates acos.  It works on `a`, on which
 * the distance is a non-increasing function, and the library's host side turns every
 * comparison the reference makes on distances into an exact comparison on `a` by searching the
 * f64 bit patterns with the SAME libm `acos` the reference calls (Rust's f64::acos is the
 * platform libm's):
 *
 *     "distance < r"        <=>  a >= min{x : f(x) <  r}
 *     "distance > angle"    <=>  a <  min{x : f(x) <= angle}
 *     arg-min of distance    =   lowest index with a >= min{x : f(x) <= f(max a)}
 *
 * (several adjacent `a` can share one acos value, so the arg-max of `a` alone is not the
 * reference's strict-`<` lowest-index winner).  Returned distances are f(a) evaluated on the host,
 * i.e. the reference's own numbers.  The only assumption is that libm's acos is monotone, which
 * the parity tests exercise against the brute-force oracle.
 *
 * A store holding SO3 states is created with oxg_space_desc{n_real = 4} — coordinates (x, y, z, w).
 * Error codes, threading and ownership rules are those of oxg.h.
 */
#ifndef OXG_SO3_H
#define OXG_SO3_H

#include "oxg.h"

#ifdef __cplusplus
extern "C" {
#endif

/* 1-NN scan of rrt.rs:187-196 with SO3StateSpace::distance(stored, query): idx[q] is the lowest
 * index attaining the minimum distance, dist[q] that distance. */
int oxg_so3_nn1(oxg_store store, const double* queries /* row-major Q x 4 */, uint64_t q,
                uint32_t* idx, double* dist);

/* Radius scan of rrt_star.rs:121-131 / prm.rs:131-138 with distance(query, stored) < r: CSR output,
 * ascending index per query.  offsets has q + 1 entries; idx / dist (dist nullable) hold up to cap
 * entries (OXG_E_CAPACITY if more are needed; offsets is still filled in). */
int oxg_so3_radius(oxg_store store, const double* queries, uint64_t q, double r, uint64_t* offsets,
                   uint32_t* idx, uint64_t cap, double* dist);

/* k nearest stored quaternions by (distance(query, stored), index) ascending -- oxg_knn's contract
 * (oxg.h; not in the reference, SURVEY F1, semantics pinned by oracle/) under the SO3 metric.
 * The device finds the k-th largest a of every query; the host turns it into the low end of its
 * acos plateau; a second device pass gathers the states at or above it (normally exactly k) with
 * their a; distances and the final order come from the host's libm.  k <= 32.  exclude_idx
 * (nullable, one per query) removes that stored index; states whose a is NaN never enter; rows
 * are padded with OXG_INVALID_INDEX / +inf. */
int oxg_so3_knn(oxg_store store, const double* queries /* row-major Q x 4 */, uint64_t q, uint32_t k,
                const uint32_t* exclude_idx, uint32_t* idx /* q x k */, double* dist /* q x k */);

/* Declarative form of the reference's SO3 validity fixture (ForbiddenConeChecker,
 * oxmpl/tests/rrt_so3ss_tests.rs:46-56): a state is valid iff distance(centre_c, state) > angle_c
 * for every cone c.  cones is row-major n x 5: (x, y, z, w, angle). */
typedef struct oxg_so3_cones_s* oxg_so3_cones;
int oxg_so3_cones_create(const double* cones, uint64_t n_cones, oxg_so3_cones* out);
int oxg_so3_cones_destroy(oxg_so3_cones cones);
int oxg_so3_is_valid(oxg_so3_cones cones, const double* states /* row-major Q x 4 */, uint64_t q,
                     uint8_t* out);

/* check_motion (rrt.rs:90-116) over SO3: the interpolated states (SLERP / LERP of
 * so3_state_space.rs:114-159, sin/acos from libm) are produced on the host so that they are the
 * reference's own bits, then every state of every edge is tested against every cone in one launch.
 * valid[e] = 1/0; first_invalid (nullable) = 0 when valid, else the 1-based failing step. */
int oxg_so3_check_motion(oxg_so3_cones cones, const double* from, const double* to, uint64_t e,
                         double resolution, uint8_t* valid, uint32_t* first_invalid);

/* RRT::solve (rrt.rs:158-227) over SO3 for a whole pre-drawn sample stream -- the SO3 counterpart
 * of oxg_rrt_solve (oxg.h; same params / result / status codes; search_radius and stop_at_goal are
 * ignored).  `tree` holds the start state at index 0 (or a tree grown by an earlier call, with its
 * parents[] passed back in); the goal is the ball distance(state, goal_target) <= goal_radius
 * (SO3GoalRegion, oxmpl/tests/rrt_so3ss_tests.rs:58-68; NULL target or radius < 0: no goal test).
 * SLERP needs libm, so the loop's steering runs on the host, and the scans and cone tests run on
 * the device for a CHUNK of samples at a time against the tree as it stood when the chunk began;
 * samples are committed in order and a chunk is cut at the first sample whose nearest node would
 * have been a node added inside the chunk (checked exactly), so the tree, parents and iteration
 * count are those of the one-sample-at-a-time loop.  out->ctas reports the number of chunks,
 * out->motion_checks_executed the motion checks launched (speculative ones included),
 * out->device_ms the wall time of the call.  parents: in/out, max_nodes entries. */
int oxg_so3_rrt_solve(oxg_store tree, oxg_so3_cones cones, const oxg_plan_params* params,
                      const double* goal_target /* 4 */, const double* samples /* n x 4 */,
                      uint64_t n_samples, int64_t* parents, oxg_plan_result* out);

/* RRTConnect::solve (rrt_connect.rs:227-309, extend :121-159) over SO3 -- the counterpart of
 * oxg_rrt_connect_solve, run in speculative chunks like oxg_so3_rrt_solve.  Both growth cases
 * (start tree or goal tree extends towards q_rand; the other tree then extends towards the new
 * node) are prepared for every sample of a chunk in six launches; samples are committed in order
 * with the case the tree sizes dictate, and a chunk is cut at the first sample one of whose scans
 * would have found a node added inside the chunk.  goal_tree holds the goal state at index 0
 * (rrt_connect.rs:201-216).  out->goal_node = the joined nodes (start tree, goal tree), [1] = -1
 * when the start tree reached the goal ball itself (:271-274); n_nodes[0/1] = tree sizes.
 * start_parents / goal_parents: in/out, max_nodes entries each. */
int oxg_so3_rrt_connect_solve(oxg_store start_tree, oxg_store goal_tree, oxg_so3_cones cones,
                              const oxg_plan_params* params, const double* goal_target,
                              const double* samples, uint64_t n_samples, int64_t* start_parents,
                              int64_t* goal_parents, oxg_plan_result* out);

#ifdef __cplusplus
}
#endif
#endif /* OXG_SO3_H */
