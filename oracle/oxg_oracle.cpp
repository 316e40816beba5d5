// oxg_oracle.cpp — CPU restatement of oxmpl v0.4.1's nearest-neighbour / validity hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / `--impl reference` legs may load it, and only as the
// checker or the timed CPU baseline.  liboxg (the product) never links or calls it.
//
// Parity status.  The Rust reference cannot be compiled in this image (no cargo/rustc,
// SURVEY F5), so there is no oracle/_ref build.  Every function below restates the cited
// Rust source operation by operation (IEEE-754 binary64, round-to-nearest, NO fused
// multiply-add, left-to-right sums) and must be compiled with
//     g++ -O2 -ffp-contract=off            (never -ffast-math)
// It is pinned by the few exact known answers the reference's own tests hold for this path
// (tests/test_oracle_pins.py): SO2 normalise(3pi/2) == -pi/2 exactly
// (oxmpl/src/base/states/so2_state.rs:80-87), the SO2State::new doctests (so2_state.rs:28-32),
// the config-1 resolution constants derived from real_vector_state_space.rs:103-118,251-253
// and rrt.rs:96-97, and the fixture semantics of oxmpl/tests/*.rs.  The reference has no
// golden vector for a distance value, an NN index, a check_motion result, a tree or a
// roadmap (its integration tests assert path properties only), so for those rows parity is
// pinned by the source text restated here.  Three things have NO reference form and are
// "parity unpinned" (defined here, and the CUDA path must match this definition):
//   * k-NN (orc_knn)            — the reference has only 1-NN and radius scans
//   * compound R^n x SO2 spaces — the reference has no CompoundStateSpace
//   * the capsule obstacle      — the reference's fixtures have spheres/AABBs/intervals only
//
// All paths in comments are relative to the oxmpl v0.4.1 source tree.

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#include <algorithm>

#if defined(__FAST_MATH__)
#error "the oracle must not be compiled with -ffast-math"
#endif
#if defined(__FP_FAST_FMA) && !defined(ORC_ALLOW_FMA_HW)
// Hardware FMA being available is fine as long as -ffp-contract=off is passed; the
// Makefile does so and tests/test_oracle_pins.py checks a contraction-sensitive vector.
#endif

extern "C" {

// ---------------------------------------------------------------------------------------
// Space descriptor (mirrors include/oxg.h; duplicated on purpose — the oracle is independent)
// ---------------------------------------------------------------------------------------
struct orc_space {
  uint32_t n_real;
  uint32_t n_so2;
  double w_real;
  double w_so2;
};

static const double ORC_PI = 3.14159265358979323846264338327950288;  // std::f64::consts::PI

// Rust f64::rem_euclid(rhs): `let r = self % rhs; if r < 0.0 { r + rhs.abs() } else { r }`
// (`%` on f64 is C fmod, exact).
static inline double rem_euclid(double v, double rhs) {
  double r = std::fmod(v, rhs);
  if (r < 0.0) r = r + std::fabs(rhs);
  return r;
}

// SO2State::normalise — oxmpl/src/base/states/so2_state.rs:55-59 (and ::new, :34-38):
//   (value + PI).rem_euclid(2.0 * PI) - PI
double orc_so2_normalise(double v) { return rem_euclid(v + ORC_PI, 2.0 * ORC_PI) - ORC_PI; }

// RealVectorStateSpace::distance — oxmpl/src/base/spaces/real_vector_state_space.rs:137-155
//   zip -> map(|(v1,v2)| (v1 - v2).powi(2)) -> sum::<f64>() -> sqrt()
// powi(2) is x*x; sum is a left fold from 0.0 (0.0 + x*x == x*x bitwise for x*x >= +0).
static inline double rv_sumsq(const double* a, const double* b, uint32_t n) {
  double s = 0.0;
  for (uint32_t i = 0; i < n; ++i) {
    double d = a[i] - b[i];
    s = s + d * d;
  }
  return s;
}
static inline double rv_distance(const double* a, const double* b, uint32_t n) {
  return std::sqrt(rv_sumsq(a, b, n));
}

// SO2StateSpace::distance — oxmpl/src/base/spaces/so2_state_space.rs:97-101
//   diff = s1 - s2; diff = (diff + PI).rem_euclid(2.0*PI) - PI; diff.abs()
static inline double so2_distance(double a, double b) {
  double diff = a - b;
  diff = rem_euclid(diff + ORC_PI, 2.0 * ORC_PI) - ORC_PI;
  return std::fabs(diff);
}

// distance(a, b) for the described space.  Pure RealVector and pure single-SO2 spaces follow
// the reference exactly; the compound rule (weights, left-to-right) is defined HERE
// (parity unpinned).
double orc_distance(const orc_space* sp, const double* a, const double* b) {
  if (sp->n_so2 == 0) return rv_distance(a, b, sp->n_real);
  if (sp->n_real == 0 && sp->n_so2 == 1) return so2_distance(a[0], b[0]);
  double acc = 0.0;
  bool have = false;
  if (sp->n_real > 0) {
    acc = sp->w_real * rv_distance(a, b, sp->n_real);
    have = true;
  }
  for (uint32_t c = 0; c < sp->n_so2; ++c) {
    double term = sp->w_so2 * so2_distance(a[sp->n_real + c], b[sp->n_real + c]);
    acc = have ? acc + term : term;
    have = true;
  }
  return acc;
}

// RealVectorStateSpace::interpolate — real_vector_state_space.rs:161-186
//   out[i] = from[i] + (to[i] - from[i]) * t
// SO2StateSpace::interpolate — so2_state_space.rs:107-122
//   diff = normalise(to) - normalise(from); if diff > PI {diff -= 2PI} else if diff < -PI {diff += 2PI}
//   out = from + diff * t   (un-normalised `from`);  out = normalise(out)
void orc_interpolate(const orc_space* sp, const double* from, const double* to, double t,
                     double* out) {
  for (uint32_t i = 0; i < sp->n_real; ++i) out[i] = from[i] + (to[i] - from[i]) * t;
  for (uint32_t c = 0; c < sp->n_so2; ++c) {
    uint32_t i = sp->n_real + c;
    double diff = orc_so2_normalise(to[i]) - orc_so2_normalise(from[i]);
    if (diff > ORC_PI) {
      diff -= 2.0 * ORC_PI;
    } else if (diff < -ORC_PI) {
      diff += 2.0 * ORC_PI;
    }
    double v = from[i] + diff * t;
    out[i] = orc_so2_normalise(v);
  }
}

// RealVectorStateSpace::get_maximum_extent — real_vector_state_space.rs:103-118
//   1.0 if any bound is non-finite, else sqrt(sum((high - low).powi(2)))
double orc_rv_max_extent(const double* lo, const double* hi, uint32_t dim) {
  for (uint32_t i = 0; i < dim; ++i)
    if (!std::isfinite(lo[i]) || !std::isfinite(hi[i])) return 1.0;
  double s = 0.0;
  for (uint32_t i = 0; i < dim; ++i) {
    double d = hi[i] - lo[i];
    s = s + d * d;
  }
  return std::sqrt(s);
}
// get_longest_valid_segment_length — real_vector_state_space.rs:251-253 / so2_state_space.rs:168-170
//   extent * longest_valid_segment_fraction  (fraction default 0.05, :98)
double orc_longest_valid_segment_length(double extent, double fraction) { return extent * fraction; }
// SO2StateSpace::get_maximum_extent — so2_state_space.rs:77-79
double orc_so2_max_extent(void) { return ORC_PI; }
// The planners' motion resolution: `space.get_longest_valid_segment_length() * 0.1` (rrt.rs:97)
double orc_motion_resolution(double lvsl) { return lvsl * 0.1; }
// set_longest_valid_segment_fraction clamp — real_vector_state_space.rs:121-129
double orc_clamp_fraction(double f) {
  if (f > 0.0 && f <= 1.0) return f;
  if (f <= 0.0) return 0.0;
  return 1.0;
}

// ---------------------------------------------------------------------------------------
// Declarative world (obstacles over the first `dim` coordinates).  Fixture semantics:
//   sphere : valid iff sqrt(sum (x_i-c_i)^2) > r               README.md:147-150 (strict)
//   aabb   : invalid iff all_i lo_i <= x_i <= hi_i             oxmpl/tests/rrt_rvss_tests.rs:24-36
//   so2 interval: invalid iff lo <= normalise(theta) <= hi      oxmpl/tests/rrt_so2ss_tests.rs:17-27
//   capsule: valid iff dist(x, segment[a,b]) > r — DEFINED HERE (parity unpinned):
//       ab_i = b_i - a_i ; den = ab_0*ab_0 + ab_1*ab_1 + ...   (left to right)
//       ap_i = x_i - a_i ; num = ap_0*ab_0 + ap_1*ab_1 + ...   (left to right)
//       t = den > 0 ? num / den : 0 ; if t < 0 t = 0 ; if t > 1 t = 1
//       e_i = x_i - (a_i + ab_i * t) ; d2 = e_0*e_0 + ... ; valid iff sqrt(d2) > r
// ---------------------------------------------------------------------------------------
struct orc_world {
  uint32_t dim;
  uint64_t n_spheres;
  const double* spheres;  // n x (dim+1): c..., r
  uint64_t n_aabbs;
  const double* aabbs;  // n x 2dim: lo..., hi...
  uint64_t n_capsules;
  const double* capsules;  // n x (2dim+1): a..., b..., r
  uint64_t n_so2;
  const uint32_t* so2_comp;  // component index (0-based among SO2 components)
  const double* so2_lohi;    // n x 2
};

struct orc_counters {
  uint64_t states;      // states evaluated by is_valid
  uint64_t prim_tests;  // obstacle primitives evaluated (sequential, early exit at first hit)
  double flops;         // algorithmic FP64 operations (SURVEY §8d cost model)
};

static inline bool capsule_valid(const double* x, const double* cap, uint32_t dim) {
  const double* a = cap;
  const double* b = cap + dim;
  double r = cap[2 * dim];
  double den = 0.0, num = 0.0;
  for (uint32_t i = 0; i < dim; ++i) {
    double ab = b[i] - a[i];
    double ap = x[i] - a[i];
    double m1 = ab * ab;
    double m2 = ap * ab;
    den = (i == 0) ? m1 : den + m1;
    num = (i == 0) ? m2 : num + m2;
  }
  double t = (den > 0.0) ? num / den : 0.0;
  if (t < 0.0) t = 0.0;
  if (t > 1.0) t = 1.0;
  double d2 = 0.0;
  for (uint32_t i = 0; i < dim; ++i) {
    double ab = b[i] - a[i];
    double c = a[i] + ab * t;
    double e = x[i] - c;
    double m = e * e;
    d2 = (i == 0) ? m : d2 + m;
  }
  return std::sqrt(d2) > r;
}

// StateValidityChecker::is_valid — oxmpl/src/base/validity.rs:39-48, declarative form.
// Evaluation order (only matters for the work counters): spheres, aabbs, capsules, so2.
int orc_is_valid(const orc_world* w, const orc_space* sp, const double* x, orc_counters* cnt) {
  uint32_t dim = w->dim;
  if (cnt) cnt->states += 1;
  for (uint64_t s = 0; s < w->n_spheres; ++s) {
    const double* c = w->spheres + s * (dim + 1);
    if (cnt) {
      cnt->prim_tests += 1;
      cnt->flops += 3.0 * dim;
    }
    if (!(rv_distance(x, c, dim) > c[dim])) return 0;
  }
  for (uint64_t s = 0; s < w->n_aabbs; ++s) {
    const double* lo = w->aabbs + s * 2 * dim;
    const double* hi = lo + dim;
    if (cnt) {
      cnt->prim_tests += 1;
      cnt->flops += 2.0 * dim;
    }
    bool inside = true;
    for (uint32_t i = 0; i < dim; ++i)
      if (!(x[i] >= lo[i] && x[i] <= hi[i])) {
        inside = false;
        break;
      }
    if (inside) return 0;
  }
  for (uint64_t s = 0; s < w->n_capsules; ++s) {
    if (cnt) {
      cnt->prim_tests += 1;
      cnt->flops += 8.0 * dim + 2.0;
    }
    if (!capsule_valid(x, w->capsules + s * (2 * dim + 1), dim)) return 0;
  }
  for (uint64_t s = 0; s < w->n_so2; ++s) {
    double v = orc_so2_normalise(x[sp->n_real + w->so2_comp[s]]);
    if (cnt) {
      cnt->prim_tests += 1;
      cnt->flops += 6.0;
    }
    if (v >= w->so2_lohi[2 * s] && v <= w->so2_lohi[2 * s + 1]) return 0;
  }
  return 1;
}

void orc_is_valid_batch(const orc_world* w, const orc_space* sp, const double* states,
                        uint64_t Q, uint8_t* out) {
  uint32_t dim = sp->n_real + sp->n_so2;
  for (uint64_t q = 0; q < Q; ++q) out[q] = (uint8_t)orc_is_valid(w, sp, states + q * dim, nullptr);
}

// Rust `x.ceil() as usize`: saturating, NaN -> 0.
static inline uint64_t ceil_as_usize(double x) {
  double c = std::ceil(x);
  if (!(c == c)) return 0;
  if (c <= 0.0) return 0;
  if (c >= 18446744073709551615.0) return UINT64_MAX;
  return (uint64_t)c;
}

// check_motion — rrt.rs:90-116 (identical copies rrt_connect.rs:166-189, rrt_star.rs:76-104,
// prm.rs:161-187).  Returns 1 valid / 0 invalid / -1 if num_steps is absurd (> 2^32; the
// reference would loop "forever").  *first_invalid = 0 when valid, else the 1-based step.
int orc_check_motion(const orc_world* w, const orc_space* sp, const double* from,
                     const double* to, double resolution, uint32_t* first_invalid,
                     orc_counters* cnt) {
  uint32_t dim = sp->n_real + sp->n_so2;
  double dist = orc_distance(sp, from, to);
  uint64_t num_steps = ceil_as_usize(dist / resolution);
  if (first_invalid) *first_invalid = 0;
  if (num_steps <= 1) {
    int v = orc_is_valid(w, sp, to, cnt);
    if (!v && first_invalid) *first_invalid = 1;
    return v;
  }
  if (num_steps > (1ull << 32)) return -1;
  double buf[64];
  std::vector<double> big;
  double* st = buf;
  if (dim > 64) {
    big.resize(dim);
    st = big.data();
  }
  for (uint64_t i = 1; i <= num_steps; ++i) {
    double t = (double)i / (double)num_steps;
    orc_interpolate(sp, from, to, t, st);
    if (cnt) cnt->flops += 3.0 * dim + 1.0;
    if (!orc_is_valid(w, sp, st, cnt)) {
      if (first_invalid) *first_invalid = (uint32_t)i;
      return 0;
    }
  }
  return 1;
}

void orc_check_motion_batch(const orc_world* w, const orc_space* sp, const double* from,
                            const double* to, uint64_t E, double resolution, uint8_t* valid,
                            uint32_t* first_invalid, orc_counters* cnt) {
  uint32_t dim = sp->n_real + sp->n_so2;
  for (uint64_t e = 0; e < E; ++e) {
    uint32_t fi = 0;
    int v = orc_check_motion(w, sp, from + e * dim, to + e * dim, resolution, &fi, cnt);
    valid[e] = (uint8_t)(v == 1);
    if (first_invalid) first_invalid[e] = fi;
  }
}

// ---------------------------------------------------------------------------------------
// Neighbour scans over a row-major store (rows[i*dim + d])
// ---------------------------------------------------------------------------------------

// 1-NN — rrt.rs:187-196 (= rrt_connect.rs:128-136, rrt_star.rs:199-208):
//   best = 0; min = distance(tree[0], q); for i in 1..N { d = distance(tree[i], q); if d < min {..} }
void orc_nn1(const orc_space* sp, const double* rows, uint64_t N, const double* q, uint32_t* idx,
             double* dist) {
  uint32_t dim = sp->n_real + sp->n_so2;
  uint64_t best = 0;
  double min_dist = orc_distance(sp, rows, q);
  for (uint64_t i = 1; i < N; ++i) {
    double d = orc_distance(sp, rows + i * dim, q);
    if (d < min_dist) {
      min_dist = d;
      best = i;
    }
  }
  *idx = (uint32_t)best;
  *dist = min_dist;
}

void orc_nn1_batch(const orc_space* sp, const double* rows, uint64_t N, const double* queries,
                   uint64_t Q, uint32_t* idx, double* dist) {
  uint32_t dim = sp->n_real + sp->n_so2;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t q = 0; q < (int64_t)Q; ++q)
    orc_nn1(sp, rows, N, queries + q * dim, idx + q, dist + q);
}

// Radius scan — rrt_star.rs:121-131, prm.rs:131-138, prm.rs:249-256:
//   for i in 0..N { if distance(q, store[i]) < r { out.push(i) } }      (limit: i < idx_limit)
// Returns the count; writes at most cap entries.
uint64_t orc_radius(const orc_space* sp, const double* rows, uint64_t N, const double* q, double r,
                    uint64_t idx_limit, uint32_t* out_idx, double* out_dist, uint64_t cap) {
  uint32_t dim = sp->n_real + sp->n_so2;
  uint64_t n = 0;
  uint64_t lim = idx_limit < N ? idx_limit : N;
  for (uint64_t i = 0; i < lim; ++i) {
    double d = orc_distance(sp, q, rows + i * dim);
    if (d < r) {
      if (n < cap) {
        if (out_idx) out_idx[n] = (uint32_t)i;
        if (out_dist) out_dist[n] = d;
      }
      ++n;
    }
  }
  return n;
}

// k-NN — DEFINED HERE (the reference has none): the k smallest stored states by
// (distance(q, store[i]), i) lexicographically, `exclude` removed by index; output ascending;
// padded with 0xFFFFFFFF / +inf.
void orc_knn(const orc_space* sp, const double* rows, uint64_t N, const double* q, uint32_t k,
             uint32_t exclude, uint32_t* out_idx, double* out_dist) {
  uint32_t dim = sp->n_real + sp->n_so2;
  uint32_t have = 0;
  for (uint32_t e = 0; e < k; ++e) {
    out_idx[e] = 0xFFFFFFFFu;
    out_dist[e] = std::numeric_limits<double>::infinity();
  }
  if (k == 0) return;
  for (uint64_t i = 0; i < N; ++i) {
    if ((uint32_t)i == exclude) continue;
    double d = orc_distance(sp, q, rows + i * dim);
    if (!(d == d)) continue;  // NaN never orders below anything
    // ascending scan in i: a later index only enters on strictly smaller distance
    if (have == k && !(d < out_dist[k - 1])) continue;
    uint32_t pos = have < k ? have : k - 1;
    while (pos > 0 && d < out_dist[pos - 1]) {
      out_dist[pos] = out_dist[pos - 1];
      out_idx[pos] = out_idx[pos - 1];
      --pos;
    }
    out_dist[pos] = d;
    out_idx[pos] = (uint32_t)i;
    if (have < k) ++have;
  }
}

void orc_knn_batch(const orc_space* sp, const double* rows, uint64_t N, const double* queries,
                   uint64_t Q, uint32_t k, const uint32_t* exclude, uint32_t* idx, double* dist) {
  uint32_t dim = sp->n_real + sp->n_so2;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t q = 0; q < (int64_t)Q; ++q)
    orc_knn(sp, rows, N, queries + q * dim, k, exclude ? exclude[q] : 0xFFFFFFFFu, idx + q * k,
            dist + q * k);
}

// ---------------------------------------------------------------------------------------
// Planner loops driven by a pre-generated sample stream (the reference's RNG is unseedable,
// the stream replaces `rng.random_bool(goal_bias) ? sample_goal : sample_uniform`).
// Goal = ball: is_satisfied(s) = distance(s, target) <= radius
//        (CircularGoalRegion, oxmpl/tests/rrt_rvss_tests.rs:46-50).
// ---------------------------------------------------------------------------------------
struct orc_goal {
  const double* target;
  double radius;
};
static inline bool goal_satisfied(const orc_space* sp, const orc_goal* g, const double* s) {
  return orc_distance(sp, s, g->target) <= g->radius;
}

// steering — rrt.rs:200-208
static inline void steer(const orc_space* sp, const double* q_near, const double* q_rand,
                         double min_dist, double max_distance, double* q_new, bool* reached) {
  uint32_t dim = sp->n_real + sp->n_so2;
  if (min_dist > max_distance) {
    double t = max_distance / min_dist;
    orc_interpolate(sp, q_near, q_rand, t, q_new);
    *reached = false;
  } else {
    std::memcpy(q_new, q_rand, sizeof(double) * dim);
    *reached = true;
  }
}

// RRT::solve — rrt.rs:158-227.  tree_states must hold max_nodes rows and be pre-filled with
// the start state in row 0 (setup, rrt.rs:139-156).  Returns the number of nodes; *solved,
// *iters (samples consumed).  Stops when solved, samples run out or the tree is full.
uint64_t orc_rrt(const orc_world* w, const orc_space* sp, const orc_goal* goal, double max_distance,
                 double resolution, const double* samples, uint64_t n_samples, uint64_t max_nodes,
                 double* tree_states, int64_t* parents, int* solved, uint64_t* iters) {
  uint32_t dim = sp->n_real + sp->n_so2;
  uint64_t n = 1;
  parents[0] = -1;
  *solved = 0;
  std::vector<double> q_new(dim);
  uint64_t it = 0;
  for (; it < n_samples && n < max_nodes; ++it) {
    const double* q_rand = samples + it * dim;
    uint32_t near;
    double min_dist;
    orc_nn1(sp, tree_states, n, q_rand, &near, &min_dist);
    const double* q_near = tree_states + (uint64_t)near * dim;
    bool reached;
    steer(sp, q_near, q_rand, min_dist, max_distance, q_new.data(), &reached);
    if (orc_check_motion(w, sp, q_near, q_new.data(), resolution, nullptr, nullptr) == 1) {
      std::memcpy(tree_states + n * dim, q_new.data(), sizeof(double) * dim);
      parents[n] = near;
      ++n;
      if (goal && goal_satisfied(sp, goal, q_new.data())) {
        *solved = 1;
        ++it;
        break;
      }
    }
  }
  *iters = it;
  return n;
}

// RRTConnect::solve — rrt_connect.rs:227-309 with extend rrt_connect.rs:121-159.
// Trees a (start) and b (goal) are pre-filled with their roots.  Returns 1 if solved.
// On a tree-tree connection *conn_start / *conn_goal are the joined node indices; on
// "start tree reached the goal directly" *conn_goal = -1.
static int rc_extend(const orc_world* w, const orc_space* sp, double* tree, int64_t* parents,
                     uint64_t* n, uint64_t cap, const double* q_target, double max_distance,
                     double resolution, bool* reached, uint64_t* new_idx) {
  uint32_t dim = sp->n_real + sp->n_so2;
  uint32_t near;
  double min_dist;
  orc_nn1(sp, tree, *n, q_target, &near, &min_dist);
  std::vector<double> q_near(tree + (uint64_t)near * dim, tree + (uint64_t)near * dim + dim);
  std::vector<double> q_new(dim);
  steer(sp, q_near.data(), q_target, min_dist, max_distance, q_new.data(), reached);
  if (*n >= cap) return 0;
  if (orc_check_motion(w, sp, q_near.data(), q_new.data(), resolution, nullptr, nullptr) == 1) {
    std::memcpy(tree + (*n) * dim, q_new.data(), sizeof(double) * dim);
    parents[*n] = near;
    *new_idx = *n;
    ++(*n);
    return 1;
  }
  return 0;
}

int orc_rrt_connect(const orc_world* w, const orc_space* sp, const orc_goal* goal,
                    double max_distance, double resolution, const double* samples,
                    uint64_t n_samples, uint64_t max_nodes, double* start_tree,
                    int64_t* start_parents, uint64_t* n_start, double* goal_tree,
                    int64_t* goal_parents, uint64_t* n_goal, int64_t* conn_start,
                    int64_t* conn_goal, uint64_t* iters) {
  uint32_t dim = sp->n_real + sp->n_so2;
  *n_start = 1;
  *n_goal = 1;
  start_parents[0] = -1;
  goal_parents[0] = -1;
  *conn_start = -1;
  *conn_goal = -1;
  uint64_t it = 0;
  int solved = 0;
  for (; it < n_samples && *n_start < max_nodes && *n_goal < max_nodes; ++it) {
    bool growing_start = *n_start <= *n_goal;  // rrt_connect.rs:249-254
    double* ta = growing_start ? start_tree : goal_tree;
    int64_t* pa = growing_start ? start_parents : goal_parents;
    uint64_t* na = growing_start ? n_start : n_goal;
    double* tb = growing_start ? goal_tree : start_tree;
    int64_t* pb = growing_start ? goal_parents : start_parents;
    uint64_t* nb = growing_start ? n_goal : n_start;
    const double* q_rand = samples + it * dim;
    bool reached_a;
    uint64_t idx_a;
    if (rc_extend(w, sp, ta, pa, na, max_nodes, q_rand, max_distance, resolution, &reached_a,
                  &idx_a)) {
      const double* q_new = ta + idx_a * dim;
      if (growing_start && goal && goal_satisfied(sp, goal, q_new)) {  // :271-274
        *conn_start = (int64_t)idx_a;
        solved = 1;
        ++it;
        break;
      }
      bool reached_b;
      uint64_t idx_b;
      if (rc_extend(w, sp, tb, pb, nb, max_nodes, q_new, max_distance, resolution, &reached_b,
                    &idx_b)) {
        if (reached_b) {  // :280
          *conn_start = (int64_t)(growing_start ? idx_a : idx_b);
          *conn_goal = (int64_t)(growing_start ? idx_b : idx_a);
          solved = 1;
          ++it;
          break;
        }
      }
    }
  }
  *iters = it;
  return solved;
}

// RRTStar::solve — rrt_star.rs:171-290 (find_neighbours :121-131, cost :107-115,
// choose-parent :233-249, rewire :261-282).  stop_at_goal = 0 runs count-bounded.
// motion_checks counts the check_motion calls the sequential reference makes.
uint64_t orc_rrt_star(const orc_world* w, const orc_space* sp, const orc_goal* goal,
                      double max_distance, double search_radius, double resolution,
                      const double* samples, uint64_t n_samples, uint64_t max_nodes,
                      int stop_at_goal, double* tree_states, int64_t* parents, double* costs,
                      int* solved, uint64_t* iters, uint64_t* motion_checks) {
  uint32_t dim = sp->n_real + sp->n_so2;
  uint64_t n = 1;
  parents[0] = -1;
  costs[0] = 0.0;
  *solved = 0;
  uint64_t mc = 0;
  std::vector<double> q_new(dim);
  std::vector<uint32_t> nbrs;
  uint64_t it = 0;
  for (; it < n_samples && n < max_nodes; ++it) {
    const double* q_rand = samples + it * dim;
    uint32_t near;
    double min_dist;
    orc_nn1(sp, tree_states, n, q_rand, &near, &min_dist);
    const double* q_near = tree_states + (uint64_t)near * dim;
    bool reached;
    steer(sp, q_near, q_rand, min_dist, max_distance, q_new.data(), &reached);
    ++mc;
    if (orc_check_motion(w, sp, q_near, q_new.data(), resolution, nullptr, nullptr) != 1) continue;
    // find_neighbours(temp_node): distance(node.state, tree[i].state) < search_radius
    nbrs.clear();
    for (uint64_t i = 0; i < n; ++i)
      if (orc_distance(sp, q_new.data(), tree_states + i * dim) < search_radius)
        nbrs.push_back((uint32_t)i);
    // choose parent: cost(temp, nbr) = nbr.cost + distance(temp.state, nbr.state)
    uint32_t best_parent = near;
    double min_cost = costs[near] + orc_distance(sp, q_new.data(), q_near);
    for (uint32_t j : nbrs) {
      const double* sj = tree_states + (uint64_t)j * dim;
      double c = costs[j] + orc_distance(sp, q_new.data(), sj);
      if (c < min_cost) {
        ++mc;
        if (orc_check_motion(w, sp, sj, q_new.data(), resolution, nullptr, nullptr) == 1) {
          min_cost = c;
          best_parent = j;
        }
      }
    }
    std::memcpy(tree_states + n * dim, q_new.data(), sizeof(double) * dim);
    parents[n] = best_parent;
    costs[n] = min_cost;
    uint64_t new_idx = n;
    ++n;
    // rewire: cost(nbr, new) = new.cost + distance(nbr.state, new.state)
    for (uint32_t j : nbrs) {
      if ((int64_t)j == parents[new_idx]) continue;
      const double* sj = tree_states + (uint64_t)j * dim;
      double c = costs[new_idx] + orc_distance(sp, sj, q_new.data());
      if (c < costs[j]) {
        ++mc;
        if (orc_check_motion(w, sp, q_new.data(), sj, resolution, nullptr, nullptr) == 1) {
          parents[j] = (int64_t)new_idx;
          costs[j] = c;
        }
      }
    }
    if (stop_at_goal && goal && goal_satisfied(sp, goal, q_new.data())) {
      *solved = 1;
      ++it;
      break;
    }
  }
  *iters = it;
  if (motion_checks) *motion_checks = mc;
  return n;
}

// PRM::construct_roadmap — prm.rs:117-146, count-bounded over a sample stream instead of a
// wall-clock timeout.  Emits the accepted states and, per node, `new_node.edges` at push time
// (the ascending lower-index neighbours) as CSR.  Returns the number of nodes.
uint64_t orc_prm_radius(const orc_world* w, const orc_space* sp, double connection_radius,
                        double resolution, const double* samples, uint64_t n_samples,
                        double* node_states, uint64_t* offsets, uint32_t* idx, uint64_t cap,
                        uint64_t* n_dist_pairs, uint64_t* n_motion_checks) {
  uint32_t dim = sp->n_real + sp->n_so2;
  uint64_t n = 0, ne = 0, pairs = 0, mc = 0;
  offsets[0] = 0;
  for (uint64_t it = 0; it < n_samples; ++it) {
    const double* q = samples + it * dim;
    if (!orc_is_valid(w, sp, q, nullptr)) continue;  // prm.rs:123
    for (uint64_t i = 0; i < n; ++i) {
      const double* other = node_states + i * dim;
      double d = orc_distance(sp, q, other);  // prm.rs:133
      ++pairs;
      if (d < connection_radius) {
        ++mc;
        if (orc_check_motion(w, sp, q, other, resolution, nullptr, nullptr) == 1) {  // :134
          if (ne < cap) idx[ne] = (uint32_t)i;
          ++ne;
        }
      }
    }
    std::memcpy(node_states + n * dim, q, sizeof(double) * dim);
    ++n;
    offsets[n] = ne;
  }
  if (n_dist_pairs) *n_dist_pairs = pairs;
  if (n_motion_checks) *n_motion_checks = mc;
  return n;
}

// k-NN PRM (defined here, the reference has no k-NN PRM): rows
// [row_begin,row_end): k nearest other nodes, each candidate {i,j} checked as
// check_motion(from = max(i,j), to = min(i,j)).  valid = 0 for padded slots.
void orc_prm_knn(const orc_world* w, const orc_space* sp, const double* rows, uint64_t N, uint32_t k,
                 double resolution, uint64_t row_begin, uint64_t row_end, uint32_t* idx,
                 double* dist, uint8_t* valid, orc_counters* cnt) {
  uint32_t dim = sp->n_real + sp->n_so2;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t i = (int64_t)row_begin; i < (int64_t)row_end; ++i) {
    uint64_t o = (uint64_t)(i - (int64_t)row_begin) * k;
    orc_knn(sp, rows, N, rows + (uint64_t)i * dim, k, (uint32_t)i, idx + o, dist + o);
  }
  for (uint64_t i = row_begin; i < row_end; ++i) {
    uint64_t o = (i - row_begin) * k;
    for (uint32_t e = 0; e < k; ++e) {
      uint32_t j = idx[o + e];
      if (j == 0xFFFFFFFFu) {
        valid[o + e] = 0;
        continue;
      }
      uint64_t hi = i > j ? i : j, lo = i > j ? j : i;
      valid[o + e] = (uint8_t)(orc_check_motion(w, sp, rows + hi * dim, rows + lo * dim, resolution,
                                                nullptr, cnt) == 1);
    }
  }
}

// ---------------------------------------------------------------------------------------
// SO(3) — unit quaternions (x, y, z, w), rows of 4 doubles.  acos / sin are the platform libm's,
// which is also what Rust's f64::acos / f64::sin call on linux-gnu, so on one machine these are
// the reference's own bits.  No golden value exists in the reference for this path (its SO3
// tests assert path properties only): parity is pinned by the source text restated here.
// ---------------------------------------------------------------------------------------

// SO3StateSpace::distance — oxmpl/src/base/spaces/so3_state_space.rs:102-112
//   abs_dot = (s1.x*s2.x + s1.y*s2.y + s1.z*s2.z + s1.w*s2.w).abs()
//   if abs_dot > 1.0 - 1e-9 { 0.0 } else { abs_dot.acos() }
double orc_so3_distance(const double* s1, const double* s2) {
  double abs_dot = std::fabs(s1[0] * s2[0] + s1[1] * s2[1] + s1[2] * s2[2] + s1[3] * s2[3]);
  if (abs_dot > 1.0 - 1e-9) return 0.0;
  return std::acos(abs_dot);
}

// SO3StateSpace::interpolate — so3_state_space.rs:114-159: LERP + renormalise when the
// (sign-corrected) dot exceeds 0.9995, SLERP otherwise.
void orc_so3_interpolate(const double* from, const double* to, double t, double* out) {
  double dot = from[0] * to[0] + from[1] * to[1] + from[2] * to[2] + from[3] * to[3];
  double sign = dot < 0.0 ? -1.0 : 1.0;
  dot *= sign;
  if (dot > 0.9995) {
    double v[4];
    for (int c = 0; c < 4; ++c) v[c] = from[c] + t * (to[c] * sign - from[c]);
    double norm = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
    for (int c = 0; c < 4; ++c) out[c] = v[c] / norm;
  } else {
    double theta = std::acos(dot);
    double sin_theta = std::sin(theta);
    double s0 = std::sin((1.0 - t) * theta) / sin_theta;
    double s1 = std::sin(t * theta) / sin_theta * sign;
    for (int c = 0; c < 4; ++c) out[c] = from[c] * s0 + to[c] * s1;
  }
}

// 1-NN loop of rrt.rs:187-196 with distance(tree[i], q)
void orc_so3_nn1_batch(const double* rows, uint64_t N, const double* queries, uint64_t Q,
                       uint32_t* idx, double* dist) {
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < (int64_t)Q; ++q) {
    const double* qq = queries + (uint64_t)q * 4;
    uint32_t best = 0;
    double mn = orc_so3_distance(rows, qq);
    for (uint64_t i = 1; i < N; ++i) {
      double d = orc_so3_distance(rows + i * 4, qq);
      if (d < mn) {
        mn = d;
        best = (uint32_t)i;
      }
    }
    idx[q] = best;
    dist[q] = mn;
  }
}

// k-NN over SO3 -- orc_knn's definition (the reference has no k-NN) with
// SO3StateSpace::distance(q, store[i]), so3_state_space.rs:102-112.
void orc_so3_knn_batch(const double* rows, uint64_t N, const double* queries, uint64_t Q, uint32_t k,
                       const uint32_t* exclude, uint32_t* idx, double* dist) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t qi = 0; qi < (int64_t)Q; ++qi) {
    const double* q = queries + (uint64_t)qi * 4;
    uint32_t* out_idx = idx + (uint64_t)qi * k;
    double* out_dist = dist + (uint64_t)qi * k;
    const uint32_t ex = exclude ? exclude[qi] : 0xFFFFFFFFu;
    uint32_t have = 0;
    for (uint32_t e = 0; e < k; ++e) {
      out_idx[e] = 0xFFFFFFFFu;
      out_dist[e] = std::numeric_limits<double>::infinity();
    }
    if (k == 0) continue;
    for (uint64_t i = 0; i < N; ++i) {
      if ((uint32_t)i == ex) continue;
      double d = orc_so3_distance(q, rows + i * 4);
      if (!(d == d)) continue;
      if (have == k && !(d < out_dist[k - 1])) continue;
      uint32_t pos = have < k ? have : k - 1;
      while (pos > 0 && d < out_dist[pos - 1]) {
        out_dist[pos] = out_dist[pos - 1];
        out_idx[pos] = out_idx[pos - 1];
        --pos;
      }
      out_dist[pos] = d;
      out_idx[pos] = (uint32_t)i;
      if (have < k) ++have;
    }
  }
}

// radius loop of rrt_star.rs:121-131 with distance(q, store[i]) < r; returns the count, writes
// at most cap entries (ascending index).
uint64_t orc_so3_radius(const double* rows, uint64_t N, const double* q, double r, uint32_t* idx,
                        double* dist, uint64_t cap) {
  uint64_t n = 0;
  for (uint64_t i = 0; i < N; ++i) {
    double d = orc_so3_distance(q, rows + i * 4);
    if (d < r) {
      if (n < cap) {
        if (idx) idx[n] = (uint32_t)i;
        if (dist) dist[n] = d;
      }
      ++n;
    }
  }
  return n;
}

// ForbiddenConeChecker — oxmpl/tests/rrt_so3ss_tests.rs:46-56, any number of cones:
//   valid iff distance(centre, state) > angle for every cone.  cones: rows (x, y, z, w, angle)
int orc_so3_is_valid(const double* cones, uint64_t n_cones, const double* state) {
  for (uint64_t c = 0; c < n_cones; ++c)
    if (!(orc_so3_distance(cones + c * 5, state) > cones[c * 5 + 4])) return 0;
  return 1;
}

// check_motion (rrt.rs:90-116) over SO3
int orc_so3_check_motion(const double* cones, uint64_t n_cones, const double* from, const double* to,
                         double resolution, uint32_t* first_invalid) {
  double dist = orc_so3_distance(from, to);
  uint64_t num_steps = ceil_as_usize(dist / resolution);
  if (first_invalid) *first_invalid = 0;
  if (num_steps <= 1) {
    int v = orc_so3_is_valid(cones, n_cones, to);
    if (!v && first_invalid) *first_invalid = 1;
    return v;
  }
  if (num_steps > (1ull << 32)) return -1;
  double st[4];
  for (uint64_t i = 1; i <= num_steps; ++i) {
    double t = (double)i / (double)num_steps;
    orc_so3_interpolate(from, to, t, st);
    if (!orc_so3_is_valid(cones, n_cones, st)) {
      if (first_invalid) *first_invalid = (uint32_t)i;
      return 0;
    }
  }
  return 1;
}

// PRM::solve — prm.rs:227-307, on a roadmap given as each node's `new_node.edges` at push time
// (the lower-index adjacency, CSR, what orc_prm_radius / the library's radius build emit).
// The roadmap's edge lists are rebuilt the way construct_roadmap grows them: node i takes its
// lower neighbours in ascending order (prm.rs:135) and is appended to each of them afterwards
// (prm.rs:143-145).  Goal = ball (CircularGoalRegion).  Writes the roadmap indices of the path
// (the reference's Path is the start state followed by these nodes' states, prm.rs:189-208).
// Returns 0 = Ok(path), 1 = NoSolutionFound, 2 = InvalidStartState, 3 = UnsampledStateSpace.
int orc_prm_solve(const orc_world* w, const orc_space* sp, const double* rows, uint64_t N,
                  const uint64_t* lower_off, const uint32_t* lower_idx, double connection_radius,
                  double resolution, const double* start, const orc_goal* goal, uint32_t* path,
                  uint32_t cap, uint32_t* path_len) {
  uint32_t dim = sp->n_real + sp->n_so2;
  *path_len = 0;
  if (N == 0) return 3;                                   // prm.rs:239-241
  if (!orc_is_valid(w, sp, start, nullptr)) return 2;     // prm.rs:243-246
  std::vector<std::vector<uint32_t>> edges(N);
  for (uint64_t i = 0; i < N; ++i) {
    for (uint64_t t = lower_off[i]; t < lower_off[i + 1]; ++t) edges[i].push_back(lower_idx[t]);
    for (uint64_t t = lower_off[i]; t < lower_off[i + 1]; ++t) edges[lower_idx[t]].push_back((uint32_t)i);
  }
  // start connection — prm.rs:249-256
  std::vector<uint32_t> start_connections;
  for (uint64_t i = 0; i < N; ++i)
    if (orc_distance(sp, start, rows + i * dim) < connection_radius &&
        orc_check_motion(w, sp, start, rows + i * dim, resolution, nullptr, nullptr) == 1)
      start_connections.push_back((uint32_t)i);
  // goal nodes — prm.rs:259-264
  std::vector<uint8_t> is_goal(N, 0);
  bool any_goal = false;
  for (uint64_t i = 0; i < N; ++i)
    if (goal_satisfied(sp, goal, rows + i * dim)) is_goal[i] = 1, any_goal = true;
  if (start_connections.empty() || !any_goal) return 1;   // prm.rs:266-268
  // BFS — prm.rs:271-301.  The reference seeds the queue with the start connections AND pushes
  // each of them again while filling parent_map; the duplicates are restated (they pop after the
  // first copies and find every neighbour visited).
  const uint32_t NONE = 0xFFFFFFFFu;
  std::vector<uint32_t> queue(start_connections.begin(), start_connections.end());
  std::vector<uint32_t> parent(N, NONE);
  std::vector<uint8_t> visited(N, 0);
  for (uint32_t idx : start_connections) {
    queue.push_back(idx);
    parent[idx] = NONE;
    visited[idx] = 1;
  }
  int64_t goal_reached = -1;
  for (size_t head = 0; head < queue.size(); ++head) {
    uint32_t cur = queue[head];
    if (is_goal[cur]) {
      goal_reached = cur;
      break;
    }
    for (uint32_t nb : edges[cur])
      if (!visited[nb]) {
        visited[nb] = 1;
        parent[nb] = cur;
        queue.push_back(nb);
      }
  }
  if (goal_reached < 0) return 1;                          // prm.rs:304
  // reconstruct_path — prm.rs:189-208
  std::vector<uint32_t> chain;
  for (uint32_t v = (uint32_t)goal_reached; v != NONE; v = parent[v]) chain.push_back(v);
  std::reverse(chain.begin(), chain.end());
  *path_len = (uint32_t)chain.size();
  for (size_t i = 0; i < chain.size() && i < cap; ++i) path[i] = chain[i];
  return 0;
}

// Build identification, so tests can assert the flags the oracle was compiled with.
int orc_build_info(void) {
  int v = 1;
#if defined(_OPENMP)
  v |= 2;
#endif
  return v;
}

// A contraction-sensitive probe: returns a*b+c computed with separate roundings.  With
// a = 1+2^-27, b = 1+2^-27, c = -1 an FMA gives 2^-26 + 2^-54, separate ops give 2^-26.
double orc_probe_mul_add(double a, double b, double c) { return a * b + c; }

// RRT::solve over SO3 -- rrt.rs:158-227 with SO3StateSpace::distance / interpolate and the
// ForbiddenConeChecker / SO3GoalRegion fixtures (oxmpl/tests/rrt_so3ss_tests.rs:46-68); same
// conventions as orc_rrt.  goal_target NULL: no goal test.
uint64_t orc_so3_rrt(const double* cones, uint64_t n_cones, const double* goal_target,
                     double goal_radius, double max_distance, double resolution,
                     const double* samples, uint64_t n_samples, uint64_t max_nodes,
                     double* tree_states, int64_t* parents, int* solved, uint64_t* iters) {
  uint64_t n = 1;
  parents[0] = -1;
  *solved = 0;
  uint64_t it = 0;
  for (; it < n_samples && n < max_nodes; ++it) {
    const double* q_rand = samples + it * 4;
    uint32_t near = 0;
    double min_dist = orc_so3_distance(tree_states, q_rand);  // rrt.rs:187-196
    for (uint64_t i = 1; i < n; ++i) {
      double d = orc_so3_distance(tree_states + i * 4, q_rand);
      if (d < min_dist) {
        min_dist = d;
        near = (uint32_t)i;
      }
    }
    const double* q_near = tree_states + (uint64_t)near * 4;
    double q_new[4];
    if (min_dist > max_distance) {  // rrt.rs:200-208
      orc_so3_interpolate(q_near, q_rand, max_distance / min_dist, q_new);
    } else {
      std::memcpy(q_new, q_rand, sizeof(q_new));
    }
    if (orc_so3_check_motion(cones, n_cones, q_near, q_new, resolution, nullptr) == 1) {
      std::memcpy(tree_states + n * 4, q_new, sizeof(q_new));
      parents[n] = near;
      ++n;
      if (goal_target && orc_so3_distance(q_new, goal_target) <= goal_radius) {
        *solved = 1;
        ++it;
        break;
      }
    }
  }
  *iters = it;
  return n;
}

// RRTConnect::solve over SO3 -- rrt_connect.rs:227-309 with extend :121-159; conventions of
// orc_rrt_connect, fixtures of orc_so3_rrt.
static int so3_rc_extend(const double* cones, uint64_t n_cones, double* tree, int64_t* parents,
                         uint64_t* n, uint64_t cap, const double* q_target, double max_distance,
                         double resolution, bool* reached, uint64_t* new_idx) {
  uint32_t near = 0;
  double min_dist = orc_so3_distance(tree, q_target);  // rrt_connect.rs:128-136
  for (uint64_t i = 1; i < *n; ++i) {
    double d = orc_so3_distance(tree + i * 4, q_target);
    if (d < min_dist) {
      min_dist = d;
      near = (uint32_t)i;
    }
  }
  double q_near[4], q_new[4];
  std::memcpy(q_near, tree + (uint64_t)near * 4, sizeof(q_near));
  if (min_dist > max_distance) {
    orc_so3_interpolate(q_near, q_target, max_distance / min_dist, q_new);
    *reached = false;
  } else {
    std::memcpy(q_new, q_target, sizeof(q_new));
    *reached = true;
  }
  if (*n >= cap) return 0;
  if (orc_so3_check_motion(cones, n_cones, q_near, q_new, resolution, nullptr) == 1) {
    std::memcpy(tree + (*n) * 4, q_new, sizeof(q_new));
    parents[*n] = near;
    *new_idx = *n;
    ++(*n);
    return 1;
  }
  return 0;
}

int orc_so3_rrt_connect(const double* cones, uint64_t n_cones, const double* goal_target,
                        double goal_radius, double max_distance, double resolution,
                        const double* samples, uint64_t n_samples, uint64_t max_nodes,
                        double* start_tree, int64_t* start_parents, uint64_t* n_start,
                        double* goal_tree, int64_t* goal_parents, uint64_t* n_goal,
                        int64_t* conn_start, int64_t* conn_goal, uint64_t* iters) {
  *n_start = 1;
  *n_goal = 1;
  start_parents[0] = -1;
  goal_parents[0] = -1;
  *conn_start = -1;
  *conn_goal = -1;
  uint64_t it = 0;
  int solved = 0;
  for (; it < n_samples && *n_start < max_nodes && *n_goal < max_nodes; ++it) {
    bool growing_start = *n_start <= *n_goal;  // rrt_connect.rs:249-254
    double* ta = growing_start ? start_tree : goal_tree;
    int64_t* pa = growing_start ? start_parents : goal_parents;
    uint64_t* na = growing_start ? n_start : n_goal;
    double* tb = growing_start ? goal_tree : start_tree;
    int64_t* pb = growing_start ? goal_parents : start_parents;
    uint64_t* nb = growing_start ? n_goal : n_start;
    bool reached_a, reached_b;
    uint64_t idx_a, idx_b;
    if (!so3_rc_extend(cones, n_cones, ta, pa, na, max_nodes, samples + it * 4, max_distance,
                       resolution, &reached_a, &idx_a))
      continue;
    const double* q_new = ta + idx_a * 4;
    if (growing_start && goal_target && orc_so3_distance(q_new, goal_target) <= goal_radius) {
      *conn_start = (int64_t)idx_a;  // :271-274
      solved = 1;
      ++it;
      break;
    }
    if (so3_rc_extend(cones, n_cones, tb, pb, nb, max_nodes, q_new, max_distance, resolution,
                      &reached_b, &idx_b) &&
        reached_b) {  // :280
      *conn_start = (int64_t)(growing_start ? idx_a : idx_b);
      *conn_goal = (int64_t)(growing_start ? idx_b : idx_a);
      solved = 1;
      ++it;
      break;
    }
  }
  *iters = it;
  return solved;
}

}  // extern "C"
