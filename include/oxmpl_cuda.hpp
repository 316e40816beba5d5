// oxmpl_cuda.hpp — host-side mirror of oxmpl's planner-facing interface on top of the liboxg
// C ABI (include/oxg.h).  Header-only C++17; link with -loxg.
//
// The reference is Rust and this image has no cargo/rustc, so the compiled host side above the
// C ABI is written in C++ with the reference's names, argument meaning and error behaviour
// (INTEGRATION.md shows the equivalent `oxmpl-cuda` Rust crate).  What maps to what:
//
//   oxmpl::cuda::RealVectorStateSpace   oxmpl/src/base/spaces/real_vector_state_space.rs
//   oxmpl::cuda::GpuValidityChecker     impl StateValidityChecker (oxmpl/src/base/validity.rs:39-48)
//   oxmpl::cuda::GoalRegion             the CircularGoalRegion of oxmpl/tests/rrt_rvss_tests.rs:39-69
//   oxmpl::cuda::ProblemDefinition      oxmpl/src/base/problem_definition.rs:16-20
//   oxmpl::cuda::RRT / RRTConnect / RRTStar / PRM
//                                       oxmpl/src/geometric/planners/{rrt,rrt_connect,rrt_star,prm}.rs
//                                       same constructors: RRT(max_distance, goal_bias),
//                                       RRTConnect(max_distance, goal_bias),
//                                       RRTStar(max_distance, goal_bias, search_radius),
//                                       PRM(timeout, connection_radius); setup(); solve(timeout)
//   oxmpl::cuda::PlanningError          oxmpl/src/base/error.rs:97-108
//
// The planners keep the reference's control flow (sample -> nearest -> steer -> check_motion ->
// push -> goal test) but run it on the GPU: states live in an oxg_store, the per-iteration loop
// is one persistent kernel per chunk of pre-drawn samples (oxg_rrt_solve & co.).  Sampling is
// host-side (std::mt19937_64 here, `rand` in the reference — inputs to the path, not part of
// it), drawn in the reference's order: `random_bool(goal_bias) ? sample_goal : sample_uniform`.
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <deque>
#include <limits>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "oxg.h"

namespace oxmpl {
namespace cuda {

using State = std::vector<double>;  // RealVectorState::values
using Path = std::vector<State>;    // Path<S>(pub Vec<S>)

// oxmpl/src/base/error.rs:97-108
struct PlanningError : std::runtime_error {
  enum Kind { Timeout, NoSolutionFound, PlannerUninitialised, InvalidStartState, UnsampledStateSpace };
  Kind kind;
  PlanningError(Kind k, const char* msg) : std::runtime_error(msg), kind(k) {}
  static PlanningError timeout() { return {Timeout, "No solution found within timeout."}; }
  static PlanningError no_solution() { return {NoSolutionFound, "No solution found."}; }
  static PlanningError uninitialised() {
    return {PlannerUninitialised, "<Planner>.setup() was not called, thus Planner is uninitialised."};
  }
  static PlanningError invalid_start() {
    return {InvalidStartState, "Start state is not valid in the current StateSpace."};
  }
  static PlanningError unsampled() {
    return {UnsampledStateSpace, "StateSpace is not sampled. Either Tree or Roadmap is empty."};
  }
};

struct BackendError : std::runtime_error {
  int code;
  BackendError(int c) : std::runtime_error(std::string("liboxg: ") + oxg_last_error()), code(c) {}
};
inline void check(int code) {
  if (code != OXG_OK) throw BackendError(code);
}

// ------------------------------------------------------------------------------------------
// RealVectorStateSpace — host arithmetic identical to the reference (IEEE f64, no FMA: compile
// this header with -ffp-contract=off), used for the pieces that stay on the host.
// ------------------------------------------------------------------------------------------
class RealVectorStateSpace {
 public:
  // RealVectorStateSpace::new(dimension, bounds) (real_vector_state_space.rs:65-100)
  RealVectorStateSpace(std::size_t dimension, std::vector<std::pair<double, double>> bounds)
      : dimension_(dimension), bounds_(std::move(bounds)) {
    if (bounds_.empty())
      bounds_.assign(dimension_, {-std::numeric_limits<double>::infinity(),
                                  std::numeric_limits<double>::infinity()});
    if (bounds_.size() != dimension_) throw std::invalid_argument("DimensionMismatch");
    for (auto& b : bounds_)
      if (b.first >= b.second) throw std::invalid_argument("InvalidBound");
  }
  std::size_t dimension() const { return dimension_; }
  const std::vector<std::pair<double, double>>& bounds() const { return bounds_; }

  double distance(const State& a, const State& b) const {  // :137-155
    double s = 0.0;
    for (std::size_t i = 0; i < dimension_; ++i) {
      double d = a[i] - b[i];
      s = s + d * d;
    }
    return std::sqrt(s);
  }
  void interpolate(const State& from, const State& to, double t, State& out) const {  // :161-186
    out.resize(dimension_);
    for (std::size_t i = 0; i < dimension_; ++i) out[i] = from[i] + (to[i] - from[i]) * t;
  }
  double get_maximum_extent() const {  // :103-118
    for (auto& b : bounds_)
      if (!std::isfinite(b.first) || !std::isfinite(b.second)) return 1.0;
    double s = 0.0;
    for (auto& b : bounds_) {
      double d = b.second - b.first;
      s = s + d * d;
    }
    return std::sqrt(s);
  }
  void set_longest_valid_segment_fraction(double f) {  // :121-129
    fraction_ = (f > 0.0 && f <= 1.0) ? f : (f <= 0.0 ? 0.0 : 1.0);
  }
  double get_longest_valid_segment_length() const { return get_maximum_extent() * fraction_; }  // :251-253
  bool satisfies_bounds(const State& s) const {  // :209-224
    for (std::size_t i = 0; i < dimension_; ++i)
      if (s[i] < bounds_[i].first || s[i] > bounds_[i].second) return false;
    return true;
  }
  template <class Rng>
  State sample_uniform(Rng& rng) const {  // :233-249 (needs finite bounds)
    State s(dimension_);
    for (std::size_t i = 0; i < dimension_; ++i) {
      if (!std::isfinite(bounds_[i].first) || !std::isfinite(bounds_[i].second))
        throw std::runtime_error("UnboundedDimension");
      s[i] = std::uniform_real_distribution<double>(bounds_[i].first, bounds_[i].second)(rng);
    }
    return s;
  }
  oxg_space_desc desc() const { return oxg_space_desc{(uint32_t)dimension_, 0u, 1.0, 1.0}; }

 private:
  std::size_t dimension_;
  std::vector<std::pair<double, double>> bounds_;
  double fraction_ = 0.05;  // :98
};

// ------------------------------------------------------------------------------------------
// GpuValidityChecker — a declarative StateValidityChecker (validity.rs:39-48)
// ------------------------------------------------------------------------------------------
class GpuValidityChecker {
 public:
  struct Obstacles {
    std::vector<double> spheres;   // rows [c_0..c_{d-1}, r]      valid iff dist > r   (README.md:147-150)
    std::vector<double> aabbs;     // rows [lo_0.., hi_0..]       invalid iff inside, faces inclusive
    std::vector<double> capsules;  // rows [a.., b.., r]
  };
  GpuValidityChecker(uint32_t dim, const Obstacles& o) : dim_(dim) {
    check(oxg_world_create(dim, o.spheres.data(), o.spheres.size() / (dim + 1), o.aabbs.data(),
                           o.aabbs.size() / (2 * dim), o.capsules.data(),
                           o.capsules.size() / (2 * dim + 1), &w_));
  }
  ~GpuValidityChecker() { oxg_world_destroy(w_); }
  GpuValidityChecker(const GpuValidityChecker&) = delete;
  GpuValidityChecker& operator=(const GpuValidityChecker&) = delete;

  // fn is_valid(&self, state: &S) -> bool — the trait method (one state per call: correct, slow;
  // the planners use the batched paths below)
  bool is_valid(const State& s) const {
    uint8_t out = 0;
    oxg_space_desc sp{dim_, 0u, 1.0, 1.0};
    check(oxg_is_valid(w_, &sp, s.data(), 1, &out));
    return out != 0;
  }
  std::vector<uint8_t> is_valid_batch(const std::vector<double>& rows) const {
    std::vector<uint8_t> out(rows.size() / dim_);
    oxg_space_desc sp{dim_, 0u, 1.0, 1.0};
    check(oxg_is_valid(w_, &sp, rows.data(), out.size(), out.data()));
    return out;
  }
  // check_motion(from, to) of rrt.rs:90-116 for one edge
  bool check_motion(const State& from, const State& to, double resolution) const {
    uint8_t v = 0;
    oxg_space_desc sp{dim_, 0u, 1.0, 1.0};
    check(oxg_check_motion(w_, &sp, from.data(), to.data(), 1, resolution, &v, nullptr));
    return v != 0;
  }
  oxg_world handle() const { return w_; }
  uint32_t dim() const { return dim_; }

 private:
  uint32_t dim_;
  oxg_world w_ = nullptr;
};

// The goal of the reference's tests: a ball around a target (rrt_rvss_tests.rs:39-69).
struct GoalRegion {
  State target;
  double radius;
  bool is_satisfied(const RealVectorStateSpace& sp, const State& s) const {
    return sp.distance(s, target) <= radius;  // :46-50
  }
  double distance_goal(const RealVectorStateSpace& sp, const State& s) const {
    return std::max(0.0, sp.distance(s, target) - radius);
  }
  template <class Rng>
  State sample_goal(Rng& rng) const {  // 2-D disc sampling of :58-68; other dims: the centre
    if (target.size() != 2) return target;
    const double angle = std::uniform_real_distribution<double>(0.0, 2.0 * M_PI)(rng);
    const double r = radius * std::sqrt(std::uniform_real_distribution<double>(0.0, 1.0)(rng));
    return State{target[0] + r * std::cos(angle), target[1] + r * std::sin(angle)};
  }
};

struct ProblemDefinition {  // problem_definition.rs:16-20
  std::shared_ptr<RealVectorStateSpace> space;
  std::vector<State> start_states;
  std::shared_ptr<GoalRegion> goal;
};

namespace detail {
class Store {
 public:
  explicit Store(const oxg_space_desc& sp, uint64_t cap = 0) { check(oxg_store_create(&sp, cap, &s_)); }
  ~Store() { oxg_store_destroy(s_); }
  Store(const Store&) = delete;
  Store& operator=(const Store&) = delete;
  oxg_store h() const { return s_; }
  void clear() { check(oxg_store_clear(s_)); }
  uint64_t size() const {
    uint64_t n = 0;
    check(oxg_store_size(s_, &n));
    return n;
  }
  uint64_t append(const double* rows, uint64_t n) {
    uint64_t first = 0;
    check(oxg_store_append(s_, rows, n, &first));
    return first;
  }
  State read(uint64_t i, uint32_t dim) const {
    State out(dim);
    check(oxg_store_read(s_, i, 1, out.data()));
    return out;
  }

 private:
  oxg_store s_ = nullptr;
};

using Clock = std::chrono::steady_clock;
inline double seconds_since(Clock::time_point t0) {
  return std::chrono::duration<double>(Clock::now() - t0).count();
}

// The reference's per-iteration draw (rrt.rs:177-184) for a chunk of iterations.
template <class Rng>
std::vector<double> draw_stream(const ProblemDefinition& pd, double goal_bias, uint64_t n, Rng& rng) {
  const std::size_t dim = pd.space->dimension();
  std::vector<double> out(n * dim);
  std::bernoulli_distribution coin(goal_bias);
  for (uint64_t i = 0; i < n; ++i) {
    State q = coin(rng) ? pd.goal->sample_goal(rng) : pd.space->sample_uniform(rng);
    std::copy(q.begin(), q.end(), out.begin() + i * dim);
  }
  return out;
}

inline Path trace(const Store& st, const std::vector<int64_t>& parents, int64_t from, uint32_t dim) {
  Path p;  // reconstruct_path (rrt.rs:118-128): walk the parents, then reverse
  for (int64_t i = from; i >= 0; i = parents[(std::size_t)i]) p.push_back(st.read((uint64_t)i, dim));
  std::reverse(p.begin(), p.end());
  return p;
}
}  // namespace detail

// ------------------------------------------------------------------------------------------
// Planners.  Shared shape: Planner::setup(problem_def, validity_checker); solve(timeout).
// `chunk` = iterations per kernel launch; `max_nodes` = capacity of a tree (the reference is
// bounded by the timeout only).
// ------------------------------------------------------------------------------------------
class RRT {
 public:
  double max_distance, goal_bias;
  uint64_t chunk = 1u << 14, max_nodes = 1u << 20;
  uint64_t seed = 0x9E3779B97F4A7C15ull;

  RRT(double max_distance_, double goal_bias_) : max_distance(max_distance_), goal_bias(goal_bias_) {}

  void setup(std::shared_ptr<ProblemDefinition> pd, std::shared_ptr<GpuValidityChecker> vc) {  // rrt.rs:139-156
    pd_ = std::move(pd);
    vc_ = std::move(vc);
    tree_.reset(new detail::Store(pd_->space->desc(), 1024));
    tree_->append(pd_->start_states.at(0).data(), 1);
    parents_.clear();
  }

  Path solve(double timeout_s) {  // rrt.rs:158-227
    if (!pd_ || !vc_) throw PlanningError::uninitialised();
    std::mt19937_64 rng(seed);
    const auto t0 = detail::Clock::now();
    const uint32_t dim = (uint32_t)pd_->space->dimension();
    oxg_plan_params prm{};
    prm.max_distance = max_distance;
    prm.resolution = pd_->space->get_longest_valid_segment_length() * 0.1;  // rrt.rs:97
    prm.goal_radius = pd_->goal->radius;
    prm.max_nodes = max_nodes;
    prm.stop_at_goal = 1;
    parents_.assign(max_nodes, -1);
    for (;;) {
      const double left = timeout_s - detail::seconds_since(t0);
      if (left <= 0.0) throw PlanningError::timeout();
      prm.timeout_s = left;
      const std::vector<double> stream = detail::draw_stream(*pd_, goal_bias, chunk, rng);
      oxg_plan_result r{};
      check(oxg_rrt_solve(vc_->handle(), tree_->h(), &prm, pd_->goal->target.data(), stream.data(),
                          chunk, parents_.data(), &r));
      last_ = r;
      if (r.solved) return detail::trace(*tree_, parents_, r.goal_node[0], dim);
      if (r.status == 3 || r.status == 2) throw PlanningError::timeout();
    }
  }
  uint64_t tree_size() const { return tree_ ? tree_->size() : 0; }
  const oxg_plan_result& last_result() const { return last_; }

 private:
  std::shared_ptr<ProblemDefinition> pd_;
  std::shared_ptr<GpuValidityChecker> vc_;
  std::unique_ptr<detail::Store> tree_;
  std::vector<int64_t> parents_;
  oxg_plan_result last_{};
};

class RRTStar {
 public:
  double max_distance, goal_bias, search_radius;
  uint64_t chunk = 1u << 14, max_nodes = 1u << 20;
  uint64_t seed = 0x9E3779B97F4A7C15ull;

  RRTStar(double max_distance_, double goal_bias_, double search_radius_)
      : max_distance(max_distance_), goal_bias(goal_bias_), search_radius(search_radius_) {}

  void setup(std::shared_ptr<ProblemDefinition> pd, std::shared_ptr<GpuValidityChecker> vc) {
    pd_ = std::move(pd);
    vc_ = std::move(vc);
    tree_.reset(new detail::Store(pd_->space->desc(), 1024));
    tree_->append(pd_->start_states.at(0).data(), 1);
  }

  Path solve(double timeout_s) {  // rrt_star.rs:171-290
    if (!pd_ || !vc_) throw PlanningError::uninitialised();
    std::mt19937_64 rng(seed);
    const auto t0 = detail::Clock::now();
    const uint32_t dim = (uint32_t)pd_->space->dimension();
    oxg_plan_params prm{};
    prm.max_distance = max_distance;
    prm.search_radius = search_radius;
    prm.resolution = pd_->space->get_longest_valid_segment_length() * 0.1;
    prm.goal_radius = pd_->goal->radius;
    prm.max_nodes = max_nodes;
    prm.stop_at_goal = 1;
    parents_.assign(max_nodes, -1);
    costs_.assign(max_nodes, 0.0);
    for (;;) {
      const double left = timeout_s - detail::seconds_since(t0);
      if (left <= 0.0) throw PlanningError::timeout();
      prm.timeout_s = left;
      const std::vector<double> stream = detail::draw_stream(*pd_, goal_bias, chunk, rng);
      oxg_plan_result r{};
      const int rc = oxg_rrt_star_solve(vc_->handle(), tree_->h(), &prm, pd_->goal->target.data(),
                                        stream.data(), chunk, parents_.data(), costs_.data(), &r);
      if (rc == OXG_E_CAPACITY) throw PlanningError::no_solution();  // OXG_RRT_STAR_MAX_NEIGHBOURS (oxg.h)
      check(rc);
      last_ = r;
      if (r.solved) return detail::trace(*tree_, parents_, r.goal_node[0], dim);
      if (r.status == 3 || r.status == 2) throw PlanningError::timeout();
    }
  }
  double cost_of(uint64_t node) const { return costs_.at(node); }
  const oxg_plan_result& last_result() const { return last_; }

 private:
  std::shared_ptr<ProblemDefinition> pd_;
  std::shared_ptr<GpuValidityChecker> vc_;
  std::unique_ptr<detail::Store> tree_;
  std::vector<int64_t> parents_;
  std::vector<double> costs_;
  oxg_plan_result last_{};
};

class RRTConnect {
 public:
  double max_distance, goal_bias;
  uint64_t chunk = 1u << 14, max_nodes = 1u << 20;
  uint64_t seed = 0x9E3779B97F4A7C15ull;

  RRTConnect(double max_distance_, double goal_bias_) : max_distance(max_distance_), goal_bias(goal_bias_) {}

  void setup(std::shared_ptr<ProblemDefinition> pd, std::shared_ptr<GpuValidityChecker> vc) {  // rrt_connect.rs:191-216
    pd_ = std::move(pd);
    vc_ = std::move(vc);
    start_.reset(new detail::Store(pd_->space->desc(), 1024));
    goal_.reset(new detail::Store(pd_->space->desc(), 1024));
    start_->append(pd_->start_states.at(0).data(), 1);
    // the goal tree is rooted at a sampled goal state (rrt_connect.rs:207-214)
    std::mt19937_64 rng(seed ^ 0xD1B54A32D192ED03ull);
    const State g = pd_->goal->sample_goal(rng);
    goal_->append(g.data(), 1);
  }

  Path solve(double timeout_s) {  // rrt_connect.rs:227-309
    if (!pd_ || !vc_) throw PlanningError::uninitialised();
    std::mt19937_64 rng(seed);
    const auto t0 = detail::Clock::now();
    const uint32_t dim = (uint32_t)pd_->space->dimension();
    oxg_plan_params prm{};
    prm.max_distance = max_distance;
    prm.resolution = pd_->space->get_longest_valid_segment_length() * 0.1;
    prm.goal_radius = pd_->goal->radius;
    prm.max_nodes = max_nodes;
    std::vector<int64_t> pa(max_nodes, -1), pb(max_nodes, -1);
    for (;;) {
      const double left = timeout_s - detail::seconds_since(t0);
      if (left <= 0.0) throw PlanningError::timeout();
      prm.timeout_s = left;
      const std::vector<double> stream = detail::draw_stream(*pd_, goal_bias, chunk, rng);
      oxg_plan_result r{};
      check(oxg_rrt_connect_solve(vc_->handle(), start_->h(), goal_->h(), &prm,
                                  pd_->goal->target.data(), stream.data(), chunk, pa.data(),
                                  pb.data(), &r));
      last_ = r;
      if (r.solved) {
        Path p = detail::trace(*start_, pa, r.goal_node[0], dim);
        if (r.goal_node[1] >= 0) {  // merge (rrt_connect.rs:288-304): reversed goal path, skip the duplicate
          Path g = detail::trace(*goal_, pb, r.goal_node[1], dim);
          std::reverse(g.begin(), g.end());
          p.insert(p.end(), g.begin() + 1, g.end());
        }
        return p;
      }
      if (r.status == 3 || r.status == 2) throw PlanningError::timeout();
    }
  }
  const oxg_plan_result& last_result() const { return last_; }

 private:
  std::shared_ptr<ProblemDefinition> pd_;
  std::shared_ptr<GpuValidityChecker> vc_;
  std::unique_ptr<detail::Store> start_, goal_;
  oxg_plan_result last_{};
};

// PRM (prm.rs): construct_roadmap() samples, keeps the valid states and connects every new node
// to the earlier nodes within connection_radius whose motion is valid — in batches through
// oxg_prm_build_radius (the batched build is identical to the incremental one because node i
// only ever looks at nodes j < i).  solve() connects the start, scans for goal nodes and runs
// the reference's BFS.
class PRM {
 public:
  double timeout, connection_radius;
  uint64_t batch = 1u << 13, max_nodes = 1u << 22;
  uint64_t seed = 0x9E3779B97F4A7C15ull;

  PRM(double timeout_, double connection_radius_) : timeout(timeout_), connection_radius(connection_radius_) {}

  void setup(std::shared_ptr<ProblemDefinition> pd, std::shared_ptr<GpuValidityChecker> vc) {  // prm.rs:212-225
    pd_ = std::move(pd);
    vc_ = std::move(vc);
    store_.reset(new detail::Store(pd_->space->desc(), 1024));
    edges_.clear();
  }
  void set_problem_definition(std::shared_ptr<ProblemDefinition> pd) { pd_ = std::move(pd); }  // prm.rs:88-90
  const std::vector<std::vector<uint32_t>>& get_roadmap() const { return edges_; }            // prm.rs:82-84
  uint64_t roadmap_size() const { return edges_.size(); }

  // prm.rs:96-154; `max_samples` bounds the build by count as well as by `timeout`.
  void construct_roadmap(uint64_t max_samples = std::numeric_limits<uint64_t>::max()) {
    if (!pd_ || !vc_) throw PlanningError::uninitialised();
    if (!edges_.empty()) return;  // prm.rs:106-113: an existing roadmap is reused
    std::mt19937_64 rng(seed);
    const auto t0 = detail::Clock::now();
    const uint32_t dim = (uint32_t)pd_->space->dimension();
    const double res = pd_->space->get_longest_valid_segment_length() * 0.1;
    uint64_t drawn = 0;
    std::vector<uint64_t> off;
    std::vector<uint32_t> idx;
    while (detail::seconds_since(t0) <= timeout && drawn < max_samples && edges_.size() < max_nodes) {
      const uint64_t nb = std::min<uint64_t>(batch, max_samples - drawn);
      std::vector<double> rows(nb * dim);
      for (uint64_t i = 0; i < nb; ++i) {
        State q = pd_->space->sample_uniform(rng);  // prm.rs:122
        std::copy(q.begin(), q.end(), rows.begin() + i * dim);
      }
      drawn += nb;
      const std::vector<uint8_t> ok = vc_->is_valid_batch(rows);  // prm.rs:123
      std::vector<double> keep;
      keep.reserve(rows.size());
      for (uint64_t i = 0; i < nb; ++i)
        if (ok[i]) keep.insert(keep.end(), rows.begin() + i * dim, rows.begin() + (i + 1) * dim);
      const uint64_t nk = keep.size() / dim;
      if (nk == 0) continue;
      const uint64_t first = store_->append(keep.data(), nk);
      off.assign(nk + 1, 0);
      uint64_t cap = std::max<uint64_t>(idx.size(), 64 * nk);
      oxg_prm_stats st{};
      for (;;) {
        idx.resize(cap);
        const int rc = oxg_prm_build_radius(vc_->handle(), store_->h(), connection_radius, res, first,
                                            first + nk, off.data(), idx.data(), cap, &st);
        if (rc == OXG_E_CAPACITY) {
          cap = off[nk];
          continue;
        }
        check(rc);
        break;
      }
      edges_.resize(first + nk);
      for (uint64_t i = 0; i < nk; ++i)
        for (uint64_t e = off[i]; e < off[i + 1]; ++e) {
          edges_[first + i].push_back(idx[e]);                  // new_node.edges (prm.rs:135)
          edges_[idx[e]].push_back((uint32_t)(first + i));       // prm.rs:143-145
        }
    }
  }

  // prm.rs:227-307 through oxg_prm_solve: start validity, start connection, goal scan and the
  // reference-order BFS all run in the library; the roadmap goes in as each node's lower-index
  // neighbours (new_node.edges at push time, prm.rs:135).
  Path solve(double timeout_s) {
    (void)timeout_s;
    if (!pd_ || !vc_) throw PlanningError::uninitialised();
    if (edges_.empty()) throw PlanningError::unsampled();
    const uint32_t dim = (uint32_t)pd_->space->dimension();
    const double res = pd_->space->get_longest_valid_segment_length() * 0.1;
    const State& start = pd_->start_states.at(0);
    const uint64_t n = edges_.size();
    std::vector<uint64_t> off(n + 1, 0);
    std::vector<uint32_t> lower;
    for (uint64_t i = 0; i < n; ++i) {
      for (uint32_t j : edges_[i])
        if (j < i) lower.push_back(j);
      off[i + 1] = lower.size();
    }
    std::vector<uint32_t> nodes(n);
    oxg_prm_query_result r{};
    check(oxg_prm_solve(vc_->handle(), store_->h(), off.data(), lower.data(), connection_radius, res,
                        start.data(), pd_->goal->target.data(), pd_->goal->radius, nodes.data(),
                        (uint32_t)nodes.size(), &r));
    last_query_ = r;
    if (r.status == 2) throw PlanningError::invalid_start();   // prm.rs:243-246
    if (r.status == 3) throw PlanningError::unsampled();       // prm.rs:239-241
    if (r.status != 0) throw PlanningError::no_solution();     // prm.rs:266-268, 304
    Path p;                                                    // reconstruct_path, prm.rs:189-208
    p.push_back(start);
    for (uint32_t t = 0; t < r.path_len; ++t) p.push_back(store_->read(nodes[t], dim));
    return p;
  }
  const oxg_prm_query_result& last_query() const { return last_query_; }

 private:
  std::shared_ptr<ProblemDefinition> pd_;
  std::shared_ptr<GpuValidityChecker> vc_;
  std::unique_ptr<detail::Store> store_;
  std::vector<std::vector<uint32_t>> edges_;
  oxg_prm_query_result last_query_{};
};

}  // namespace cuda
}  // namespace oxmpl
