// oxmpl_cuda_so3.hpp — SO(3) part of the C++ host mirror (include/oxg_so3.h underneath).
//
//   oxmpl::cuda::SO3State        oxmpl/src/base/states/so3_state.rs (x, y, z, w)
//   oxmpl::cuda::SO3StateSpace   oxmpl/src/base/spaces/so3_state_space.rs: distance :102-112,
//                                interpolate :114-159, get_maximum_extent :62-64, sample_uniform :196-228
//   oxmpl::cuda::GpuConeChecker  impl StateValidityChecker<SO3State>: the ForbiddenConeChecker fixture
//                                of oxmpl/tests/rrt_so3ss_tests.rs:46-56, any number of cones
//   oxmpl::cuda::SO3RRT          RRT::solve (rrt.rs:158-227) over SO3, unfused: oxg_so3_nn1 ->
//                                steer on the host (SLERP needs libm) -> oxg_so3_check_motion -> append
//
// Host arithmetic is the reference's, operation by operation (compile with -ffp-contract=off);
// acos / sin are the platform libm's, which is what Rust's f64::acos / f64::sin call.
#pragma once

#include <array>
#include <functional>

#include "oxg_so3.h"
#include "oxmpl_cuda.hpp"

namespace oxmpl {
namespace cuda {

struct SO3State {
  double x = 0.0, y = 0.0, z = 0.0, w = 1.0;
  static SO3State identity() { return SO3State{0.0, 0.0, 0.0, 1.0}; }
  std::array<double, 4> values() const { return {x, y, z, w}; }
};
using SO3Path = std::vector<SO3State>;

class SO3StateSpace {
 public:
  SO3StateSpace() : center_(SO3State::identity()), max_angle_(kPi) {}
  SO3StateSpace(const SO3State& center, double max_angle) : center_(center) {
    if (max_angle < 0.0) throw std::invalid_argument("InvalidAngularDistance");
    max_angle_ = std::min(max_angle, kPi);
  }

  double distance(const SO3State& a, const SO3State& b) const {
    const double abs_dot = std::fabs(a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w);
    return abs_dot > 1.0 - 1e-9 ? 0.0 : std::acos(abs_dot);
  }

  SO3State interpolate(const SO3State& from, const SO3State& to, double t) const {
    double dot = from.x * to.x + from.y * to.y + from.z * to.z + from.w * to.w;
    const double sign = dot < 0.0 ? -1.0 : 1.0;
    dot *= sign;
    SO3State o;
    if (dot > 0.9995) {
      o.x = from.x + t * (to.x * sign - from.x);
      o.y = from.y + t * (to.y * sign - from.y);
      o.z = from.z + t * (to.z * sign - from.z);
      o.w = from.w + t * (to.w * sign - from.w);
      const double norm = std::sqrt(o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w);
      o.x /= norm;
      o.y /= norm;
      o.z /= norm;
      o.w /= norm;
    } else {
      const double theta = std::acos(dot);
      const double sin_theta = std::sin(theta);
      const double s0 = std::sin((1.0 - t) * theta) / sin_theta;
      const double s1 = std::sin(t * theta) / sin_theta * sign;
      o.x = from.x * s0 + to.x * s1;
      o.y = from.y * s0 + to.y * s1;
      o.z = from.z * s0 + to.z * s1;
      o.w = from.w * s0 + to.w * s1;
    }
    return o;
  }

  double get_maximum_extent() const { return 0.5 * kPi; }
  void set_longest_valid_segment_fraction(double f) {
    fraction_ = (f > 0.0 && f <= 1.0) ? f : (f <= 0.0 ? 0.0 : 1.0);
  }
  double get_longest_valid_segment_length() const { return get_maximum_extent() * fraction_; }
  bool satisfies_bounds(const SO3State& s) const { return distance(center_, s) <= max_angle_; }

  template <class Rng>
  SO3State sample_uniform(Rng& rng) const {  // rejection sampling in the unit 4-ball
    if (max_angle_ < 1e-9) return center_;
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    for (;;) {
      const double x = u(rng), y = u(rng), z = u(rng), w = u(rng);
      const double norm_sq = x * x + y * y + z * z + w * w;
      if (norm_sq > 1e-9 && norm_sq < 1.0) {
        const double norm = std::sqrt(norm_sq);
        const SO3State q{x / norm, y / norm, z / norm, w / norm};
        if (distance(center_, q) <= max_angle_) return q;
      }
    }
  }

 private:
  static constexpr double kPi = 3.14159265358979323846;
  SO3State center_;
  double max_angle_;
  double fraction_ = 0.05;
};

// Forbidden cones (centre, angle): valid iff distance(centre, state) > angle for every cone.
class GpuConeChecker {
 public:
  explicit GpuConeChecker(const std::vector<std::pair<SO3State, double>>& cones) {
    std::vector<double> rows;
    for (const auto& c : cones) {
      const auto v = c.first.values();
      rows.insert(rows.end(), v.begin(), v.end());
      rows.push_back(c.second);
    }
    check(oxg_so3_cones_create(rows.data(), cones.size(), &h_));
  }
  ~GpuConeChecker() { oxg_so3_cones_destroy(h_); }
  GpuConeChecker(const GpuConeChecker&) = delete;
  GpuConeChecker& operator=(const GpuConeChecker&) = delete;

  bool is_valid(const SO3State& s) const {
    const auto v = s.values();
    uint8_t ok = 0;
    check(oxg_so3_is_valid(h_, v.data(), 1, &ok));
    return ok != 0;
  }
  bool check_motion(const SO3State& from, const SO3State& to, double resolution) const {
    const auto a = from.values(), b = to.values();
    uint8_t ok = 0;
    check(oxg_so3_check_motion(h_, a.data(), b.data(), 1, resolution, &ok, nullptr));
    return ok != 0;
  }
  oxg_so3_cones handle() const { return h_; }

 private:
  oxg_so3_cones h_ = nullptr;
};

struct SO3ProblemDefinition {
  std::shared_ptr<SO3StateSpace> space;
  std::vector<SO3State> start_states;
  std::function<bool(const SO3State&)> goal_is_satisfied;   // Goal::is_satisfied
  std::function<SO3State(std::mt19937_64&)> sample_goal;    // GoalSampleableRegion::sample_goal
  // Optional: the goal as the ball distance(state, goal_target) <= goal_radius (SO3GoalRegion,
  // oxmpl/tests/rrt_so3ss_tests.rs:58-68).  When set, SO3RRT::solve runs the library's chunked
  // loop (oxg_so3_rrt_solve) instead of one call per step of every iteration.
  bool ball_goal = false;
  SO3State goal_target;
  double goal_radius = 0.0;
};

class SO3RRT {
 public:
  double max_distance, goal_bias;
  uint64_t seed = 0x9E3779B97F4A7C15ull;

  SO3RRT(double max_distance_, double goal_bias_) : max_distance(max_distance_), goal_bias(goal_bias_) {}

  void setup(std::shared_ptr<SO3ProblemDefinition> pd, std::shared_ptr<GpuConeChecker> vc) {
    pd_ = std::move(pd);
    vc_ = std::move(vc);
    const oxg_space_desc sp{4, 0, 1.0, 1.0};
    tree_.reset(new detail::Store(sp, 1024));
    nodes_.assign(1, pd_->start_states.at(0));
    parents_.assign(1, -1);
    const auto v = nodes_[0].values();
    tree_->append(v.data(), 1);
  }

  SO3Path solve(double timeout_s) {  // rrt.rs:158-227
    if (!pd_ || !vc_) throw PlanningError::uninitialised();
    std::mt19937_64 rng(seed);
    std::bernoulli_distribution coin(goal_bias);
    const auto t0 = detail::Clock::now();
    const double res = pd_->space->get_longest_valid_segment_length() * 0.1;  // rrt.rs:97
    if (pd_->ball_goal) return solve_chunked(timeout_s, res, rng, coin);
    while (detail::seconds_since(t0) <= timeout_s) {
      const SO3State q_rand = coin(rng) ? pd_->sample_goal(rng) : pd_->space->sample_uniform(rng);
      const auto q = q_rand.values();
      uint32_t near = 0;
      double min_dist = 0.0;
      check(oxg_so3_nn1(tree_->h(), q.data(), 1, &near, &min_dist));
      const SO3State& q_near = nodes_[near];
      const SO3State q_new = min_dist > max_distance
                                 ? pd_->space->interpolate(q_near, q_rand, max_distance / min_dist)
                                 : q_rand;
      if (!vc_->check_motion(q_near, q_new, res)) continue;
      const auto v = q_new.values();
      tree_->append(v.data(), 1);
      nodes_.push_back(q_new);
      parents_.push_back((int64_t)near);
      if (pd_->goal_is_satisfied(q_new)) {
        SO3Path p;
        for (int64_t i = (int64_t)nodes_.size() - 1; i >= 0; i = parents_[(std::size_t)i])
          p.push_back(nodes_[(std::size_t)i]);
        std::reverse(p.begin(), p.end());
        return p;
      }
    }
    throw PlanningError::timeout();
  }
  uint64_t tree_size() const { return nodes_.size(); }
  uint64_t chunks() const { return chunks_; }      // device round trips of the last chunked solve
  uint64_t iterations() const { return iters_; }

 private:
  // rrt.rs:158-227 on oxg_so3_rrt_solve: the per-iteration draw (rrt.rs:177-184) does not depend
  // on the tree, so `kStream` iterations are drawn ahead and handed to the library at once.
  SO3Path solve_chunked(double timeout_s, double res, std::mt19937_64& rng,
                        std::bernoulli_distribution& coin) {
    constexpr uint64_t kStream = 4096, kMaxNodes = 1u << 20;
    const auto t0 = detail::Clock::now();
    std::vector<int64_t> parents(kMaxNodes, -1);
    std::copy(parents_.begin(), parents_.end(), parents.begin());
    oxg_plan_params prm{};
    prm.max_distance = max_distance;
    prm.resolution = res;
    prm.goal_radius = pd_->goal_radius;
    prm.max_nodes = kMaxNodes;
    prm.stop_at_goal = 1;
    const auto target = pd_->goal_target.values();
    std::vector<double> stream(kStream * 4);
    chunks_ = iters_ = 0;
    for (;;) {
      const double left = timeout_s - detail::seconds_since(t0);
      if (left <= 0.0) throw PlanningError::timeout();
      prm.timeout_s = left;
      for (uint64_t i = 0; i < kStream; ++i) {
        const SO3State q = coin(rng) ? pd_->sample_goal(rng) : pd_->space->sample_uniform(rng);
        const auto v = q.values();
        std::copy(v.begin(), v.end(), stream.begin() + i * 4);
      }
      oxg_plan_result r{};
      check(oxg_so3_rrt_solve(tree_->h(), vc_->handle(), &prm, target.data(), stream.data(), kStream,
                              parents.data(), &r));
      chunks_ += (uint64_t)r.ctas;
      iters_ += r.iters;
      std::vector<double> rows((r.n_nodes[0] - nodes_.size()) * 4);
      if (!rows.empty())
        check(oxg_store_read(tree_->h(), nodes_.size(), rows.size() / 4, rows.data()));
      for (std::size_t i = 0; i < rows.size(); i += 4)
        nodes_.push_back(SO3State{rows[i], rows[i + 1], rows[i + 2], rows[i + 3]});
      parents_.assign(parents.begin(), parents.begin() + (std::ptrdiff_t)nodes_.size());
      if (r.solved) {
        SO3Path p;
        for (int64_t i = r.goal_node[0]; i >= 0; i = parents_[(std::size_t)i])
          p.push_back(nodes_[(std::size_t)i]);
        std::reverse(p.begin(), p.end());
        return p;
      }
      if (r.status == 2 || r.status == 3) throw PlanningError::timeout();
    }
  }

  uint64_t chunks_ = 0, iters_ = 0;
  std::shared_ptr<SO3ProblemDefinition> pd_;
  std::shared_ptr<GpuConeChecker> vc_;
  std::unique_ptr<detail::Store> tree_;
  std::vector<SO3State> nodes_;
  std::vector<int64_t> parents_;
};

}  // namespace cuda
}  // namespace oxmpl
