// host_mirror_test.cpp — the reference's own integration tests, re-expressed against the C++
// host mirror (include/oxmpl_cuda.hpp).  Worlds, planner parameters and acceptance properties
// are those of oxmpl/tests/{rrt,rrt_connect,rrt_star,prm}_rvss_tests.rs (wall fixture,
// :111-188) and of the README example (README.md:97-150): the planner returns Ok within the
// time budget, the path is non-empty, starts at the start state (distance < 1e-9), ends inside
// the goal region, and every segment re-walked at the space's longest valid segment length is
// collision-free (is_path_valid, rrt_rvss_tests.rs:72-107).
//
// Runs on a CUDA device only (liboxg has no CPU fallback); tests/test_gpu_host_mirror.py builds
// and runs it.
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>

#include "../../include/oxmpl_cuda.hpp"
#include "../../include/oxmpl_cuda_so3.hpp"

using namespace oxmpl::cuda;

static int g_failures = 0;
#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);       \
      ++g_failures;                                                         \
    }                                                                       \
  } while (0)

// WallObstacleChecker::is_valid (rrt_rvss_tests.rs:24-36), evaluated on the host as the
// independent re-walk the reference's test does.
struct Wall {
  double x_center = 5.0, y_min = 2.0, y_max = 8.0, thickness = 0.5;
  bool is_valid(const State& s) const {
    const bool in_x = s[0] >= x_center - thickness / 2.0 && s[0] <= x_center + thickness / 2.0;
    const bool in_y = s[1] >= y_min && s[1] <= y_max;
    return !(in_x && in_y);
  }
};

// is_path_valid (rrt_rvss_tests.rs:72-107)
static bool is_path_valid(const Path& path, const RealVectorStateSpace& space,
                          const std::function<bool(const State&)>& valid) {
  for (std::size_t k = 0; k + 1 < path.size(); ++k) {
    const State &a = path[k], &b = path[k + 1];
    if (!valid(a)) return false;
    const double dist = space.distance(a, b);
    const std::size_t num_steps = (std::size_t)std::ceil(dist / space.get_longest_valid_segment_length());
    if (num_steps > 1) {
      State s;
      for (std::size_t i = 1; i <= num_steps; ++i) {
        space.interpolate(a, b, (double)i / (double)num_steps, s);
        if (!valid(s)) return false;
      }
    }
  }
  return path.empty() || valid(path.back());
}

static void check_path(const char* name, const Path& path, const ProblemDefinition& pd,
                       const std::function<bool(const State&)>& valid) {
  EXPECT(!path.empty());
  if (path.empty()) return;
  EXPECT(pd.space->distance(path.front(), pd.start_states[0]) < 1e-9);
  EXPECT(pd.goal->is_satisfied(*pd.space, path.back()));
  EXPECT(is_path_valid(path, *pd.space, valid));
  std::printf("  %s: path of %zu states\n", name, path.size());
}

int main() {
  int ndev = 0;
  if (oxg_device_count(&ndev) != OXG_OK || ndev < 1) {
    std::printf("no CUDA device: %s\n", oxg_last_error());
    return 77;
  }
  // ---- wall world (rrt_rvss_tests.rs:111-143)
  auto space = std::make_shared<RealVectorStateSpace>(2, std::vector<std::pair<double, double>>{{0, 10}, {0, 10}});
  auto goal = std::make_shared<GoalRegion>(GoalRegion{{9.0, 5.0}, 0.5});
  auto pd = std::make_shared<ProblemDefinition>(ProblemDefinition{space, {{1.0, 5.0}}, goal});
  Wall wall;
  GpuValidityChecker::Obstacles ob;
  ob.aabbs = {wall.x_center - wall.thickness / 2.0, wall.y_min, wall.x_center + wall.thickness / 2.0, wall.y_max};
  auto vc = std::make_shared<GpuValidityChecker>(2, ob);
  auto wall_valid = [&](const State& s) { return wall.is_valid(s); };

  // the trait method agrees with the fixture, faces included
  EXPECT(vc->is_valid({1.0, 5.0}) && !vc->is_valid({5.0, 5.0}) && !vc->is_valid({4.75, 2.0}) &&
         !vc->is_valid({5.25, 8.0}) && vc->is_valid({5.26, 8.0}) && vc->is_valid({5.0, 8.0000001}));

  std::printf("RRT(0.5, 0.0) on the wall world\n");  // rrt_rvss_tests.rs:154-158
  {
    RRT planner(0.5, 0.0);
    planner.setup(pd, vc);
    try {
      check_path("rrt", planner.solve(5.0), *pd, wall_valid);
    } catch (const std::exception& e) {
      std::printf("  FAILED: %s\n", e.what());
      ++g_failures;
    }
  }
  std::printf("RRTConnect(0.5, 0.0)\n");  // rrt_connect_rvss_tests.rs
  {
    RRTConnect planner(0.5, 0.0);
    planner.setup(pd, vc);
    try {
      check_path("rrt_connect", planner.solve(5.0), *pd, wall_valid);
    } catch (const std::exception& e) {
      std::printf("  FAILED: %s\n", e.what());
      ++g_failures;
    }
  }
  std::printf("RRTStar(0.5, 0.0, 0.25)\n");  // rrt_star_rvss_tests.rs:148
  {
    RRTStar planner(0.5, 0.0, 0.25);
    planner.setup(pd, vc);
    try {
      check_path("rrt_star", planner.solve(5.0), *pd, wall_valid);
    } catch (const std::exception& e) {
      std::printf("  FAILED: %s\n", e.what());
      ++g_failures;
    }
  }
  std::printf("PRM(5.0, 0.5)\n");  // prm_rvss_tests.rs:154 (count-bounded here: 12000 samples)
  {
    PRM planner(5.0, 0.5);
    planner.setup(pd, vc);
    try {
      bool threw = false;
      try {
        planner.solve(5.0);
      } catch (const PlanningError& e) {
        threw = e.kind == PlanningError::UnsampledStateSpace;  // prm.rs:238-240
      }
      EXPECT(threw);
      planner.construct_roadmap(12000);
      std::printf("  roadmap: %llu nodes\n", (unsigned long long)planner.roadmap_size());
      EXPECT(planner.roadmap_size() > 5000);
      // adjacency is symmetric (prm.rs:143-145)
      const auto& rm = planner.get_roadmap();
      std::size_t asym = 0;
      for (std::size_t i = 0; i < rm.size(); ++i)
        for (uint32_t j : rm[i])
          if (std::find(rm[j].begin(), rm[j].end(), (uint32_t)i) == rm[j].end()) ++asym;
      EXPECT(asym == 0);
      check_path("prm", planner.solve(5.0), *pd, wall_valid);
    } catch (const std::exception& e) {
      std::printf("  FAILED: %s\n", e.what());
      ++g_failures;
    }
  }
  // ---- README example (README.md:97-150): circle obstacle r=2 at the origin, strict boundary
  std::printf("README RRT(0.5, 0.05)\n");
  {
    auto sp2 = std::make_shared<RealVectorStateSpace>(2, std::vector<std::pair<double, double>>{{-10, 10}, {-10, 10}});
    auto g2 = std::make_shared<GoalRegion>(GoalRegion{{5.0, 5.0}, 0.5});
    auto pd2 = std::make_shared<ProblemDefinition>(ProblemDefinition{sp2, {{-5.0, -5.0}}, g2});
    GpuValidityChecker::Obstacles o2;
    o2.spheres = {0.0, 0.0, 2.0};
    auto vc2 = std::make_shared<GpuValidityChecker>(2, o2);
    EXPECT(!vc2->is_valid({2.0, 0.0}) && vc2->is_valid({2.0000000000000004, 0.0}));
    auto circle_valid = [](const State& s) { return std::sqrt(s[0] * s[0] + s[1] * s[1]) > 2.0; };
    RRT planner(0.5, 0.05);
    planner.setup(pd2, vc2);
    try {
      check_path("readme_rrt", planner.solve(5.0), *pd2, circle_valid);
    } catch (const std::exception& e) {
      std::printf("  FAILED: %s\n", e.what());
      ++g_failures;
    }
    // error behaviour: solve before setup (rrt.rs:160-163)
    RRT fresh(0.5, 0.05);
    bool threw = false;
    try {
      fresh.solve(1.0);
    } catch (const PlanningError& e) {
      threw = e.kind == PlanningError::PlannerUninitialised;
    }
    EXPECT(threw);
    // a start walled in by obstacles times out (PlanningError::Timeout)
    GpuValidityChecker::Obstacles o3;
    o3.aabbs = {-6.0, -6.0, -4.0, -5.2, -6.0, -4.8, -4.0, -4.0, -6.0, -6.0, -5.2, -4.0, -4.8, -6.0, -4.0, -4.0};
    auto vc3 = std::make_shared<GpuValidityChecker>(2, o3);
    RRT boxed(0.5, 0.05);
    boxed.chunk = 2048;
    boxed.setup(pd2, vc3);
    threw = false;
    try {
      boxed.solve(0.3);
    } catch (const PlanningError& e) {
      threw = e.kind == PlanningError::Timeout;
    }
    EXPECT(threw);
  }
  // ---- SO(3): oxmpl/tests/rrt_so3ss_tests.rs:146-229 — rotate from +90 deg to -90 deg about y
  // around a forbidden cone of 44.9 deg at the identity; goal region of 10 deg; RRT(0.5, 0.0).
  std::printf("SO3 RRT(0.5, 0.0)\n");
  {
    const double kPi = 3.14159265358979323846;
    auto axis_angle = [](double ax, double ay, double az, double angle) {
      const double s = std::sin(angle * 0.5), c = std::cos(angle * 0.5);
      return SO3State{ax * s, ay * s, az * s, c};
    };
    auto space = std::make_shared<SO3StateSpace>();
    const SO3State start = axis_angle(0.0, 1.0, 0.0, kPi / 2.0);
    const SO3State target = axis_angle(0.0, 1.0, 0.0, -kPi / 2.0);
    const double goal_radius = 10.0 * kPi / 180.0, cone = 44.9 * kPi / 180.0;
    auto pd = std::make_shared<SO3ProblemDefinition>();
    pd->space = space;
    pd->start_states = {start};
    pd->goal_is_satisfied = [=](const SO3State& s) { return space->distance(s, target) <= goal_radius; };
    pd->sample_goal = [=](std::mt19937_64& rng) {
      SO3StateSpace around(target, goal_radius);
      return around.sample_uniform(rng);
    };
    auto vc = std::make_shared<GpuConeChecker>(
        std::vector<std::pair<SO3State, double>>{{SO3State::identity(), cone}});
    auto host_valid = [&](const SO3State& s) { return space->distance(SO3State::identity(), s) > cone; };
    EXPECT(vc->is_valid(start) && vc->is_valid(target) && !vc->is_valid(SO3State::identity()));
    const SO3State rim = axis_angle(0.0, 1.0, 0.0, 2.0 * cone);         // on the rim of the cone
    EXPECT(vc->is_valid(rim) == host_valid(rim));
    SO3RRT planner(0.5, 0.0);
    planner.setup(pd, vc);
    try {
      const SO3Path path = planner.solve(20.0);
      EXPECT(path.size() > 1);
      EXPECT(space->distance(path.front(), start) < 1e-9);
      EXPECT(pd->goal_is_satisfied(path.back()));
      bool ok = true;  // is_path_valid, rrt_so3ss_tests.rs:106-143
      for (std::size_t i = 0; i + 1 < path.size(); ++i) {
        ok = ok && host_valid(path[i]) && host_valid(path[i + 1]);
        const double dist = space->distance(path[i], path[i + 1]);
        const std::size_t n = (std::size_t)std::ceil(dist / space->get_longest_valid_segment_length());
        for (std::size_t j = 1; n > 1 && j <= n; ++j)
          ok = ok && host_valid(space->interpolate(path[i], path[i + 1], (double)j / (double)n));
      }
      EXPECT(ok);
      std::printf("  so3_rrt: %zu states, tree %llu\n", path.size(), (unsigned long long)planner.tree_size());
    } catch (const std::exception& e) {
      std::printf("  FAILED: %s\n", e.what());
      ++g_failures;
    }
    // the same problem with the goal declared as a ball: the library's chunked loop, same draws
    // -> same tree as the per-iteration loop above
    {
      auto pd2 = std::make_shared<SO3ProblemDefinition>(*pd);
      pd2->ball_goal = true;
      pd2->goal_target = target;
      pd2->goal_radius = goal_radius;
      SO3RRT chunked(0.5, 0.0);
      chunked.setup(pd2, vc);
      try {
        const SO3Path path = chunked.solve(20.0);
        EXPECT(path.size() > 1 && pd->goal_is_satisfied(path.back()));
        EXPECT(chunked.tree_size() == planner.tree_size());
        EXPECT(chunked.chunks() < chunked.iterations());
        std::printf("  so3_rrt chunked: %zu states, tree %llu, %llu iterations in %llu chunks\n", path.size(),
                    (unsigned long long)chunked.tree_size(), (unsigned long long)chunked.iterations(),
                    (unsigned long long)chunked.chunks());
      } catch (const std::exception& e) {
        std::printf("  FAILED: %s\n", e.what());
        ++g_failures;
      }
    }
    SO3RRT fresh(0.5, 0.0);
    bool threw = false;
    try {
      fresh.solve(1.0);
    } catch (const PlanningError& e) {
      threw = e.kind == PlanningError::PlannerUninitialised;
    }
    EXPECT(threw);
  }
  std::printf(g_failures ? "FAILED (%d)\n" : "ALL OK\n", g_failures);
  return g_failures ? 1 : 0;
}
