// oxg_planner.cu — the sequential planners' loops fused into one persistent kernel each.
//
// Replaces the bodies of RRT::solve (oxmpl/src/geometric/planners/rrt.rs:158-227),
// RRTConnect::solve + extend (rrt_connect.rs:121-159,227-309) and RRTStar::solve
// (rrt_star.rs:171-290 with find_neighbours :121-131, cost :107-115) for a pre-generated
// sample stream (the reference's RNG draws do not depend on the tree, so the host can draw
// the stream with the reference's own call sequence).
//
// Why one kernel.  One iteration is a 1-NN scan (A9), a steer (A13), a motion check (A7) and
// an append (A12); issued as separate launches with a D2H sync per iteration it is
// launch-latency-bound.  Here a cooperative grid (one CTA per SM at most)
// runs the whole loop: every CTA executes the sequential logic redundantly and
// deterministically (identical inputs -> identical decisions, so control flow is uniform
// without communication), while the O(N) scans and the steps x obstacles validity work are
// partitioned over the grid and joined by grid barriers:
//
//   phase A  scan this CTA's slice of the tree for the nearest state        -> cta_best[]
//   ---- grid barrier
//   phase B  reduce cta_best, steer, motion-check tasks (state x obstacle-chunk per warp),
//            RRT*: radius scan around q_new with ordered per-CTA hit lists
//   ---- grid barrier
//   phase C  append / goal test;  RRT*: choose-parent candidates S0 = {j : cost_j < min_cost}
//            motion-checked in parallel (a superset of what the lazy sequential loop checks;
//            check_motion is pure), ---- barrier, sequential replay of the selection, append,
//            rewire candidates checked in parallel, ---- barrier, apply.
//
// The boolean of check_motion does not depend on evaluation order, distances are computed
// with the reference's operation order and argument order, selections replay the reference's
// strict-< / lowest-index rules, so trees, parents and costs are bit-identical to the CPU loop.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include <chrono>

#include "oxg_internal.cuh"
#include "oxg_math.cuh"

namespace cg = cooperative_groups;

namespace oxg {

constexpr int PL_THREADS = 512;
constexpr int PL_WARPS = PL_THREADS / 32;
constexpr int NBR_CAP = OXG_RRT_STAR_MAX_NEIGHBOURS;  // neighbours of one RRT* iteration (OXG_E_CAPACITY beyond; include/oxg.h)
constexpr unsigned long long MAX_STEPS = 1ull << 24;  // steps of one motion (reference: unbounded)

enum { MODE_RRT = 0, MODE_CONNECT = 1, MODE_STAR = 2 };
enum {
  ST_SOLVED = 0,
  ST_SAMPLES = 1,   // sample stream exhausted
  ST_FULL = 2,      // a tree reached max_nodes
  ST_TIMEOUT = 3,
  ST_E_NEIGHBOURS = -1,  // more than NBR_CAP neighbours inside search_radius
  ST_E_STEPS = -2        // a motion needs more than MAX_STEPS interpolation steps
};

struct TreeView {
  double* data;
  uint64_t pitch;
  long long* parent;
  int cg;  // 1: several CTAs append to the tree, read it with ld.global.cg (L1 is not coherent
           // across SMs); 0: a single CTA runs the loop and its own L1 serves the reads
};
template <class V>
__device__ __forceinline__ V ld_tree(const TreeView& T, const V* p) {
  return T.cg ? __ldcg(p) : *p;
}

struct PlanCtrl {
  unsigned long long n[2];
  unsigned long long iters;
  unsigned long long mc_ref;   // check_motion calls the sequential reference would make
  unsigned long long mc_exec;  // check_motion evaluations executed here (superset)
  long long node[2];
  int solved;
  int status;
  int stop;
  int pad;
  unsigned long long phase_ns[8];  // time attribution (CTA 0, thread 0), see oxg_plan_result
};

struct PlanParams {
  TreeView tree[2];
  double* cost;
  uint32_t n_init[2];
  uint32_t max_nodes;
  uint32_t n_real, n_so2;
  double w_real, w_so2;
  WorldView w;
  const double* samples;
  unsigned long long n_samples;
  double max_distance, search_radius, resolution, goal_radius;
  const double* goal;  // device pointer or nullptr
  int stop_at_goal;
  unsigned long long timeout_ns;
  // scratch (device)
  double* cta_best_dist;
  uint32_t* cta_best_idx;
  uint32_t* hit_ext;  // [2]
  uint32_t* hit_cp;   // [NBR_CAP]
  uint32_t* hit_rw;   // [NBR_CAP]
  uint32_t* nbr_cnt;  // [G]
  uint32_t* nbr_idx;  // [G][nbr_capc]
  double* nbr_dist;
  uint32_t nbr_capc;
  PlanCtrl* ctrl;
  unsigned long long* barrier;  // grid barrier arrival counter (zeroed per launch)
};

struct PlanShared {
  double q_rand[kMaxDim], q_near[kMaxDim], q_new[kMaxDim], pend[kMaxDim];
  double red_d[PL_WARPS];
  double bc_d[4];
  uint32_t red_i[PL_WARPS];
  uint32_t warp_cnt[PL_WARPS];
  uint32_t bc_u[8];
  int bc_i[4];
  // RRT* neighbour / edge lists
  uint32_t nb_idx[NBR_CAP];
  double nb_dist[NBR_CAP];
  double nb_c[NBR_CAP];      // candidate cost of each neighbour (choose-parent, then rewire)
  uint32_t nb_ns[NBR_CAP];   // steps of the motion to/from each neighbour
  uint32_t e_nb[NBR_CAP];    // edge -> neighbour slot
  uint32_t e_pre[NBR_CAP + 1];
  uint32_t cta_off[160];
  // single-CTA runs keep the hit flags here instead of in global memory
  uint32_t hit_ext[2];
  uint32_t hit_cp[NBR_CAP];
  uint32_t hit_rw[NBR_CAP];
};

// ------------------------------------------------------------------------------- arithmetic
// Accessors: a state in shared memory, or a tree / goal state in global memory.  With several
// CTAs global states are read with ld.global.cg: other CTAs append to the tree during the kernel,
// and L1 is not coherent across SMs.
struct SmState {
  const double* p;
  __device__ __forceinline__ double operator[](uint32_t d) const { return p[d]; }
};
struct GlState {
  const double* p;
  uint64_t stride;
  bool cg;
  __device__ __forceinline__ double operator[](uint32_t d) const {
    return cg ? __ldcg(p + d * stride) : p[d * stride];
  }
};
__device__ __forceinline__ SmState sm(const double* p) { return SmState{p}; }
__device__ __forceinline__ GlState node(const TreeView& T, uint32_t i) {
  return GlState{T.data + i, T.pitch, T.cg != 0};
}

// distance(a, b) in the reference's operation and argument order (real_vector_state_space.rs:
// 137-155, so2_state_space.rs:97-101; compound rule of include/oxg.h).
template <class A, class B>
__device__ __forceinline__ double dist_ab(const PlanParams& P, A a, B b) {
  double acc = 0.0;
  bool have = false;
  if (P.n_real > 0) {
    double s = 0.0;
    for (uint32_t d = 0; d < P.n_real; ++d) {
      const double df = __dsub_rn(a[d], b[d]);
      const double m = __dmul_rn(df, df);
      s = (d == 0) ? m : __dadd_rn(s, m);
    }
    const double r = sqrt(s);
    if (P.n_so2 == 0) return r;
    acc = __dmul_rn(P.w_real, r);
    have = true;
  } else if (P.n_so2 == 1) {
    return so2_distance_dev(a[0], b[0]);
  }
  for (uint32_t c = 0; c < P.n_so2; ++c) {
    const uint32_t d = P.n_real + c;
    const double term = __dmul_rn(P.w_so2, so2_distance_dev(a[d], b[d]));
    acc = have ? __dadd_rn(acc, term) : term;
    have = true;
  }
  return acc;
}

__device__ __forceinline__ double interp_coord(const PlanParams& P, uint32_t d, double f, double t,
                                               double u) {
  if (d < P.n_real) return __dadd_rn(f, __dmul_rn(__dsub_rn(t, f), u));
  return so2_interpolate_dev(f, t, u);
}

__device__ __forceinline__ uint32_t world_chunks(const WorldView& w) {
  return (w.n_spheres + 31) / 32 + (w.n_aabbs + 31) / 32 + (w.n_capsules + 31) / 32 +
         (w.n_so2 + 31) / 32;
}

// Does primitive `lane` of obstacle chunk c invalidate state x (shared memory, dim values)?
__device__ __forceinline__ bool chunk_hit(const PlanParams& P, uint32_t c, uint32_t lane,
                                          const double* x) {
  const WorldView& w = P.w;
  const int dim = (int)w.dim;
  uint32_t cs = (w.n_spheres + 31) / 32, ca = (w.n_aabbs + 31) / 32, cc = (w.n_capsules + 31) / 32;
  if (c < cs) {
    const uint32_t p = c * 32 + lane;
    return p < w.n_spheres && sphere_hit_dev<0>(x, w.spheres + (size_t)p * w.sphere_stride, dim);
  }
  c -= cs;
  if (c < ca) {
    const uint32_t p = c * 32 + lane;
    return p < w.n_aabbs && aabb_hit_dev<0>(x, w.aabbs + (size_t)p * w.aabb_stride, dim);
  }
  c -= ca;
  if (c < cc) {
    const uint32_t p = c * 32 + lane;
    return p < w.n_capsules && capsule_hit_dev<0>(x, w.capsules + (size_t)p * w.capsule_stride, dim);
  }
  c -= cc;
  const uint32_t p = c * 32 + lane;
  if (p >= w.n_so2) return false;
  const double v = so2_normalise_dev(x[P.n_real + w.so2_comp[p]]);
  return v >= w.so2_lohi[2 * p] && v <= w.so2_lohi[2 * p + 1];
}

// Grid lookup (oxg_world_s::d_grid_*): the cell of a state in the first grid_dims coordinates,
// clamped to the grid — the host registered every obstacle in all cells within one cell of its
// bounding box with the same formula, so a state that an obstacle invalidates always finds it.
__device__ __forceinline__ uint32_t grid_cell(const WorldView& w, const double* x) {
  uint32_t cell = 0, mul = 1;
  for (uint32_t d = 0; d < w.grid_dims; ++d) {
    double c = floor((x[d] - w.grid_lo[d]) * w.grid_inv[d]);
    if (!(c == c)) c = 0.0;
    c = fmin(fmax(c, 0.0), (double)(w.grid_n[d] - 1));
    cell += (uint32_t)c * mul;
    mul *= w.grid_n[d];
  }
  return cell;
}

// Exact test of obstacle `id` (spheres, then aabbs, then capsules) against state x.
__device__ __forceinline__ bool obstacle_hit(const WorldView& w, uint32_t id, const double* x) {
  const int dim = (int)w.dim;
  if (id < w.n_spheres) return sphere_hit_dev<0>(x, w.spheres + (size_t)id * w.sphere_stride, dim);
  id -= w.n_spheres;
  if (id < w.n_aabbs) return aabb_hit_dev<0>(x, w.aabbs + (size_t)id * w.aabb_stride, dim);
  id -= w.n_aabbs;
  return capsule_hit_dev<0>(x, w.capsules + (size_t)id * w.capsule_stride, dim);
}

// num_steps of check_motion (rrt.rs:95-97); 0/1 both mean "check `to` only".
__device__ __forceinline__ unsigned long long motion_steps(double dist, double resolution) {
  return ceil_as_usize_dev(__ddiv_rn(dist, resolution));
}

// ------------------------------------------------------------------------------- barriers
// Grid barrier of the cooperative launch (all CTAs are co-resident).  One monotonically
// increasing arrival counter: barrier number g completes when it reaches g * gridDim.x.
// Every datum that crosses CTAs in this kernel is read with ld.global.cg / volatile (L2), so
// the barrier needs no L1 invalidation — which also keeps the obstacle set warm in L1 across
// the several barriers of an iteration (cooperative_groups' grid.sync() flushes L1 each time).
template <bool MULTI>
__device__ __forceinline__ void grid_barrier(unsigned long long* counter, unsigned long long& gen) {
  if (MULTI) {
    __syncthreads();
    if (threadIdx.x == 0) {
      ++gen;
      __threadfence();
      atomicAdd(counter, 1ull);
      const unsigned long long target = gen * gridDim.x;
      while (*reinterpret_cast<volatile unsigned long long*>(counter) < target) {
      }
      __threadfence();
    }
    __syncthreads();
  } else {
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------- building blocks
struct ScanBest {
  double d;
  uint32_t i;
};
__device__ __forceinline__ bool better(double d, uint32_t i, double bd, uint32_t bi) {
  return d < bd || (d == bd && i < bi);
}

// Slice of [0, n) owned by this CTA (contiguous, so per-CTA hit lists concatenate in order).
__device__ __forceinline__ void cta_slice(uint32_t n, uint32_t* lo, uint32_t* hi) {
  const uint32_t G = gridDim.x;
  uint32_t chunk = (n + G - 1) / G;
  chunk = (chunk + 31u) & ~31u;
  const uint64_t a = (uint64_t)blockIdx.x * chunk;
  const uint64_t b = a + chunk;
  *lo = (uint32_t)(a < n ? a : n);
  *hi = (uint32_t)(b < n ? b : n);
}

// Phase A: nearest stored state to `target` (shared memory) over this CTA's slice, as
// distance(stored, target) with strict-< / lowest-index selection (rrt.rs:187-196).
// `pend_idx` (if < n) is the index whose state is still only in sh.pend (appended after the last
// barrier).
template <bool MULTI>
__device__ void phase_scan(const PlanParams& P, PlanShared& sh, const TreeView& T, uint32_t n,
                           const double* target, uint32_t pend_idx) {
  uint32_t lo, hi;
  cta_slice(n, &lo, &hi);
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  double bd = inf;
  uint32_t bi = OXG_INVALID_INDEX;
#pragma unroll 4
  for (uint32_t i = lo + threadIdx.x; i < hi; i += PL_THREADS) {
    const double d = (i == pend_idx) ? dist_ab(P, sm(sh.pend), sm(target))
                                     : dist_ab(P, node(T, i), sm(target));
    if (better(d, i, bd, bi)) {
      bd = d;
      bi = i;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_down_sync(0xFFFFFFFFu, bd, o);
    const uint32_t oi = __shfl_down_sync(0xFFFFFFFFu, bi, o);
    if (better(od, oi, bd, bi)) {
      bd = od;
      bi = oi;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh.red_d[warp] = bd;
    sh.red_i[warp] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int wv = 1; wv < PL_WARPS; ++wv)
      if (better(sh.red_d[wv], sh.red_i[wv], bd, bi)) {
        bd = sh.red_d[wv];
        bi = sh.red_i[wv];
      }
    if (MULTI) {
      P.cta_best_dist[blockIdx.x] = bd;
      P.cta_best_idx[blockIdx.x] = bi;
    } else {
      sh.bc_d[0] = bd;
      sh.bc_u[0] = bi;
    }
  }
}

// Phase B head: reduce the per-CTA bests, load q_near, steer (rrt.rs:200-208).  Leaves
// sh.q_near, sh.q_new; returns (near, min_dist, reached).
template <bool MULTI>
__device__ void phase_steer(const PlanParams& P, PlanShared& sh, const TreeView& T,
                            const double* target, uint32_t pend_idx, uint32_t* near_out,
                            double* min_dist_out, bool* reached_out) {
  const int dim = (int)(P.n_real + P.n_so2);
  if (!MULTI) {
    if (threadIdx.x == 0 && sh.bc_u[0] == OXG_INVALID_INDEX) {
      sh.bc_u[0] = 0;  // every distance was NaN: the reference keeps index 0
      sh.bc_d[0] = __longlong_as_double(0x7FF8000000000000ll);
    }
  } else if (gridDim.x > 32) {
    // one parallel round of L2 loads (a thread per CTA), then warp 0 reduces from shared memory
    for (uint32_t c = threadIdx.x; c < gridDim.x; c += PL_THREADS) {
      sh.nb_dist[c] = __ldcg(&P.cta_best_dist[c]);
      sh.nb_idx[c] = __ldcg(&P.cta_best_idx[c]);
    }
    __syncthreads();
  }
  if (MULTI && threadIdx.x < 32) {
    const double inf = __longlong_as_double(0x7FF0000000000000ll);
    double bd = inf;
    uint32_t bi = OXG_INVALID_INDEX;
    const bool direct = gridDim.x <= 32;  // a single round: read L2 directly, no staging barrier
    for (uint32_t c = threadIdx.x; c < gridDim.x; c += 32) {
      const double d = direct ? __ldcg(&P.cta_best_dist[c]) : sh.nb_dist[c];
      const uint32_t i = direct ? __ldcg(&P.cta_best_idx[c]) : sh.nb_idx[c];
      if (better(d, i, bd, bi)) {
        bd = d;
        bi = i;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_down_sync(0xFFFFFFFFu, bd, o);
      const uint32_t oi = __shfl_down_sync(0xFFFFFFFFu, bi, o);
      if (better(od, oi, bd, bi)) {
        bd = od;
        bi = oi;
      }
    }
    if (threadIdx.x == 0) {
      if (bi == OXG_INVALID_INDEX) {  // every distance was NaN: the reference keeps index 0
        bi = 0;
        bd = __longlong_as_double(0x7FF8000000000000ll);
      }
      sh.bc_d[0] = bd;
      sh.bc_u[0] = bi;
    }
  }
  __syncthreads();
  const uint32_t near = sh.bc_u[0];
  const double min_dist = sh.bc_d[0];
  if ((int)threadIdx.x < dim)
    sh.q_near[threadIdx.x] = (near == pend_idx) ? sh.pend[threadIdx.x]
                                                : ld_tree(T, &T.data[(uint64_t)threadIdx.x * T.pitch + near]);
  __syncthreads();
  const bool far = min_dist > P.max_distance;
  if ((int)threadIdx.x < dim) {
    double v;
    if (far) {
      const double t = __ddiv_rn(P.max_distance, min_dist);
      v = interp_coord(P, threadIdx.x, sh.q_near[threadIdx.x], target[threadIdx.x], t);
    } else {
      v = target[threadIdx.x];
    }
    sh.q_new[threadIdx.x] = v;
  }
  __syncthreads();
  *near_out = near;
  *min_dist_out = min_dist;
  *reached_out = !far;
}

// Motion-check tasks.  Edge e (0 <= e < E) joins the fixed state `fixed` (shared memory) and
// either `other_sm` (shared memory; E must be 1) or tree state nb_idx[e_nb[e]]; `fixed_is_to`
// selects the direction (from = other, to = fixed) or (from = fixed, to = other).  Edge e has
// ns = steps[e] reference steps and states pre[e]..pre[e+1]-1.
//
// Two stages per batch of states: (1) the CTA interpolates every state of the batch into shared
// memory once (thread per coordinate; reference operation order), (2) each warp owns one chunk
// of 32 obstacles (lane = obstacle) and sweeps the states of its group — no per-task search,
// division or global load is left in the sweep.  A hit sets hit[e]; edges this warp already
// knows to be hit are skipped, the global flag is polled only every 16th state.
constexpr int PL_STATE_BYTES = 64 * 1024;

__device__ void motion_tasks(const PlanParams& P, PlanShared& sh, double* st_buf, const TreeView& T,
                             uint32_t E, const double* fixed, bool fixed_is_to,
                             const double* other_sm, const uint32_t* steps, const uint32_t* pre,
                             const uint32_t* e_nb, uint32_t* hit) {
  const uint32_t n_oc = world_chunks(P.w);
  if (n_oc == 0 || E == 0) return;
  const uint32_t NS = pre[E];
  const uint32_t dim = P.n_real + P.n_so2;
  const uint32_t cap = (uint32_t)PL_STATE_BYTES / (dim * 8u + 4u);
  uint32_t* st_edge = reinterpret_cast<uint32_t*>(st_buf + (size_t)cap * dim);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t gw = blockIdx.x * PL_WARPS + warp, GW = gridDim.x * PL_WARPS;
  uint32_t c0, cstep, g0, groups;
  if (GW >= n_oc) {
    groups = GW / n_oc;
    c0 = gw % n_oc;
    cstep = n_oc;  // one chunk per warp
    g0 = gw / n_oc;
    if (g0 >= groups) c0 = n_oc;  // idle warp (still takes part in the CTA barriers)
  } else {
    c0 = gw;
    cstep = GW;
    g0 = 0;
    groups = 1;
  }
  for (uint32_t sb = 0; sb < NS; sb += cap) {
    const uint32_t nb = min(cap, NS - sb);
    __syncthreads();  // the previous batch (or the caller's use of st_buf) is finished
    // ---- stage 1: interpolate states [sb, sb+nb)
    for (uint32_t u = threadIdx.x; u < nb * dim; u += PL_THREADS) {
      const uint32_t sl = u / dim, d = u - sl * dim, s_glob = sb + sl;
      uint32_t lo = 0, hi = E;  // largest e with pre[e] <= s_glob
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pre[mid] <= s_glob) lo = mid; else hi = mid;
      }
      const uint32_t e = lo;
      const uint32_t ns = e_nb ? steps[e_nb[e]] : steps[e];
      const uint32_t i = s_glob - pre[e] + 1;
      const double o = other_sm ? other_sm[d]
                                : ld_tree(T, &T.data[(uint64_t)d * T.pitch + sh.nb_idx[e_nb[e]]]);
      const double f = fixed_is_to ? o : fixed[d];
      const double t = fixed_is_to ? fixed[d] : o;
      // `if num_steps <= 1 { return is_valid(to) }`; step i of n: t = i / n (rrt.rs:99-105)
      st_buf[(size_t)sl * dim + d] =
          (ns <= 1) ? t : interp_coord(P, d, f, t, __ddiv_rn((double)i, (double)ns));
      if (d == 0) st_edge[sl] = e;
    }
    __syncthreads();
    // ---- stage 2 with the world's grid: a warp per state, lanes over the obstacles of the
    // state's cell (a handful instead of the whole world) and over the SO2 intervals
    if (P.w.grid_off) {
      const WorldView& w = P.w;
      for (uint32_t sl = gw; sl < nb; sl += GW) {
        const uint32_t e = st_edge[sl];
        if (*reinterpret_cast<volatile uint32_t*>(&hit[e])) continue;
        const double* x = st_buf + (size_t)sl * dim;
        const uint32_t cell = grid_cell(w, x);
        const uint32_t a = __ldg(w.grid_off + cell), b = __ldg(w.grid_off + cell + 1);
        bool h = false;
        for (uint32_t t = a + lane; t < b && !h; t += 32) h = obstacle_hit(w, __ldg(w.grid_ids + t), x);
        for (uint32_t t = lane; t < w.n_so2 && !h; t += 32) {
          const double v = so2_normalise_dev(x[P.n_real + w.so2_comp[t]]);
          h = v >= w.so2_lohi[2 * t] && v <= w.so2_lohi[2 * t + 1];
        }
        if (__any_sync(0xFFFFFFFFu, h) && lane == 0) atomicOr(&hit[e], 1u);
      }
      continue;
    }
    // ---- stage 2: obstacle chunk x state sweep
    for (uint32_t c = c0; c < n_oc; c += cstep) {
      uint32_t skip_e = OXG_INVALID_INDEX, tick = 0;
      for (uint32_t sl = g0; sl < nb; sl += groups, ++tick) {
        const uint32_t e = st_edge[sl];
        if (e == skip_e) continue;
        if ((tick & 15u) == 15u && *reinterpret_cast<volatile uint32_t*>(&hit[e])) {
          skip_e = e;
          continue;
        }
        const bool h = chunk_hit(P, c, lane, st_buf + (size_t)sl * dim);
        if (__any_sync(0xFFFFFFFFu, h)) {
          if (lane == 0) atomicOr(&hit[e], 1u);
          skip_e = e;
        }
      }
    }
  }
}

// RRT* find_neighbours (rrt_star.rs:121-131): all i < n with distance(q_new, tree[i]) <
// search_radius, ascending.  Each CTA writes the ordered hits of its slice.
template <bool MULTI>
__device__ void phase_radius(const PlanParams& P, PlanShared& sh, const TreeView& T, uint32_t n,
                             uint32_t pend_idx) {
  uint32_t lo, hi;
  cta_slice(n, &lo, &hi);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t run = 0;
  uint32_t* out_i = MULTI ? P.nbr_idx + (size_t)blockIdx.x * P.nbr_capc : sh.nb_idx;
  double* out_d = MULTI ? P.nbr_dist + (size_t)blockIdx.x * P.nbr_capc : sh.nb_dist;
  const uint32_t cap = MULTI ? P.nbr_capc : (uint32_t)NBR_CAP;
  for (uint32_t base = lo; base < hi; base += PL_THREADS) {
    const uint32_t i = base + threadIdx.x;
    bool h = false;
    double d = 0.0;
    if (i < hi) {
      d = (i == pend_idx) ? dist_ab(P, sm(sh.q_new), sm(sh.pend))
                          : dist_ab(P, sm(sh.q_new), node(T, i));
      h = d < P.search_radius;
    }
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, h);
    if (lane == 0) sh.warp_cnt[warp] = __popc(b);
    __syncthreads();
    uint32_t off = run, tot = 0;
    for (int wv = 0; wv < PL_WARPS; ++wv) {
      const uint32_t cnt = sh.warp_cnt[wv];
      if (wv < (int)warp) off += cnt;
      tot += cnt;
    }
    if (h) {
      const uint32_t pos = off + __popc(b & ((1u << lane) - 1u));
      if (pos < cap) {
        out_i[pos] = i;
        out_d[pos] = d;
      }
    }
    run += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (MULTI)
      P.nbr_cnt[blockIdx.x] = run;
    else
      sh.bc_u[1] = run;
  }
}

// Ordered compaction of the selected neighbour slots into the edge list (e_nb, e_pre) by warp 0:
// 32 slots per round, ballot for the positions, a shuffle scan for the state prefix.
// nb_ns[u] == 0xFFFFFFFE marks a motion with too many steps (-> *bad).
template <class Sel>
__device__ __forceinline__ void compact_edges(PlanShared& sh, uint32_t m, Sel sel) {
  if (threadIdx.x < 32) {
    const uint32_t lane = threadIdx.x;
    uint32_t E = 0, pre = 0;
    bool bad = false;
    for (uint32_t base = 0; base < m; base += 32) {
      const uint32_t u = base + lane;
      const bool sl = u < m && sel(u);
      const uint32_t ns = sl ? sh.nb_ns[u] : 0u;
      const uint32_t neff = sl ? ((ns <= 1u || ns == 0xFFFFFFFEu) ? 1u : ns) : 0u;
      const uint32_t b = __ballot_sync(0xFFFFFFFFu, sl);
      uint32_t incl = neff;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((int)lane >= o) incl += v;
      }
      if (sl) {
        const uint32_t pos = E + __popc(b & ((1u << lane) - 1u));
        sh.e_nb[pos] = u;
        sh.e_pre[pos] = pre + incl - neff;
      }
      bad = bad || __any_sync(0xFFFFFFFFu, sl && ns == 0xFFFFFFFEu);
      E += __popc(b);
      pre += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (lane == 0) {
      sh.e_pre[E] = pre;
      sh.bc_u[2] = E;
      sh.bc_i[0] = bad ? 1 : 0;
    }
  }
}

// ------------------------------------------------------------------------------- the kernel
template <int MODE, bool MULTI>
__global__ void __launch_bounds__(PL_THREADS, 1) planner_kernel(const PlanParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PlanShared& sh = *reinterpret_cast<PlanShared*>(smem_raw);
  double* const st_buf = reinterpret_cast<double*>(smem_raw + ((sizeof(PlanShared) + 15) & ~size_t(15)));
  unsigned long long bar_gen = 0;  // barriers passed (tracked by thread 0)
  const int dim = (int)(P.n_real + P.n_so2);
  const uint32_t tid = threadIdx.x;
  const uint32_t gtid = blockIdx.x * PL_THREADS + tid, gthreads = gridDim.x * PL_THREADS;

  uint32_t* const hit_ext = MULTI ? P.hit_ext : sh.hit_ext;
  uint32_t* const hit_cp = MULTI ? P.hit_cp : sh.hit_cp;
  uint32_t* const hit_rw = MULTI ? P.hit_rw : sh.hit_rw;
  if (!MULTI) {
    if (tid < 2) sh.hit_ext[tid] = 0u;
    __syncthreads();
  }
  auto flag = [](const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); };

  uint32_t n[2] = {P.n_init[0], P.n_init[1]};
  uint32_t pend_idx = OXG_INVALID_INDEX;  // index (in pend_tree) of the state still only in sh.pend
  int pend_tree = 0;
  unsigned long long it = 0, mc_ref = 0, mc_exec = 0, ext_count = 0;
  int solved = 0, status = ST_SAMPLES;
  long long node0 = -1, node1 = -1;
  // phase attribution (CTA 0 / thread 0 only): ns spent between consecutive marks
  unsigned long long ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ph_last = 0;
  const bool timing = gtid == 0;
  auto mark = [&](int slot) {
    if (timing) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (ph_last) ph_acc[slot] += now - ph_last;
      ph_last = now;
    }
  };
  double q_next = ((int)tid < dim && P.n_samples > 0) ? P.samples[tid] : 0.0;
  unsigned long long t_start = 0;
  if (P.timeout_ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));

  // one extend (rrt.rs:187-217 / rrt_connect.rs:121-159): returns true if a node was appended to
  // tree `t`; *new_idx, *reached; the new state is left in sh.q_new (and sh.pend).
  auto extend = [&](int t, const double* target, uint32_t* near_out, double* min_dist_out,
                    bool* reached_out, bool count_ref) -> bool {
    const TreeView& T = P.tree[t];
    const uint32_t pidx = (pend_tree == t) ? pend_idx : OXG_INVALID_INDEX;
    const unsigned par = (unsigned)(ext_count & 1ull);
    ++ext_count;
    mark(7);  // everything since the previous mark that is not attributed elsewhere
    phase_scan<MULTI>(P, sh, T, n[t], target, pidx);
    mark(0);  // 1-NN scan
    grid_barrier<MULTI>(P.barrier, bar_gen);
    mark(1);  // barrier 1
    if (gtid == 0) hit_ext[par ^ 1u] = 0u;  // the other parity's flag: next extend's
    phase_steer<MULTI>(P, sh, T, target, pidx, near_out, min_dist_out, reached_out);
    // check_motion(q_near, q_new)
    const double dist = dist_ab(P, sm(sh.q_near), sm(sh.q_new));
    const unsigned long long ns64 = motion_steps(dist, P.resolution);
    if (ns64 > MAX_STEPS) {
      status = ST_E_STEPS;
      return false;
    }
    if (tid == 0) {
      sh.e_pre[0] = 0;
      sh.e_pre[1] = ns64 <= 1 ? 1u : (uint32_t)ns64;
      sh.nb_ns[0] = (uint32_t)ns64;
    }
    __syncthreads();
    motion_tasks(P, sh, st_buf, T, 1, sh.q_new, true, sh.q_near, sh.nb_ns, sh.e_pre, nullptr,
                 hit_ext + par);
    if (count_ref) {
      ++mc_ref;
      ++mc_exec;
    }
    if (MODE == MODE_STAR) {
      // the radius scan and the flag reset ride in the same phase
      for (uint32_t u = gtid; u < (uint32_t)NBR_CAP; u += gthreads) {
        hit_cp[u] = 0u;
        hit_rw[u] = 0u;
      }
      phase_radius<MULTI>(P, sh, T, n[t], pidx);
    }
    mark(2);  // steer + motion check (+ radius scan)
    grid_barrier<MULTI>(P.barrier, bar_gen);
    mark(3);  // barrier 2
    const bool valid = flag(hit_ext + par) == 0u;
    return valid;
  };

  auto append = [&](int t, long long parent, double cost) -> uint32_t {
    const TreeView& T = P.tree[t];
    const uint32_t idx = n[t];
    if (blockIdx.x == 0) {
      if ((int)tid < dim) T.data[(uint64_t)tid * T.pitch + idx] = sh.q_new[tid];
      if (tid == 0) {
        T.parent[idx] = parent;
        if (MODE == MODE_STAR) P.cost[idx] = cost;
      }
    }
    __syncthreads();  // every reader of sh.pend is done
    if ((int)tid < dim) sh.pend[tid] = sh.q_new[tid];
    __syncthreads();
    pend_idx = idx;
    pend_tree = t;
    n[t] = idx + 1;
    return idx;
  };

  auto goal_ok = [&]() -> bool {
    if (!P.goal || P.goal_radius < 0.0) return false;
    return dist_ab(P, sm(sh.q_new), GlState{P.goal, 1, false}) <= P.goal_radius;
  };

  for (;;) {
    // loop bounds (the reference has only a wall-clock timeout; these make runs reproducible)
    if (it >= P.n_samples) {
      status = ST_SAMPLES;
      break;
    }
    if (n[0] >= P.max_nodes || (MODE == MODE_CONNECT && n[1] >= P.max_nodes)) {
      status = ST_FULL;
      break;
    }
    if (P.timeout_ns) {
      if (gtid == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t_start > P.timeout_ns) P.ctrl->stop = 1;
      }
    }
    if ((int)tid < dim) sh.q_rand[tid] = q_next;
    __syncthreads();
    ++it;
    if ((int)tid < dim && it < P.n_samples)  // prefetch the next sample off the critical path
      q_next = P.samples[it * (unsigned long long)dim + tid];

    uint32_t near;
    double min_dist;
    bool reached;

    if (MODE == MODE_RRT) {
      const bool ok = extend(0, sh.q_rand, &near, &min_dist, &reached, true);
      if (status < 0) break;
      if (P.timeout_ns && __ldcg(&P.ctrl->stop)) {
        status = ST_TIMEOUT;
        break;
      }
      if (ok) {
        const uint32_t idx = append(0, (long long)near, 0.0);
        if (goal_ok()) {  // rrt.rs:220-223
          solved = 1;
          node0 = idx;
          status = ST_SOLVED;
          break;
        }
      }
    } else if (MODE == MODE_CONNECT) {
      const int a = (n[0] <= n[1]) ? 0 : 1;  // rrt_connect.rs:249-254
      const int b = a ^ 1;
      const bool ok_a = extend(a, sh.q_rand, &near, &min_dist, &reached, true);
      if (status < 0) break;
      if (P.timeout_ns && __ldcg(&P.ctrl->stop)) {
        status = ST_TIMEOUT;
        break;
      }
      if (ok_a) {
        const uint32_t idx_a = append(a, (long long)near, 0.0);
        if (a == 0 && goal_ok()) {  // rrt_connect.rs:271-274
          solved = 1;
          node0 = idx_a;
          node1 = -1;
          status = ST_SOLVED;
          break;
        }
        if (n[b] < P.max_nodes) {
          // connect: ONE extend of the other tree towards q_new (rrt_connect.rs:276-279)
          if ((int)tid < dim) sh.q_rand[tid] = sh.q_new[tid];
          __syncthreads();
          const bool ok_b = extend(b, sh.q_rand, &near, &min_dist, &reached, true);
          if (status < 0) break;
          if (ok_b) {
            const uint32_t idx_b = append(b, (long long)near, 0.0);
            if (reached) {  // rrt_connect.rs:280
              solved = 1;
              node0 = (a == 0) ? idx_a : idx_b;
              node1 = (a == 0) ? idx_b : idx_a;
              status = ST_SOLVED;
              break;
            }
          }
        }
      }
    } else {  // MODE_STAR
      const TreeView& T = P.tree[0];
      const bool ok = extend(0, sh.q_rand, &near, &min_dist, &reached, true);
      if (status < 0) break;
      if (P.timeout_ns && __ldcg(&P.ctrl->stop)) {
        status = ST_TIMEOUT;
        break;
      }
      if (!ok) continue;  // rrt_star.rs:221-223
      // ---- gather the ordered neighbour list (a single CTA wrote it in place)
      if (MULTI) {
        for (uint32_t c = tid; c < gridDim.x; c += PL_THREADS) sh.e_nb[c] = __ldcg(&P.nbr_cnt[c]);
        __syncthreads();
      }
      if (MULTI && tid < 32) {
        uint32_t run = 0;
        for (uint32_t c0 = 0; c0 < gridDim.x; c0 += 32) {
          const uint32_t c = c0 + tid;
          const uint32_t cnt = c < gridDim.x ? sh.e_nb[c] : 0u;
          uint32_t incl = cnt;
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if ((int)tid >= o) incl += v;
          }
          if (c < gridDim.x) sh.cta_off[c] = run + incl - cnt;
          run += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        if (tid == 0) sh.bc_u[1] = run;
      }
      __syncthreads();
      const uint32_t m = sh.bc_u[1];
      if (m > (uint32_t)NBR_CAP) {
        status = ST_E_NEIGHBOURS;
        break;
      }
      if (MULTI) {
        // one parallel round of L2 loads: slot u comes from the CTA c with cta_off[c] <= u
        for (uint32_t u = tid; u < m; u += PL_THREADS) {
          uint32_t lo = 0, hi = gridDim.x;
          while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (sh.cta_off[mid] <= u) lo = mid; else hi = mid;
          }
          const uint32_t off = u - sh.cta_off[lo];
          sh.nb_idx[u] = __ldcg(&P.nbr_idx[(size_t)lo * P.nbr_capc + off]);
          sh.nb_dist[u] = __ldcg(&P.nbr_dist[(size_t)lo * P.nbr_capc + off]);
        }
      }
      __syncthreads();
      // ---- choose parent (rrt_star.rs:233-249)
      // cost(temp, nbr) = nbr.cost + distance(temp.state, nbr.state); the distance is the radius
      // scan's value (same argument order).
      const double cost_near = ld_tree(T, &P.cost[near]);
      const double min_cost0 = __dadd_rn(cost_near, dist_ab(P, sm(sh.q_new), sm(sh.q_near)));
      for (uint32_t u = tid; u < m; u += PL_THREADS) {
        const uint32_t j = sh.nb_idx[u];
        sh.nb_c[u] = __dadd_rn(ld_tree(T, &P.cost[j]), sh.nb_dist[u]);
        // check_motion(neighbour.state, q_new): dist = distance(from = nbr, to = new)
        const double dd = dist_ab(P, node(T, j), sm(sh.q_new));
        const unsigned long long ns64 = motion_steps(dd, P.resolution);
        sh.nb_ns[u] = ns64 > MAX_STEPS ? 0xFFFFFFFEu : (uint32_t)ns64;
      }
      __syncthreads();
      compact_edges(sh, m, [&](uint32_t u) { return sh.nb_c[u] < min_cost0; });  // superset of the lazy loop's
      __syncthreads();
      const uint32_t E0 = sh.bc_u[2];
      if (sh.bc_i[0]) {
        status = ST_E_STEPS;
        break;
      }
      if (E0 > 0) {
        motion_tasks(P, sh, st_buf, T, E0, sh.q_new, true, nullptr, sh.nb_ns, sh.e_pre, sh.e_nb, hit_cp);
        mc_exec += E0;
        mark(4);  // neighbour gather + choose-parent candidates + their motion checks
        grid_barrier<MULTI>(P.barrier, bar_gen);
        mark(5);  // barrier 3 / 4
      }
      // sequential replay of the lazy selection (hit flags fetched in parallel first)
      uint32_t* const hit_sm = reinterpret_cast<uint32_t*>(st_buf);
      for (uint32_t e = tid; e < E0; e += PL_THREADS) hit_sm[e] = flag(hit_cp + e);
      __syncthreads();
      if (tid == 0) {
        double min_cost = min_cost0;
        uint32_t best = near;
        unsigned long long ref = 0;
        for (uint32_t e = 0; e < E0; ++e) {
          const uint32_t u = sh.e_nb[e];
          if (sh.nb_c[u] < min_cost) {
            ++ref;
            if (hit_sm[e] == 0u) {
              min_cost = sh.nb_c[u];
              best = sh.nb_idx[u];
            }
          }
        }
        sh.bc_d[1] = min_cost;
        sh.bc_u[3] = best;
        sh.bc_u[4] = (uint32_t)ref;
      }
      __syncthreads();
      const double min_cost = sh.bc_d[1];
      const uint32_t best = sh.bc_u[3];
      mc_ref += sh.bc_u[4];
      const uint32_t new_idx = append(0, (long long)best, min_cost);
      // ---- rewire (rrt_star.rs:261-282): cost(nbr, new) = new.cost + distance(nbr.state, new.state)
      for (uint32_t u = tid; u < m; u += PL_THREADS) {
        const uint32_t j = sh.nb_idx[u];
        const double dn = dist_ab(P, node(T, j), sm(sh.q_new));
        const double c = __dadd_rn(min_cost, dn);
        const bool cand = (j != best) && (c < ld_tree(T, &P.cost[j]));
        sh.nb_c[u] = c;
        // check_motion(new, nbr): dist = distance(from = new, to = nbr)
        const double dd = dist_ab(P, sm(sh.q_new), node(T, j));
        const unsigned long long ns64 = motion_steps(dd, P.resolution);
        sh.nb_ns[u] = cand ? (ns64 > MAX_STEPS ? 0xFFFFFFFEu : (uint32_t)ns64) : 0xFFFFFFFFu;
      }
      __syncthreads();
      compact_edges(sh, m, [&](uint32_t u) { return sh.nb_ns[u] != 0xFFFFFFFFu; });
      __syncthreads();
      const uint32_t E1 = sh.bc_u[2];
      if (sh.bc_i[0]) {
        status = ST_E_STEPS;
        break;
      }
      if (E1 > 0) {
        motion_tasks(P, sh, st_buf, T, E1, sh.q_new, false, nullptr, sh.nb_ns, sh.e_pre, sh.e_nb, hit_rw);
        mc_exec += E1;
        mc_ref += E1;
        mark(6);  // replay, append, rewire candidates + their motion checks
        grid_barrier<MULTI>(P.barrier, bar_gen);
        mark(5);
        if (blockIdx.x == 0) {
          for (uint32_t e = tid; e < E1; e += PL_THREADS) {
            if (flag(hit_rw + e) == 0u) {
              const uint32_t u = sh.e_nb[e];
              const uint32_t j = sh.nb_idx[u];
              T.parent[j] = (long long)new_idx;
              P.cost[j] = sh.nb_c[u];
            }
          }
        }
      }
      if (goal_ok() && !solved) {  // rrt_star.rs:285-288
        solved = 1;
        node0 = new_idx;
        if (P.stop_at_goal) {
          status = ST_SOLVED;
          break;
        }
      }
    }
  }

  if (gtid == 0) {
    PlanCtrl* c = P.ctrl;
    c->n[0] = n[0];
    c->n[1] = n[1];
    c->iters = it;
    c->mc_ref = mc_ref;
    c->mc_exec = mc_exec;
    c->node[0] = node0;
    c->node[1] = node1;
    c->solved = solved;
    c->status = status;
    for (int i = 0; i < 8; ++i) c->phase_ns[i] = ph_acc[i];
  }
}

}  // namespace oxg

using namespace oxg;

namespace {

int env_int_pl(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// How many CTAs run the loop.  A single CTA needs no grid barrier and wins on small trees with
// small worlds; large trees / worlds use up to one CTA per SM.  OXG_PLAN_CTAS overrides.
int choose_ctas(uint64_t max_nodes, uint32_t dim, uint64_t n_obstacles, int sm_count) {
  const int forced = env_int_pl("OXG_PLAN_CTAS", 0);
  if (forced > 0) return std::min(forced, sm_count);
  uint64_t by_tree = (max_nodes * dim + 32767) / 32768;
  // below ~2000 obstacles one CTA sweeps a motion in less time than two grid barriers cost
  // (config 2, 1000 boxes: 7.0 us per iteration on 1 CTA, 8.7 on 16)
  uint64_t by_world = n_obstacles >= 2048 ? (n_obstacles + 63) / 64 : 1;
  uint64_t g = std::max<uint64_t>(by_tree, by_world);
  if (g < 1) g = 1;
  if (g > (uint64_t)sm_count) g = (uint64_t)sm_count;
  return (int)g;
}

template <int MODE>
int launch_planner(const PlanParams& P, int G, cudaStream_t st) {
  const size_t smem = ((sizeof(PlanShared) + 15) & ~size_t(15)) + PL_STATE_BYTES + 64;
  PlanParams pp = P;
  void* args[] = {(void*)&pp};
  if (G > 1) {
    OXG_CUDA_TRY(cudaFuncSetAttribute(planner_kernel<MODE, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OXG_CUDA_TRY(cudaLaunchCooperativeKernel((void*)planner_kernel<MODE, true>, dim3(G),
                                             dim3(PL_THREADS), args, smem, st));
  } else {
    OXG_CUDA_TRY(cudaFuncSetAttribute(planner_kernel<MODE, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OXG_CUDA_TRY(cudaLaunchCooperativeKernel((void*)planner_kernel<MODE, false>, dim3(1),
                                             dim3(PL_THREADS), args, smem, st));
  }
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

// parents / costs of nodes [n_planned, size) default to "root" (-1, 0.0): the caller appended
// them (the start / goal states, rrt.rs:139-156) rather than a planner kernel.
int prepare_tree(oxg_store_s* s, uint64_t max_nodes, bool want_cost) {
  if (max_nodes > 0xFFFFFFF0ull) {
    set_error("max_nodes exceeds the u32 index range");
    return OXG_E_ARG;
  }
  OXG_TRY(oxg_store_reserve(s, max_nodes));
  OXG_TRY(s->buf[B_PARENT].ensure_keep(max_nodes * sizeof(long long), s->n_planned * sizeof(long long),
                                       s->stream));
  OXG_TRY(s->buf[B_COST].ensure_keep(max_nodes * sizeof(double), s->n_planned * sizeof(double),
                                     s->stream));
  if (s->size > s->n_planned) {
    const uint64_t k = s->size - s->n_planned;
    OXG_CUDA_TRY(cudaMemsetAsync(s->buf[B_PARENT].as<long long>() + s->n_planned, 0xFF,
                                 k * sizeof(long long), s->stream));
    OXG_CUDA_TRY(cudaMemsetAsync(s->buf[B_COST].as<double>() + s->n_planned, 0, k * sizeof(double),
                                 s->stream));
    s->n_planned = s->size;
  }
  (void)want_cost;
  return OXG_OK;
}

int check_plan_args(oxg_world w, oxg_store tree, const oxg_plan_params* p, const double* samples,
                    uint64_t n_samples, oxg_plan_result* out) {
  if (!w || !tree || !p || !out || (n_samples > 0 && !samples)) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  memset(out, 0, sizeof(*out));
  out->goal_node[0] = out->goal_node[1] = -1;
  if (!(p->resolution > 0.0)) {
    set_error("motion resolution must be > 0 (got %g): the reference loop would not terminate",
              p->resolution);
    return OXG_E_RESOLUTION;
  }
  if (tree->size == 0) {
    set_error("the tree store must hold the start state (Planner::setup pushes it, rrt.rs:147-155)");
    return OXG_E_EMPTY;
  }
  if (p->max_nodes < tree->size) {
    set_error("max_nodes (%llu) is smaller than the tree (%llu)", (unsigned long long)p->max_nodes,
              (unsigned long long)tree->size);
    return OXG_E_ARG;
  }
  const oxg_space_desc& sp = tree->space;
  const bool has_rn = w->n_spheres || w->n_aabbs || w->n_capsules;
  if (has_rn && w->dim != sp.n_real) {
    set_error("world dimension %u does not match the space's RealVector block (%u)", w->dim, sp.n_real);
    return OXG_E_DIM;
  }
  for (uint32_t i = 0; i < w->n_so2; ++i)
    if (w->h_so2_comp[i] >= sp.n_so2) {
      set_error("SO2 interval refers to component %u but the space has %u", w->h_so2_comp[i], sp.n_so2);
      return OXG_E_DIM;
    }
  return OXG_OK;
}

// One launch of the planner kernel on G CTAs, growing the tree(s) up to stage_cap nodes
// (<= prm->max_nodes, which sizes the allocations).
template <int MODE>
int plan_stage(oxg_world w, oxg_store t0, oxg_store t1, const oxg_plan_params* prm,
               const double* goal_target, const double* samples, uint64_t n_samples,
               int64_t* parents0, int64_t* parents1, double* costs, oxg_plan_result* out, int G,
               uint64_t stage_cap) {
  OXG_TRY(check_plan_args(w, t0, prm, samples, n_samples, out));
  if (MODE == MODE_CONNECT) {
    if (!t1 || t1->size == 0) {
      set_error("RRT-Connect needs a goal tree holding the goal state (rrt_connect.rs:201-216)");
      return OXG_E_EMPTY;
    }
    if (t1->space.n_real != t0->space.n_real || t1->space.n_so2 != t0->space.n_so2 || t1 == t0) {
      set_error("start and goal trees must be two stores of the same space");
      return OXG_E_DIM;
    }
    if (prm->max_nodes < t1->size) {
      set_error("max_nodes is smaller than the goal tree");
      return OXG_E_ARG;
    }
  }
  if (MODE == MODE_STAR && !(prm->search_radius == prm->search_radius)) {
    set_error("search_radius is NaN");
    return OXG_E_ARG;
  }
  oxg_store_s* s = t0;
  const uint32_t dim = s->dim;
  cudaStream_t st = s->stream;
  OXG_TRY(prepare_tree(t0, prm->max_nodes, MODE == MODE_STAR));
  if (MODE == MODE_CONNECT) {
    OXG_CUDA_TRY(cudaStreamSynchronize(t1->stream));
    OXG_TRY(prepare_tree(t1, prm->max_nodes, false));
    OXG_CUDA_TRY(cudaStreamSynchronize(t1->stream));
  }

  // scratch layout in B_PLAN
  const uint32_t capc = (uint32_t)(((stage_cap + G - 1) / G + 31) / 32 * 32 + 32);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = (off + bytes + 255) / 256 * 256;
    return o;
  };
  const size_t o_ctrl = take(sizeof(PlanCtrl));
  const size_t o_bd = take(sizeof(double) * G);
  const size_t o_bi = take(sizeof(uint32_t) * G);
  const size_t o_hext = take(sizeof(uint32_t) * 2);
  const size_t o_hcp = take(sizeof(uint32_t) * NBR_CAP);
  const size_t o_hrw = take(sizeof(uint32_t) * NBR_CAP);
  const size_t o_goal = take(sizeof(double) * kMaxDim);
  const size_t o_cnt = take(sizeof(uint32_t) * G);
  const size_t o_bar = take(sizeof(unsigned long long));
  const size_t zero_bytes = off;  // everything above starts zeroed
  size_t o_nidx = 0, o_ndist = 0;
  if (MODE == MODE_STAR) {
    o_nidx = take(sizeof(uint32_t) * (size_t)G * capc);
    o_ndist = take(sizeof(double) * (size_t)G * capc);
  }
  OXG_TRY(s->buf[B_PLAN].ensure(off));
  unsigned char* base = s->buf[B_PLAN].as<unsigned char>();
  OXG_CUDA_TRY(cudaMemsetAsync(base, 0, zero_bytes, st));
  const bool has_goal = goal_target != nullptr && prm->goal_radius >= 0.0;
  if (has_goal)
    OXG_CUDA_TRY(cudaMemcpyAsync(base + o_goal, goal_target, sizeof(double) * dim,
                                 cudaMemcpyHostToDevice, st));
  const size_t sample_bytes = sizeof(double) * n_samples * dim;
  OXG_TRY(s->buf[B_QUERY].ensure(std::max<size_t>(sample_bytes, 8)));
  if (sample_bytes)
    OXG_CUDA_TRY(cudaMemcpyAsync(s->buf[B_QUERY].p, samples, sample_bytes, cudaMemcpyHostToDevice, st));

  PlanParams P;
  memset(&P, 0, sizeof(P));
  P.tree[0] = TreeView{t0->d_data, t0->pitch, t0->buf[B_PARENT].as<long long>(), G > 1 ? 1 : 0};
  P.n_init[0] = (uint32_t)t0->size;
  if (MODE == MODE_CONNECT) {
    P.tree[1] = TreeView{t1->d_data, t1->pitch, t1->buf[B_PARENT].as<long long>(), G > 1 ? 1 : 0};
    P.n_init[1] = (uint32_t)t1->size;
  }
  P.cost = t0->buf[B_COST].as<double>();
  P.max_nodes = (uint32_t)stage_cap;
  P.n_real = s->space.n_real;
  P.n_so2 = s->space.n_so2;
  P.w_real = s->space.w_real;
  P.w_so2 = s->space.w_so2;
  P.w = world_view(w);
  P.samples = s->buf[B_QUERY].as<double>();
  P.n_samples = n_samples;
  P.max_distance = prm->max_distance;
  P.search_radius = prm->search_radius;
  P.resolution = prm->resolution;
  P.goal_radius = has_goal ? prm->goal_radius : -1.0;
  P.goal = has_goal ? reinterpret_cast<const double*>(base + o_goal) : nullptr;
  P.stop_at_goal = prm->stop_at_goal;
  P.timeout_ns = prm->timeout_s > 0.0 ? (unsigned long long)(prm->timeout_s * 1e9) : 0ull;
  P.cta_best_dist = reinterpret_cast<double*>(base + o_bd);
  P.cta_best_idx = reinterpret_cast<uint32_t*>(base + o_bi);
  P.hit_ext = reinterpret_cast<uint32_t*>(base + o_hext);
  P.hit_cp = reinterpret_cast<uint32_t*>(base + o_hcp);
  P.hit_rw = reinterpret_cast<uint32_t*>(base + o_hrw);
  P.nbr_cnt = reinterpret_cast<uint32_t*>(base + o_cnt);
  P.nbr_idx = reinterpret_cast<uint32_t*>(base + o_nidx);
  P.nbr_dist = reinterpret_cast<double*>(base + o_ndist);
  P.nbr_capc = capc;
  P.ctrl = reinterpret_cast<PlanCtrl*>(base + o_ctrl);
  P.barrier = reinterpret_cast<unsigned long long*>(base + o_bar);

  OXG_CUDA_TRY(cudaStreamSynchronize(w->stream));  // the world upload is complete
  OXG_CUDA_TRY(cudaEventRecord(s->ev0, st));
  OXG_TRY(launch_planner<MODE>(P, G, st));
  OXG_CUDA_TRY(cudaEventRecord(s->ev1, st));
  PlanCtrl h;
  OXG_CUDA_TRY(cudaMemcpyAsync(&h, P.ctrl, sizeof(h), cudaMemcpyDeviceToHost, st));
  OXG_CUDA_TRY(cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, s->ev0, s->ev1);

  // the kernel appended on the device: publish the new sizes and refresh the f32 mirror
  const uint64_t old0 = t0->size;
  t0->size = h.n[0];
  t0->n_planned = h.n[0];
  t0->epoch += 1;
  OXG_TRY(store_refresh_mirror(t0, old0, h.n[0] - old0));
  if (MODE == MODE_CONNECT) {
    const uint64_t old1 = t1->size;
    t1->size = h.n[1];
    t1->n_planned = h.n[1];
    t1->epoch += 1;
    // t1's columns were written on t0's stream; its own stream sees them after the sync above
    OXG_TRY(store_refresh_mirror(t1, old1, h.n[1] - old1));
    OXG_CUDA_TRY(cudaStreamSynchronize(t1->stream));
  }
  if (parents0)
    OXG_CUDA_TRY(cudaMemcpyAsync(parents0, t0->buf[B_PARENT].p, h.n[0] * sizeof(long long),
                                 cudaMemcpyDeviceToHost, st));
  if (MODE == MODE_CONNECT && parents1)
    OXG_CUDA_TRY(cudaMemcpyAsync(parents1, t1->buf[B_PARENT].p, h.n[1] * sizeof(long long),
                                 cudaMemcpyDeviceToHost, st));
  if (MODE == MODE_STAR && costs)
    OXG_CUDA_TRY(cudaMemcpyAsync(costs, t0->buf[B_COST].p, h.n[0] * sizeof(double),
                                 cudaMemcpyDeviceToHost, st));
  OXG_CUDA_TRY(cudaStreamSynchronize(st));

  out->n_nodes[0] = h.n[0];
  out->n_nodes[1] = h.n[1];
  out->iters = h.iters;
  out->motion_checks = h.mc_ref;
  out->motion_checks_executed = h.mc_exec;
  out->goal_node[0] = h.node[0];
  out->goal_node[1] = h.node[1];
  out->solved = h.solved;
  out->status = h.status;
  out->device_ms = ms;
  out->ctas = G;
  for (int i = 0; i < 8; ++i) out->phase_ms[i] = (double)h.phase_ns[i] * 1e-6;
  if (h.status == ST_E_NEIGHBOURS) {
    set_error("more than %d tree nodes inside search_radius in one RRT* iteration "
              "(OXG_RRT_STAR_MAX_NEIGHBOURS, include/oxg.h): reduce search_radius", NBR_CAP);
    return OXG_E_CAPACITY;
  }
  if (h.status == ST_E_STEPS) {
    set_error("a motion needs more than %llu interpolation steps at resolution %g",
              MAX_STEPS, prm->resolution);
    return OXG_E_RESOLUTION;
  }
  return OXG_OK;
}

// Shared driver of the three planners.  The grid grows with the tree: a run starts on the CTA
// count that suits a tree of 4096 nodes and is relaunched (the kernels resume from the trees in
// the stores) when a tree reaches 4096 / 32 768 / 262 144 / 2 097 152 nodes and a larger tree wants
// more CTAs — a small tree on many CTAs only pays grid barriers.  Worlds of >= 2048 obstacles ask
// for their CTAs from the start (choose_ctas), so those runs are a single launch.
template <int MODE>
int plan_solve(oxg_world w, oxg_store t0, oxg_store t1, const oxg_plan_params* prm,
               const double* goal_target, const double* samples, uint64_t n_samples,
               int64_t* parents0, int64_t* parents1, double* costs, oxg_plan_result* out) {
  OXG_TRY(check_plan_args(w, t0, prm, samples, n_samples, out));
  if (MODE == MODE_CONNECT && !t1) {
    set_error("RRT-Connect needs a goal tree holding the goal state (rrt_connect.rs:201-216)");
    return OXG_E_EMPTY;
  }
  int dev = 0, sm_count = 148;
  OXG_CUDA_TRY(cudaGetDevice(&dev));
  OXG_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  const uint64_t n_obst = w->n_spheres + w->n_aabbs + w->n_capsules + w->n_so2;
  const uint32_t dim = t0->dim;
  static const uint64_t kCaps[] = {4096ull, 32768ull, 262144ull, 2097152ull};
  auto grid_for = [&](uint64_t cap) { return choose_ctas(cap, dim, n_obst, sm_count); };

  oxg_plan_result acc;
  memset(&acc, 0, sizeof(acc));
  uint64_t consumed = 0;
  const auto t_begin = std::chrono::steady_clock::now();
  for (;;) {
    uint64_t cur = t0->size;
    if (MODE == MODE_CONNECT && t1->size > cur) cur = t1->size;
    uint64_t cap = prm->max_nodes;
    for (uint64_t c : kCaps)
      if (c > cur && c < prm->max_nodes) {
        cap = c;
        break;
      }
    const int G = grid_for(cap);
    for (uint64_t c : kCaps)  // later stages that want the same grid join this one
      if (c > cap && c < prm->max_nodes && grid_for(c) == G) cap = c;
    if (grid_for(prm->max_nodes) == G) {
      bool last = true;
      for (uint64_t c : kCaps)
        if (c > cap && c < prm->max_nodes) last = false;
      if (last) cap = prm->max_nodes;
    }
    oxg_plan_params sp = *prm;
    if (prm->timeout_s > 0.0) {
      const double used =
          std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
      if (consumed > 0 && used >= prm->timeout_s) {
        acc.status = ST_TIMEOUT;
        break;
      }
      sp.timeout_s = std::max(prm->timeout_s - used, 1e-6);
    }
    oxg_plan_result r;
    OXG_TRY((plan_stage<MODE>(w, t0, t1, &sp, goal_target, samples + consumed * dim,
                              n_samples - consumed, parents0, parents1, costs, &r, G, cap)));
    consumed += r.iters;
    acc.iters += r.iters;
    acc.motion_checks += r.motion_checks;
    acc.motion_checks_executed += r.motion_checks_executed;
    acc.device_ms += r.device_ms;
    for (int i = 0; i < 8; ++i) acc.phase_ms[i] += r.phase_ms[i];
    acc.n_nodes[0] = r.n_nodes[0];
    acc.n_nodes[1] = r.n_nodes[1];
    acc.goal_node[0] = r.goal_node[0];
    acc.goal_node[1] = r.goal_node[1];
    acc.solved = r.solved;
    acc.status = r.status;
    acc.ctas = r.ctas;
    if (r.status == ST_FULL && cap < prm->max_nodes && consumed < n_samples) continue;
    break;
  }
  *out = acc;
  return OXG_OK;
}

}  // namespace

extern "C" {

int oxg_rrt_solve(oxg_world w, oxg_store tree, const oxg_plan_params* params,
                  const double* goal_target, const double* samples, uint64_t n_samples,
                  int64_t* parents, oxg_plan_result* out) {
  OXG_API_BEGIN
  return plan_solve<MODE_RRT>(w, tree, nullptr, params, goal_target, samples, n_samples, parents,
                              nullptr, nullptr, out);
  OXG_API_END
}

int oxg_rrt_connect_solve(oxg_world w, oxg_store start_tree, oxg_store goal_tree,
                          const oxg_plan_params* params, const double* goal_target,
                          const double* samples, uint64_t n_samples, int64_t* start_parents,
                          int64_t* goal_parents, oxg_plan_result* out) {
  OXG_API_BEGIN
  return plan_solve<MODE_CONNECT>(w, start_tree, goal_tree, params, goal_target, samples, n_samples,
                                  start_parents, goal_parents, nullptr, out);
  OXG_API_END
}

int oxg_rrt_star_solve(oxg_world w, oxg_store tree, const oxg_plan_params* params,
                       const double* goal_target, const double* samples, uint64_t n_samples,
                       int64_t* parents, double* costs, oxg_plan_result* out) {
  OXG_API_BEGIN
  return plan_solve<MODE_STAR>(w, tree, nullptr, params, goal_target, samples, n_samples, parents,
                               nullptr, costs, out);
  OXG_API_END
}

}  // extern "C"
