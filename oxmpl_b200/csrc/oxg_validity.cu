// oxg_validity.cu — declarative StateValidityChecker and the interpolated motion check.
//
// is_valid   : oxmpl/src/base/validity.rs:39-48 with the obstacle semantics of the reference's
//              own fixtures (README.md:147-150 sphere, oxmpl/tests/rrt_rvss_tests.rs:24-36 AABB,
//              oxmpl/tests/rrt_so2ss_tests.rs:17-27 SO2 interval) plus the oracle-defined capsule.
// check_motion: rrt.rs:90-116 (= rrt_connect.rs:166-189, rrt_star.rs:76-104, prm.rs:161-187).
//
// Mapping.  A group of L lanes owns one edge; lane g evaluates steps g+1, g+1+L, ... so the
// interpolation steps of an edge run in parallel, and a group ballot after each round of L
// steps gives the early exit and the exact index of the first invalid step (all earlier rounds
// were valid, and within a round the lowest set bit is the first).  Every lane walks the
// obstacle list in the same order, so obstacle reads are warp-uniform (one burst per primitive,
// served by L1/L2) and the FP64 pipe is the bound.  The boolean result does not depend on the
// evaluation order, so it is bit-identical to the sequential reference loop.
#include <algorithm>
#include <cstdlib>

#include "oxg_internal.cuh"
#include "oxg_math.cuh"

namespace oxg {

struct SpaceP {
  uint32_t n_real, n_so2;
  double w_real, w_so2;
};

// valid iff no obstacle marks the state invalid.  x holds n_real coordinates then n_so2 angles.
template <int DT>
__device__ __forceinline__ bool state_valid_dev(const WorldView& w, const double* x, int n_real_rt) {
  const int dim = DimOf<DT>::get(n_real_rt);
  for (uint32_t i = 0; i < w.n_spheres; ++i)
    if (sphere_hit_dev<DT>(x, w.spheres + (size_t)i * w.sphere_stride, dim)) return false;
  for (uint32_t i = 0; i < w.n_aabbs; ++i)
    if (aabb_hit_dev<DT>(x, w.aabbs + (size_t)i * w.aabb_stride, dim)) return false;
  for (uint32_t i = 0; i < w.n_capsules; ++i)
    if (capsule_hit_dev<DT>(x, w.capsules + (size_t)i * w.capsule_stride, dim)) return false;
  if (DT == 0) {  // fast RealVector instantiations have no SO2 components
    for (uint32_t i = 0; i < w.n_so2; ++i) {
      double v = so2_normalise_dev(x[dim + w.so2_comp[i]]);
      if (v >= w.so2_lohi[2 * i] && v <= w.so2_lohi[2 * i + 1]) return false;
    }
  }
  return true;
}

// Same predicate with the certified FP32 pre-filter in front of every exact test (RealVector
// fast paths).  xf = fl32(x); tol covers the magnitude of x.
template <int DT>
__device__ __forceinline__ bool state_valid_filtered_dev(const WorldView& w, const double* x,
                                                         const float* xf, const EdgeTol& tol) {
  for (uint32_t i = 0; i < w.n_spheres; ++i)
    if (sphere_maybe_hit32<DT>(xf, w.spheres32 + (size_t)i * w.sphere32_stride, tol.a2_s, tol.ce_s, DT) &&
        sphere_hit_dev<DT>(x, w.spheres + (size_t)i * w.sphere_stride, DT))
      return false;
  for (uint32_t i = 0; i < w.n_aabbs; ++i)
    if (aabb_maybe_hit32<DT>(xf, w.aabbs32 + (size_t)i * w.aabb_stride, DT) &&
        aabb_hit_dev<DT>(x, w.aabbs + (size_t)i * w.aabb_stride, DT))
      return false;
  for (uint32_t i = 0; i < w.n_capsules; ++i)
    if (capsule_maybe_hit32<DT>(xf, w.capsules32 + (size_t)i * w.capsule32_stride, tol.a2_c, tol.ce_c, DT) &&
        capsule_hit_dev<DT>(x, w.capsules + (size_t)i * w.capsule_stride, DT))
      return false;
  return true;
}

// Same, with the sphere / capsule mirrors staged in shared memory (sm_sph / sm_cap: 32-bit
// shared addresses of the first rows; 0 = not staged, read the global mirror).
template <int DT>
__device__ __forceinline__ bool state_valid_staged_dev(const WorldView& w, const double* x,
                                                       const float* xf, const EdgeTol& tol,
                                                       uint32_t sm_sph, uint32_t sm_cap) {
  for (uint32_t i = 0; i < w.n_spheres; ++i)
    if (sphere_maybe_hit32_s<DT>(xf, sm_sph + i * w.sphere32_stride * 4u, tol.a2_s, tol.ce_s) &&
        sphere_hit_dev<DT>(x, w.spheres + (size_t)i * w.sphere_stride, DT))
      return false;
  for (uint32_t i = 0; i < w.n_aabbs; ++i)
    if (aabb_maybe_hit32<DT>(xf, w.aabbs32 + (size_t)i * w.aabb_stride, DT) &&
        aabb_hit_dev<DT>(x, w.aabbs + (size_t)i * w.aabb_stride, DT))
      return false;
  for (uint32_t i = 0; i < w.n_capsules; i += 2) {
    const unsigned mh = capsule_pair_maybe_hit32_s<DT>(xf, sm_cap + (i >> 1) * w.capsule32p_stride * 4u,
                                                       tol.a2_c, tol.ce_c);
    if (mh) {
      if ((mh & 1u) && capsule_hit_dev<DT>(x, w.capsules + (size_t)i * w.capsule_stride, DT)) return false;
      if ((mh & 2u) && i + 1 < w.n_capsules &&
          capsule_hit_dev<DT>(x, w.capsules + (size_t)(i + 1) * w.capsule_stride, DT))
        return false;
    }
  }
  return true;
}

// Copies the sphere and capsule f32 mirrors into dynamic shared memory (whole CTA); returns
// false when the kernel was launched without the staging area.
__device__ __forceinline__ bool stage_world32(const WorldView& w, uint32_t smem_bytes,
                                              unsigned char* smem, uint32_t* sm_sph, uint32_t* sm_cap) {
  const uint32_t sph_bytes = w.n_spheres * w.sphere32_stride * 4u;
  const uint32_t cap_bytes = ((w.n_capsules + 1) / 2) * w.capsule32p_stride * 4u;
  if (smem_bytes < sph_bytes + cap_bytes || sph_bytes + cap_bytes == 0) return false;
  float4* dst = reinterpret_cast<float4*>(smem);
  const float4* s1 = reinterpret_cast<const float4*>(w.spheres32);
  const float4* s2 = reinterpret_cast<const float4*>(w.capsules32p);
  for (uint32_t i = threadIdx.x; i < sph_bytes / 16; i += blockDim.x) dst[i] = s1[i];
  for (uint32_t i = threadIdx.x; i < cap_bytes / 16; i += blockDim.x) dst[sph_bytes / 16 + i] = s2[i];
  __syncthreads();
  *sm_sph = (uint32_t)__cvta_generic_to_shared(smem);
  *sm_cap = *sm_sph + sph_bytes;
  return true;
}

template <int DT>
__global__ void __launch_bounds__(128) is_valid_kernel(const WorldView w, const SpaceP sp,
                                                       const StateView st, uint64_t Q,
                                                       uint8_t* __restrict__ out) {
  constexpr int XN = DT > 0 ? DT : kMaxDim;
  const int n_real = DimOf<DT>::get((int)sp.n_real);
  const int dim = DT > 0 ? DT : (int)(sp.n_real + sp.n_so2);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Q;
       i += (uint64_t)gridDim.x * blockDim.x) {
    double x[XN];
#pragma unroll
    for (int d = 0; d < dim; ++d) x[d] = st.p[i * st.stride_i + (uint64_t)d * st.stride_d];
    if (DT > 0) {
      float xf[XN];
      double ms = 0.0;
#pragma unroll
      for (int d = 0; d < dim; ++d) {
        xf[d] = __double2float_rn(x[d]);
        ms = fmax(ms, fabs(x[d]));
        if (!(x[d] == x[d])) ms = x[d];
      }
      const EdgeTol tol = edge_tol_dev(ms, w.wmax_spheres, w.wmax_capsules, DT);
      out[i] = state_valid_filtered_dev<DT>(w, x, xf, tol) ? 1 : 0;
    } else {
      out[i] = state_valid_dev<DT>(w, x, n_real) ? 1 : 0;
    }
  }
}

struct MotionParams {
  WorldView w;
  SpaceP sp;
  StateView from, to;
  const uint32_t* from_idx;  // nullable: endpoint e is row from_idx[e] of `from`
  const uint32_t* to_idx;
  uint64_t E;
  const uint64_t* E_dev;  // nullable: actual edge count in device memory
  double resolution;
  uint8_t* valid;
  uint32_t* first_invalid;  // nullable
  uint32_t world_smem_bytes;  // dynamic shared memory available for the staged f32 mirrors
  RowCull rc;                 // motion_cull_kernel: per-row survivor lists (all null: full pass)
};

template <int DT, int L>
__global__ void __launch_bounds__(128, 6) motion_kernel(const MotionParams p) {
  extern __shared__ __align__(16) unsigned char world_smem[];
  constexpr int XN = DT > 0 ? DT : kMaxDim;
  uint32_t sm_sph = 0, sm_cap = 0;
  const bool staged = DT > 0 && stage_world32(p.w, p.world_smem_bytes, world_smem, &sm_sph, &sm_cap);
  const int n_real = DimOf<DT>::get((int)p.sp.n_real);
  const int n_so2 = DT > 0 ? 0 : (int)p.sp.n_so2;
  const int dim = n_real + n_so2;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lane_g = lane % L;
  const unsigned group_shift = lane - lane_g;
  const unsigned group_mask = (L == 32) ? 0xFFFFFFFFu : (((1u << L) - 1u) << group_shift);
  const uint64_t gid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) / L;

  const uint64_t E = p.E_dev ? *p.E_dev : p.E;
  for (uint64_t e = gid; e < E; e += ngroups) {
    const uint64_t fi = p.from_idx ? p.from_idx[e] : e;
    const uint64_t ti = p.to_idx ? p.to_idx[e] : e;
    double from[XN], to[XN], x[XN];
#pragma unroll
    for (int d = 0; d < dim; ++d) {
      from[d] = p.from.p[fi * p.from.stride_i + (uint64_t)d * p.from.stride_d];
      to[d] = p.to.p[ti * p.to.stride_i + (uint64_t)d * p.to.stride_d];
    }
    // dist = space.distance(from, to); num_steps = ceil(dist / resolution) as usize  (rrt.rs:95-97)
    double dist;
    if (DT > 0)
      dist = sqrt(rv_sumsq_dev<DT>(from, to, n_real));
    else
      dist = space_distance_dev(p.sp.n_real, p.sp.n_so2, p.sp.w_real, p.sp.w_so2, from, to);
    unsigned long long n = ceil_as_usize_dev(__ddiv_rn(dist, p.resolution));
    unsigned long long first_bad = 0;
    EdgeTol tol;
    if (DT > 0) {
      // every interpolated coordinate lies between the endpoints' (up to one rounding, which the
      // 1.0001 factors of the tolerance absorb)
      double ms = 0.0;
#pragma unroll
      for (int d = 0; d < dim; ++d) {
        ms = fmax(ms, fmax(fabs(from[d]), fabs(to[d])));
        if (!(from[d] == from[d]) || !(to[d] == to[d])) ms = __longlong_as_double(0x7FF8000000000000ll);
      }
      tol = edge_tol_dev(ms * 1.0000001, p.w.wmax_spheres, p.w.wmax_capsules, DT);
    }
    if (n > (1ull << 32)) {
      first_bad = 0xFFFFFFFFull;  // the reference loop would not terminate in practice
    } else {
      const bool single = n <= 1;   // `if num_steps <= 1 { return vc.is_valid(to); }` (rrt.rs:99-101)
      const unsigned long long nn = single ? 1ull : n;
      const double nf = (double)n;
      for (unsigned long long base = 0; base < nn; base += L) {
        const unsigned long long i = base + lane_g + 1;
        bool bad = false;
        if (i <= nn) {
          if (single) {
#pragma unroll
            for (int d = 0; d < dim; ++d) x[d] = to[d];
          } else {
            const double t = __ddiv_rn((double)i, nf);  // `i as f64 / num_steps as f64` (rrt.rs:105)
#pragma unroll
            for (int d = 0; d < n_real; ++d)
              x[d] = __dadd_rn(from[d], __dmul_rn(__dsub_rn(to[d], from[d]), t));
            for (int c = 0; c < n_so2; ++c)
              x[n_real + c] = so2_interpolate_dev(from[n_real + c], to[n_real + c], t);
          }
          if (DT > 0) {
            float xf[XN];
#pragma unroll
            for (int d = 0; d < dim; ++d) xf[d] = __double2float_rn(x[d]);
            bad = staged ? !state_valid_staged_dev<DT>(p.w, x, xf, tol, sm_sph, sm_cap)
                         : !state_valid_filtered_dev<DT>(p.w, x, xf, tol);
          } else {
            bad = !state_valid_dev<DT>(p.w, x, n_real);
          }
        }
        const unsigned b = (__ballot_sync(group_mask, bad) & group_mask) >> group_shift;
        if (b) {
          first_bad = base + (unsigned long long)(__ffs(b) - 1) + 1ull;
          break;
        }
      }
    }
    if (lane_g == 0) {
      p.valid[e] = first_bad == 0 ? 1 : 0;
      if (p.first_invalid)
        p.first_invalid[e] = first_bad > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)first_bad;
    }
  }
}

// check_motion with a whole-edge cull in front of the per-step tests (RealVector fast paths,
// world mirrors staged in shared memory).  Phase 1: the L lanes of an edge share out the
// obstacles and test each ONCE against the ball around the edge's midpoint that contains every
// interpolated state (certified, *_fat_maybe_hit32*): an obstacle that fails it cannot
// invalidate any step.  Phase 2: the reference's step loop (rrt.rs:104-109), lanes over steps
// as in motion_kernel, against the few survivors only — per-state FP32 filter, then the exact
// FP64 predicate.  Results (validity, first invalid step) are those of the sequential loop.
constexpr int CULL_CAP = 40;  // survivors kept per edge; more than that falls back to the full walk

template <int DT, int L>
__global__ void __launch_bounds__(128, 6) motion_cull_kernel(const MotionParams p) {
  extern __shared__ __align__(16) unsigned char world_smem[];
  __shared__ uint16_t s_list[128 / L][CULL_CAP];
  uint32_t sm_sph = 0, sm_cap = 0;
  stage_world32(p.w, p.world_smem_bytes, world_smem, &sm_sph, &sm_cap);  // the launcher guarantees it fits
  // Control flow is kept warp-uniform so that every vote is a plain full-warp ballot (a vote
  // over a 4-lane mask is emulated by one pass per group): edges past the end and finished
  // step loops ride along predicated off.
  constexpr unsigned FULL = 0xFFFFFFFFu;
  constexpr unsigned GMASK = (L == 32) ? 0xFFFFFFFFu : ((1u << L) - 1u);
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lane_g = lane % L;
  const unsigned group_shift = lane - lane_g;
  const unsigned below = (1u << lane_g) - 1u;  // lanes of my group before me
  const uint64_t gid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) / L;
  uint16_t* my_list = s_list[threadIdx.x / L];
  const WorldView& w = p.w;
  const uint32_t n_sph = w.n_spheres, n_box = w.n_aabbs, n_cap = w.n_capsules;
  FatTol ft;
  ft.sqrt_d = __double2float_ru(sqrt((double)DT) * 1.000001);
  ft.inv_sqrt_d = __double2float_ru(1.000001 / sqrt((double)DT));

  const uint64_t E = p.E_dev ? *p.E_dev : p.E;
  const uint64_t e_rounds = (E + ngroups - 1) / ngroups;
  for (uint64_t round = 0; round < e_rounds; ++round) {
    const uint64_t e = gid + round * ngroups;
    const bool live = e < E;
    const uint64_t fi = live ? (p.from_idx ? p.from_idx[e] : e) : 0;
    const uint64_t ti = live ? (p.to_idx ? p.to_idx[e] : e) : 0;
    double from[DT], to[DT], x[DT];
#pragma unroll
    for (int d = 0; d < DT; ++d) {
      from[d] = live ? p.from.p[fi * p.from.stride_i + (uint64_t)d * p.from.stride_d] : 0.0;
      to[d] = live ? p.to.p[ti * p.to.stride_i + (uint64_t)d * p.to.stride_d] : 0.0;
    }
    const double dist = sqrt(rv_sumsq_dev<DT>(from, to, DT));
    const unsigned long long n = ceil_as_usize_dev(__ddiv_rn(dist, p.resolution));
    unsigned long long first_bad = 0;
    double ms = 0.0;
    float mf[DT];
#pragma unroll
    for (int d = 0; d < DT; ++d) {
      ms = fmax(ms, fmax(fabs(from[d]), fabs(to[d])));
      if (!(from[d] == from[d]) || !(to[d] == to[d])) ms = __longlong_as_double(0x7FF8000000000000ll);
      mf[d] = __double2float_rn(0.5 * from[d] + 0.5 * to[d]);
    }
    const EdgeTol tol = edge_tol_dev(ms * 1.0000001, w.wmax_spheres, w.wmax_capsules, DT);
    // every state of the edge lies within dist/2 (+ roundings) of the midpoint; NaN / inf make
    // rho NaN / inf and then nothing is culled
    ft.rho = __double2float_ru(0.5 * dist * 1.00001 + ms * 1.0e-6 + 1e-37);

    // ---- phase 1: cull
    unsigned cnt = 0;
    for (uint32_t i0 = 0; i0 < n_sph; i0 += L) {
      const uint32_t i = i0 + lane_g;
      const bool keep = live && i < n_sph &&
                        sphere_fat_maybe_hit32_s<DT>(mf, sm_sph + i * w.sphere32_stride * 4u, tol.a2_s, tol.ce_s, ft);
      const unsigned b = (__ballot_sync(FULL, keep) >> group_shift) & GMASK;
      if (keep) {
        const unsigned pos = cnt + __popc(b & below);
        if (pos < CULL_CAP) my_list[pos] = (uint16_t)i;
      }
      cnt += __popc(b);
    }
    for (uint32_t i0 = 0; i0 < n_box; i0 += L) {
      const uint32_t i = i0 + lane_g;
      const bool keep = live && i < n_box && aabb_fat_maybe_hit32<DT>(mf, w.aabbs32 + (size_t)i * w.aabb_stride, ft);
      const unsigned b = (__ballot_sync(FULL, keep) >> group_shift) & GMASK;
      if (keep) {
        const unsigned pos = cnt + __popc(b & below);
        if (pos < CULL_CAP) my_list[pos] = (uint16_t)(n_sph + i);
      }
      cnt += __popc(b);
    }
    // capsules: the pairs the owner row's ball could not rule out (row cull), else all of them
    bool all_caps = live;
    if (p.rc.row_cnt) {
      const uint32_t row = live ? __ldg(p.rc.e_row + e) : 0u;
      const uint32_t rcnt = live ? (uint32_t)__ldg(p.rc.row_cnt + row) : 0u;
      all_caps = live && rcnt == kRowCullOverflow;
      const uint32_t mine = all_caps ? 0u : rcnt;
      const uint16_t* rl = p.rc.row_list + (size_t)row * kRowCullCap;
      const uint32_t most = __reduce_max_sync(FULL, mine);
      for (uint32_t s0 = 0; s0 < most; s0 += L) {
        const uint32_t sidx = s0 + lane_g;
        unsigned mh = 0;
        uint32_t i = 0;
        if (sidx < mine) {
          i = 2u * (uint32_t)__ldg(rl + sidx);
          mh = capsule_pair_fat_maybe_hit32_s<DT>(mf, sm_cap + (i >> 1) * w.capsule32p_stride * 4u, tol.a2_c,
                                                  tol.ce_c, ft);
          if (i + 1 >= n_cap) mh &= 1u;
        }
        const unsigned b0 = (__ballot_sync(FULL, (mh & 1u) != 0) >> group_shift) & GMASK;
        const unsigned b1 = (__ballot_sync(FULL, (mh & 2u) != 0) >> group_shift) & GMASK;
        if (mh & 1u) {
          const unsigned pos = cnt + __popc(b0 & below);
          if (pos < CULL_CAP) my_list[pos] = (uint16_t)(n_sph + n_box + i);
        }
        cnt += __popc(b0);
        if (mh & 2u) {
          const unsigned pos = cnt + __popc(b1 & below);
          if (pos < CULL_CAP) my_list[pos] = (uint16_t)(n_sph + n_box + i + 1);
        }
        cnt += __popc(b1);
      }
    }
    const uint32_t n_cap_walk = __any_sync(FULL, all_caps) ? n_cap : 0u;
    for (uint32_t i0 = 0; i0 < n_cap_walk; i0 += 2 * L) {
      const uint32_t i = i0 + 2 * lane_g;
      unsigned mh = 0;
      if (all_caps && i < n_cap)
        mh = capsule_pair_fat_maybe_hit32_s<DT>(mf, sm_cap + (i >> 1) * w.capsule32p_stride * 4u, tol.a2_c,
                                                tol.ce_c, ft);
      if (i + 1 >= n_cap) mh &= 1u;
      const unsigned b0 = (__ballot_sync(FULL, (mh & 1u) != 0) >> group_shift) & GMASK;
      const unsigned b1 = (__ballot_sync(FULL, (mh & 2u) != 0) >> group_shift) & GMASK;
      // list order inside a pair block: all first halves, then all second halves (any order is
      // fine: the step loop takes the earliest step with ANY hit)
      if (mh & 1u) {
        const unsigned pos = cnt + __popc(b0 & below);
        if (pos < CULL_CAP) my_list[pos] = (uint16_t)(n_sph + n_box + i);
      }
      cnt += __popc(b0);
      if (mh & 2u) {
        const unsigned pos = cnt + __popc(b1 & below);
        if (pos < CULL_CAP) my_list[pos] = (uint16_t)(n_sph + n_box + i + 1);
      }
      cnt += __popc(b1);
    }
    __syncwarp();
    const bool full_walk = cnt > CULL_CAP;

    // ---- phase 2: the step loop over the survivors
    const bool endless = n > (1ull << 32);  // the reference loop would not terminate in practice
    if (endless) first_bad = 0xFFFFFFFFull;
    const bool single = n <= 1;   // `if num_steps <= 1 { return vc.is_valid(to); }` (rrt.rs:99-101)
    const unsigned long long nn = single ? 1ull : n;
    const double nf = (double)n;
    bool running = live && !endless && cnt > 0;
    for (unsigned long long base = 0; __ballot_sync(FULL, running && base < nn) != 0u; base += L) {
      const unsigned long long i = base + lane_g + 1;
      bool bad = false;
      if (running && i <= nn) {
        if (single) {
#pragma unroll
          for (int d = 0; d < DT; ++d) x[d] = to[d];
        } else {
          const double t = __ddiv_rn((double)i, nf);  // `i as f64 / num_steps as f64` (rrt.rs:105)
#pragma unroll
          for (int d = 0; d < DT; ++d) x[d] = __dadd_rn(from[d], __dmul_rn(__dsub_rn(to[d], from[d]), t));
        }
        float xf[DT];
#pragma unroll
        for (int d = 0; d < DT; ++d) xf[d] = __double2float_rn(x[d]);
        if (full_walk) {
          bad = !state_valid_staged_dev<DT>(w, x, xf, tol, sm_sph, sm_cap);
        } else {
          for (unsigned sidx = 0; sidx < cnt && !bad; ++sidx) {
            uint32_t o = my_list[sidx];
            if (o < n_sph) {
              bad = sphere_maybe_hit32_s<DT>(xf, sm_sph + o * w.sphere32_stride * 4u, tol.a2_s, tol.ce_s) &&
                    sphere_hit_dev<DT>(x, w.spheres + (size_t)o * w.sphere_stride, DT);
            } else if ((o -= n_sph) < n_box) {
              bad = aabb_hit_dev<DT>(x, w.aabbs + (size_t)o * w.aabb_stride, DT);
            } else {
              o -= n_box;
              bad = capsule_half_maybe_hit32_s<DT>(xf, sm_cap + (o >> 1) * w.capsule32p_stride * 4u, o & 1u,
                                                   tol.a2_c, tol.ce_c) &&
                    capsule_hit_dev<DT>(x, w.capsules + (size_t)o * w.capsule_stride, DT);
            }
          }
        }
      }
      const unsigned b = (__ballot_sync(FULL, bad) >> group_shift) & GMASK;
      if (running && b) {
        first_bad = base + (unsigned long long)(__ffs(b) - 1) + 1ull;
        running = false;
      }
      if (base + L >= nn) running = false;
    }
    __syncwarp();  // the list is reused by the group's next edge
    if (live && lane_g == 0) {
      p.valid[e] = first_bad == 0 ? 1 : 0;
      if (p.first_invalid)
        p.first_invalid[e] = first_bad > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)first_bad;
    }
  }
}


// Row cull of the PRM k-NN build (see RowCull in oxg_internal.cuh).  The ball is
// centred on the row's state; its radius is the longest of the row's candidate edges, computed
// with the same reference-order distance the edge kernel uses, so every interpolated state of
// every owned edge lies inside it and the certified fat test (FatTol) may discard for all of them.
struct RowCullParams {
  WorldView w;
  StateView st;
  const uint32_t* idx;
  uint32_t k;
  uint64_t base_row, row_begin, row_end;
  double fixed_radius;  // idx == nullptr: every candidate edge of every row is shorter than this
  uint16_t* row_cnt;
  uint16_t* row_list;
  uint32_t world_smem_bytes;
};

template <int DT>
__global__ void __launch_bounds__(128, 6) prm_row_cull_kernel(const RowCullParams p) {
  extern __shared__ __align__(16) unsigned char world_smem[];
  uint32_t sm_sph = 0, sm_cap = 0;
  stage_world32(p.w, p.world_smem_bytes, world_smem, &sm_sph, &sm_cap);
  // One row per LANE: the 32 rows of a warp walk the capsule pairs together, so every pair row is
  // one broadcast read of shared memory for 32 tests (a warp per row reads the whole staged world
  // once per row and is bound by shared-memory bandwidth: 0.90 ms against 0.3x ms this way).
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  const WorldView& w = p.w;
  const uint32_t n_cap = w.n_capsules, n_pairs = (n_cap + 1) / 2;
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  FatTol ft;
  ft.sqrt_d = __double2float_ru(sqrt((double)DT) * 1.000001);
  ft.inv_sqrt_d = __double2float_ru(1.000001 / sqrt((double)DT));
  for (uint64_t r = p.row_begin + tid; r < p.row_end; r += nthreads) {
    double x[DT];
#pragma unroll
    for (int d = 0; d < DT; ++d) x[d] = p.st.p[r * p.st.stride_i + (uint64_t)d * p.st.stride_d];
    // longest candidate edge of the row
    double R = p.idx ? 0.0 : p.fixed_radius;
    bool odd = !(R == R);
    const uint32_t* slots = p.idx ? p.idx + (r - p.base_row) * p.k : nullptr;
    for (uint32_t sl = 0; slots && sl < p.k; ++sl) {
      const uint32_t j = __ldg(slots + sl);
      if (j == OXG_INVALID_INDEX) continue;
      double y[DT];
#pragma unroll
      for (int d = 0; d < DT; ++d) y[d] = p.st.p[(uint64_t)j * p.st.stride_i + (uint64_t)d * p.st.stride_d];
      const double dj = sqrt(rv_sumsq_dev<DT>(x, y, DT));
      odd = odd || !(dj == dj);
      R = fmax(R, dj);
    }
    if (odd) R = inf;
    double ms = 0.0;
    float mf[DT];
#pragma unroll
    for (int d = 0; d < DT; ++d) {
      ms = fmax(ms, fabs(x[d]));
      if (!(x[d] == x[d])) ms = __longlong_as_double(0x7FF8000000000000ll);
      mf[d] = __double2float_rn(x[d]);
    }
    ms = ms + R;  // every coordinate of every state of the row's edges is within R of x
    const EdgeTol tol = edge_tol_dev(ms * 1.0000001, w.wmax_spheres, w.wmax_capsules, DT);
    ft.rho = __double2float_ru(R * 1.00001 + ms * 1.0e-6 + 1e-37);  // NaN / inf: nothing is culled
    uint16_t* out = p.row_list + (size_t)(r - p.row_begin) * kRowCullCap;
    unsigned cnt = 0;
    for (uint32_t q = 0; q < n_pairs; ++q) {
      unsigned mh = capsule_pair_fat_maybe_hit32_s<DT>(mf, sm_cap + q * w.capsule32p_stride * 4u, tol.a2_c,
                                                      tol.ce_c, ft);
      if (2 * q + 1 >= n_cap) mh &= 1u;
      if (mh) {
        if (cnt < kRowCullCap) out[cnt] = (uint16_t)q;
        ++cnt;
      }
    }
    p.row_cnt[r - p.row_begin] = cnt > kRowCullCap ? kRowCullOverflow : (uint16_t)cnt;
  }
}

// Few edges against a large world: one CTA per edge, threads spread over (step, obstacle) pairs
// so that a single check_motion / is_valid call (the trait path, or an unfused planner
// iteration) is not one thread walking thousands of obstacles.  first_invalid = the smallest
// step with any hit (shared-memory atomicMin), identical to the sequential loop's answer.
template <int DT>
__global__ void __launch_bounds__(256) motion_small_kernel(const MotionParams p) {
  constexpr int XN = DT > 0 ? DT : kMaxDim;
  __shared__ unsigned int s_first;
  const int n_real = DimOf<DT>::get((int)p.sp.n_real);
  const int n_so2 = DT > 0 ? 0 : (int)p.sp.n_so2;
  const int dim = n_real + n_so2;
  const uint64_t e = blockIdx.x;
  const uint64_t fi = p.from_idx ? p.from_idx[e] : e;
  const uint64_t ti = p.to_idx ? p.to_idx[e] : e;
  double from[XN], to[XN], x[XN];
#pragma unroll
  for (int d = 0; d < dim; ++d) {
    from[d] = p.from.p[fi * p.from.stride_i + (uint64_t)d * p.from.stride_d];
    to[d] = p.to.p[ti * p.to.stride_i + (uint64_t)d * p.to.stride_d];
  }
  double dist;
  if (DT > 0)
    dist = sqrt(rv_sumsq_dev<DT>(from, to, n_real));
  else
    dist = space_distance_dev(p.sp.n_real, p.sp.n_so2, p.sp.w_real, p.sp.w_so2, from, to);
  const unsigned long long n = ceil_as_usize_dev(__ddiv_rn(dist, p.resolution));
  if (threadIdx.x == 0) s_first = 0xFFFFFFFFu;
  __syncthreads();
  const WorldView& w = p.w;
  const unsigned long long C = (unsigned long long)w.n_spheres + w.n_aabbs + w.n_capsules + w.n_so2;
  if (n > (1ull << 32)) {
    if (threadIdx.x == 0) s_first = 0xFFFFFFFEu;  // reported as 0xFFFFFFFF below
  } else if (C > 0) {
    const bool single = n <= 1;
    const unsigned long long nn = single ? 1ull : n;
    const double nf = (double)n;
    const unsigned long long total = nn * C;
    unsigned long long cur_step = 0;  // step whose state is in x[]
    for (unsigned long long t = threadIdx.x; t < total; t += blockDim.x) {
      const unsigned long long step = t / C + 1;
      unsigned long long o = t - (step - 1) * C;
      if (step >= s_first) continue;  // a hit at an earlier step is already known (benign race)
      if (step != cur_step) {
        cur_step = step;
        if (single) {
#pragma unroll
          for (int d = 0; d < dim; ++d) x[d] = to[d];
        } else {
          const double tt = __ddiv_rn((double)step, nf);
#pragma unroll
          for (int d = 0; d < n_real; ++d)
            x[d] = __dadd_rn(from[d], __dmul_rn(__dsub_rn(to[d], from[d]), tt));
          for (int c = 0; c < n_so2; ++c)
            x[n_real + c] = so2_interpolate_dev(from[n_real + c], to[n_real + c], tt);
        }
      }
      bool hit;
      if (o < w.n_spheres) {
        hit = sphere_hit_dev<DT>(x, w.spheres + (size_t)o * w.sphere_stride, n_real);
      } else if ((o -= w.n_spheres) < w.n_aabbs) {
        hit = aabb_hit_dev<DT>(x, w.aabbs + (size_t)o * w.aabb_stride, n_real);
      } else if ((o -= w.n_aabbs) < w.n_capsules) {
        hit = capsule_hit_dev<DT>(x, w.capsules + (size_t)o * w.capsule_stride, n_real);
      } else {
        o -= w.n_capsules;
        const double v = so2_normalise_dev(x[n_real + w.so2_comp[o]]);
        hit = v >= w.so2_lohi[2 * o] && v <= w.so2_lohi[2 * o + 1];
      }
      if (hit) atomicMin(&s_first, (unsigned int)(step > 0xFFFFFFFDull ? 0xFFFFFFFDull : step));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int f = s_first;
    p.valid[e] = f == 0xFFFFFFFFu ? 1 : 0;
    if (p.first_invalid) p.first_invalid[e] = f == 0xFFFFFFFFu ? 0u : (f == 0xFFFFFFFEu ? 0xFFFFFFFFu : f);
  }
}

#define OXG_DISPATCH_DT(D_RT, CALL)                     \
  switch (D_RT) {                                       \
    case 1: { constexpr int DD = 1; CALL; } break;      \
    case 2: { constexpr int DD = 2; CALL; } break;      \
    case 3: { constexpr int DD = 3; CALL; } break;      \
    case 4: { constexpr int DD = 4; CALL; } break;      \
    case 5: { constexpr int DD = 5; CALL; } break;      \
    case 6: { constexpr int DD = 6; CALL; } break;      \
    case 7: { constexpr int DD = 7; CALL; } break;      \
    case 8: { constexpr int DD = 8; CALL; } break;      \
    case 9: { constexpr int DD = 9; CALL; } break;      \
    case 10: { constexpr int DD = 10; CALL; } break;    \
    case 11: { constexpr int DD = 11; CALL; } break;    \
    case 12: { constexpr int DD = 12; CALL; } break;    \
    case 13: { constexpr int DD = 13; CALL; } break;    \
    case 14: { constexpr int DD = 14; CALL; } break;    \
    case 15: { constexpr int DD = 15; CALL; } break;    \
    case 16: { constexpr int DD = 16; CALL; } break;    \
    default: { constexpr int DD = 0; CALL; } break;     \
  }

#define OXG_DISPATCH_FAST_DT(D_RT, CALL)                \
  switch (D_RT) {                                       \
    case 1: { constexpr int DD = 1; CALL; } break;      \
    case 2: { constexpr int DD = 2; CALL; } break;      \
    case 3: { constexpr int DD = 3; CALL; } break;      \
    case 4: { constexpr int DD = 4; CALL; } break;      \
    case 5: { constexpr int DD = 5; CALL; } break;      \
    case 6: { constexpr int DD = 6; CALL; } break;      \
    case 7: { constexpr int DD = 7; CALL; } break;      \
    case 8: { constexpr int DD = 8; CALL; } break;      \
    case 9: { constexpr int DD = 9; CALL; } break;      \
    case 10: { constexpr int DD = 10; CALL; } break;    \
    case 11: { constexpr int DD = 11; CALL; } break;    \
    case 12: { constexpr int DD = 12; CALL; } break;    \
    case 13: { constexpr int DD = 13; CALL; } break;    \
    case 14: { constexpr int DD = 14; CALL; } break;    \
    case 15: { constexpr int DD = 15; CALL; } break;    \
    case 16: { constexpr int DD = 16; CALL; } break;    \
    default: break;                                     \
  }

static int check_world_space(const oxg_world_s* w, const oxg_space_desc& sp) {
  const bool has_rn = w->n_spheres || w->n_aabbs || w->n_capsules;
  if (has_rn && w->dim != sp.n_real) {
    set_error("world dimension %u does not match the space's RealVector block (%u)", w->dim,
              sp.n_real);
    return OXG_E_DIM;
  }
  for (uint32_t i = 0; i < w->n_so2; ++i)
    if (w->h_so2_comp[i] >= sp.n_so2) {
      set_error("SO2 interval refers to component %u but the space has %u", w->h_so2_comp[i],
                sp.n_so2);
      return OXG_E_DIM;
    }
  return OXG_OK;
}

static inline int fast_dim(const oxg_space_desc& sp) {
  return (sp.n_so2 == 0 && sp.n_real <= (uint32_t)kMaxFastDim) ? (int)sp.n_real : 0;
}

int is_valid_device(const oxg_world_s* w, const oxg_space_desc& sp, cudaStream_t stream,
                    StateView st, uint64_t Q, uint8_t* d_out) {
  OXG_TRY(check_world_space(w, sp));
  if (Q == 0) return OXG_OK;
  const uint64_t n_obst = w->n_spheres + w->n_aabbs + w->n_capsules + w->n_so2;
  if (Q <= 148ull * 4 && n_obst >= 64) {
    // few states, large world: is_valid(x) == check_motion(x, x) (zero length: `to` only)
    MotionParams p;
    p.w = world_view(w);
    p.sp = SpaceP{sp.n_real, sp.n_so2, sp.w_real, sp.w_so2};
    p.from = st;
    p.to = st;
    p.from_idx = nullptr;
    p.to_idx = nullptr;
    p.E = Q;
    p.E_dev = nullptr;
    p.resolution = 1.0;
    p.valid = d_out;
    p.first_invalid = nullptr;
    p.world_smem_bytes = 0;
    OXG_DISPATCH_DT(fast_dim(sp), (motion_small_kernel<DD><<<(unsigned)Q, 256, 0, stream>>>(p)));
    OXG_LAUNCH_CHECK();
    return OXG_OK;
  }
  WorldView wv = world_view(w);
  SpaceP spp{sp.n_real, sp.n_so2, sp.w_real, sp.w_so2};
  int threads = 128;
  unsigned blocks = (unsigned)std::min<uint64_t>((Q + threads - 1) / threads, 148ull * 32);
  OXG_DISPATCH_DT(fast_dim(sp),
                  (is_valid_kernel<DD><<<blocks, threads, 0, stream>>>(wv, spp, st, Q, d_out)));
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

int motion_check_device(const oxg_world_s* w, const oxg_space_desc& sp, cudaStream_t stream,
                        StateView from, StateView to, const uint32_t* d_from_idx,
                        const uint32_t* d_to_idx, uint64_t E, const uint64_t* E_dev,
                        double resolution, uint8_t* d_valid, uint32_t* d_first_invalid,
                        const RowCull* row_cull) {
  OXG_TRY(check_world_space(w, sp));
  if (!(resolution > 0.0)) {
    set_error("motion resolution must be > 0 (got %g): the reference loop would not terminate",
              resolution);
    return OXG_E_RESOLUTION;
  }
  if (E == 0) return OXG_OK;
  MotionParams p;
  p.w = world_view(w);
  p.sp = SpaceP{sp.n_real, sp.n_so2, sp.w_real, sp.w_so2};
  p.from = from;
  p.to = to;
  p.from_idx = d_from_idx;
  p.to_idx = d_to_idx;
  p.E = E;
  p.E_dev = E_dev;
  p.resolution = resolution;
  p.valid = d_valid;
  p.first_invalid = d_first_invalid;
  const uint64_t n_obst = w->n_spheres + w->n_aabbs + w->n_capsules + w->n_so2;
  if (E_dev == nullptr && E <= 148ull * 4 && n_obst >= 64) {
    // few edges, large world: a CTA per edge over (step, obstacle) pairs
    OXG_DISPATCH_DT(fast_dim(sp), (motion_small_kernel<DD><<<(unsigned)E, 256, 0, stream>>>(p)));
    OXG_LAUNCH_CHECK();
    return OXG_OK;
  }
  const int threads = 128;
  // few edges: a whole warp per edge; many edges: 4 lanes per edge keeps lanes busy at the
  // typical 4-12 steps per edge.
  const bool wide = E < 148ull * 64;
  const int L = wide ? 32 : 4;
  unsigned blocks = (unsigned)std::min<uint64_t>((E * L + threads - 1) / threads, 148ull * 16);
  // stage the sphere / capsule f32 mirrors in shared memory when they fit beside 5 CTAs per SM
  size_t stage = (size_t)w->n_spheres * w->sphere32_stride * 4 +
                 (size_t)((w->n_capsules + 1) / 2) * w->capsule32p_stride * 4;
  if (stage > 40 * 1024 || fast_dim(sp) == 0) stage = 0;
  p.world_smem_bytes = (uint32_t)stage;
  // Whole-edge cull (motion_cull_kernel): staged RealVector worlds with enough obstacles for
  // the cull pass to pay.  OXG_MOTION_CULL=0 runs the plain per-step walk for A/B measurements.
  static const bool cull_on = [] {
    const char* v = getenv("OXG_MOTION_CULL");
    return !(v && v[0] == '0');
  }();
  if (cull_on && stage > 0 && !wide && w->n_so2 == 0 && n_obst >= 16 && n_obst <= 60000) {
    if (row_cull) p.rc = *row_cull;
    OXG_DISPATCH_FAST_DT(fast_dim(sp), (motion_cull_kernel<DD, 4><<<blocks, threads, stage, stream>>>(p)));
    OXG_LAUNCH_CHECK();
    return OXG_OK;
  }
  if (wide) {
    OXG_DISPATCH_DT(fast_dim(sp), (motion_kernel<DD, 32><<<blocks, threads, stage, stream>>>(p)));
  } else {
    OXG_DISPATCH_DT(fast_dim(sp), (motion_kernel<DD, 4><<<blocks, threads, stage, stream>>>(p)));
  }
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

// The row lists are only consumed by motion_cull_kernel: the conditions below are the ones under
// which motion_check_device takes that kernel for the same world / space / edge count.
int prm_row_cull_device(const oxg_world_s* w, const oxg_space_desc& sp, cudaStream_t stream, StateView st,
                        const uint32_t* d_idx, uint32_t k, uint64_t base_row, uint64_t row_begin,
                        uint64_t row_end, double fixed_radius, uint64_t n_edges, uint16_t* row_cnt,
                        uint16_t* row_list, bool* enabled) {
  *enabled = false;
  static const bool on = [] {
    const char* v = getenv("OXG_ROW_CULL");
    const char* c = getenv("OXG_MOTION_CULL");
    return !(v && v[0] == '0') && !(c && c[0] == '0');
  }();
  const uint64_t rows = row_end - row_begin;
  const uint64_t n_obst = w->n_spheres + w->n_aabbs + w->n_capsules + w->n_so2;
  size_t stage = (size_t)w->n_spheres * w->sphere32_stride * 4 +
                 (size_t)((w->n_capsules + 1) / 2) * w->capsule32p_stride * 4;
  if (!on || g_tune_row_cull.load() == 0 || rows == 0 || (d_idx && (k == 0 || k > 32)) || fast_dim(sp) == 0 ||
      fast_dim(sp) > kMaxFastDim || stage == 0 || stage > 40 * 1024 || w->n_so2 != 0 || w->n_capsules < 64 ||
      n_obst > 60000 || n_edges < 148ull * 64)
    return OXG_OK;
  RowCullParams p;
  p.w = world_view(w);
  p.st = st;
  p.idx = d_idx;
  p.k = k;
  p.base_row = base_row;
  p.row_begin = row_begin;
  p.row_end = row_end;
  p.fixed_radius = fixed_radius;
  p.row_cnt = row_cnt;
  p.row_list = row_list;
  p.world_smem_bytes = (uint32_t)stage;
  const unsigned blocks = (unsigned)std::min<uint64_t>((rows + 127) / 128, 148ull * 16);
  OXG_DISPATCH_FAST_DT(fast_dim(sp), (prm_row_cull_kernel<DD><<<blocks, 128, stage, stream>>>(p)));
  OXG_LAUNCH_CHECK();
  *enabled = true;
  return OXG_OK;
}

}  // namespace oxg

using namespace oxg;

extern "C" {

int oxg_is_valid(oxg_world w, const oxg_space_desc* space, const double* states, uint64_t Q,
                 uint8_t* out) {
  OXG_API_BEGIN
  if (!w || (Q > 0 && (!states || !out))) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  OXG_TRY(validate_space(space));
  if (Q == 0) return OXG_OK;
  uint32_t dim = space->n_real + space->n_so2;
  size_t bytes = sizeof(double) * Q * dim;
  OXG_TRY(w->buf[0].ensure(bytes));
  OXG_TRY(w->buf[1].ensure(Q));
  OXG_CUDA_TRY(cudaMemcpyAsync(w->buf[0].p, states, bytes, cudaMemcpyHostToDevice, w->stream));
  OXG_TRY(is_valid_device(w, *space, w->stream, rows_view(w->buf[0].as<double>(), dim), Q,
                          w->buf[1].as<uint8_t>()));
  OXG_CUDA_TRY(cudaMemcpyAsync(out, w->buf[1].p, Q, cudaMemcpyDeviceToHost, w->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(w->stream));
  return OXG_OK;
  OXG_API_END
}

int oxg_check_motion(oxg_world w, const oxg_space_desc* space, const double* from, const double* to,
                     uint64_t E, double resolution, uint8_t* valid, uint32_t* first_invalid_step) {
  OXG_API_BEGIN
  if (!w || (E > 0 && (!from || !to || !valid))) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  OXG_TRY(validate_space(space));
  if (!(resolution > 0.0)) {
    set_error("motion resolution must be > 0 (got %g): the reference loop would not terminate",
              resolution);
    return OXG_E_RESOLUTION;
  }
  if (E == 0) return OXG_OK;
  uint32_t dim = space->n_real + space->n_so2;
  size_t bytes = sizeof(double) * E * dim;
  OXG_TRY(w->buf[0].ensure(bytes));
  OXG_TRY(w->buf[1].ensure(bytes));
  OXG_TRY(w->buf[2].ensure(E));
  OXG_TRY(w->buf[3].ensure(E * sizeof(uint32_t)));
  OXG_CUDA_TRY(cudaMemcpyAsync(w->buf[0].p, from, bytes, cudaMemcpyHostToDevice, w->stream));
  OXG_CUDA_TRY(cudaMemcpyAsync(w->buf[1].p, to, bytes, cudaMemcpyHostToDevice, w->stream));
  OXG_TRY(motion_check_device(w, *space, w->stream, rows_view(w->buf[0].as<double>(), dim),
                              rows_view(w->buf[1].as<double>(), dim), nullptr, nullptr, E, nullptr,
                              resolution, w->buf[2].as<uint8_t>(),
                              first_invalid_step ? w->buf[3].as<uint32_t>() : nullptr));
  OXG_CUDA_TRY(cudaMemcpyAsync(valid, w->buf[2].p, E, cudaMemcpyDeviceToHost, w->stream));
  if (first_invalid_step)
    OXG_CUDA_TRY(cudaMemcpyAsync(first_invalid_step, w->buf[3].p, E * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, w->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(w->stream));
  return OXG_OK;
  OXG_API_END
}

int oxg_check_edges(oxg_world w, oxg_store s, const uint32_t* from_idx, const uint32_t* to_idx,
                    uint64_t E, double resolution, uint8_t* valid) {
  OXG_API_BEGIN
  if (!w || !s || (E > 0 && (!from_idx || !to_idx || !valid))) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (!(resolution > 0.0)) {
    set_error("motion resolution must be > 0 (got %g): the reference loop would not terminate",
              resolution);
    return OXG_E_RESOLUTION;
  }
  if (E == 0) return OXG_OK;
  for (uint64_t e = 0; e < E; ++e)
    if (from_idx[e] >= s->size || to_idx[e] >= s->size) {
      set_error("edge %llu refers to a state outside the store (size %llu)", (unsigned long long)e,
                (unsigned long long)s->size);
      return OXG_E_ARG;
    }
  OXG_TRY(s->buf[B_AUX].ensure(E * sizeof(uint32_t)));
  OXG_TRY(s->buf[B_AUX2].ensure(E * sizeof(uint32_t)));
  OXG_TRY(s->buf[B_AUX3].ensure(E));
  OXG_CUDA_TRY(cudaMemcpyAsync(s->buf[B_AUX].p, from_idx, E * sizeof(uint32_t), cudaMemcpyHostToDevice,
                               s->stream));
  OXG_CUDA_TRY(cudaMemcpyAsync(s->buf[B_AUX2].p, to_idx, E * sizeof(uint32_t), cudaMemcpyHostToDevice,
                               s->stream));
  OXG_TRY(motion_check_device(w, s->space, s->stream, store_view(s), store_view(s),
                              s->buf[B_AUX].as<uint32_t>(), s->buf[B_AUX2].as<uint32_t>(), E,
                              nullptr, resolution, s->buf[B_AUX3].as<uint8_t>(), nullptr));
  OXG_CUDA_TRY(cudaMemcpyAsync(valid, s->buf[B_AUX3].p, E, cudaMemcpyDeviceToHost, s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  return OXG_OK;
  OXG_API_END
}

}  // extern "C"
