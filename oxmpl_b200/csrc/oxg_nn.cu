// oxg_nn.cu — nearest / k-nearest / radius scans over the SoA store.
//
// Replaces the planners' inline linear scans: 1-NN rrt.rs:187-196, rrt_connect.rs:128-136,
// rrt_star.rs:199-208; radius rrt_star.rs:121-131, prm.rs:131-138, prm.rs:249-256.
//
// Batched kernels ("query-per-thread"): each thread keeps QPT queries in registers and every
// stored state of a tile is broadcast to the whole CTA from shared memory, where the tile was
// put by the bulk-copy engine (D contiguous runs of the SoA columns).  The FP64 pipe is the
// bound: per (query, state) pair D subtractions, D multiplications and D-1 additions, each
// separately rounded in the reference's order, and ONE integer compare.
//
// Exactness without a sqrt per pair.  The reference compares sqrt(d2) values (strict '<',
// lowest index wins ties) and sqrt is monotone but not injective, so comparing d2 alone can
// pick another index.  The kernels filter on  d2 <= U  with  U = max{x : sqrt(x) <= s_worst}
// (recomputed only when a query's k-th best changes), which is EXACTLY  sqrt(d2) <= s_worst;
// the rare survivors take the slow path: sqrt, then the lexicographic (distance, index) test.
// d2 >= +0 always, so the filter is a signed 64-bit integer compare of the bit patterns
// (NaN patterns compare above +inf or negative and are rejected in the slow path).
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <vector>

#include "oxg_internal.cuh"
#include "oxg_math.cuh"
#include "oxg_tma.cuh"

namespace oxg {

constexpr int NN_THREADS = 128;
constexpr int NN_TILE = 128;   // stored states per pipeline stage
constexpr int NN_STAGES = 4;

enum { EXCL_NONE = 0, EXCL_ARRAY = 1, EXCL_SELF = 2 };

struct ScanParams {
  const double* db;   // SoA store
  uint64_t pitch;
  uint32_t n_db;
  StateView q;        // queries
  uint32_t Q;
  uint32_t k;
  int exclude_mode;
  const uint32_t* exclude;
  uint32_t self_base;
  uint32_t split_len;  // stored states per blockIdx.y slice (multiple of NN_TILE)
  uint32_t nsplit;
  uint32_t* out_idx;   // knn: [nsplit][Q][k]
  double* out_dist;
  // radius
  long long r_bits;            // bits of max{x : sqrt(x) < r}, or -1
  const uint32_t* idx_limit;   // nullable
  int limit_self;              // idx_limit[q] = self_base + q
  uint64_t* counts;            // [Q][nsplit] (count pass) / exclusive offsets (fill pass)
  uint32_t* hit_idx;
  double* hit_dist;            // nullable
};

// Sorted insert into a per-thread list laid out with stride `stride` (entry e at [e*stride]).
// Order: ascending (dist, idx).  Returns true if the list changed.
__device__ __forceinline__ bool list_insert(double* sl, uint32_t* il, int stride, int k, int& have,
                                            double s, uint32_t j) {
  if (have == k) {
    double sw = sl[(k - 1) * stride];
    uint32_t iw = il[(k - 1) * stride];
    if (!(s < sw || (s == sw && j < iw))) return false;
  }
  int pos = (have < k) ? have : k - 1;
  while (pos > 0) {
    double sp = sl[(pos - 1) * stride];
    uint32_t ip = il[(pos - 1) * stride];
    if (sp < s || (sp == s && ip < j)) break;
    sl[pos * stride] = sp;
    il[pos * stride] = ip;
    --pos;
  }
  sl[pos * stride] = s;
  il[pos * stride] = j;
  if (have < k) ++have;
  return true;
}

// Slow path of the k-NN filter: exact test + insert + threshold update.
__device__ __noinline__ void knn_slow(double* sl, uint32_t* il, int stride, int k, int* have_p,
                                      long long* ubits_p, double d2, uint32_t j, uint32_t excl) {
  if (j == excl || !(d2 == d2)) return;
  int have = *have_p;
  double s = sqrt(d2);
  if (list_insert(sl, il, stride, k, have, s, j)) {
    *have_p = have;
    if (have == k) *ubits_p = __double_as_longlong(sqrt_le_threshold_dev(sl[(k - 1) * stride]));
  }
}

template <int D>
__device__ __forceinline__ double sumsq_q(const double (&q)[D], const double (&p)[D]) {
  double s = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double df = __dsub_rn(q[d], p[d]);
    double m = __dmul_rn(df, df);
    s = (d == 0) ? m : __dadd_rn(s, m);
  }
  return s;
}

// Issues the D bulk copies of one tile (called by one thread).
template <int D>
__device__ __forceinline__ void issue_tile(const ScanParams& p, double* stage, uint64_t* bar,
                                           uint32_t base, uint32_t cnt) {
  uint32_t cnt_pad = (cnt + 1u) & ~1u;  // 16-byte granules; the pad element is never read
  uint32_t bytes = cnt_pad * 8u;
  mbar_arrive_expect_tx(bar, bytes * D);
#pragma unroll
  for (int d = 0; d < D; ++d)
    bulk_copy_g2s(stage + d * NN_TILE, p.db + (uint64_t)d * p.pitch + base, bytes, bar);
}

// ---------------------------------------------------------------------------------------
// k-NN (k = 1 is the planners' 1-NN).  grid = (ceil(Q / (NN_THREADS*QPT)), nsplit)
// ---------------------------------------------------------------------------------------
template <int D, int QPT>
__global__ void __launch_bounds__(NN_THREADS) knn_rv_kernel(const ScanParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NQ = NN_THREADS * QPT;
  double* tiles = reinterpret_cast<double*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + NN_STAGES * D * NN_TILE);
  double* s_list = reinterpret_cast<double*>(bars + NN_STAGES);
  uint32_t* i_list = reinterpret_cast<uint32_t*>(s_list + (size_t)p.k * NQ);

  const int tid = threadIdx.x;
  const int k = (int)p.k;
  double q[QPT][D];
  uint32_t qi[QPT], excl[QPT];
  long long ubits[QPT];
  int have[QPT];
#pragma unroll
  for (int a = 0; a < QPT; ++a) {
    qi[a] = blockIdx.x * NQ + a * NN_THREADS + tid;
    bool ok = qi[a] < p.Q;
    uint32_t qq = ok ? qi[a] : 0u;
#pragma unroll
    for (int d = 0; d < D; ++d) q[a][d] = p.q.p[(uint64_t)qq * p.q.stride_i + (uint64_t)d * p.q.stride_d];
    excl[a] = OXG_INVALID_INDEX;
    if (p.exclude_mode == EXCL_ARRAY) excl[a] = p.exclude[qq];
    if (p.exclude_mode == EXCL_SELF) excl[a] = p.self_base + qq;
    ubits[a] = ok ? 0x7FF0000000000000ll : -1ll;  // +inf: accept everything / never
    have[a] = 0;
  }

  const uint32_t n0 = blockIdx.y * p.split_len;
  const uint32_t n1 = min(p.n_db, n0 + p.split_len);
  const uint32_t ntiles = (n1 > n0) ? (n1 - n0 + NN_TILE - 1) / NN_TILE : 0;

  if (tid == 0) {
    for (int s = 0; s < NN_STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (uint32_t t = 0; t < (uint32_t)(NN_STAGES - 1) && t < ntiles; ++t) {
      uint32_t base = n0 + t * NN_TILE;
      issue_tile<D>(p, tiles + (t % NN_STAGES) * D * NN_TILE, &bars[t % NN_STAGES], base,
                    min((uint32_t)NN_TILE, n1 - base));
    }
  }

  for (uint32_t t = 0; t < ntiles; ++t) {
    if (tid == 0) {
      uint32_t tn = t + NN_STAGES - 1;
      if (tn < ntiles) {
        uint32_t base = n0 + tn * NN_TILE;
        issue_tile<D>(p, tiles + (tn % NN_STAGES) * D * NN_TILE, &bars[tn % NN_STAGES], base,
                      min((uint32_t)NN_TILE, n1 - base));
      }
    }
    mbar_wait(&bars[t % NN_STAGES], (t / NN_STAGES) & 1u);
    const double* sp = tiles + (t % NN_STAGES) * D * NN_TILE;
    const uint32_t base = n0 + t * NN_TILE;
    const int cnt = (int)min((uint32_t)NN_TILE, n1 - base);
    int j = 0;
    for (; j + 2 <= cnt; j += 2) {
      double p0[D], p1[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        double2 v = *reinterpret_cast<const double2*>(sp + d * NN_TILE + j);
        p0[d] = v.x;
        p1[d] = v.y;
      }
      double s0[QPT], s1[QPT];
      bool any = false;
#pragma unroll
      for (int a = 0; a < QPT; ++a) {
        s0[a] = sumsq_q<D>(q[a], p0);
        s1[a] = sumsq_q<D>(q[a], p1);
        any = any || (__double_as_longlong(s0[a]) <= ubits[a]) ||
              (__double_as_longlong(s1[a]) <= ubits[a]);
      }
      if (any) {
#pragma unroll
        for (int a = 0; a < QPT; ++a) {
          if (__double_as_longlong(s0[a]) <= ubits[a])
            knn_slow(s_list + a * NN_THREADS + tid, i_list + a * NN_THREADS + tid, NQ, k, &have[a],
                     &ubits[a], s0[a], base + j, excl[a]);
          if (__double_as_longlong(s1[a]) <= ubits[a])
            knn_slow(s_list + a * NN_THREADS + tid, i_list + a * NN_THREADS + tid, NQ, k, &have[a],
                     &ubits[a], s1[a], base + j + 1, excl[a]);
        }
      }
    }
    if (j < cnt) {
      double p0[D];
#pragma unroll
      for (int d = 0; d < D; ++d) p0[d] = sp[d * NN_TILE + j];
#pragma unroll
      for (int a = 0; a < QPT; ++a) {
        double s0 = sumsq_q<D>(q[a], p0);
        if (__double_as_longlong(s0) <= ubits[a])
          knn_slow(s_list + a * NN_THREADS + tid, i_list + a * NN_THREADS + tid, NQ, k, &have[a],
                   &ubits[a], s0, base + j, excl[a]);
      }
    }
    __syncthreads();  // everyone is done with this stage before it is refilled
  }

#pragma unroll
  for (int a = 0; a < QPT; ++a) {
    if (qi[a] >= p.Q) continue;
    uint64_t o = ((uint64_t)blockIdx.y * p.Q + qi[a]) * (uint64_t)k;
    const double* sl = s_list + a * NN_THREADS + tid;
    const uint32_t* il = i_list + a * NN_THREADS + tid;
    for (int e = 0; e < k; ++e) {
      bool f = e < have[a];
      p.out_idx[o + e] = f ? il[e * NQ] : OXG_INVALID_INDEX;
      p.out_dist[o + e] = f ? sl[e * NQ] : __longlong_as_double(0x7FF0000000000000ll);
    }
  }
}

// Merges the nsplit partial lists of each query (each ascending by (dist, idx)).
__global__ void knn_merge_kernel(const uint32_t* __restrict__ part_idx,
                                 const double* __restrict__ part_dist, uint32_t Q, uint32_t k,
                                 uint32_t nsplit, uint32_t* __restrict__ out_idx,
                                 double* __restrict__ out_dist) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* s_list = reinterpret_cast<double*>(smem_raw);
  uint32_t* i_list = reinterpret_cast<uint32_t*>(s_list + (size_t)k * blockDim.x);
  uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= Q) return;
  double* sl = s_list + threadIdx.x;
  uint32_t* il = i_list + threadIdx.x;
  int have = 0;
  for (uint32_t s = 0; s < nsplit; ++s) {
    uint64_t o = ((uint64_t)s * Q + qi) * k;
    for (uint32_t e = 0; e < k; ++e) {
      uint32_t j = part_idx[o + e];
      if (j == OXG_INVALID_INDEX) break;
      if (!list_insert(sl, il, (int)blockDim.x, (int)k, have, part_dist[o + e], j)) break;
    }
  }
  uint64_t o = (uint64_t)qi * k;
  for (uint32_t e = 0; e < k; ++e) {
    bool f = (int)e < have;
    out_idx[o + e] = f ? il[e * blockDim.x] : OXG_INVALID_INDEX;
    out_dist[o + e] = f ? sl[e * blockDim.x] : __longlong_as_double(0x7FF0000000000000ll);
  }
}

// ---------------------------------------------------------------------------------------
// Radius scan, two passes (count, then fill at the scanned offsets) so that the CSR output
// is ascending in index and deterministic.  One query per thread.
// ---------------------------------------------------------------------------------------
template <int D, bool FILL>
__global__ void __launch_bounds__(NN_THREADS) radius_rv_kernel(const ScanParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* tiles = reinterpret_cast<double*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + NN_STAGES * D * NN_TILE);
  __shared__ uint32_t s_maxlim;

  const int tid = threadIdx.x;
  const uint32_t qi = blockIdx.x * NN_THREADS + tid;
  const bool ok = qi < p.Q;
  const uint32_t qq = ok ? qi : 0u;
  double q[D];
#pragma unroll
  for (int d = 0; d < D; ++d) q[d] = p.q.p[(uint64_t)qq * p.q.stride_i + (uint64_t)d * p.q.stride_d];
  uint32_t lim = p.n_db;
  if (p.idx_limit) lim = min(lim, p.idx_limit[qq]);
  if (p.limit_self) lim = min(lim, p.self_base + qq);
  if (!ok) lim = 0;
  const long long rbits = ok ? p.r_bits : -1ll;

  if (tid == 0) s_maxlim = 0;
  __syncthreads();
  atomicMax(&s_maxlim, lim);
  if (tid == 0) {
    for (int s = 0; s < NN_STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const uint32_t n0 = blockIdx.y * p.split_len;
  const uint32_t n1 = min(min(p.n_db, s_maxlim), n0 + p.split_len);
  const uint32_t ntiles = (n1 > n0) ? (n1 - n0 + NN_TILE - 1) / NN_TILE : 0;

  uint64_t cursor = 0;
  if (FILL && ok) cursor = p.counts[(uint64_t)qi * p.nsplit + blockIdx.y];
  uint32_t count = 0;

  if (tid == 0) {
    for (uint32_t t = 0; t < (uint32_t)(NN_STAGES - 1) && t < ntiles; ++t) {
      uint32_t base = n0 + t * NN_TILE;
      issue_tile<D>(p, tiles + (t % NN_STAGES) * D * NN_TILE, &bars[t % NN_STAGES], base,
                    min((uint32_t)NN_TILE, n1 - base));
    }
  }
  for (uint32_t t = 0; t < ntiles; ++t) {
    if (tid == 0) {
      uint32_t tn = t + NN_STAGES - 1;
      if (tn < ntiles) {
        uint32_t base = n0 + tn * NN_TILE;
        issue_tile<D>(p, tiles + (tn % NN_STAGES) * D * NN_TILE, &bars[tn % NN_STAGES], base,
                      min((uint32_t)NN_TILE, n1 - base));
      }
    }
    mbar_wait(&bars[t % NN_STAGES], (t / NN_STAGES) & 1u);
    const double* sp = tiles + (t % NN_STAGES) * D * NN_TILE;
    const uint32_t base = n0 + t * NN_TILE;
    const int cnt = (int)min((uint32_t)NN_TILE, n1 - base);
    for (int j = 0; j < cnt; ++j) {
      double pt[D];
#pragma unroll
      for (int d = 0; d < D; ++d) pt[d] = sp[d * NN_TILE + j];
      double s = sumsq_q<D>(q, pt);
      if (__double_as_longlong(s) <= rbits) {
        uint32_t jj = base + j;
        if (jj < lim && s == s) {
          if (FILL) {
            p.hit_idx[cursor + count] = jj;
            if (p.hit_dist) p.hit_dist[cursor + count] = sqrt(s);
          }
          ++count;
        }
      }
    }
    __syncthreads();
  }
  if (!FILL && ok) p.counts[(uint64_t)qi * p.nsplit + blockIdx.y] = count;
}

// ---------------------------------------------------------------------------------------
// Generic spaces (SO2, compound R^n x SO2, or RealVector with D > 16): runtime dimensions,
// full distance per pair in the reference's order, stored states read through the L1/L2 as
// warp-uniform loads.  MODE 0 = k-NN, 1 = radius count, 2 = radius fill.
// ---------------------------------------------------------------------------------------
struct GenericParams {
  ScanParams s;
  uint32_t n_real, n_so2;
  double w_real, w_so2;
  int stored_first;  // distance(stored, query) (1-NN call sites) vs distance(query, stored)
  double radius;
};

template <int MODE>
__global__ void __launch_bounds__(NN_THREADS) scan_generic_kernel(const GenericParams g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const ScanParams& p = g.s;
  const int tid = threadIdx.x;
  const int dim = (int)(g.n_real + g.n_so2);
  const uint32_t qi = blockIdx.x * NN_THREADS + tid;
  const bool ok = qi < p.Q;
  const uint32_t qq = ok ? qi : 0u;
  double q[kMaxDim], pt[kMaxDim];
  for (int d = 0; d < dim; ++d) q[d] = p.q.p[(uint64_t)qq * p.q.stride_i + (uint64_t)d * p.q.stride_d];

  const uint32_t n0 = blockIdx.y * p.split_len;
  uint32_t n1 = min(p.n_db, n0 + p.split_len);

  if (MODE == 0) {
    double* s_list = reinterpret_cast<double*>(smem_raw);
    uint32_t* i_list = reinterpret_cast<uint32_t*>(s_list + (size_t)p.k * NN_THREADS);
    double* sl = s_list + tid;
    uint32_t* il = i_list + tid;
    const int k = (int)p.k;
    int have = 0;
    uint32_t excl = OXG_INVALID_INDEX;
    if (p.exclude_mode == EXCL_ARRAY) excl = p.exclude[qq];
    if (p.exclude_mode == EXCL_SELF) excl = p.self_base + qq;
    double worst = __longlong_as_double(0x7FF0000000000000ll);
    if (ok) {
      for (uint32_t j = n0; j < n1; ++j) {
        for (int d = 0; d < dim; ++d) pt[d] = __ldg(p.db + (uint64_t)d * p.pitch + j);
        double dist = g.stored_first
                          ? space_distance_dev(g.n_real, g.n_so2, g.w_real, g.w_so2, pt, q)
                          : space_distance_dev(g.n_real, g.n_so2, g.w_real, g.w_so2, q, pt);
        if (dist <= worst && j != excl) {
          if (list_insert(sl, il, NN_THREADS, k, have, dist, j) && have == k)
            worst = sl[(k - 1) * NN_THREADS];
        }
      }
      uint64_t o = ((uint64_t)blockIdx.y * p.Q + qi) * (uint64_t)k;
      for (int e = 0; e < k; ++e) {
        bool f = e < have;
        p.out_idx[o + e] = f ? il[e * NN_THREADS] : OXG_INVALID_INDEX;
        p.out_dist[o + e] = f ? sl[e * NN_THREADS] : __longlong_as_double(0x7FF0000000000000ll);
      }
    }
  } else {
    if (!ok) return;
    uint32_t lim = p.n_db;
    if (p.idx_limit) lim = min(lim, p.idx_limit[qq]);
    if (p.limit_self) lim = min(lim, p.self_base + qq);
    n1 = min(n1, lim);
    uint64_t cursor = (MODE == 2) ? p.counts[(uint64_t)qi * p.nsplit + blockIdx.y] : 0;
    uint32_t count = 0;
    for (uint32_t j = n0; j < n1; ++j) {
      for (int d = 0; d < dim; ++d) pt[d] = __ldg(p.db + (uint64_t)d * p.pitch + j);
      double dist = space_distance_dev(g.n_real, g.n_so2, g.w_real, g.w_so2, q, pt);
      if (dist < g.radius) {
        if (MODE == 2) {
          p.hit_idx[cursor + count] = j;
          if (p.hit_dist) p.hit_dist[cursor + count] = dist;
        }
        ++count;
      }
    }
    if (MODE == 1) p.counts[(uint64_t)qi * p.nsplit + blockIdx.y] = count;
  }
}

// ---------------------------------------------------------------------------------------
// Host-side launch logic
// ---------------------------------------------------------------------------------------
template <int D, int QPT>
static size_t knn_smem_bytes(uint32_t k) {
  return sizeof(double) * NN_STAGES * D * NN_TILE + sizeof(uint64_t) * NN_STAGES +
         (size_t)k * NN_THREADS * QPT * (sizeof(double) + sizeof(uint32_t));
}

template <int D, int QPT>
static int launch_knn_rv(const ScanParams& p, dim3 grid, cudaStream_t st) {
  size_t smem = knn_smem_bytes<D, QPT>(p.k);
  OXG_CUDA_TRY(cudaFuncSetAttribute(knn_rv_kernel<D, QPT>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  knn_rv_kernel<D, QPT><<<grid, NN_THREADS, smem, st>>>(p);
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

template <int D>
static int launch_radius_rv(const ScanParams& p, dim3 grid, bool fill, cudaStream_t st) {
  size_t smem = sizeof(double) * NN_STAGES * D * NN_TILE + sizeof(uint64_t) * NN_STAGES;
  if (fill) {
    OXG_CUDA_TRY(cudaFuncSetAttribute(radius_rv_kernel<D, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    radius_rv_kernel<D, true><<<grid, NN_THREADS, smem, st>>>(p);
  } else {
    OXG_CUDA_TRY(cudaFuncSetAttribute(radius_rv_kernel<D, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    radius_rv_kernel<D, false><<<grid, NN_THREADS, smem, st>>>(p);
  }
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

#define OXG_DISPATCH_D(D_RT, CALL)  \
  switch (D_RT) {                   \
    case 1: { constexpr int DD = 1; CALL; } break;   \
    case 2: { constexpr int DD = 2; CALL; } break;   \
    case 3: { constexpr int DD = 3; CALL; } break;   \
    case 4: { constexpr int DD = 4; CALL; } break;   \
    case 5: { constexpr int DD = 5; CALL; } break;   \
    case 6: { constexpr int DD = 6; CALL; } break;   \
    case 7: { constexpr int DD = 7; CALL; } break;   \
    case 8: { constexpr int DD = 8; CALL; } break;   \
    case 9: { constexpr int DD = 9; CALL; } break;   \
    case 10: { constexpr int DD = 10; CALL; } break; \
    case 11: { constexpr int DD = 11; CALL; } break; \
    case 12: { constexpr int DD = 12; CALL; } break; \
    case 13: { constexpr int DD = 13; CALL; } break; \
    case 14: { constexpr int DD = 14; CALL; } break; \
    case 15: { constexpr int DD = 15; CALL; } break; \
    case 16: { constexpr int DD = 16; CALL; } break; \
    default: break;                 \
  }

static uint32_t round_up_u32(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

// Chooses how many slices of the store each query tile is split into, so that small query
// batches still fill the GPU.
static void choose_split(uint32_t n_db, uint64_t qtiles, uint32_t* nsplit, uint32_t* split_len) {
  const uint64_t target_ctas = 148ull * 6;
  uint64_t want = (target_ctas + qtiles - 1) / qtiles;
  uint64_t max_split = std::max<uint64_t>(1, n_db / (NN_TILE * 4));
  uint64_t ns = std::min<uint64_t>(std::min<uint64_t>(want, max_split), 2048);
  if (ns < 1) ns = 1;
  uint32_t len = round_up_u32((uint32_t)((n_db + ns - 1) / ns), NN_TILE);
  if (len == 0) len = NN_TILE;
  *split_len = len;
  *nsplit = std::max<uint32_t>(1, (n_db + len - 1) / len);
}

// k-NN of Q device-resident queries against the store; results in device memory.
int nn_knn_device(oxg_store_s* s, StateView q, uint64_t Q, uint32_t k, int exclude_mode,
                  const uint32_t* d_exclude, uint64_t self_base, bool stored_first, uint32_t* d_idx,
                  double* d_dist) {
  if (Q == 0) return OXG_OK;
  if (s->size == 0) {
    set_error("nearest-neighbour query against an empty store");
    return OXG_E_EMPTY;
  }
  if (k == 0 || k > kMaxK) {
    set_error("k must be in [1, %u]", kMaxK);
    return OXG_E_ARG;
  }
  if (Q > 0xFFFFFFF0ull) {
    set_error("too many queries in one batch");
    return OXG_E_ARG;
  }
  ScanParams p;
  memset(&p, 0, sizeof(p));
  p.db = s->d_data;
  p.pitch = s->pitch;
  p.n_db = (uint32_t)s->size;
  p.q = q;
  p.Q = (uint32_t)Q;
  p.k = k;
  p.exclude_mode = exclude_mode;
  p.exclude = d_exclude;
  p.self_base = (uint32_t)self_base;

  const bool fast = is_pure_rv(s->space) && s->dim <= (uint32_t)kMaxFastDim;
  // two queries per thread once there is enough work to fill the machine that way
  const bool qpt2 = fast && Q >= 2ull * 148 * NN_THREADS * 2 && k <= 32;
  const uint32_t nq_per_cta = NN_THREADS * (qpt2 ? 2 : 1);
  const uint64_t qtiles = (Q + nq_per_cta - 1) / nq_per_cta;
  choose_split(p.n_db, qtiles, &p.nsplit, &p.split_len);

  uint32_t* part_idx = d_idx;
  double* part_dist = d_dist;
  if (p.nsplit > 1) {
    size_t n = (size_t)p.nsplit * Q * k;
    OXG_TRY(s->buf[B_PIDX].ensure(n * sizeof(uint32_t)));
    OXG_TRY(s->buf[B_PDIST].ensure(n * sizeof(double)));
    part_idx = s->buf[B_PIDX].as<uint32_t>();
    part_dist = s->buf[B_PDIST].as<double>();
  }
  p.out_idx = part_idx;
  p.out_dist = part_dist;
  dim3 grid((unsigned)qtiles, p.nsplit);

  if (fast) {
    int r = OXG_E_INTERNAL;
    if (qpt2) {
      OXG_DISPATCH_D(s->dim, r = (launch_knn_rv<DD, 2>(p, grid, s->stream)));
    } else {
      OXG_DISPATCH_D(s->dim, r = (launch_knn_rv<DD, 1>(p, grid, s->stream)));
    }
    OXG_TRY(r);
  } else {
    GenericParams g;
    g.s = p;
    g.n_real = s->space.n_real;
    g.n_so2 = s->space.n_so2;
    g.w_real = s->space.w_real;
    g.w_so2 = s->space.w_so2;
    g.stored_first = stored_first ? 1 : 0;
    g.radius = 0.0;
    size_t smem = (size_t)k * NN_THREADS * (sizeof(double) + sizeof(uint32_t));
    OXG_CUDA_TRY(cudaFuncSetAttribute(scan_generic_kernel<0>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    scan_generic_kernel<0><<<grid, NN_THREADS, smem, s->stream>>>(g);
    OXG_LAUNCH_CHECK();
  }
  if (p.nsplit > 1) {
    int threads = 64;
    size_t smem = (size_t)k * threads * (sizeof(double) + sizeof(uint32_t));
    OXG_CUDA_TRY(cudaFuncSetAttribute(knn_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    knn_merge_kernel<<<(unsigned)((Q + threads - 1) / threads), threads, smem, s->stream>>>(
        part_idx, part_dist, (uint32_t)Q, k, p.nsplit, d_idx, d_dist);
    OXG_LAUNCH_CHECK();
  }
  return OXG_OK;
}

// Radius scan of Q device-resident queries; results left in the store's scratch arenas:
//   per-(query, slice) exclusive offsets in B_OFFSETS, idx in B_PIDX, dist in B_PDIST.
// *total receives offsets[Q].  Synchronises the stream (the total is needed on the host).
int nn_radius_device(oxg_store_s* s, StateView q, uint64_t Q, double r, const uint32_t* d_idx_limit,
                     bool limit_self, uint64_t self_base, bool want_dist, uint64_t* total) {
  *total = 0;
  if (Q > 0xFFFFFFF0ull) {
    set_error("too many queries in one batch");
    return OXG_E_ARG;
  }
  ScanParams p;
  memset(&p, 0, sizeof(p));
  p.db = s->d_data;
  p.pitch = s->pitch;
  p.n_db = (uint32_t)s->size;
  p.q = q;
  p.Q = (uint32_t)Q;
  p.k = 0;
  p.self_base = (uint32_t)self_base;
  p.idx_limit = d_idx_limit;
  p.limit_self = limit_self ? 1 : 0;
  p.r_bits = (long long)0;
  double thr = sqrt_lt_threshold(r);
  if (thr < 0.0) {
    p.r_bits = -1ll;
  } else {
    memcpy(&p.r_bits, &thr, sizeof(double));
  }
  const uint64_t qtiles = (Q + NN_THREADS - 1) / NN_THREADS;
  choose_split(std::max<uint32_t>(p.n_db, 1), std::max<uint64_t>(qtiles, 1), &p.nsplit, &p.split_len);

  size_t ncounts = (size_t)Q * p.nsplit + 1;
  OXG_TRY(s->buf[B_OFFSETS].ensure(ncounts * sizeof(uint64_t)));
  OXG_TRY(s->buf[B_COUNTS].ensure(ncounts * sizeof(uint64_t)));
  uint64_t* counts = s->buf[B_COUNTS].as<uint64_t>();
  uint64_t* offsets = s->buf[B_OFFSETS].as<uint64_t>();
  OXG_CUDA_TRY(cudaMemsetAsync(counts, 0, ncounts * sizeof(uint64_t), s->stream));
  if (Q == 0 || p.n_db == 0) {
    OXG_CUDA_TRY(cudaMemsetAsync(offsets, 0, ncounts * sizeof(uint64_t), s->stream));
    OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    return OXG_OK;
  }
  p.counts = counts;
  dim3 grid((unsigned)qtiles, p.nsplit);
  const bool fast = is_pure_rv(s->space) && s->dim <= (uint32_t)kMaxFastDim;
  GenericParams g;
  g.n_real = s->space.n_real;
  g.n_so2 = s->space.n_so2;
  g.w_real = s->space.w_real;
  g.w_so2 = s->space.w_so2;
  g.stored_first = 0;
  g.radius = r;
  if (fast) {
    int rr = OXG_E_INTERNAL;
    OXG_DISPATCH_D(s->dim, rr = (launch_radius_rv<DD>(p, grid, false, s->stream)));
    OXG_TRY(rr);
  } else {
    g.s = p;
    scan_generic_kernel<1><<<grid, NN_THREADS, 0, s->stream>>>(g);
    OXG_LAUNCH_CHECK();
  }
  // exclusive scan of the (query-major, slice-minor) counts; the trailing zero makes
  // offsets[Q*nsplit] the total.
  size_t tmp_bytes = 0;
  OXG_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, offsets, (int)ncounts,
                                             s->stream));
  OXG_TRY(s->buf[B_TMP].ensure(tmp_bytes));
  OXG_CUDA_TRY(cub::DeviceScan::ExclusiveSum(s->buf[B_TMP].p, tmp_bytes, counts, offsets,
                                             (int)ncounts, s->stream));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  uint64_t tot = 0;
  OXG_CUDA_TRY(cudaMemcpyAsync(&tot, offsets + (ncounts - 1), sizeof(uint64_t),
                               cudaMemcpyDeviceToHost, s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  *total = tot;
  if (tot == 0) return OXG_OK;
  OXG_TRY(s->buf[B_PIDX].ensure(tot * sizeof(uint32_t)));
  if (want_dist) OXG_TRY(s->buf[B_PDIST].ensure(tot * sizeof(double)));
  p.counts = offsets;
  p.hit_idx = s->buf[B_PIDX].as<uint32_t>();
  p.hit_dist = want_dist ? s->buf[B_PDIST].as<double>() : nullptr;
  if (fast) {
    int rr = OXG_E_INTERNAL;
    OXG_DISPATCH_D(s->dim, rr = (launch_radius_rv<DD>(p, grid, true, s->stream)));
    OXG_TRY(rr);
  } else {
    g.s = p;
    scan_generic_kernel<2><<<grid, NN_THREADS, 0, s->stream>>>(g);
    OXG_LAUNCH_CHECK();
  }
  return OXG_OK;
}

// offsets_out[q] = scanned[q * nsplit]  (q = 0..Q)
__global__ void pick_offsets_kernel(const uint64_t* __restrict__ scanned, uint64_t Q, uint32_t nsplit,
                                    uint64_t* __restrict__ out) {
  uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q <= Q) out[q] = scanned[q * nsplit];
}

int nn_radius_nsplit(const oxg_store_s* s, uint64_t Q) {
  uint32_t nsplit, split_len;
  const uint64_t qtiles = (Q + NN_THREADS - 1) / NN_THREADS;
  choose_split(std::max<uint32_t>((uint32_t)s->size, 1), std::max<uint64_t>(qtiles, 1), &nsplit,
               &split_len);
  return (int)nsplit;
}

// Compacts B_OFFSETS (per-slice offsets) into per-query CSR offsets in `d_out` (Q+1 entries).
int nn_radius_offsets(oxg_store_s* s, uint64_t Q, uint64_t* d_out) {
  uint32_t nsplit = (uint32_t)nn_radius_nsplit(s, Q);
  int threads = 256;
  pick_offsets_kernel<<<(unsigned)((Q + 1 + threads - 1) / threads), threads, 0, s->stream>>>(
      s->buf[B_OFFSETS].as<uint64_t>(), Q, nsplit, d_out);
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

}  // namespace oxg

using namespace oxg;

extern "C" {

static int upload_queries(oxg_store s, const double* queries, uint64_t Q) {
  size_t bytes = sizeof(double) * Q * s->dim;
  OXG_TRY(s->buf[B_QUERY].ensure(bytes));
  OXG_CUDA_TRY(cudaMemcpyAsync(s->buf[B_QUERY].p, queries, bytes, cudaMemcpyHostToDevice, s->stream));
  return OXG_OK;
}

static int knn_host(oxg_store s, const double* queries, uint64_t Q, uint32_t k,
                    const uint32_t* exclude_idx, bool stored_first, uint32_t* idx, double* dist) {
  if (!s || !queries || !idx || !dist) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (Q == 0) return OXG_OK;
  if (s->size == 0) {
    set_error("nearest-neighbour query against an empty store");
    return OXG_E_EMPTY;
  }
  if (k == 0 || k > kMaxK) {
    set_error("k must be in [1, %u]", kMaxK);
    return OXG_E_ARG;
  }
  OXG_TRY(upload_queries(s, queries, Q));
  size_t n = (size_t)Q * k;
  OXG_TRY(s->buf[B_OIDX].ensure(n * sizeof(uint32_t)));
  OXG_TRY(s->buf[B_ODIST].ensure(n * sizeof(double)));
  uint32_t* d_idx = s->buf[B_OIDX].as<uint32_t>();
  double* d_dist = s->buf[B_ODIST].as<double>();
  uint32_t* d_excl = nullptr;
  if (exclude_idx) {
    OXG_TRY(s->buf[B_AUX].ensure(Q * sizeof(uint32_t)));
    d_excl = s->buf[B_AUX].as<uint32_t>();
    OXG_CUDA_TRY(cudaMemcpyAsync(d_excl, exclude_idx, Q * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                 s->stream));
  }
  OXG_TRY(nn_knn_device(s, rows_view(s->buf[B_QUERY].as<double>(), s->dim), Q, k,
                        exclude_idx ? EXCL_ARRAY : EXCL_NONE, d_excl, 0, stored_first, d_idx, d_dist));
  OXG_CUDA_TRY(cudaMemcpyAsync(idx, d_idx, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  OXG_CUDA_TRY(cudaMemcpyAsync(dist, d_dist, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  return OXG_OK;
}

int oxg_nn1(oxg_store s, const double* queries, uint64_t Q, uint32_t* idx, double* dist) {
  OXG_API_BEGIN
  return knn_host(s, queries, Q, 1, nullptr, /*stored_first=*/true, idx, dist);
  OXG_API_END
}

int oxg_knn(oxg_store s, const double* queries, uint64_t Q, uint32_t k, const uint32_t* exclude_idx,
            uint32_t* idx, double* dist) {
  OXG_API_BEGIN
  return knn_host(s, queries, Q, k, exclude_idx, /*stored_first=*/false, idx, dist);
  OXG_API_END
}

int oxg_radius(oxg_store s, const double* queries, uint64_t Q, double r, const uint32_t* idx_limit,
               uint64_t* offsets, uint32_t* idx, uint64_t cap, double* dist) {
  OXG_API_BEGIN
  if (!s || !offsets || (Q > 0 && !queries)) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  offsets[0] = 0;
  if (Q == 0) return OXG_OK;
  OXG_TRY(upload_queries(s, queries, Q));
  uint32_t* d_lim = nullptr;
  if (idx_limit) {
    OXG_TRY(s->buf[B_AUX].ensure(Q * sizeof(uint32_t)));
    d_lim = s->buf[B_AUX].as<uint32_t>();
    OXG_CUDA_TRY(cudaMemcpyAsync(d_lim, idx_limit, Q * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                 s->stream));
  }
  uint64_t total = 0;
  OXG_TRY(nn_radius_device(s, rows_view(s->buf[B_QUERY].as<double>(), s->dim), Q, r, d_lim, false, 0,
                           dist != nullptr, &total));
  // per-query offsets
  OXG_TRY(s->buf[B_AUX2].ensure((Q + 1) * sizeof(uint64_t)));
  OXG_TRY(nn_radius_offsets(s, Q, s->buf[B_AUX2].as<uint64_t>()));
  OXG_CUDA_TRY(cudaMemcpyAsync(offsets, s->buf[B_AUX2].p, (Q + 1) * sizeof(uint64_t),
                               cudaMemcpyDeviceToHost, s->stream));
  if (total > cap || (total > 0 && !idx)) {
    OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    set_error("radius result needs %llu entries, capacity is %llu", (unsigned long long)total,
              (unsigned long long)cap);
    return OXG_E_CAPACITY;
  }
  if (total > 0) {
    OXG_CUDA_TRY(cudaMemcpyAsync(idx, s->buf[B_PIDX].p, total * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                 s->stream));
    if (dist)
      OXG_CUDA_TRY(cudaMemcpyAsync(dist, s->buf[B_PDIST].p, total * sizeof(double),
                                   cudaMemcpyDeviceToHost, s->stream));
  }
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  return OXG_OK;
  OXG_API_END
}

int oxg_knn_rows_dev(oxg_store s, uint64_t row_begin, uint64_t row_end, uint32_t k, uint32_t* d_idx,
                     double* d_dist) {
  OXG_API_BEGIN
  if (!s || !d_idx || !d_dist) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (row_begin > row_end || row_end > s->size) {
    set_error("row range [%llu, %llu) outside the store (size %llu)", (unsigned long long)row_begin,
              (unsigned long long)row_end, (unsigned long long)s->size);
    return OXG_E_ARG;
  }
  StateView q = store_view(s);
  q.p += row_begin;
  return nn_knn_device(s, q, row_end - row_begin, k, EXCL_SELF, nullptr, row_begin, false, d_idx,
                       d_dist);
  OXG_API_END
}

}  // extern "C"
