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
//
// Certified FP32 pre-filter (knn_rv_f32_kernel, radius_rv_f32_kernel).  3D FP64 operations per
// pair at 64 lanes/clk/SM is the floor of the all-FP64 kernel, and all but ~1e-4 of the pairs
// are rejected by a wide margin.  So each pair is first evaluated in FP32 from the store's f32
// mirror (D packed subtracts + D packed FMAs, FADD2/FFMA2) and compared with a per-query
// threshold T32 that dominates U by a rigorous bound on the FP32 evaluation error
// (filter_threshold_dev).  A pair is discarded only if  a32 > T32, which PROVES d2 > U; every
// other pair (including anything NaN/overflowed) is recomputed from the f64 store in the
// reference's operation order and takes the exact path above.  Results are unchanged bit for
// bit; only certain rejections get cheaper.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdlib>
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
  const float* db32;  // f32 mirror of the store (same pitch), filter input only
  const unsigned long long* maxabs;  // bits of max |coordinate| over the store
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

// ---------------------------------------------------------------------------------------
// Certified FP32 pre-filter
// ---------------------------------------------------------------------------------------
constexpr int F_TILE = 256;   // stored states per pipeline stage (f32 mirror)
constexpr int F_STAGES = 3;

// Threshold for the FP32 evaluation a32 = fl32-chain( sum_d fma(df_d, df_d, .) ), df_d =
// fl32(fl32(q_d) - fl32(p_d)), such that   d2 <= U  (the exact FP64 admission test)  implies
// a32 <= T32.   With u = 2^-24, M >= every |q_d|, |p_d|, t_d = q_d - p_d, S = sum t_d^2:
//   * inputs: |fl32(x) - x| <= u|x| (+ 2^-150 when subnormal), so
//     df_d = t_d + e_d with |e_d| <= 2uM(1+u) + u|t_d| (1+2u)  <=  a + 2u|t_d|,  a := 4uM + 1e-37
//   * sum df_d^2 <= S(1 + 4u + 8u^2) + 2a sqrt(D S) + 2 D a^2          (Cauchy-Schwarz)
//   * the FP32 FMA chain adds at most a factor (1+u)^D <= 1 + D 2^-22 and 2^-149 per step
//   * the FP64 chain loses at most a factor (1 - (D+2) 2^-53):  S <= U (1 + 1e-12)
// The constants below are deliberately generous (1e-5 relative); the filter's selectivity is
// decided by U, which is thousands of times smaller than a typical d2.
__device__ __forceinline__ float filter_threshold_dev(long long ubits, double M, int D) {
  const float inf = __int_as_float(0x7F800000);
  if (ubits < 0) return -1.0f;  // nothing is admitted; a32 >= 0 > -1 discards, NaN still passes
  const double U = __longlong_as_double(ubits);
  if (!(U < 1.0e300)) return inf;
  const double a = M * 2.384185791015625e-07 + 1e-37;  // 2^-22 M
  const double S = U * 1.00001;
  double T = S + 2.0 * a * sqrt((double)D * S) + 2.0 * (double)D * a * a + 1e-35;
  T = T * (1.0 + (double)D * 2.384185791015625e-07);
  if (!(T == T)) return inf;
  return __double2float_ru(T);  // rounds up; overflows to +inf
}

__device__ __forceinline__ float2 sub2(float q, float2 p) {  // (q - p.x, q - p.y): one FADD2
  float2 qq = make_float2(q, q), r;
  asm("sub.f32x2 %0, %1, %2;"
      : "=l"(*reinterpret_cast<unsigned long long*>(&r))
      : "l"(*reinterpret_cast<unsigned long long*>(&qq)),
        "l"(*reinterpret_cast<unsigned long long*>(&p)));
  return r;
}

template <int D>
__device__ __forceinline__ double exact_sumsq_global(const ScanParams& p, uint32_t qi, uint32_t j,
                                                     double* maxq) {
  double s = 0.0, mq = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double qd = p.q.p[(uint64_t)qi * p.q.stride_i + (uint64_t)d * p.q.stride_d];
    const double pd = p.db[(uint64_t)d * p.pitch + j];
    mq = fmax(mq, fabs(qd));  // fmax drops NaN; a NaN query is caught by the d2 == d2 test
    const double df = __dsub_rn(qd, pd);
    const double m = __dmul_rn(df, df);
    s = (d == 0) ? m : __dadd_rn(s, m);
  }
  *maxq = mq;
  return s;
}

// Upper bound of max{x : sqrt(x) <= sw} (sw >= 0): sqrt(x) <= sw after rounding implies
// x < sw^2 (1 + 2^-53)^2, so fl_up(sw*sw) (1 + 2^-50) dominates it.  Tiny values (where sw*sw
// is subnormal and the relative bound fails) take the exact routine.  Used only to gate the
// exact (sqrt, index) comparison, so any upper bound is correct.
__device__ __forceinline__ double sqrt_le_upper_dev(double sw) {
  const double u = __dmul_ru(sw, sw);
  if (!(u > 1.0e-290)) return sqrt_le_threshold_dev(sw);
  return __dadd_ru(u, __dmul_ru(u, 8.8817841970012523e-16));  // 2^-50
}

constexpr int CAND_CAP = 16;  // deferred candidates per query (u32 indices, shared memory)
constexpr int DRAIN_EVERY = 16;  // tiles between scheduled warp-wide drains
constexpr int DOT_STAGES = 4;   // tile ring of the dot kernel (full/empty mbarrier pairs)
#ifndef OXG_DOT_MIN_CTAS
#define OXG_DOT_MIN_CTAS 3
#endif
constexpr int DOT_MIN_CTAS = OXG_DOT_MIN_CTAS; // CTAs per SM the dot kernel is compiled for (caps registers)

// Dot-product form of the pre-filter.  With g_d = -2 fl32(q_d), pn = fl32(|p|^2) and the FP32
// FMA chain  A = fma(g_D, p_D, ... fma(g_1, p_1, pn)),  A approximates |p|^2 - 2 q.p = d2 - |q|^2
// with  |A - (d2 - |q|^2)| <= gamma_{D+2} (2|p|^2 + |q|^2)  (one rounding per input, per FMA and
// of pn; 2|q.p| <= |q|^2 + |p|^2).  So d2 <= U implies  A <= U - |q|^2 + B  with
// B = gamma_{D+2} (2 Pmax + |q|^2),  Pmax = max squared norm over the store.  Half the FP32
// work of the difference form (D FMAs per pair), but the bound grows with |p|^2 instead of
// |p| sqrt(U): each warp watches its pass rate and falls back to the difference form when the
// bound is too loose for its data (adaptive, results unaffected either way).
__device__ __forceinline__ float filter_threshold_dot_dev(long long ubits, double Qn, double Pmax,
                                                          int D) {
  const float inf = __int_as_float(0x7F800000);
  if (ubits < 0) return -inf;  // nothing is admitted (NaN evaluations still pass to the exact path)
  const double U = __longlong_as_double(ubits);
  const double W = 2.0 * Pmax + Qn;
  if (!(U < 1.0e300) || !(W < 1.0e30)) return inf;
  const double B = W * (double)(D + 2) * 1.1920928955078125e-07 * 1.0001 + 1e-30;  // 2^-23: 2x margin
  const double T = U * (1.0 + 1e-9) - Qn + B + Qn * 1e-13;
  if (!(T == T)) return inf;
  return __double2float_ru(T);
}

struct FilterThr {
  float diff;  // difference-form threshold (filter_threshold_dev)
  float dot;   // dot-form threshold (filter_threshold_dot_dev)
};

// Exact path of the pairs the FP32 filter could not discard, run for all of a query's deferred
// candidates at once (the whole warp drains together, so the lanes stay converged).  Each
// candidate is recomputed from the f64 store in the reference's operation order and goes
// through the exact (distance, index) comparison.  Returns the query's new FP32 thresholds.
template <int D>
__device__ __noinline__ FilterThr knn_drain_f32(const ScanParams& p, uint32_t qi, uint32_t n_end,
                                                uint32_t excl, const uint32_t* cand, int ncand,
                                                int cstride, int stride, double* sl, uint32_t* il,
                                                int* have_p, long long* ubits_p, FilterThr thr) {
  if (ncand == 0) return thr;
  const int k = (int)p.k;
  double q[D], mq = 0.0, qn = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    q[d] = p.q.p[(uint64_t)qi * p.q.stride_i + (uint64_t)d * p.q.stride_d];
    mq = fmax(mq, fabs(q[d]));  // fmax drops NaN; a NaN query fails the d2 == d2 test below
    qn += q[d] * q[d];
  }
  int have = *have_p;
  long long ub = *ubits_p;
  bool changed = false;
  for (int c = 0; c < ncand; ++c) {
    const uint32_t j = cand[c * cstride];
    if (j >= n_end || j == excl) continue;
    double d2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double df = __dsub_rn(q[d], p.db[(uint64_t)d * p.pitch + j]);
      const double m = __dmul_rn(df, df);
      d2 = (d == 0) ? m : __dadd_rn(d2, m);
    }
    if (!(__double_as_longlong(d2) <= ub) || !(d2 == d2)) continue;
    if (list_insert(sl, il, stride, k, have, sqrt(d2), j)) {
      changed = true;
      if (have == k) ub = __double_as_longlong(sqrt_le_upper_dev(sl[(k - 1) * stride]));
    }
  }
  if (changed) {
    *have_p = have;
    if (have == k) {
      *ubits_p = ub;
      const double ms = __longlong_as_double((long long)p.maxabs[0]);
      const double M = (ms > mq || !(ms == ms)) ? ms : mq;
      thr.diff = filter_threshold_dev(ub, M, D);
      thr.dot = filter_threshold_dot_dev(ub, qn, __longlong_as_double((long long)p.maxabs[1]), D);
    }
  }
  return thr;
}

// Issues the D bulk copies of one f32 tile (called by one thread).
template <int D>
__device__ __forceinline__ void issue_tile32(const ScanParams& p, float* stage, uint64_t* bar,
                                             uint32_t base, uint32_t cnt) {
  uint32_t cnt_pad = (cnt + 3u) & ~3u;  // 16-byte granules; pad elements are index-checked
  uint32_t bytes = cnt_pad * 4u;
  mbar_arrive_expect_tx(bar, bytes * D);
#pragma unroll
  for (int d = 0; d < D; ++d)
    bulk_copy_g2s(stage + d * F_TILE, p.db32 + (uint64_t)d * p.pitch + base, bytes, bar);
}

// k-NN (k = 1 is the planners' 1-NN) with the FP32 pre-filter.
// grid = (ceil(Q / (NN_THREADS*QPT)), nsplit)
template <int D, int QPT>
__global__ void __launch_bounds__(NN_THREADS) knn_rv_f32_kernel(const ScanParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NQ = NN_THREADS * QPT;
  float* tiles = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + F_STAGES * D * F_TILE);
  double* s_list = reinterpret_cast<double*>(bars + F_STAGES);
  long long* s_ubits = reinterpret_cast<long long*>(s_list + (size_t)p.k * NQ);
  uint32_t* i_list = reinterpret_cast<uint32_t*>(s_ubits + NQ);
  int* s_have = reinterpret_cast<int*>(i_list + (size_t)p.k * NQ);
  uint32_t* s_cand = reinterpret_cast<uint32_t*>(s_have + NQ);  // [CAND_CAP][NQ]

  const int tid = threadIdx.x;
  const int k = (int)p.k;
  const float inf32 = __int_as_float(0x7F800000);
  float qf[QPT][D];
  float t32[QPT];
  uint32_t qi[QPT], excl[QPT];
  int ncand[QPT];
#pragma unroll
  for (int a = 0; a < QPT; ++a) {
    qi[a] = blockIdx.x * NQ + a * NN_THREADS + tid;
    const bool ok = qi[a] < p.Q;
    const uint32_t qq = ok ? qi[a] : 0u;
#pragma unroll
    for (int d = 0; d < D; ++d)
      qf[a][d] = __double2float_rn(p.q.p[(uint64_t)qq * p.q.stride_i + (uint64_t)d * p.q.stride_d]);
    excl[a] = OXG_INVALID_INDEX;
    if (p.exclude_mode == EXCL_ARRAY) excl[a] = p.exclude[qq];
    if (p.exclude_mode == EXCL_SELF) excl[a] = p.self_base + qq;
    t32[a] = ok ? inf32 : -1.0f;
    s_ubits[a * NN_THREADS + tid] = ok ? 0x7FF0000000000000ll : -1ll;
    s_have[a * NN_THREADS + tid] = 0;
    qi[a] = qq;
    ncand[a] = 0;
  }

  const uint32_t n0 = blockIdx.y * p.split_len;
  const uint32_t n1 = min(p.n_db, n0 + p.split_len);
  const uint32_t ntiles = (n1 > n0) ? (n1 - n0 + F_TILE - 1) / F_TILE : 0;

  if (tid == 0) {
    for (int s = 0; s < F_STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (uint32_t t = 0; t < (uint32_t)(F_STAGES - 1) && t < ntiles; ++t) {
      uint32_t base = n0 + t * F_TILE;
      issue_tile32<D>(p, tiles + (t % F_STAGES) * D * F_TILE, &bars[t % F_STAGES], base,
                      min((uint32_t)F_TILE, n1 - base));
    }
  }

  for (uint32_t t = 0; t < ntiles; ++t) {
    if (tid == 0) {
      uint32_t tn = t + F_STAGES - 1;
      if (tn < ntiles) {
        uint32_t base = n0 + tn * F_TILE;
        issue_tile32<D>(p, tiles + (tn % F_STAGES) * D * F_TILE, &bars[tn % F_STAGES], base,
                        min((uint32_t)F_TILE, n1 - base));
      }
    }
    mbar_wait(&bars[t % F_STAGES], (t / F_STAGES) & 1u);
    const float* sp = tiles + (t % F_STAGES) * D * F_TILE;
    const uint32_t base = n0 + t * F_TILE;
    const int cnt = (int)min((uint32_t)F_TILE, n1 - base);
#pragma unroll 1
    for (int j = 0; j < cnt; j += 4) {
      float4 pv[D];
#pragma unroll
      for (int d = 0; d < D; ++d) pv[d] = *reinterpret_cast<const float4*>(sp + d * F_TILE + j);
      float2 acc[QPT][2];
      bool any = false;
#pragma unroll
      for (int a = 0; a < QPT; ++a) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float2 d01 = sub2(qf[a][d], make_float2(pv[d].x, pv[d].y));
          const float2 d23 = sub2(qf[a][d], make_float2(pv[d].z, pv[d].w));
          if (d == 0) {
            acc[a][0] = __fmul2_rn(d01, d01);
            acc[a][1] = __fmul2_rn(d23, d23);
          } else {
            acc[a][0] = __ffma2_rn(d01, d01, acc[a][0]);
            acc[a][1] = __ffma2_rn(d23, d23, acc[a][1]);
          }
        }
        any = any || !(acc[a][0].x > t32[a]) || !(acc[a][0].y > t32[a]) ||
              !(acc[a][1].x > t32[a]) || !(acc[a][1].y > t32[a]);
      }
      if (any) {
#pragma unroll
        for (int a = 0; a < QPT; ++a) {
          const float av[4] = {acc[a][0].x, acc[a][0].y, acc[a][1].x, acc[a][1].y};
          const int li = a * NN_THREADS + tid;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (av[e] > t32[a]) continue;
            if (ncand[a] == CAND_CAP) {  // buffer full: drain this query now
              t32[a] = knn_drain_f32<D>(p, qi[a], n1, excl[a], s_cand + li, ncand[a], NQ, NQ,
                                        s_list + li, i_list + li, &s_have[li], &s_ubits[li],
                                        FilterThr{t32[a], 0.f}).diff;
              ncand[a] = 0;
              if (av[e] > t32[a]) continue;
            }
            s_cand[li + ncand[a] * NQ] = base + j + e;
            ++ncand[a];
          }
        }
      }
    }
    // tile boundary: if any lane of the warp holds deferred candidates, the warp drains together
    {
      bool pend = false;
#pragma unroll
      for (int a = 0; a < QPT; ++a) pend = pend || ncand[a] > 0;
      if (__any_sync(0xFFFFFFFFu, pend)) {
#pragma unroll
        for (int a = 0; a < QPT; ++a) {
          const int li = a * NN_THREADS + tid;
          t32[a] = knn_drain_f32<D>(p, qi[a], n1, excl[a], s_cand + li, ncand[a], NQ, NQ, s_list + li,
                                    i_list + li, &s_have[li], &s_ubits[li],
                                    FilterThr{t32[a], 0.f}).diff;
          ncand[a] = 0;
        }
      }
    }
    __syncthreads();  // everyone is done with this stage before it is refilled
  }

#pragma unroll
  for (int a = 0; a < QPT; ++a) {
    const uint32_t q_out = blockIdx.x * NQ + a * NN_THREADS + tid;
    if (q_out >= p.Q) continue;
    const int li = a * NN_THREADS + tid;
    const int have = s_have[li];
    uint64_t o = ((uint64_t)blockIdx.y * p.Q + q_out) * (uint64_t)k;
    const double* sl = s_list + li;
    const uint32_t* il = i_list + li;
    for (int e = 0; e < k; ++e) {
      bool f = e < have;
      p.out_idx[o + e] = f ? il[e * NQ] : OXG_INVALID_INDEX;
      p.out_dist[o + e] = f ? sl[e * NQ] : __longlong_as_double(0x7FF0000000000000ll);
    }
  }
}

// Issues the D+1 bulk copies (coordinates + squared norms) of one f32 tile.
template <int D>
__device__ __forceinline__ void issue_tile32n(const ScanParams& p, float* stage, uint64_t* bar,
                                              uint32_t base, uint32_t cnt) {
  uint32_t cnt_pad = (cnt + 3u) & ~3u;
  uint32_t bytes = cnt_pad * 4u;
  mbar_arrive_expect_tx(bar, bytes * (D + 1));
#pragma unroll
  for (int d = 0; d <= D; ++d)
    bulk_copy_g2s(stage + d * F_TILE, p.db32 + (uint64_t)d * p.pitch + base, bytes, bar);
}

// min of three with NaN propagation (FMNMX3.NAN): a NaN evaluation must reach the exact path.
__device__ __forceinline__ float min3_nan(float a, float b, float c) {
  float r;
  asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__device__ __forceinline__ float2 fma2_bcast(float g, float2 pv, float2 acc) {
  return __ffma2_rn(make_float2(g, g), pv, acc);
}

// k-NN with the dot-form pre-filter and per-warp fallback to the difference form.
// grid = (ceil(Q / (NN_THREADS*QPT)), nsplit)
// One tile of the filtered k-NN scan: 8 stored states per step, DOT selects the dot form
// (columns 0..D-1 = coordinates, column D = squared norms) or the difference form.  Returns
// the number of candidates this thread deferred.
template <int D, int QPT, bool DOT>
__device__ __forceinline__ int knn_scan_tile(const ScanParams& p, uint32_t sp_addr, int cnt,
                                             uint32_t base, uint32_t n1, const float (&qv)[QPT][D],
                                             FilterThr (&thr)[QPT], const uint32_t (&qi)[QPT],
                                             const uint32_t (&excl)[QPT], int (&ncand)[QPT],
                                             uint32_t* s_cand, double* const (&g_list)[QPT],
                                             uint32_t* const (&g_ilist)[QPT], int* s_have,
                                             long long* s_ubits, int tid) {
  constexpr int NQ = NN_THREADS * QPT;
  int passed = 0;
#pragma unroll 1
  for (int j = 0; j < cnt; j += 8) {
    // A step can defer at most 8 candidates per query: make room BEFORE the step (one rarely taken
    // branch with the only call sites of the exact path), so the step's body is pure arithmetic
    // and stores and ptxas does not have to keep the call's operands live across it.
    {
      bool full = false;
#pragma unroll
      for (int a = 0; a < QPT; ++a) full = full || ncand[a] > CAND_CAP - 8;
      if (full) {
#pragma unroll
        for (int a = 0; a < QPT; ++a) {
          const int li = a * NN_THREADS + tid;
          thr[a] = knn_drain_f32<D>(p, qi[a], n1, excl[a], s_cand + li, ncand[a], NQ, 1, g_list[a],
                                    g_ilist[a], &s_have[li], &s_ubits[li], thr[a]);
          ncand[a] = 0;
        }
      }
    }
    const uint32_t row = sp_addr + (uint32_t)j * 4u;
    float2 acc[QPT][4];
    if (DOT) {
      const float4 pn0 = lds128(row + D * F_TILE * 4);
      const float4 pn1 = lds128(row + D * F_TILE * 4 + 16);
#pragma unroll
      for (int a = 0; a < QPT; ++a) {
        acc[a][0] = make_float2(pn0.x, pn0.y);
        acc[a][1] = make_float2(pn0.z, pn0.w);
        acc[a][2] = make_float2(pn1.x, pn1.y);
        acc[a][3] = make_float2(pn1.z, pn1.w);
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float4 pv0 = lds128(row + d * F_TILE * 4);
        const float4 pv1 = lds128(row + d * F_TILE * 4 + 16);
#pragma unroll
        for (int a = 0; a < QPT; ++a) {
          acc[a][0] = fma2_bcast(qv[a][d], make_float2(pv0.x, pv0.y), acc[a][0]);
          acc[a][1] = fma2_bcast(qv[a][d], make_float2(pv0.z, pv0.w), acc[a][1]);
          acc[a][2] = fma2_bcast(qv[a][d], make_float2(pv1.x, pv1.y), acc[a][2]);
          acc[a][3] = fma2_bcast(qv[a][d], make_float2(pv1.z, pv1.w), acc[a][3]);
        }
      }
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float4 pv0 = lds128(row + d * F_TILE * 4);
        const float4 pv1 = lds128(row + d * F_TILE * 4 + 16);
#pragma unroll
        for (int a = 0; a < QPT; ++a) {
          const float2 df[4] = {sub2(qv[a][d], make_float2(pv0.x, pv0.y)),
                                sub2(qv[a][d], make_float2(pv0.z, pv0.w)),
                                sub2(qv[a][d], make_float2(pv1.x, pv1.y)),
                                sub2(qv[a][d], make_float2(pv1.z, pv1.w))};
#pragma unroll
          for (int g = 0; g < 4; ++g)
            acc[a][g] = (d == 0) ? __fmul2_rn(df[g], df[g]) : __ffma2_rn(df[g], df[g], acc[a][g]);
        }
      }
    }
    bool any = false;
#pragma unroll
    for (int a = 0; a < QPT; ++a) {
      const float tcur = DOT ? thr[a].dot : thr[a].diff;
      // smallest of the 8 evaluations (FMNMX3 with NaN propagation), one compare per query
      const float m0 = min3_nan(acc[a][0].x, acc[a][0].y, acc[a][1].x);
      const float m1 = min3_nan(acc[a][1].y, acc[a][2].x, acc[a][2].y);
      const float m2 = min3_nan(acc[a][3].x, acc[a][3].y, m0);
      any = any || !(min3_nan(m1, m2, m2) > tcur);
    }
    if (any) {
#pragma unroll
      for (int a = 0; a < QPT; ++a) {
        const float av[8] = {acc[a][0].x, acc[a][0].y, acc[a][1].x, acc[a][1].y,
                             acc[a][2].x, acc[a][2].y, acc[a][3].x, acc[a][3].y};
        const int li = a * NN_THREADS + tid;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (av[e] > (DOT ? thr[a].dot : thr[a].diff)) continue;
          s_cand[li + ncand[a] * NQ] = base + j + e;  // room for 8 was made at the top of the step
          ++ncand[a];
          ++passed;
        }
      }
    }
  }
  return passed;
}

template <int D, int QPT>
__global__ void __launch_bounds__(NN_THREADS, DOT_MIN_CTAS) knn_rv_dot_kernel(const ScanParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NQ = NN_THREADS * QPT;
  constexpr int COLS = D + 1;
  float* tiles = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + DOT_STAGES * COLS * F_TILE);  // full[S]
  uint64_t* empty = bars + DOT_STAGES;  // empty[S]: one arrival per warp
  // The sorted (dist, idx) lists live in the output buffers themselves (global memory, one
  // contiguous run per query, touched only by the rare exact path): shared memory is left to
  // the tiles and the candidate buffers, which is what bounds the CTAs per SM.
  long long* s_ubits = reinterpret_cast<long long*>(bars + 2 * DOT_STAGES);
  int* s_have = reinterpret_cast<int*>(s_ubits + NQ);
  uint32_t* s_cand = reinterpret_cast<uint32_t*>(s_have + NQ);  // [CAND_CAP][NQ]
  double* g_list[QPT];
  uint32_t* g_ilist[QPT];

  const int tid = threadIdx.x;
  const int k = (int)p.k;
  const float inf32 = __int_as_float(0x7F800000);
  float qv[QPT][D];  // dot mode: -2 fl32(q_d); difference mode: fl32(q_d)
  FilterThr thr[QPT];
  uint32_t qi[QPT], excl[QPT];
  int ncand[QPT];
  bool dot_mode = true;  // warp-uniform
#pragma unroll
  for (int a = 0; a < QPT; ++a) {
    qi[a] = blockIdx.x * NQ + a * NN_THREADS + tid;
    const bool ok = qi[a] < p.Q;
    const uint32_t qq = ok ? qi[a] : 0u;
#pragma unroll
    for (int d = 0; d < D; ++d)
      qv[a][d] = -2.0f * __double2float_rn(p.q.p[(uint64_t)qq * p.q.stride_i + (uint64_t)d * p.q.stride_d]);
    excl[a] = OXG_INVALID_INDEX;
    if (p.exclude_mode == EXCL_ARRAY) excl[a] = p.exclude[qq];
    if (p.exclude_mode == EXCL_SELF) excl[a] = p.self_base + qq;
    thr[a].diff = ok ? inf32 : -1.0f;
    thr[a].dot = ok ? inf32 : -inf32;
    s_ubits[a * NN_THREADS + tid] = ok ? 0x7FF0000000000000ll : -1ll;
    s_have[a * NN_THREADS + tid] = 0;
    {
      const uint64_t o = ((uint64_t)blockIdx.y * p.Q + qq) * (uint64_t)p.k;
      g_list[a] = p.out_dist + o;
      g_ilist[a] = p.out_idx + o;
    }
    qi[a] = qq;
    ncand[a] = 0;
  }

  const uint32_t n0 = blockIdx.y * p.split_len;
  const uint32_t n1 = min(p.n_db, n0 + p.split_len);
  const uint32_t ntiles = (n1 > n0) ? (n1 - n0 + F_TILE - 1) / F_TILE : 0;

  // Producer / consumer pipeline without a CTA-wide barrier per tile: full[s] is completed by
  // the bulk copies of a tile, empty[s] by one arrival per warp once the warp has finished the
  // tile; the producer (thread 0) refills a stage only when all warps have left it.  Warps can
  // drift up to DOT_STAGES-1 tiles apart, which absorbs the stalls of the exact path.
  if (tid == 0) {
    for (int s = 0; s < DOT_STAGES; ++s) {
      mbar_init(&bars[s], 1);
      mbar_init(&empty[s], NN_THREADS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (uint32_t t = 0; t < (uint32_t)(DOT_STAGES - 1) && t < ntiles; ++t) {
      uint32_t base = n0 + t * F_TILE;
      issue_tile32n<D>(p, tiles + (t % DOT_STAGES) * COLS * F_TILE, &bars[t % DOT_STAGES], base,
                       min((uint32_t)F_TILE, n1 - base));
    }
  }

  for (uint32_t t = 0; t < ntiles; ++t) {
    if (tid == 0) {
      uint32_t tn = t + DOT_STAGES - 1;
      if (tn < ntiles) {
        const uint32_t st = tn % DOT_STAGES;
        if (tn >= (uint32_t)DOT_STAGES)  // the stage's previous tile (tn - S) has been left by all warps
          mbar_wait(&empty[st], ((tn / DOT_STAGES) - 1u) & 1u);
        uint32_t base = n0 + tn * F_TILE;
        issue_tile32n<D>(p, tiles + st * COLS * F_TILE, &bars[st], base,
                         min((uint32_t)F_TILE, n1 - base));
      }
    }
    __syncwarp();
    mbar_wait(&bars[t % DOT_STAGES], (t / DOT_STAGES) & 1u);
    const float* sp = tiles + (t % DOT_STAGES) * COLS * F_TILE;
    const uint32_t base = n0 + t * F_TILE;
    const int cnt = (int)min((uint32_t)F_TILE, n1 - base);
    int passed = 0;  // candidates this thread deferred in this tile
    const uint32_t sp_addr = smem_u32(sp);
    if (dot_mode)
      passed = knn_scan_tile<D, QPT, true>(p, sp_addr, cnt, base, n1, qv, thr, qi, excl, ncand, s_cand,
                                           g_list, g_ilist, s_have, s_ubits, tid);
    else
      passed = knn_scan_tile<D, QPT, false>(p, sp_addr, cnt, base, n1, qv, thr, qi, excl, ncand, s_cand,
                                            g_list, g_ilist, s_have, s_ubits, tid);
    // Tile boundary.  The warp drains together (lanes stay converged) when some lane's buffer
    // is getting full, every DRAIN_EVERY-th tile, and at the end: thresholds may lag a few
    // tiles behind, which only lets a few more pairs through the filter.
    {
      bool pend = false, urgent = false;
#pragma unroll
      for (int a = 0; a < QPT; ++a) {
        pend = pend || ncand[a] > 0;
        urgent = urgent || ncand[a] > CAND_CAP / 2;
      }
      const bool due = (t % DRAIN_EVERY) == DRAIN_EVERY - 1 || t + 1 == ntiles || t < 8;
      if (__any_sync(0xFFFFFFFFu, urgent) || (due && __any_sync(0xFFFFFFFFu, pend))) {
#pragma unroll
        for (int a = 0; a < QPT; ++a) {
          const int li = a * NN_THREADS + tid;
          thr[a] = knn_drain_f32<D>(p, qi[a], n1, excl[a], s_cand + li, ncand[a], NQ, 1, g_list[a],
                                    g_ilist[a], &s_have[li], &s_ubits[li], thr[a]);
          ncand[a] = 0;
        }
      }
      // Adaptive fallback: once the lists have settled (a few tiles in), a warp that still
      // defers more than ~2% of its pairs has a dot-form bound too loose for its data.
      if (dot_mode && t >= 4) {
        const int warp_passed = __reduce_add_sync(0xFFFFFFFFu, passed);
        if (warp_passed > (32 * QPT * F_TILE) / 48) {
          dot_mode = false;
#pragma unroll
          for (int a = 0; a < QPT; ++a)
#pragma unroll
            for (int d = 0; d < D; ++d) qv[a][d] = -0.5f * qv[a][d];  // exact: back to fl32(q_d)
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[t % DOT_STAGES]);  // this warp has left the stage
  }

  // pad the unused tail of each list (fewer than k candidates in the store)
#pragma unroll
  for (int a = 0; a < QPT; ++a) {
    const uint32_t q_out = blockIdx.x * NQ + a * NN_THREADS + tid;
    if (q_out >= p.Q) continue;
    const int have = s_have[a * NN_THREADS + tid];
    for (int e = have; e < k; ++e) {
      g_ilist[a][e] = OXG_INVALID_INDEX;
      g_list[a][e] = __longlong_as_double(0x7FF0000000000000ll);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Few-query 1-NN scan (the planners' per-iteration query through the C ABI, Q <= 8):
// state-per-thread.  Every thread streams stored states straight from the f64 SoA columns
// (coalesced, each byte of the store read exactly once: 8*D*N algorithmic bytes per batch) and
// evaluates the QN queries held in registers.  This is the HBM-bound regime of SURVEY §8(d);
// when the store fits the 126 MB L2 it is L2-bound.  Exactness: a state is admitted by
// d2 <= U(best) (bit compare) and then compared as (sqrt(d2), index), exactly like the batched
// kernel; per-CTA winners are merged by knn_merge_kernel.
// ---------------------------------------------------------------------------------------
constexpr int S1_THREADS = 256;
constexpr int RS_MAXQ = 8;  // most queries a few-query (state-per-thread) kernel takes

template <int D, int QN>
__global__ void __launch_bounds__(S1_THREADS) nn1_stream_kernel(const ScanParams p) {
  __shared__ double s_d[QN][S1_THREADS / 32];
  __shared__ uint32_t s_i[QN][S1_THREADS / 32];
  double q[QN][D];
  double best[QN];
  uint32_t bi[QN], excl1[QN];
  long long ub[QN];
#pragma unroll
  for (int a = 0; a < QN; ++a) {
    const uint32_t qq = (uint32_t)a < p.Q ? a : 0u;
    excl1[a] = OXG_INVALID_INDEX;
    if (p.exclude_mode == EXCL_ARRAY) excl1[a] = p.exclude[qq];
    if (p.exclude_mode == EXCL_SELF) excl1[a] = p.self_base + qq;
#pragma unroll
    for (int d = 0; d < D; ++d) q[a][d] = p.q.p[(uint64_t)qq * p.q.stride_i + (uint64_t)d * p.q.stride_d];
    best[a] = __longlong_as_double(0x7FF0000000000000ll);
    bi[a] = OXG_INVALID_INDEX;
    ub[a] = 0x7FF0000000000000ll;
  }
  const uint32_t stride = gridDim.x * S1_THREADS;
#pragma unroll 2
  for (uint32_t i = blockIdx.x * S1_THREADS + threadIdx.x; i < p.n_db; i += stride) {
    double pt[D];
#pragma unroll
    for (int d = 0; d < D; ++d) pt[d] = __ldcs(p.db + (uint64_t)d * p.pitch + i);  // streamed once
#pragma unroll
    for (int a = 0; a < QN; ++a) {
      // distance(stored, query): (stored - query)^2, the 1-NN call sites' argument order
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double df = __dsub_rn(pt[d], q[a][d]);
        const double m = __dmul_rn(df, df);
        s = (d == 0) ? m : __dadd_rn(s, m);
      }
      if (__double_as_longlong(s) <= ub[a] && s == s && i != excl1[a]) {
        const double r = sqrt(s);
        if (r < best[a] || (r == best[a] && i < bi[a])) {
          best[a] = r;
          bi[a] = i;
          ub[a] = __double_as_longlong(sqrt_le_upper_dev(r));
        }
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int a = 0; a < QN; ++a) {
    double bd = best[a];
    uint32_t bx = bi[a];
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_down_sync(0xFFFFFFFFu, bd, o);
      const uint32_t oi = __shfl_down_sync(0xFFFFFFFFu, bx, o);
      if (od < bd || (od == bd && oi < bx)) {
        bd = od;
        bx = oi;
      }
    }
    if (lane == 0) {
      s_d[a][warp] = bd;
      s_i[a][warp] = bx;
    }
  }
  __syncthreads();
  if (threadIdx.x < QN && threadIdx.x < p.Q) {
    const int a = threadIdx.x;
    double bd = s_d[a][0];
    uint32_t bx = s_i[a][0];
    for (int wv = 1; wv < S1_THREADS / 32; ++wv)
      if (s_d[a][wv] < bd || (s_d[a][wv] == bd && s_i[a][wv] < bx)) {
        bd = s_d[a][wv];
        bx = s_i[a][wv];
      }
    // partial list layout of knn_merge_kernel with k = 1: [slice][Q]
    p.out_idx[(uint64_t)blockIdx.x * p.Q + a] = bx;
    p.out_dist[(uint64_t)blockIdx.x * p.Q + a] = bd;
  }
}

// ---------------------------------------------------------------------------------------
// Few-query k-NN (Q <= 8, 2 <= k <= 16): state-per-thread streaming like nn1_stream_kernel.
// One pass over the CTA's slice per query (queries are few; the slice is re-read from L2 / HBM).
// Every thread keeps a sorted k-list in shared memory (stride = threads); the CTA then extracts
// its k best by k rounds of a block-wide (dist, idx) arg-min over the threads' list heads and
// writes them as one partial list per (CTA, query); knn_select_kernel merges the partial lists
// the same way.  Exactness as everywhere: admission by d2 <= U(kth), order by (sqrt(d2), idx).
// ---------------------------------------------------------------------------------------
struct ArgMin {
  double d;
  uint32_t i;
  uint32_t who;
};
__device__ __forceinline__ ArgMin argmin_better(ArgMin a, ArgMin b) {
  return (b.d < a.d || (b.d == a.d && b.i < a.i)) ? b : a;
}
// Block-wide arg-min of (d, i); result valid in every thread.  s_red: 3 x 32 words of scratch.
__device__ __forceinline__ ArgMin block_argmin(ArgMin v, double* s_d, uint32_t* s_i, uint32_t* s_w) {
  for (int o = 16; o > 0; o >>= 1) {
    ArgMin u;
    u.d = __shfl_xor_sync(0xFFFFFFFFu, v.d, o);
    u.i = __shfl_xor_sync(0xFFFFFFFFu, v.i, o);
    u.who = __shfl_xor_sync(0xFFFFFFFFu, v.who, o);
    v = argmin_better(v, u);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();  // scratch free
  if (lane == 0) {
    s_d[warp] = v.d;
    s_i[warp] = v.i;
    s_w[warp] = v.who;
  }
  __syncthreads();
  ArgMin r;
  r.d = lane < nw ? s_d[lane] : __longlong_as_double(0x7FF0000000000000ll);
  r.i = lane < nw ? s_i[lane] : OXG_INVALID_INDEX;
  r.who = lane < nw ? s_w[lane] : 0xFFFFFFFFu;
  for (int o = 16; o > 0; o >>= 1) {
    ArgMin u;
    u.d = __shfl_xor_sync(0xFFFFFFFFu, r.d, o);
    u.i = __shfl_xor_sync(0xFFFFFFFFu, r.i, o);
    u.who = __shfl_xor_sync(0xFFFFFFFFu, r.who, o);
    r = argmin_better(r, u);
  }
  return r;
}

template <int D>
__global__ void __launch_bounds__(S1_THREADS) knn_stream_kernel(const ScanParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int k = (int)p.k;
  double* s_list = reinterpret_cast<double*>(smem_raw);                 // [k][S1_THREADS]
  uint32_t* i_list = reinterpret_cast<uint32_t*>(s_list + (size_t)k * S1_THREADS);
  __shared__ double s_q[RS_MAXQ][D];
  __shared__ double s_rd[32];
  __shared__ uint32_t s_ri[32], s_rw[32];
  if (threadIdx.x < p.Q * D) {
    const uint32_t a = threadIdx.x / D, d = threadIdx.x % D;
    s_q[a][d] = p.q.p[(uint64_t)a * p.q.stride_i + (uint64_t)d * p.q.stride_d];
  }
  __syncthreads();
  const uint32_t n0 = blockIdx.x * p.split_len;
  const uint32_t n1 = min(p.n_db, n0 + p.split_len);
  double* sl = s_list + threadIdx.x;
  uint32_t* il = i_list + threadIdx.x;
  for (uint32_t a = 0; a < p.Q; ++a) {
    uint32_t excl = OXG_INVALID_INDEX;
    if (p.exclude_mode == EXCL_ARRAY) excl = p.exclude[a];
    if (p.exclude_mode == EXCL_SELF) excl = p.self_base + a;
    int have = 0;
    long long ub = 0x7FF0000000000000ll;
    for (uint32_t i = n0 + threadIdx.x; i < n1; i += S1_THREADS) {
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double df = __dsub_rn(s_q[a][d], __ldcs(p.db + (uint64_t)d * p.pitch + i));  // (query, stored)
        const double m = __dmul_rn(df, df);
        s = (d == 0) ? m : __dadd_rn(s, m);
      }
      if (__double_as_longlong(s) <= ub && s == s && i != excl) {
        if (list_insert(sl, il, S1_THREADS, k, have, sqrt(s), i) && have == k)
          ub = __double_as_longlong(sqrt_le_upper_dev(sl[(k - 1) * S1_THREADS]));
      }
    }
    // the CTA's k best: k rounds of arg-min over the list heads
    int head = 0;
    const uint64_t o = ((uint64_t)blockIdx.x * p.Q + a) * (uint64_t)k;
    for (int e = 0; e < k; ++e) {
      ArgMin v;
      v.d = head < have ? sl[head * S1_THREADS] : __longlong_as_double(0x7FF0000000000000ll);
      v.i = head < have ? il[head * S1_THREADS] : OXG_INVALID_INDEX;
      v.who = threadIdx.x;
      const ArgMin w = block_argmin(v, s_rd, s_ri, s_rw);
      if (w.i != OXG_INVALID_INDEX && w.who == threadIdx.x) ++head;
      if (threadIdx.x == 0) {
        p.out_idx[o + e] = w.i;
        p.out_dist[o + e] = w.i == OXG_INVALID_INDEX ? __longlong_as_double(0x7FF0000000000000ll) : w.d;
      }
    }
    __syncthreads();
  }
}

// Merges nparts sorted partial k-lists per query ([part][Q][k]) by the same k rounds of
// block-wide arg-min over list heads.  One CTA per query.
__global__ void __launch_bounds__(1024) knn_select_kernel(const uint32_t* __restrict__ part_idx,
                                                          const double* __restrict__ part_dist,
                                                          uint32_t Q, uint32_t k, uint32_t nparts,
                                                          uint32_t* __restrict__ out_idx,
                                                          double* __restrict__ out_dist) {
  __shared__ double s_rd[32];
  __shared__ uint32_t s_ri[32], s_rw[32];
  const uint32_t qi = blockIdx.x;
  // thread t owns partial list t (the launcher keeps nparts <= blockDim)
  const uint32_t t = threadIdx.x;
  uint32_t part = t, head = 0;
  auto cur = [&](double* d, uint32_t* i) {
    if (part < nparts && head < k) {
      const uint64_t o = ((uint64_t)part * Q + qi) * k + head;
      *i = part_idx[o];
      *d = part_dist[o];
    } else {
      *i = OXG_INVALID_INDEX;
      *d = __longlong_as_double(0x7FF0000000000000ll);
    }
  };
  for (uint32_t e = 0; e < k; ++e) {
    ArgMin v;
    cur(&v.d, &v.i);
    if (v.i == OXG_INVALID_INDEX) v.d = __longlong_as_double(0x7FF0000000000000ll);
    v.who = t;
    const ArgMin w = block_argmin(v, s_rd, s_ri, s_rw);
    if (w.i != OXG_INVALID_INDEX && w.who == t) ++head;
    if (t == 0) {
      out_idx[(uint64_t)qi * k + e] = w.i;
      out_dist[(uint64_t)qi * k + e] =
          w.i == OXG_INVALID_INDEX ? __longlong_as_double(0x7FF0000000000000ll) : w.d;
    }
  }
}

// The reference's 1-NN loop starts from min = distance(tree[0], q) and replaces it only when
// d < min (rrt.rs:187-196).  If that first distance is NaN nothing ever replaces it and the
// answer is (0, NaN), whatever the other states are; the scan kernels treat NaN as "never a
// winner" for every index, so this fix-up restores the reference's answer for index 0.
__global__ void nn1_first_nan_kernel(StateView db, StateView q, uint64_t Q, uint32_t n_real,
                                     uint32_t n_so2, double w_real, double w_so2,
                                     uint32_t* __restrict__ idx, double* __restrict__ dist) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q) return;
  double a[kMaxDim], b[kMaxDim];
  const uint32_t dim = n_real + n_so2;
  for (uint32_t d = 0; d < dim; ++d) {
    a[d] = db.p[d * db.stride_d];
    b[d] = q.p[i * q.stride_i + d * q.stride_d];
  }
  const double d0 = space_distance_dev(n_real, n_so2, w_real, w_so2, a, b);
  if (d0 != d0) {
    idx[i] = 0;
    dist[i] = d0;
  }
}

// Final reduction of the stream kernel's per-CTA winners: one CTA per query, (dist, idx) order.
__global__ void __launch_bounds__(256) nn1_final_kernel(const uint32_t* __restrict__ part_idx,
                                                        const double* __restrict__ part_dist,
                                                        uint32_t Q, uint32_t nparts,
                                                        uint32_t* __restrict__ out_idx,
                                                        double* __restrict__ out_dist,
                                                        StateView db, StateView q, uint32_t dim,
                                                        int nan_first) {
  __shared__ double s_d[8];
  __shared__ uint32_t s_i[8];
  const uint32_t qi = blockIdx.x;
  double bd = __longlong_as_double(0x7FF0000000000000ll);
  uint32_t bx = OXG_INVALID_INDEX;
  for (uint32_t c = threadIdx.x; c < nparts; c += blockDim.x) {
    const double d = part_dist[(uint64_t)c * Q + qi];
    const uint32_t i = part_idx[(uint64_t)c * Q + qi];
    if (d < bd || (d == bd && i < bx)) {
      bd = d;
      bx = i;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_down_sync(0xFFFFFFFFu, bd, o);
    const uint32_t oi = __shfl_down_sync(0xFFFFFFFFu, bx, o);
    if (od < bd || (od == bd && oi < bx)) {
      bd = od;
      bx = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    s_d[threadIdx.x >> 5] = bd;
    s_i[threadIdx.x >> 5] = bx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int wv = 1; wv < (int)(blockDim.x >> 5); ++wv)
      if (s_d[wv] < bd || (s_d[wv] == bd && s_i[wv] < bx)) {
        bd = s_d[wv];
        bx = s_i[wv];
      }
    if (nan_first) {
      // rrt.rs:187-196 starts from min = distance(tree[0], q): a NaN there is never replaced
      // (same rule as nn1_first_nan_kernel, fused here for the few-query path; RealVector only)
      double acc = 0.0;
      for (uint32_t d = 0; d < dim; ++d) {
        const double df = __dsub_rn(db.p[d * db.stride_d], q.p[qi * q.stride_i + d * q.stride_d]);
        const double m = __dmul_rn(df, df);
        acc = d == 0 ? m : __dadd_rn(acc, m);
      }
      if (acc != acc) {
        bx = 0;
        bd = acc;
      }
    }
    out_idx[qi] = bx;
    out_dist[qi] = bd;
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

// Radius scan with the FP32 pre-filter; same two-pass scheme as radius_rv_kernel.  The radius
// threshold is fixed, so T32 is computed once per query.  QPT queries per thread.
template <int D, bool FILL, int QPT>
__global__ void __launch_bounds__(NN_THREADS) radius_rv_f32_kernel(const ScanParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NQ = NN_THREADS * QPT;
  float* tiles = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + F_STAGES * D * F_TILE);
  __shared__ uint32_t s_maxlim;

  const int tid = threadIdx.x;
  float qf[QPT][D];
  float t32[QPT];
  uint32_t qi[QPT], lim[QPT], count[QPT];
  uint64_t cursor[QPT];
  uint32_t maxlim = 0;
  const double ms = __longlong_as_double((long long)p.maxabs[0]);
#pragma unroll
  for (int a = 0; a < QPT; ++a) {
    qi[a] = blockIdx.x * NQ + a * NN_THREADS + tid;
    const bool ok = qi[a] < p.Q;
    const uint32_t qq = ok ? qi[a] : 0u;
    double mq = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double v = p.q.p[(uint64_t)qq * p.q.stride_i + (uint64_t)d * p.q.stride_d];
      mq = fmax(mq, fabs(v));
      qf[a][d] = __double2float_rn(v);
    }
    lim[a] = p.n_db;
    if (p.idx_limit) lim[a] = min(lim[a], p.idx_limit[qq]);
    if (p.limit_self) lim[a] = min(lim[a], p.self_base + qq);
    if (!ok) lim[a] = 0;
    maxlim = max(maxlim, lim[a]);
    const double M = (ms > mq || !(ms == ms)) ? ms : mq;
    t32[a] = ok ? filter_threshold_dev(p.r_bits, M, D) : -1.0f;
    cursor[a] = 0;
    if (FILL && ok) cursor[a] = p.counts[(uint64_t)qi[a] * p.nsplit + blockIdx.y];
    count[a] = 0;
    qi[a] = qq;
  }

  if (tid == 0) s_maxlim = 0;
  __syncthreads();
  atomicMax(&s_maxlim, maxlim);
  if (tid == 0) {
    for (int s = 0; s < F_STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const uint32_t n0 = blockIdx.y * p.split_len;
  const uint32_t n1 = min(min(p.n_db, s_maxlim), n0 + p.split_len);
  const uint32_t ntiles = (n1 > n0) ? (n1 - n0 + F_TILE - 1) / F_TILE : 0;

  if (tid == 0) {
    for (uint32_t t = 0; t < (uint32_t)(F_STAGES - 1) && t < ntiles; ++t) {
      uint32_t base = n0 + t * F_TILE;
      issue_tile32<D>(p, tiles + (t % F_STAGES) * D * F_TILE, &bars[t % F_STAGES], base,
                      min((uint32_t)F_TILE, n1 - base));
    }
  }
  for (uint32_t t = 0; t < ntiles; ++t) {
    if (tid == 0) {
      uint32_t tn = t + F_STAGES - 1;
      if (tn < ntiles) {
        uint32_t base = n0 + tn * F_TILE;
        issue_tile32<D>(p, tiles + (tn % F_STAGES) * D * F_TILE, &bars[tn % F_STAGES], base,
                        min((uint32_t)F_TILE, n1 - base));
      }
    }
    mbar_wait(&bars[t % F_STAGES], (t / F_STAGES) & 1u);
    const float* sp = tiles + (t % F_STAGES) * D * F_TILE;
    const uint32_t base = n0 + t * F_TILE;
    const int cnt = (int)min((uint32_t)F_TILE, n1 - base);
#pragma unroll 1
    for (int j = 0; j < cnt; j += 4) {
      float4 pv[D];
#pragma unroll
      for (int d = 0; d < D; ++d) pv[d] = *reinterpret_cast<const float4*>(sp + d * F_TILE + j);
      float2 acc[QPT][2];
      bool any = false;
#pragma unroll
      for (int a = 0; a < QPT; ++a) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float2 d01 = sub2(qf[a][d], make_float2(pv[d].x, pv[d].y));
          const float2 d23 = sub2(qf[a][d], make_float2(pv[d].z, pv[d].w));
          if (d == 0) {
            acc[a][0] = __fmul2_rn(d01, d01);
            acc[a][1] = __fmul2_rn(d23, d23);
          } else {
            acc[a][0] = __ffma2_rn(d01, d01, acc[a][0]);
            acc[a][1] = __ffma2_rn(d23, d23, acc[a][1]);
          }
        }
        any = any || !(acc[a][0].x > t32[a]) || !(acc[a][0].y > t32[a]) ||
              !(acc[a][1].x > t32[a]) || !(acc[a][1].y > t32[a]);
      }
      if (any) {
#pragma unroll
        for (int a = 0; a < QPT; ++a) {
          const float av[4] = {acc[a][0].x, acc[a][0].y, acc[a][1].x, acc[a][1].y};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (av[e] > t32[a]) continue;
            const uint32_t jj = base + j + e;
            if (jj >= n1 || jj >= lim[a]) continue;
            double mq;
            const double s = exact_sumsq_global<D>(p, qi[a], jj, &mq);
            if (__double_as_longlong(s) <= p.r_bits && s == s) {
              if (FILL) {
                p.hit_idx[cursor[a] + count[a]] = jj;
                if (p.hit_dist) p.hit_dist[cursor[a] + count[a]] = sqrt(s);
              }
              ++count[a];
            }
          }
        }
      }
    }
    __syncthreads();
  }
  if (!FILL) {
#pragma unroll
    for (int a = 0; a < QPT; ++a) {
      const uint32_t q_out = blockIdx.x * NQ + a * NN_THREADS + tid;
      if (q_out < p.Q) p.counts[(uint64_t)q_out * p.nsplit + blockIdx.y] = count[a];
    }
  }
}

// Radius scan with the dot-form pre-filter (per-warp fallback to the difference form), 8 stored
// states per step; same two-pass scheme and output contract as radius_rv_f32_kernel.
template <int D, bool FILL, int QPT>
__global__ void __launch_bounds__(NN_THREADS) radius_rv_dot_kernel(const ScanParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NQ = NN_THREADS * QPT;
  constexpr int COLS = D + 1;
  float* tiles = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + F_STAGES * COLS * F_TILE);
  __shared__ uint32_t s_maxlim;

  const int tid = threadIdx.x;
  float qv[QPT][D];  // dot mode: -2 fl32(q_d); difference mode: fl32(q_d)
  FilterThr thr[QPT];
  uint32_t qi[QPT], lim[QPT], count[QPT];
  uint64_t cursor[QPT];
  uint32_t maxlim = 0;
  bool dot_mode = true;  // warp-uniform
  const double ms = __longlong_as_double((long long)p.maxabs[0]);
  const double pmax = __longlong_as_double((long long)p.maxabs[1]);
  const float inf32 = __int_as_float(0x7F800000);
#pragma unroll
  for (int a = 0; a < QPT; ++a) {
    qi[a] = blockIdx.x * NQ + a * NN_THREADS + tid;
    const bool ok = qi[a] < p.Q;
    const uint32_t qq = ok ? qi[a] : 0u;
    double mq = 0.0, qn = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double v = p.q.p[(uint64_t)qq * p.q.stride_i + (uint64_t)d * p.q.stride_d];
      mq = fmax(mq, fabs(v));
      qn += v * v;
      qv[a][d] = -2.0f * __double2float_rn(v);
    }
    lim[a] = p.n_db;
    if (p.idx_limit) lim[a] = min(lim[a], p.idx_limit[qq]);
    if (p.limit_self) lim[a] = min(lim[a], p.self_base + qq);
    if (!ok) lim[a] = 0;
    maxlim = max(maxlim, lim[a]);
    const double M = (ms > mq || !(ms == ms)) ? ms : mq;
    thr[a].diff = ok ? filter_threshold_dev(p.r_bits, M, D) : -1.0f;
    thr[a].dot = ok ? filter_threshold_dot_dev(p.r_bits, qn, pmax, D) : -inf32;
    cursor[a] = 0;
    if (FILL && ok) cursor[a] = p.counts[(uint64_t)qi[a] * p.nsplit + blockIdx.y];
    count[a] = 0;
    qi[a] = qq;
  }

  if (tid == 0) s_maxlim = 0;
  __syncthreads();
  atomicMax(&s_maxlim, maxlim);
  if (tid == 0) {
    for (int s = 0; s < F_STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const uint32_t n0 = blockIdx.y * p.split_len;
  const uint32_t n1 = min(min(p.n_db, s_maxlim), n0 + p.split_len);
  const uint32_t ntiles = (n1 > n0) ? (n1 - n0 + F_TILE - 1) / F_TILE : 0;

  if (tid == 0) {
    for (uint32_t t = 0; t < (uint32_t)(F_STAGES - 1) && t < ntiles; ++t) {
      uint32_t base = n0 + t * F_TILE;
      issue_tile32n<D>(p, tiles + (t % F_STAGES) * COLS * F_TILE, &bars[t % F_STAGES], base,
                       min((uint32_t)F_TILE, n1 - base));
    }
  }
  for (uint32_t t = 0; t < ntiles; ++t) {
    if (tid == 0) {
      uint32_t tn = t + F_STAGES - 1;
      if (tn < ntiles) {
        uint32_t base = n0 + tn * F_TILE;
        issue_tile32n<D>(p, tiles + (tn % F_STAGES) * COLS * F_TILE, &bars[tn % F_STAGES], base,
                         min((uint32_t)F_TILE, n1 - base));
      }
    }
    mbar_wait(&bars[t % F_STAGES], (t / F_STAGES) & 1u);
    const float* sp = tiles + (t % F_STAGES) * COLS * F_TILE;
    const uint32_t base = n0 + t * F_TILE;
    const int cnt = (int)min((uint32_t)F_TILE, n1 - base);
    int passed = 0;
#pragma unroll 1
    for (int j = 0; j < cnt; j += 8) {
      float2 acc[QPT][4];
      bool any = false;
      if (dot_mode) {
        const float4 pn0 = *reinterpret_cast<const float4*>(sp + D * F_TILE + j);
        const float4 pn1 = *reinterpret_cast<const float4*>(sp + D * F_TILE + j + 4);
#pragma unroll
        for (int a = 0; a < QPT; ++a) {
          acc[a][0] = make_float2(pn0.x, pn0.y);
          acc[a][1] = make_float2(pn0.z, pn0.w);
          acc[a][2] = make_float2(pn1.x, pn1.y);
          acc[a][3] = make_float2(pn1.z, pn1.w);
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float4 pv0 = *reinterpret_cast<const float4*>(sp + d * F_TILE + j);
          const float4 pv1 = *reinterpret_cast<const float4*>(sp + d * F_TILE + j + 4);
#pragma unroll
          for (int a = 0; a < QPT; ++a) {
            acc[a][0] = fma2_bcast(qv[a][d], make_float2(pv0.x, pv0.y), acc[a][0]);
            acc[a][1] = fma2_bcast(qv[a][d], make_float2(pv0.z, pv0.w), acc[a][1]);
            acc[a][2] = fma2_bcast(qv[a][d], make_float2(pv1.x, pv1.y), acc[a][2]);
            acc[a][3] = fma2_bcast(qv[a][d], make_float2(pv1.z, pv1.w), acc[a][3]);
          }
        }
      } else {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float4 pv0 = *reinterpret_cast<const float4*>(sp + d * F_TILE + j);
          const float4 pv1 = *reinterpret_cast<const float4*>(sp + d * F_TILE + j + 4);
#pragma unroll
          for (int a = 0; a < QPT; ++a) {
            const float2 df[4] = {sub2(qv[a][d], make_float2(pv0.x, pv0.y)),
                                  sub2(qv[a][d], make_float2(pv0.z, pv0.w)),
                                  sub2(qv[a][d], make_float2(pv1.x, pv1.y)),
                                  sub2(qv[a][d], make_float2(pv1.z, pv1.w))};
#pragma unroll
            for (int g = 0; g < 4; ++g)
              acc[a][g] = (d == 0) ? __fmul2_rn(df[g], df[g]) : __ffma2_rn(df[g], df[g], acc[a][g]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < QPT; ++a) {
        const float tcur = dot_mode ? thr[a].dot : thr[a].diff;
        const float m0 = min3_nan(acc[a][0].x, acc[a][0].y, acc[a][1].x);
        const float m1 = min3_nan(acc[a][1].y, acc[a][2].x, acc[a][2].y);
        const float m2 = min3_nan(acc[a][3].x, acc[a][3].y, m0);
        any = any || !(min3_nan(m1, m2, m2) > tcur);
      }
      if (any) {
#pragma unroll
        for (int a = 0; a < QPT; ++a) {
          const float tcur = dot_mode ? thr[a].dot : thr[a].diff;
          const float av[8] = {acc[a][0].x, acc[a][0].y, acc[a][1].x, acc[a][1].y,
                               acc[a][2].x, acc[a][2].y, acc[a][3].x, acc[a][3].y};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (av[e] > tcur) continue;
            ++passed;
            const uint32_t jj = base + j + e;
            if (jj >= n1 || jj >= lim[a]) continue;
            double mq;
            const double s = exact_sumsq_global<D>(p, qi[a], jj, &mq);
            if (__double_as_longlong(s) <= p.r_bits && s == s) {
              if (FILL) {
                p.hit_idx[cursor[a] + count[a]] = jj;
                if (p.hit_dist) p.hit_dist[cursor[a] + count[a]] = sqrt(s);
              }
              ++count[a];
            }
          }
        }
      }
    }
    if (dot_mode && t >= 1) {
      const int warp_passed = __reduce_add_sync(0xFFFFFFFFu, passed);
      if (warp_passed > (32 * QPT * F_TILE) / 48) {
        dot_mode = false;
#pragma unroll
        for (int a = 0; a < QPT; ++a)
#pragma unroll
          for (int d = 0; d < D; ++d) qv[a][d] = -0.5f * qv[a][d];  // exact: back to fl32(q_d)
      }
    }
    __syncthreads();
  }
  if (!FILL) {
#pragma unroll
    for (int a = 0; a < QPT; ++a) {
      const uint32_t q_out = blockIdx.x * NQ + a * NN_THREADS + tid;
      if (q_out < p.Q) p.counts[(uint64_t)q_out * p.nsplit + blockIdx.y] = count[a];
    }
  }
}

// Few-query radius scan (Q <= 8: RRT*'s find_neighbours and PRM's start connection through the
// C ABI): state-per-thread over contiguous per-CTA slices of the f64 store, queries broadcast
// from shared memory.  Same two-pass contract as the batched kernels: the count pass writes
// counts[q * nsplit + slice], the fill pass writes the hits of each slice in ascending index
// order (ballot + block prefix per round of S1_THREADS states).  The store is streamed once per
// pass for all queries of the batch: HBM / L2 bound.

template <int D, bool FILL>
__global__ void __launch_bounds__(S1_THREADS) radius_stream_kernel(const ScanParams p) {
  __shared__ double s_q[RS_MAXQ][D];
  __shared__ uint32_t s_lim[RS_MAXQ];
  __shared__ uint32_t s_wcnt[RS_MAXQ][S1_THREADS / 32];
  const uint32_t Q = p.Q;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < Q * D) {
    const uint32_t a = threadIdx.x / D, d = threadIdx.x % D;
    s_q[a][d] = p.q.p[(uint64_t)a * p.q.stride_i + (uint64_t)d * p.q.stride_d];
  }
  if (threadIdx.x < Q) {
    uint32_t lim = p.n_db;
    if (p.idx_limit) lim = min(lim, p.idx_limit[threadIdx.x]);
    if (p.limit_self) lim = min(lim, p.self_base + threadIdx.x);
    s_lim[threadIdx.x] = lim;
  }
  __syncthreads();
  const uint32_t n0 = blockIdx.x * p.split_len;
  const uint32_t n1 = min(p.n_db, n0 + p.split_len);
  uint32_t run[RS_MAXQ];
  uint64_t cursor[RS_MAXQ];
#pragma unroll
  for (int a = 0; a < RS_MAXQ; ++a) {
    run[a] = 0;
    cursor[a] = (FILL && (uint32_t)a < Q) ? p.counts[(uint64_t)a * p.nsplit + blockIdx.x] : 0;
  }
  for (uint32_t base = n0; base < n1; base += S1_THREADS) {
    const uint32_t i = base + threadIdx.x;
    double pt[D];
    if (i < n1) {
#pragma unroll
      for (int d = 0; d < D; ++d) pt[d] = __ldcs(p.db + (uint64_t)d * p.pitch + i);
    }
    uint32_t hits = 0;  // bit a: this state is inside query a's ball
    double dist[RS_MAXQ];
#pragma unroll
    for (int a = 0; a < RS_MAXQ; ++a) {
      if ((uint32_t)a >= Q) break;
      bool h = false;
      if (i < n1 && i < s_lim[a]) {
        // distance(query, stored): radius call sites' argument order (rrt_star.rs:125, prm.rs:133)
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const double df = __dsub_rn(s_q[a][d], pt[d]);
          const double m = __dmul_rn(df, df);
          s = (d == 0) ? m : __dadd_rn(s, m);
        }
        h = __double_as_longlong(s) <= p.r_bits && s == s;  // exactly sqrt(s) < r
        dist[a] = s;
      }
      const uint32_t b = __ballot_sync(0xFFFFFFFFu, h);
      if (FILL) {
        if (lane == 0) s_wcnt[a][warp] = __popc(b);
        if (h) hits |= 1u << a;
      } else {
        run[a] += (lane == 0) ? __popc(b) : 0u;
      }
    }
    if (FILL) {
      __syncthreads();
#pragma unroll
      for (int a = 0; a < RS_MAXQ; ++a) {
        if ((uint32_t)a >= Q) break;
        const bool h = (hits >> a) & 1u;
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, h);
        uint32_t off = run[a], tot = 0;
#pragma unroll
        for (int wv = 0; wv < S1_THREADS / 32; ++wv) {
          const uint32_t c = s_wcnt[a][wv];
          if (wv < warp) off += c;
          tot += c;
        }
        if (h) {
          const uint64_t pos = cursor[a] + off + __popc(b & ((1u << lane) - 1u));
          p.hit_idx[pos] = i;
          if (p.hit_dist) p.hit_dist[pos] = sqrt(dist[a]);
        }
        run[a] += tot;
      }
      __syncthreads();
    }
  }
  if (!FILL) {
    // block total per query (lane 0 of each warp holds its warp's count)
#pragma unroll
    for (int a = 0; a < RS_MAXQ; ++a) {
      if ((uint32_t)a >= Q) break;
      if (lane == 0) s_wcnt[a][warp] = run[a];
    }
    __syncthreads();
    if (threadIdx.x < Q) {
      uint32_t tot = 0;
      for (int wv = 0; wv < S1_THREADS / 32; ++wv) tot += s_wcnt[threadIdx.x][wv];
      p.counts[(uint64_t)threadIdx.x * p.nsplit + blockIdx.x] = tot;
    }
  }
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

template <int D, int QPT>
static size_t knn_f32_smem_bytes(uint32_t k) {
  return sizeof(float) * F_STAGES * D * F_TILE + sizeof(uint64_t) * F_STAGES +
         (size_t)NN_THREADS * QPT * ((size_t)k * (sizeof(double) + sizeof(uint32_t)) +
                                     sizeof(long long) + sizeof(int) + CAND_CAP * sizeof(uint32_t));
}

template <int D, int QPT>
static int launch_knn_rv_f32(const ScanParams& p, dim3 grid, cudaStream_t st) {
  size_t smem = knn_f32_smem_bytes<D, QPT>(p.k);
  OXG_CUDA_TRY(cudaFuncSetAttribute(knn_rv_f32_kernel<D, QPT>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  knn_rv_f32_kernel<D, QPT><<<grid, NN_THREADS, smem, st>>>(p);
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

template <int D>
static int launch_knn_stream(const ScanParams& p, size_t smem, cudaStream_t st) {
  OXG_CUDA_TRY(cudaFuncSetAttribute(knn_stream_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  knn_stream_kernel<D><<<p.nsplit, S1_THREADS, smem, st>>>(p);
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

template <int D>
static int launch_nn1_stream(ScanParams p, uint32_t blocks, cudaStream_t st) {
  if (p.Q <= 1) {
    nn1_stream_kernel<D, 1><<<blocks, S1_THREADS, 0, st>>>(p);
  } else if (p.Q <= 2) {
    nn1_stream_kernel<D, 2><<<blocks, S1_THREADS, 0, st>>>(p);
  } else if (p.Q <= 4) {
    nn1_stream_kernel<D, 4><<<blocks, S1_THREADS, 0, st>>>(p);
  } else {
    nn1_stream_kernel<D, 8><<<blocks, S1_THREADS, 0, st>>>(p);
  }
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

template <int D, int QPT>
static int launch_knn_rv_dot(const ScanParams& p, dim3 grid, cudaStream_t st) {
  size_t smem = sizeof(float) * DOT_STAGES * (D + 1) * F_TILE + sizeof(uint64_t) * 2 * DOT_STAGES +
                (size_t)NN_THREADS * QPT * (sizeof(long long) + sizeof(int) + CAND_CAP * sizeof(uint32_t));
  OXG_CUDA_TRY(cudaFuncSetAttribute(knn_rv_dot_kernel<D, QPT>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  knn_rv_dot_kernel<D, QPT><<<grid, NN_THREADS, smem, st>>>(p);
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

// Tuning / A-B switches read once from the environment (measurement aids, not a fallback:
// both settings run hand-written sm_100a kernels and give bit-identical results).
//   OXG_KNN_FILTER=0  use the all-FP64 scan kernels (no FP32 pre-filter)
//   OXG_KNN_FORM=diff use the difference form of the pre-filter only (default: dot form with
//                     per-warp fallback to the difference form)
//   OXG_KNN_QPT=1|2   force the queries-per-thread of the batched k-NN kernel
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
static bool use_f32_filter() {
  static const bool on = env_int("OXG_KNN_FILTER", 1) != 0;
  return on;
}
// Smallest batch the dot-form k-NN kernel takes (below it the difference-form filter runs).  Its
// looser bound costs more exact re-tests while the lists warm up, which long slices amortise:
// measured at D = 7 it wins from 128 queries up on a 1e7-state store, from ~1000 on 1e6 states.
static uint64_t dot_min_q(uint32_t n_db) {
  static const int v = env_int("OXG_KNN_DOT_MINQ", 0);
  if (v > 0) return (uint64_t)v;
  return n_db >= 4000000u ? 65 : 1024;
}

static bool use_dot_form() {
  static const bool on = [] {
    const char* v = getenv("OXG_KNN_FORM");
    return !(v && v[0] == 'd' && v[1] == 'i');
  }();
  return on;
}
static int forced_qpt() {
  static const int q = env_int("OXG_KNN_QPT", 0);
  return q;
}

template <int D, int QPT>
static int launch_radius_rv_f32(const ScanParams& p, dim3 grid, bool fill, cudaStream_t st) {
  size_t smem = sizeof(float) * F_STAGES * D * F_TILE + sizeof(uint64_t) * F_STAGES;
  if (fill) {
    OXG_CUDA_TRY(cudaFuncSetAttribute(radius_rv_f32_kernel<D, true, QPT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    radius_rv_f32_kernel<D, true, QPT><<<grid, NN_THREADS, smem, st>>>(p);
  } else {
    OXG_CUDA_TRY(cudaFuncSetAttribute(radius_rv_f32_kernel<D, false, QPT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    radius_rv_f32_kernel<D, false, QPT><<<grid, NN_THREADS, smem, st>>>(p);
  }
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

template <int D, int QPT>
static int launch_radius_rv_dot(const ScanParams& p, dim3 grid, bool fill, cudaStream_t st) {
  size_t smem = sizeof(float) * F_STAGES * (D + 1) * F_TILE + sizeof(uint64_t) * F_STAGES;
  if (fill) {
    OXG_CUDA_TRY(cudaFuncSetAttribute(radius_rv_dot_kernel<D, true, QPT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    radius_rv_dot_kernel<D, true, QPT><<<grid, NN_THREADS, smem, st>>>(p);
  } else {
    OXG_CUDA_TRY(cudaFuncSetAttribute(radius_rv_dot_kernel<D, false, QPT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    radius_rv_dot_kernel<D, false, QPT><<<grid, NN_THREADS, smem, st>>>(p);
  }
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
static void choose_split(uint32_t n_db, uint64_t qtiles, uint32_t tile, uint32_t* nsplit,
                         uint32_t* split_len, uint32_t k = 1) {
  // six CTA waves of slices fill the machine best; with k > 1 and few query tiles every slice
  // repeats the top-k warm-up, so three waves win there (measured at N = 1e6, D = 7)
  const uint64_t target_ctas = 148ull * ((k >= 4 && qtiles <= 100) ? 3 : 6);
  uint64_t want = (target_ctas + qtiles - 1) / qtiles;
  uint64_t max_split = std::max<uint64_t>(1, n_db / (tile * 4));
  uint64_t ns = std::min<uint64_t>(std::min<uint64_t>(want, max_split), 2048);
  if (ns < 1) ns = 1;
  uint32_t len = round_up_u32((uint32_t)((n_db + ns - 1) / ns), tile);
  if (len == 0) len = tile;
  *split_len = len;
  *nsplit = std::max<uint32_t>(1, (n_db + len - 1) / len);
}

// A second stream per device for work that may overlap a store's own stream (lowest priority,
// so the store's kernels get the SMs first).
struct AuxStream {
  cudaStream_t st = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static AuxStream* aux_stream_for(int device) {
  static AuxStream table[64];
  if (device < 0 || device >= 64) return nullptr;
  AuxStream* a = &table[device];
  if (!a->st) {
    int lo = 0, hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) lo = 0;
    if (cudaStreamCreateWithPriority(&a->st, cudaStreamNonBlocking, lo) != cudaSuccess ||
        cudaEventCreateWithFlags(&a->fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&a->join, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      a->st = nullptr;
      return nullptr;
    }
  }
  return a;
}

// k-NN of Q device-resident queries against the store; results in device memory.
static int nn_knn_device_impl(oxg_store_s* s, StateView q, uint64_t Q, uint32_t k, int exclude_mode,
                              const uint32_t* d_exclude, uint64_t self_base, bool stored_first,
                              uint32_t* d_idx, double* d_dist, int force_qpt, uint32_t force_nsplit);

int nn_knn_device(oxg_store_s* s, StateView q, uint64_t Q, uint32_t k, int exclude_mode,
                  const uint32_t* d_exclude, uint64_t self_base, bool stored_first, uint32_t* d_idx,
                  double* d_dist) {
  return nn_knn_device_impl(s, q, Q, k, exclude_mode, d_exclude, self_base, stored_first, d_idx,
                            d_dist, 0, 0);
}

static int nn_knn_device_impl(oxg_store_s* s, StateView q, uint64_t Q, uint32_t k, int exclude_mode,
                              const uint32_t* d_exclude, uint64_t self_base, bool stored_first,
                              uint32_t* d_idx, double* d_dist, int force_qpt, uint32_t force_nsplit) {
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

  p.db32 = s->d_data32;
  p.maxabs = s->d_maxabs;
  const bool fast = is_pure_rv(s->space) && s->dim <= (uint32_t)kMaxFastDim;
  // Exact spatial pruning (oxg_tree.cu): low-dimensional RealVector stores with enough pairs to
  // pay for the hierarchy (or a hierarchy that is still current).  OXG_KNN_TREE=0/1 and
  // oxg_tuning_set("knn_tree", 0/1) force the choice for A/B runs; results are bit-identical.
  if (fast && force_qpt == 0 && force_nsplit == 0 && tree_supported(s, k)) {
    int mode = g_tune_knn_tree.load();
    if (mode < 0) mode = env_int("OXG_KNN_TREE", -1);
    bool use = mode > 0;
    if (mode < 0)
      // a current hierarchy answers any batch the few-query streaming kernels do not take; a fresh
      // one (0.3 ms + 0.2 ns per state in key order) pays from ~2e8 pairs, where the all-pairs
      // scan of a medium batch takes about as long (profiles/r02_medium_batches.json)
      use = p.n_db >= 16384 && Q >= 64 && (tree_is_current(s) || (double)Q * (double)p.n_db >= 2.0e8);
    if (use) {
      OXG_TRY(tree_knn_device(s, q, Q, k, exclude_mode, d_exclude, self_base, d_idx, d_dist, nullptr));
      if (k == 1 && stored_first && exclude_mode == EXCL_NONE) {
        nn1_first_nan_kernel<<<(unsigned)((Q + 127) / 128), 128, 0, s->stream>>>(
            store_view(s), q, Q, s->space.n_real, s->space.n_so2, s->space.w_real, s->space.w_so2, d_idx,
            d_dist);
        OXG_LAUNCH_CHECK();
      }
      return OXG_OK;
    }
  }
  // Few to moderately many queries (the planners' per-iteration calls, small sweeps): the
  // state-per-thread streaming kernels, RS_MAXQ queries per launch.  Beyond a few dozen queries
  // the query-per-thread kernels below win.
  // (measured crossovers at N = 1e6, D = 7: ~40 queries for k = 1, ~100 for k = 10)
  const bool stream1 = fast && k == 1 && Q <= 32;
  const bool streamk = fast && k >= 2 && k <= 16 && Q <= 64 && p.n_db >= 4096;
  if (stream1 || streamk) {
    uint32_t blocks = 0;
    size_t smem = 0;
    if (stream1) {
      blocks = (uint32_t)std::min<uint64_t>(((uint64_t)p.n_db + S1_THREADS - 1) / S1_THREADS, 148ull * 4);
    } else {
      uint32_t ns = std::min<uint32_t>(std::max<uint32_t>(1, (p.n_db + 2047) / 2048), 1024);
      p.split_len = round_up_u32((p.n_db + ns - 1) / ns, S1_THREADS);
      p.nsplit = (p.n_db + p.split_len - 1) / p.split_len;
      blocks = p.nsplit;
      smem = (size_t)k * S1_THREADS * (sizeof(double) + sizeof(uint32_t));
    }
    const size_t npart = (size_t)blocks * RS_MAXQ * k;
    OXG_TRY(s->buf[B_PIDX].ensure(npart * sizeof(uint32_t)));
    OXG_TRY(s->buf[B_PDIST].ensure(npart * sizeof(double)));
    p.out_idx = s->buf[B_PIDX].as<uint32_t>();
    p.out_dist = s->buf[B_PDIST].as<double>();
    for (uint64_t c0 = 0; c0 < Q; c0 += RS_MAXQ) {
      const uint32_t qn = (uint32_t)std::min<uint64_t>(RS_MAXQ, Q - c0);
      ScanParams pc = p;
      pc.q.p = q.p + c0 * q.stride_i;
      pc.Q = qn;
      pc.exclude = d_exclude ? d_exclude + c0 : nullptr;
      pc.self_base = (uint32_t)(self_base + c0);
      int r = OXG_E_INTERNAL;
      if (stream1) {
        // the 1-NN kernel keeps its queries in registers and takes one excluded index per query
        // through the exclusion array or the self rule
        OXG_DISPATCH_D(s->dim, r = (launch_nn1_stream<DD>(pc, blocks, s->stream)));
        OXG_TRY(r);
        nn1_final_kernel<<<qn, 256, 0, s->stream>>>(pc.out_idx, pc.out_dist, qn, blocks,
                                                   d_idx + c0, d_dist + c0, store_view(s), pc.q, s->dim,
                                                   stored_first && exclude_mode == EXCL_NONE);
      } else {
        OXG_DISPATCH_D(s->dim, r = (launch_knn_stream<DD>(pc, smem, s->stream)));
        OXG_TRY(r);
        knn_select_kernel<<<qn, (unsigned)round_up_u32(blocks, 32), 0, s->stream>>>(
            pc.out_idx, pc.out_dist, qn, k, blocks, d_idx + c0 * k, d_dist + c0 * k);
      }
      OXG_LAUNCH_CHECK();
    }
    return OXG_OK;
  }
  const bool f32 = fast && use_f32_filter();
  // several queries per thread once there is enough work to fill the machine that way
  // (two queries per thread halve the shared-memory reads per pair but also the number of
  // query tiles, so the store is cut into more slices: it pays once every slice is long —
  // measured at D = 7: from ~1e10 pairs for k = 1, ~1e11 pairs for k > 1, see tools/medq_probe.py)
  int qpt = 1;
  if (fast && Q >= 2ull * 148 * NN_THREADS * 2 && k <= 32) qpt = 2;
  if (fast && use_f32_filter() && use_dot_form() && Q >= dot_min_q(p.n_db) && k <= 32 &&
      (double)Q * (double)p.n_db >= (k == 1 ? 1e10 : 1e11))
    qpt = 2;
  if (fast && forced_qpt() > 0) {
    qpt = forced_qpt() >= 2 ? 2 : 1;
    if (qpt == 2 && k > 32) qpt = 1;
  }
  if (force_qpt > 0) qpt = force_qpt;
  const uint32_t nq_per_cta = NN_THREADS * qpt;
  const uint64_t qtiles = (Q + nq_per_cta - 1) / nq_per_cta;
  // Wave quantisation.  The batched kernels keep 3 CTAs per SM resident; a batch of W.f waves
  // with a small fraction f pays a whole extra wave for a few CTAs (the 8-GPU PRM shard is 1.1
  // waves).  Run the whole waves as they are and give the tail queries the entire machine by
  // splitting the store into slices for them only.
  if (f32 && qpt == 2 && force_nsplit == 0 && force_qpt == 0 && forced_qpt() == 0) {
    const uint64_t slots = 148ull * 3;
    const uint64_t whole = qtiles / slots, tail = qtiles % slots;
    // The tail runs on a second, lower-priority stream next to the whole waves: the block
    // scheduler drains the main kernel's CTAs first and fills the slots its last wave frees with
    // the tail's short CTAs, so neither part ends on an idle machine (main writes the outputs
    // directly, the tail goes through the partial-list scratch: no shared buffer).
    if (whole >= 1 && tail > 0 && tail * 2 <= slots && p.n_db >= 65536) {
      const uint64_t q_main = whole * slots * nq_per_cta;
      uint32_t ns = (uint32_t)std::min<uint64_t>(slots / tail, 16);
      AuxStream* ax = env_int("OXG_KNN_TAIL_CONCURRENT", 1) != 0 ? aux_stream_for(s->device) : nullptr;
      if (ax) {
        OXG_CUDA_TRY(cudaEventRecord(ax->fork, s->stream));
        OXG_CUDA_TRY(cudaStreamWaitEvent(ax->st, ax->fork, 0));
      }
      OXG_TRY(nn_knn_device_impl(s, q, q_main, k, exclude_mode, d_exclude, self_base, stored_first,
                                 d_idx, d_dist, 2, 1));
      StateView qt = q;
      qt.p = q.p + q_main * q.stride_i;
      cudaStream_t own = s->stream;
      if (ax) s->stream = ax->st;
      const int rc = nn_knn_device_impl(s, qt, Q - q_main, k, exclude_mode,
                                        d_exclude ? d_exclude + q_main : nullptr, self_base + q_main,
                                        stored_first, d_idx + q_main * k, d_dist + q_main * k, 2, ns);
      s->stream = own;
      if (ax) {
        OXG_CUDA_TRY(cudaEventRecord(ax->join, ax->st));
        OXG_CUDA_TRY(cudaStreamWaitEvent(s->stream, ax->join, 0));
      }
      return rc;
    }
  }
  if (force_nsplit > 0) {
    p.nsplit = force_nsplit;
    const uint32_t tile = f32 ? F_TILE : NN_TILE;
    p.split_len = round_up_u32((p.n_db + p.nsplit - 1) / p.nsplit, tile);
    p.nsplit = std::max<uint32_t>(1, (p.n_db + p.split_len - 1) / p.split_len);
  } else {
    choose_split(p.n_db, qtiles, f32 ? F_TILE : NN_TILE, &p.nsplit, &p.split_len, k);
  }

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

  if (f32 && use_dot_form() && Q >= dot_min_q(p.n_db)) {
    int r = OXG_E_INTERNAL;
    if (qpt == 2) {
      OXG_DISPATCH_D(s->dim, r = (launch_knn_rv_dot<DD, 2>(p, grid, s->stream)));
    } else {
      OXG_DISPATCH_D(s->dim, r = (launch_knn_rv_dot<DD, 1>(p, grid, s->stream)));
    }
    OXG_TRY(r);
  } else if (f32) {
    int r = OXG_E_INTERNAL;
    if (qpt == 2) {
      OXG_DISPATCH_D(s->dim, r = (launch_knn_rv_f32<DD, 2>(p, grid, s->stream)));
    } else {
      OXG_DISPATCH_D(s->dim, r = (launch_knn_rv_f32<DD, 1>(p, grid, s->stream)));
    }
    OXG_TRY(r);
  } else if (fast) {
    int r = OXG_E_INTERNAL;
    if (qpt == 2) {
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
  if (p.nsplit >= 16 && p.nsplit <= 1024) {
    // many slices: merge each query's partial lists in parallel (k rounds of block arg-min)
    const unsigned threads = (unsigned)round_up_u32(p.nsplit, 32);
    knn_select_kernel<<<(unsigned)Q, threads, 0, s->stream>>>(part_idx, part_dist, (uint32_t)Q, k,
                                                              p.nsplit, d_idx, d_dist);
    OXG_LAUNCH_CHECK();
  } else if (p.nsplit > 1) {
    int threads = 64;
    size_t smem = (size_t)k * threads * (sizeof(double) + sizeof(uint32_t));
    OXG_CUDA_TRY(cudaFuncSetAttribute(knn_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    knn_merge_kernel<<<(unsigned)((Q + threads - 1) / threads), threads, smem, s->stream>>>(
        part_idx, part_dist, (uint32_t)Q, k, p.nsplit, d_idx, d_dist);
    OXG_LAUNCH_CHECK();
  }
  if (k == 1 && stored_first && exclude_mode == EXCL_NONE) {
    nn1_first_nan_kernel<<<(unsigned)((Q + 127) / 128), 128, 0, s->stream>>>(
        store_view(s), q, Q, s->space.n_real, s->space.n_so2, s->space.w_real, s->space.w_so2, d_idx,
        d_dist);
    OXG_LAUNCH_CHECK();
  }
  return OXG_OK;
}

struct RadiusPlan {
  bool fast, f32, stream;
  int qpt;
  uint32_t nsplit, split_len;
  uint64_t qtiles;
};
static RadiusPlan radius_plan(const oxg_store_s* s, uint64_t Q) {
  RadiusPlan pl;
  pl.fast = is_pure_rv(s->space) && s->dim <= (uint32_t)kMaxFastDim;
  pl.f32 = pl.fast && use_f32_filter();
  pl.stream = pl.fast && Q >= 1 && Q <= (uint64_t)RS_MAXQ;
  if (pl.stream) {  // few queries: state-per-thread over contiguous per-CTA slices
    const uint32_t n = std::max<uint32_t>((uint32_t)s->size, 1);
    uint32_t ns = std::min<uint32_t>(std::max<uint32_t>(1, (n + 2047) / 2048), 148 * 4);
    pl.split_len = round_up_u32((n + ns - 1) / ns, S1_THREADS);
    pl.nsplit = (n + pl.split_len - 1) / pl.split_len;
    pl.qpt = 1;
    pl.qtiles = 1;
    return pl;
  }
  pl.qpt = (pl.f32 && Q >= 2ull * 148 * NN_THREADS * 2) ? 2 : 1;
  const uint64_t per_cta = (uint64_t)NN_THREADS * pl.qpt;
  pl.qtiles = std::max<uint64_t>((Q + per_cta - 1) / per_cta, 1);
  choose_split(std::max<uint32_t>((uint32_t)s->size, 1), pl.qtiles, pl.f32 ? F_TILE : NN_TILE,
               &pl.nsplit, &pl.split_len);
  return pl;
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
  p.db32 = s->d_data32;
  p.maxabs = s->d_maxabs;
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
  // Exact spatial pruning (oxg_tree.cu), as for k-NN: low-dimensional RealVector stores, enough
  // pairs to pay for the hierarchy or a hierarchy that is still current.
  if (tree_supported(s, 1) && Q >= 64 && p.n_db >= 16384) {
    int mode = g_tune_knn_tree.load();
    if (mode < 0) mode = env_int("OXG_KNN_TREE", -1);
    const bool use = mode > 0 || (mode < 0 && (tree_is_current(s) || (double)Q * (double)p.n_db >= 2.0e8));
    if (use) {
      s->last_radius_nsplit = 1;
      return tree_radius_device(s, q, Q, thr, d_idx_limit, limit_self, self_base, want_dist, total);
    }
  }
  const RadiusPlan pl = radius_plan(s, Q);
  p.nsplit = pl.nsplit;
  p.split_len = pl.split_len;
  s->last_radius_nsplit = pl.nsplit;

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
  dim3 grid((unsigned)pl.qtiles, p.nsplit);
  GenericParams g;
  g.n_real = s->space.n_real;
  g.n_so2 = s->space.n_so2;
  g.w_real = s->space.w_real;
  g.w_so2 = s->space.w_so2;
  g.stored_first = 0;
  g.radius = r;
  auto launch = [&](bool fill) -> int {
    if (pl.stream) {
      if (fill) {
        OXG_DISPATCH_D(s->dim, (radius_stream_kernel<DD, true><<<p.nsplit, S1_THREADS, 0, s->stream>>>(p)));
      } else {
        OXG_DISPATCH_D(s->dim, (radius_stream_kernel<DD, false><<<p.nsplit, S1_THREADS, 0, s->stream>>>(p)));
      }
      OXG_LAUNCH_CHECK();
      return OXG_OK;
    }
    if (pl.f32 && use_dot_form() && Q >= 1024) {
      int rr = OXG_E_INTERNAL;
      if (pl.qpt == 2) {
        OXG_DISPATCH_D(s->dim, rr = (launch_radius_rv_dot<DD, 2>(p, grid, fill, s->stream)));
      } else {
        OXG_DISPATCH_D(s->dim, rr = (launch_radius_rv_dot<DD, 1>(p, grid, fill, s->stream)));
      }
      return rr;
    }
    if (pl.f32) {
      int rr = OXG_E_INTERNAL;
      if (pl.qpt == 2) {
        OXG_DISPATCH_D(s->dim, rr = (launch_radius_rv_f32<DD, 2>(p, grid, fill, s->stream)));
      } else {
        OXG_DISPATCH_D(s->dim, rr = (launch_radius_rv_f32<DD, 1>(p, grid, fill, s->stream)));
      }
      return rr;
    }
    if (pl.fast) {
      int rr = OXG_E_INTERNAL;
      OXG_DISPATCH_D(s->dim, rr = (launch_radius_rv<DD>(p, grid, fill, s->stream)));
      return rr;
    }
    g.s = p;
    if (fill)
      scan_generic_kernel<2><<<grid, NN_THREADS, 0, s->stream>>>(g);
    else
      scan_generic_kernel<1><<<grid, NN_THREADS, 0, s->stream>>>(g);
    OXG_LAUNCH_CHECK();
    return OXG_OK;
  };
  OXG_TRY(launch(false));
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
  OXG_TRY(launch(true));
  return OXG_OK;
}

// offsets_out[q] = scanned[q * nsplit]  (q = 0..Q)
__global__ void pick_offsets_kernel(const uint64_t* __restrict__ scanned, uint64_t Q, uint32_t nsplit,
                                    uint64_t* __restrict__ out) {
  uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q <= Q) out[q] = scanned[q * nsplit];
}

int nn_radius_nsplit(const oxg_store_s* s, uint64_t Q) { return (int)radius_plan(s, Q).nsplit; }

// Compacts B_OFFSETS (per-slice offsets) into per-query CSR offsets in `d_out` (Q+1 entries).
int nn_radius_offsets(oxg_store_s* s, uint64_t Q, uint64_t* d_out) {
  uint32_t nsplit = s->last_radius_nsplit;
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
  OXG_CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
  OXG_TRY(nn_knn_device(s, rows_view(s->buf[B_QUERY].as<double>(), s->dim), Q, k,
                        exclude_idx ? EXCL_ARRAY : EXCL_NONE, d_excl, 0, stored_first, d_idx, d_dist));
  OXG_CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
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

int oxg_store_last_scan_ms(oxg_store s, double* out_ms) {
  OXG_API_BEGIN
  if (!s || !out_ms) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  float ms = 0.f;
  OXG_CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
  *out_ms = ms;
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
