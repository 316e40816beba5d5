// oxg_so3.cu — SO(3) scans and cone validity (include/oxg_so3.h).
//
// Reference: SO3StateSpace::distance / interpolate, oxmpl/src/base/spaces/so3_state_space.rs:99-159.
// The device works only on a = |q1 . q2| (exactly reproducible); everything that involves libm
// (acos, sin) runs on the host with the caller's libm, and every distance comparison is turned into
// an exact threshold on `a` by a monotone search over f64 bit patterns.  See the header.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <limits>
#include <vector>

#include "../../include/oxg_so3.h"
#include "oxg_internal.cuh"

using namespace oxg;

struct oxg_so3_cones_s {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint64_t n = 0;
  double* d_cones = nullptr;  // n x 5: centre (x, y, z, w), a-threshold: invalid iff a >= thr
  DevBuf states, flags;
};

namespace {

constexpr int SO3_THREADS = 256;
constexpr int SO3_WARPS = SO3_THREADS / 32;
constexpr int SO3_QC = 8;  // queries per CTA (kept in shared memory; one store pass serves all)
constexpr int MODE_BEST = 0, MODE_FIRST = 1;

struct So3Params {
  const double* col[4];  // store columns x, y, z, w
  uint32_t n;
  uint32_t split_len;  // states per split, multiple of SO3_THREADS
  uint32_t nsplit;
  const double* queries;  // device, row-major Q x 4
  uint64_t Q;
  const double* thr;  // MODE_FIRST / k-NN gather: per-query threshold on a (radius: nullable)
  double thr_all;     // radius kernels: one threshold (used when thr is NULL)
  const uint32_t* exclude;  // k-NN: per-query stored index left out (nullable)
  uint32_t k;               // k-NN: list length (<= 32)
  double* part_a;           // k-NN: [Q][nsplit][SO3_WARPS][32] descending partial lists
  long long* part_key;  // [Q][nsplit] bits of the best a (MODE_BEST)
  uint32_t* part_idx;   // [Q][nsplit]
  uint32_t* counts;         // radius: [Q][nsplit]
  const uint64_t* offsets;  // radius fill: [Q][nsplit] output position of each slice
  uint32_t* out_idx;
  double* out_a;
};

// |s . q| in the reference's order: ((sx*qx + sy*qy) + sz*qz) + sw*qw, every operation rounded.
__device__ __forceinline__ double so3_absdot(const double s[4], const double* q) {
  double d = __dmul_rn(s[0], q[0]);
  d = __dadd_rn(d, __dmul_rn(s[1], q[1]));
  d = __dadd_rn(d, __dmul_rn(s[2], q[2]));
  d = __dadd_rn(d, __dmul_rn(s[3], q[3]));
  return fabs(d);
}

__device__ __forceinline__ void so3_stage_queries(const So3Params& p, double (*s_q)[4], uint64_t q0,
                                                  int qn) {
  if (threadIdx.x < SO3_QC * 4) {
    const int a = threadIdx.x >> 2, c = threadIdx.x & 3;
    s_q[a][c] = a < qn ? p.queries[(q0 + a) * 4 + c] : 0.0;  // qn <= SO3_QC
  }
}

// Per-query threshold (p.thr, else the common p.thr_all) and excluded index (p.exclude, else none).
__device__ __forceinline__ void so3_stage_limits(const So3Params& p, double* s_thr, uint32_t* s_ex,
                                                 uint64_t q0, int qn) {
  if (threadIdx.x < SO3_QC) {
    const int a = threadIdx.x;
    s_thr[a] = a < qn && p.thr ? p.thr[q0 + a] : p.thr_all;
    s_ex[a] = a < qn && p.exclude ? p.exclude[q0 + a] : 0xFFFFFFFFu;
  }
}

// Inserts x into a warp's descending list (lane l holds the (l+1)-th largest value seen, -1 =
// empty; equal values are all kept: the list is a multiset) and returns the new k-th largest.
__device__ __forceinline__ double so3_list_insert(double& mine, double x, int lane, uint32_t k) {
  const int pos = __popc(__ballot_sync(0xFFFFFFFFu, mine >= x));  // entries >= x form a prefix
  const double up = __shfl_up_sync(0xFFFFFFFFu, mine, 1);
  if (lane == pos) mine = x;
  else if (lane > pos) mine = up;
  return __shfl_sync(0xFFFFFFFFu, mine, (int)k - 1);
}

// k-NN, pass 1: the k largest a of every (query, slice, warp), NaN and the excluded index left
// out.  Only the VALUE of the k-th largest a of a query is needed (see oxg_so3_knn): the states
// that can be among the k nearest are those with a >= the low end of its acos plateau.
template <int QN>
__global__ void __launch_bounds__(SO3_THREADS) so3_kth_kernel(const So3Params p) {
  __shared__ double s_q[SO3_QC][4];
  __shared__ double s_thr[SO3_QC];
  __shared__ uint32_t s_ex[SO3_QC];
  const uint64_t q0 = (uint64_t)blockIdx.x * QN;
  const int qn = (int)min((uint64_t)QN, p.Q - q0);
  so3_stage_queries(p, s_q, q0, qn);
  so3_stage_limits(p, s_thr, s_ex, q0, qn);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n0 = blockIdx.y * p.split_len;
  const uint32_t n1 = min(p.n, n0 + p.split_len);
  double list[QN], kth[QN];
#pragma unroll
  for (int a = 0; a < QN; ++a) list[a] = kth[a] = -1.0;
  for (uint32_t base = n0; base < n1; base += SO3_THREADS) {
    const uint32_t i = base + threadIdx.x;
    const bool live = i < n1;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    if (live) {
#pragma unroll
      for (int c = 0; c < 4; ++c) s[c] = __ldg(p.col[c] + i);
    }
#pragma unroll
    for (int a = 0; a < QN; ++a) {
      const double v = so3_absdot(s, s_q[a]);
      // strict: once the list is full an equal value cannot change the k-th largest
      uint32_t bal = __ballot_sync(0xFFFFFFFFu, live && a < qn && i != s_ex[a] && v > kth[a]);
      while (bal) {
        const int src = __ffs(bal) - 1;
        bal &= bal - 1;
        const double x = __shfl_sync(0xFFFFFFFFu, v, src);
        if (x > kth[a]) kth[a] = so3_list_insert(list[a], x, lane, p.k);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < QN; ++a) {
    if (a < qn)
      p.part_a[((((size_t)(q0 + a) * p.nsplit + blockIdx.y) * SO3_WARPS) + warp) * 32 + lane] = list[a];
  }
}

// k-NN, pass 1 merge: one warp per query folds the partial lists; out_a[q] = the k-th largest a,
// or -1 when fewer than k states qualify.
__global__ void __launch_bounds__(128) so3_kth_final_kernel(const So3Params p) {
  const uint64_t q = (uint64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (q >= p.Q) return;
  const int lane = threadIdx.x & 31;
  const double* src = p.part_a + (size_t)q * p.nsplit * SO3_WARPS * 32;
  double list = -1.0, kth = -1.0;
  for (uint32_t r = 0; r < p.nsplit * SO3_WARPS; ++r) {
    const double v = src[(size_t)r * 32 + lane];
    uint32_t bal = __ballot_sync(0xFFFFFFFFu, v > kth);
    while (bal) {
      const int from = __ffs(bal) - 1;
      bal &= bal - 1;
      const double x = __shfl_sync(0xFFFFFFFFu, v, from);
      if (x > kth) kth = so3_list_insert(list, x, lane, p.k);
    }
  }
  if (lane == 0) p.out_a[q] = kth;
}

// MODE_BEST : per (query, slice) the largest a, lowest index on ties.  The hot loop compares a as
//             doubles (a NaN never wins there); a NaN at index 0 must win like in the reference
//             loop (min starts as distance(tree[0], q) and nothing is < NaN), so the CTA that owns
//             state 0 re-tests it once and publishes the NaN's bit pattern as its partial key —
//             partial keys are compared as integers (order-preserving for a >= 0; +inf and NaN
//             sort above every finite value; "nothing found" is negative).
// MODE_FIRST: per (query, slice) the lowest index with a >= thr[query].
// QN = queries evaluated per state (1, 2, 4 or SO3_QC): a batch of fewer than SO3_QC queries runs
// the narrower instantiation so a single-query scan stays HBM-bound instead of FP64-bound.
template <int MODE, int QN>
__global__ void __launch_bounds__(SO3_THREADS) so3_scan_kernel(const So3Params p) {
  __shared__ double s_q[SO3_QC][4];
  __shared__ double s_thr[SO3_QC];
  __shared__ double s_k[SO3_QC][SO3_WARPS];
  __shared__ uint32_t s_i[SO3_QC][SO3_WARPS];
  __shared__ uint32_t s_nan0;
  const uint64_t q0 = (uint64_t)blockIdx.x * QN;
  const int qn = (int)min((uint64_t)QN, p.Q - q0);
  so3_stage_queries(p, s_q, q0, qn);
  if (MODE == MODE_FIRST && threadIdx.x < SO3_QC)
    s_thr[threadIdx.x] = (int)threadIdx.x < qn ? p.thr[q0 + threadIdx.x]
                                               : __longlong_as_double(0x7FF8000000000000ll);
  if (threadIdx.x == 0) s_nan0 = 0u;
  __syncthreads();

  double bk[QN];
  uint32_t bi[QN];
#pragma unroll
  for (int a = 0; a < QN; ++a) {
    bk[a] = -1.0;
    bi[a] = OXG_INVALID_INDEX;
  }
  const uint32_t n0 = blockIdx.y * p.split_len;
  const uint32_t n1 = min(p.n, n0 + p.split_len);
  auto eval = [&](const double* s, uint32_t i) {
#pragma unroll
    for (int a = 0; a < QN; ++a) {
      const double sv[4] = {s[0], s[1], s[2], s[3]};
      const double v = so3_absdot(sv, s_q[a]);
      if (MODE == MODE_BEST) {
        if (v > bk[a]) {  // ascending i per thread: strict > keeps the lowest index
          bk[a] = v;
          bi[a] = i;
        }
      } else {
        if (v >= s_thr[a] && bi[a] == OXG_INVALID_INDEX) bi[a] = i;
      }
    }
  };
  // U states per trip, all of their loads issued before the first use (memory-level parallelism
  // is what a single-query scan lives on)
  constexpr int U = QN <= 2 ? 4 : 2;
  uint32_t i = n0 + threadIdx.x;
  for (; i + (U - 1) * SO3_THREADS < n1; i += U * SO3_THREADS) {
    double s[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[u][c] = __ldg(p.col[c] + i + u * SO3_THREADS);
#pragma unroll
    for (int u = 0; u < U; ++u) eval(s[u], i + u * SO3_THREADS);
  }
  for (; i < n1; i += SO3_THREADS) {
    double s[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) s[c] = __ldg(p.col[c] + i);
    eval(s, i);
  }
  if (MODE == MODE_BEST && n0 == 0 && threadIdx.x == 0) {
    double s[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) s[c] = __ldg(p.col[c]);
    uint32_t m = 0;
#pragma unroll
    for (int a = 0; a < QN; ++a) {
      const double v = so3_absdot(s, s_q[a]);
      if (v != v) m |= 1u << a;
    }
    s_nan0 = m;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < QN; ++a) {
    double k = bk[a];
    uint32_t ix = bi[a];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ok = __shfl_xor_sync(0xFFFFFFFFu, k, off);
      const uint32_t oi = __shfl_xor_sync(0xFFFFFFFFu, ix, off);
      const bool take = MODE == MODE_BEST ? (ok > k || (ok == k && oi < ix)) : (oi < ix);
      if (take) {
        k = ok;
        ix = oi;
      }
    }
    if (lane == 0) {
      s_k[a][warp] = k;
      s_i[a][warp] = ix;
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < qn) {
    const int a = threadIdx.x;
    double k = s_k[a][0];
    uint32_t ix = s_i[a][0];
    for (int wv = 1; wv < SO3_WARPS; ++wv) {
      const double ok = s_k[a][wv];
      const uint32_t oi = s_i[a][wv];
      const bool take = MODE == MODE_BEST ? (ok > k || (ok == k && oi < ix)) : (oi < ix);
      if (take) {
        k = ok;
        ix = oi;
      }
    }
    const size_t o = (size_t)(q0 + a) * p.nsplit + blockIdx.y;
    if (MODE == MODE_BEST) {
      long long key = __double_as_longlong(k);  // -1.0 ("nothing") is negative as an integer
      if (s_nan0 >> a & 1u) {                   // state 0 is NaN for this query: it wins
        key = 0x7FF8000000000000ll;
        ix = 0;
      }
      p.part_key[o] = key;
    }
    p.part_idx[o] = ix;
  }
}

// Merges the per-slice partials of one query (one warp per query).
template <int MODE>
__global__ void __launch_bounds__(128) so3_final_kernel(const So3Params p) {
  const uint64_t q = (uint64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (q >= p.Q) return;
  const int lane = threadIdx.x & 31;
  long long k = -1;
  uint32_t ix = OXG_INVALID_INDEX;
  for (uint32_t s = lane; s < p.nsplit; s += 32) {
    const size_t o = (size_t)q * p.nsplit + s;
    const long long ok = MODE == MODE_BEST ? p.part_key[o] : 0;
    const uint32_t oi = p.part_idx[o];
    const bool take = MODE == MODE_BEST ? (ok > k || (ok == k && oi < ix)) : (oi < ix);
    if (take) {
      k = ok;
      ix = oi;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const long long ok = __shfl_xor_sync(0xFFFFFFFFu, k, off);
    const uint32_t oi = __shfl_xor_sync(0xFFFFFFFFu, ix, off);
    const bool take = MODE == MODE_BEST ? (ok > k || (ok == k && oi < ix)) : (oi < ix);
    if (take) {
      k = ok;
      ix = oi;
    }
  }
  if (lane == 0) {
    p.out_idx[q] = ix;
    if (MODE == MODE_BEST) p.out_a[q] = __longlong_as_double(k);
  }
}

// Radius scan on the threshold form a >= thr_all.  FILL = false counts the hits of every
// (query, slice); FILL = true writes them in ascending index order at the slice's offset
// (ballot + block prefix per round of SO3_THREADS states).
template <bool FILL, int QN>
__global__ void __launch_bounds__(SO3_THREADS) so3_radius_kernel(const So3Params p) {
  __shared__ double s_q[SO3_QC][4];
  __shared__ double s_thr[SO3_QC];
  __shared__ uint32_t s_ex[SO3_QC];
  __shared__ uint32_t s_w[SO3_QC][SO3_WARPS];
  const uint64_t q0 = (uint64_t)blockIdx.x * QN;
  const int qn = (int)min((uint64_t)QN, p.Q - q0);
  so3_stage_queries(p, s_q, q0, qn);
  so3_stage_limits(p, s_thr, s_ex, q0, qn);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n0 = blockIdx.y * p.split_len;
  const uint32_t n1 = min(p.n, n0 + p.split_len);
  uint32_t run[QN];  // hits so far (FILL: of the whole CTA; count pass: of this warp lane 0)
#pragma unroll
  for (int a = 0; a < QN; ++a) run[a] = 0;
  for (uint32_t base = n0; base < n1; base += SO3_THREADS) {
    const uint32_t i = base + threadIdx.x;
    const bool live = i < n1;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    if (live) {
#pragma unroll
      for (int c = 0; c < 4; ++c) s[c] = __ldg(p.col[c] + i);
    }
    double v[QN];
    uint32_t bal[QN];
#pragma unroll
    for (int a = 0; a < QN; ++a) {
      v[a] = so3_absdot(s, s_q[a]);
      bal[a] = __ballot_sync(0xFFFFFFFFu, live && a < qn && v[a] >= s_thr[a] && i != s_ex[a]);
    }
    if (!FILL) {
#pragma unroll
      for (int a = 0; a < QN; ++a) run[a] += __popc(bal[a]);
    } else {
      if (lane == 0) {
#pragma unroll
        for (int a = 0; a < QN; ++a) s_w[a][warp] = __popc(bal[a]);
      }
      __syncthreads();
#pragma unroll
      for (int a = 0; a < QN; ++a) {
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int wv = 0; wv < SO3_WARPS; ++wv) {
          const uint32_t c = s_w[a][wv];
          before += wv < warp ? c : 0;
          total += c;
        }
        if (bal[a] >> lane & 1u) {
          const uint64_t o = p.offsets[(size_t)(q0 + a) * p.nsplit + blockIdx.y] + run[a] + before +
                             __popc(bal[a] & ((1u << lane) - 1u));
          p.out_idx[o] = i;
          if (p.out_a) p.out_a[o] = v[a];
        }
        run[a] += total;
      }
      __syncthreads();
    }
  }
  if (!FILL) {
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < QN; ++a) s_w[a][warp] = run[a];
    }
    __syncthreads();
    if ((int)threadIdx.x < qn) {
      uint32_t tot = 0;
      for (int wv = 0; wv < SO3_WARPS; ++wv) tot += s_w[threadIdx.x][wv];
      p.counts[(size_t)(q0 + threadIdx.x) * p.nsplit + blockIdx.y] = tot;
    }
  }
}

// valid iff a(centre_c, state) < thr_c for every cone c (NaN a => invalid, as "NaN > angle" is false).
__global__ void __launch_bounds__(256) so3_cones_kernel(const double* __restrict__ cones, uint32_t n_cones,
                                                        const double* __restrict__ states, uint64_t Q,
                                                        uint8_t* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q) return;
  double s[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) s[c] = states[i * 4 + c];
  bool ok = true;
  for (uint32_t c = 0; c < n_cones && ok; ++c) {
    const double* cn = cones + (size_t)c * 5;
    // distance(centre, state): centre.x * state.x + ...  (products commute bit for bit)
    const double ctr[4] = {__ldg(cn), __ldg(cn + 1), __ldg(cn + 2), __ldg(cn + 3)};
    ok = so3_absdot(ctr, s) < __ldg(cn + 4);
  }
  out[i] = ok ? 1 : 0;
}

// ------------------------------------------------------------------ host: libm side
inline double from_bits(uint64_t b) {
  double d;
  std::memcpy(&d, &b, 8);
  return d;
}
inline uint64_t to_bits(double d) {
  uint64_t b;
  std::memcpy(&b, &d, 8);
  return b;
}
constexpr uint64_t kInfBits = 0x7FF0000000000000ull;

// distance as a function of a = |dot| (so3_state_space.rs:105-111); acos is the platform libm's
inline double so3_f(double a) { return a > 1.0 - 1e-9 ? 0.0 : std::acos(a); }

// Smallest x in [+0, +inf] (by bit pattern) with pred(so3_f(x)); pred must be monotone (false
// below, true above), which it is for "f(x) < r" and "f(x) <= r" because f is non-increasing.
// NaN if no x qualifies ("a >= NaN" is then false for every a).
template <class Pred>
double so3_threshold(Pred pred) {
  uint64_t lo = 0, hi = kInfBits;
  if (!pred(so3_f(from_bits(hi)))) return std::numeric_limits<double>::quiet_NaN();
  while (lo < hi) {
    const uint64_t mid = lo + (hi - lo) / 2;
    if (pred(so3_f(from_bits(mid))))
      hi = mid;
    else
      lo = mid + 1;
  }
  return from_bits(lo);
}

// Smallest x with f(x) == f(a) for a finite a >= 0: gallops down from a (plateaus are a few ulps
// wide except next to 0), then bisects.
double so3_plateau_low(double a) {
  const double d = so3_f(a);
  const uint64_t top = to_bits(a);
  uint64_t step = 1, hi = top;  // f(hi) == d
  uint64_t lo;
  for (;;) {
    if (hi == 0) return from_bits(0);
    lo = hi > step ? hi - step : 0;
    if (so3_f(from_bits(lo)) != d) break;  // f(lo) > d
    hi = lo;
    step *= 2;
  }
  while (hi - lo > 1) {
    const uint64_t mid = lo + (hi - lo) / 2;
    if (so3_f(from_bits(mid)) == d)
      hi = mid;
    else
      lo = mid;
  }
  return from_bits(hi);
}

// SO3StateSpace::interpolate, so3_state_space.rs:114-159 (host: sin / acos are libm's).
void so3_interpolate_host(const double* f, const double* g, double t, double* o) {
  double dot = f[0] * g[0] + f[1] * g[1] + f[2] * g[2] + f[3] * g[3];
  const double sign = dot < 0.0 ? -1.0 : 1.0;
  dot *= sign;
  if (dot > 0.9995) {
    for (int c = 0; c < 4; ++c) o[c] = f[c] + t * (g[c] * sign - f[c]);
    const double norm = std::sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
    for (int c = 0; c < 4; ++c) o[c] /= norm;
  } else {
    const double theta = std::acos(dot);
    const double sin_theta = std::sin(theta);
    const double s0 = std::sin((1.0 - t) * theta) / sin_theta;
    const double s1 = std::sin(t * theta) / sin_theta * sign;
    for (int c = 0; c < 4; ++c) o[c] = f[c] * s0 + g[c] * s1;
  }
}

inline double so3_distance_host(const double* p, const double* q) {
  const double dot = p[0] * q[0] + p[1] * q[1] + p[2] * q[2] + p[3] * q[3];
  return so3_f(std::fabs(dot));
}

int so3_check_store(oxg_store s) {
  if (!s) {
    set_error("store is NULL");
    return OXG_E_NULL;
  }
  if (s->space.n_real != 4 || s->space.n_so2 != 0) {
    set_error("SO3 scans need a store of 4 real coordinates (x, y, z, w)");
    return OXG_E_DIM;
  }
  return OXG_OK;
}

inline int so3_qn(uint64_t Q) { return Q >= 5 ? SO3_QC : Q >= 3 ? 4 : Q == 2 ? 2 : 1; }

void so3_plan(const oxg_store_s* s, uint64_t Q, So3Params& p, int qc = SO3_QC) {
  const uint64_t qchunks = (Q + qc - 1) / qc;
  const uint32_t n = (uint32_t)s->size;
  const uint32_t max_split = (n + SO3_THREADS - 1) / SO3_THREADS;
  uint64_t want = (148ull * 4 + qchunks - 1) / qchunks;
  if (want < 1) want = 1;
  if (want > max_split) want = max_split;
  if (want > 592) want = 592;
  uint32_t len = (uint32_t)((n + want - 1) / want);
  len = (len + SO3_THREADS - 1) / SO3_THREADS * SO3_THREADS;
  p.split_len = len;
  p.nsplit = (n + len - 1) / len;
  p.n = n;
  p.Q = Q;
  for (int c = 0; c < 4; ++c) p.col[c] = s->d_data + (size_t)c * s->pitch;
}

// Runs one MODE over `Q` uploaded queries (device pointer), result in d_idx / d_a (device).
template <int MODE>
int so3_scan(oxg_store_s* s, So3Params p, uint32_t* d_idx, double* d_a) {
  const size_t np = (size_t)p.Q * p.nsplit;
  OXG_TRY(s->buf[B_PIDX].ensure(np * sizeof(uint32_t)));
  OXG_TRY(s->buf[B_PDIST].ensure(np * sizeof(long long)));
  p.part_idx = s->buf[B_PIDX].as<uint32_t>();
  p.part_key = s->buf[B_PDIST].as<long long>();
  p.out_idx = d_idx;
  p.out_a = d_a;
  // the grid's x dimension carries the query chunks (up to 2^31 - 1)
  const int qn = so3_qn(p.Q);
  const uint64_t qchunks = (p.Q + qn - 1) / qn;
  dim3 grid((unsigned)qchunks, p.nsplit);
  switch (qn) {
    case 1: so3_scan_kernel<MODE, 1><<<grid, SO3_THREADS, 0, s->stream>>>(p); break;
    case 2: so3_scan_kernel<MODE, 2><<<grid, SO3_THREADS, 0, s->stream>>>(p); break;
    case 4: so3_scan_kernel<MODE, 4><<<grid, SO3_THREADS, 0, s->stream>>>(p); break;
    default: so3_scan_kernel<MODE, SO3_QC><<<grid, SO3_THREADS, 0, s->stream>>>(p); break;
  }
  OXG_LAUNCH_CHECK();
  so3_final_kernel<MODE><<<(unsigned)((p.Q + 3) / 4), 128, 0, s->stream>>>(p);
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

template <bool FILL>
void launch_so3_radius(int qn, dim3 grid, cudaStream_t st, const So3Params& p) {
  switch (qn) {
    case 1: so3_radius_kernel<FILL, 1><<<grid, SO3_THREADS, 0, st>>>(p); break;
    case 2: so3_radius_kernel<FILL, 2><<<grid, SO3_THREADS, 0, st>>>(p); break;
    case 4: so3_radius_kernel<FILL, 4><<<grid, SO3_THREADS, 0, st>>>(p); break;
    default: so3_radius_kernel<FILL, SO3_QC><<<grid, SO3_THREADS, 0, st>>>(p); break;
  }
}

// Count and fill passes of so3_radius_kernel for the thresholds / exclusions set in p.  On return
// offsets[q] .. offsets[q + 1] delimit query q's hits (ascending index) in p.out_idx = B_PIDX and,
// if want_a, their a in p.out_a = B_PDIST (both still on the device, fill enqueued on the stream).
int so3_gather(oxg_store_s* s, So3Params& p, int qn, bool want_a, uint64_t* offsets, uint64_t cap,
               uint64_t* total_out) {
  const uint64_t Q = p.Q;
  const size_t np = (size_t)Q * p.nsplit;
  OXG_TRY(s->buf[B_COUNTS].ensure(np * sizeof(uint32_t)));
  OXG_TRY(s->buf[B_OFFSETS].ensure(np * sizeof(uint64_t)));
  p.counts = s->buf[B_COUNTS].as<uint32_t>();
  const uint64_t qchunks = (Q + qn - 1) / qn;
  dim3 grid((unsigned)qchunks, p.nsplit);
  OXG_CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
  launch_so3_radius<false>(qn, grid, s->stream, p);
  OXG_LAUNCH_CHECK();
  std::vector<uint32_t> counts(np);
  OXG_CUDA_TRY(cudaMemcpyAsync(counts.data(), p.counts, np * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                               s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  std::vector<uint64_t> slice_off(np);
  uint64_t total = 0;
  offsets[0] = 0;
  for (uint64_t q = 0; q < Q; ++q) {
    for (uint32_t sl = 0; sl < p.nsplit; ++sl) {
      slice_off[q * p.nsplit + sl] = total;
      total += counts[q * p.nsplit + sl];
    }
    offsets[q + 1] = total;
  }
  *total_out = total;
  if (total > cap) {
    set_error("radius result needs %llu entries, capacity is %llu", (unsigned long long)total,
              (unsigned long long)cap);
    return OXG_E_CAPACITY;
  }
  if (total == 0) return OXG_OK;
  OXG_CUDA_TRY(cudaMemcpyAsync(s->buf[B_OFFSETS].p, slice_off.data(), np * sizeof(uint64_t),
                               cudaMemcpyHostToDevice, s->stream));
  OXG_TRY(s->buf[B_PIDX].ensure(total * sizeof(uint32_t)));
  if (want_a) OXG_TRY(s->buf[B_PDIST].ensure(total * sizeof(double)));
  p.offsets = s->buf[B_OFFSETS].as<uint64_t>();
  p.out_idx = s->buf[B_PIDX].as<uint32_t>();
  p.out_a = want_a ? s->buf[B_PDIST].as<double>() : nullptr;
  launch_so3_radius<true>(qn, grid, s->stream, p);
  OXG_LAUNCH_CHECK();
  OXG_CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
  // slice_off is pageable host memory: the copy above is staged before the call returns
  return OXG_OK;
}

int so3_upload(oxg_store_s* s, StoreBuf which, const void* src, size_t bytes) {
  OXG_TRY(s->buf[which].ensure(bytes));
  OXG_CUDA_TRY(cudaMemcpyAsync(s->buf[which].p, src, bytes, cudaMemcpyHostToDevice, s->stream));
  return OXG_OK;
}

}  // namespace

extern "C" {

int oxg_so3_nn1(oxg_store s, const double* queries, uint64_t Q, uint32_t* idx, double* dist) {
  OXG_API_BEGIN
  OXG_TRY(so3_check_store(s));
  if ((Q > 0 && !queries) || !idx || !dist) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (Q == 0) return OXG_OK;
  if (s->size == 0) {
    set_error("nearest-neighbour query against an empty store");
    return OXG_E_EMPTY;
  }
  if (Q > 0x3FFFFFFFull) {
    set_error("too many queries in one call");
    return OXG_E_ARG;
  }
  OXG_CUDA_TRY(cudaSetDevice(s->device));
  OXG_TRY(so3_upload(s, B_QUERY, queries, Q * 4 * sizeof(double)));
  OXG_TRY(s->buf[B_OIDX].ensure(Q * sizeof(uint32_t)));
  OXG_TRY(s->buf[B_ODIST].ensure(Q * sizeof(double)));
  So3Params p{};
  so3_plan(s, Q, p, so3_qn(Q));
  p.queries = s->buf[B_QUERY].as<double>();
  OXG_CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
  OXG_TRY(so3_scan<MODE_BEST>(s, p, s->buf[B_OIDX].as<uint32_t>(), s->buf[B_ODIST].as<double>()));
  OXG_CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
  std::vector<double> a(Q);
  OXG_CUDA_TRY(cudaMemcpyAsync(idx, s->buf[B_OIDX].p, Q * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                               s->stream));
  OXG_CUDA_TRY(cudaMemcpyAsync(a.data(), s->buf[B_ODIST].p, Q * sizeof(double), cudaMemcpyDeviceToHost,
                               s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));

  // Host: distances from libm, and the queries whose winner sits on an acos plateau that reaches
  // below max a — a lower index with a slightly smaller a has the same distance and wins.
  std::vector<uint64_t> redo;
  std::vector<double> redo_q, redo_thr;
  for (uint64_t q = 0; q < Q; ++q) {
    const double av = a[q];
    if (av != av) {  // NaN at index 0 (see the kernel)
      dist[q] = av;
      continue;
    }
    dist[q] = so3_f(av);
    double lo;
    if (av > 1.0 - 1e-9) {
      lo = std::nextafter(1.0 - 1e-9, 2.0);  // the whole clamp region has distance 0
    } else {
      lo = so3_plateau_low(av);
    }
    if (lo < av) {
      redo.push_back(q);
      redo_thr.push_back(lo);
      redo_q.insert(redo_q.end(), queries + q * 4, queries + q * 4 + 4);
    }
  }
  if (!redo.empty()) {
    const uint64_t R = redo.size();
    OXG_TRY(so3_upload(s, B_QUERY, redo_q.data(), R * 4 * sizeof(double)));
    OXG_TRY(so3_upload(s, B_AUX, redo_thr.data(), R * sizeof(double)));
    So3Params p2{};
    so3_plan(s, R, p2, so3_qn(R));
    p2.queries = s->buf[B_QUERY].as<double>();
    p2.thr = s->buf[B_AUX].as<double>();
    OXG_TRY(so3_scan<MODE_FIRST>(s, p2, s->buf[B_OIDX].as<uint32_t>(), nullptr));
    std::vector<uint32_t> first(R);
    OXG_CUDA_TRY(cudaMemcpyAsync(first.data(), s->buf[B_OIDX].p, R * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, s->stream));
    OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    for (uint64_t r = 0; r < R; ++r) {
      if (first[r] == OXG_INVALID_INDEX) {
        set_error("internal error: SO3 plateau pass found no state");
        return OXG_E_INTERNAL;
      }
      idx[redo[r]] = first[r];
    }
  }
  return OXG_OK;
  OXG_API_END
}

int oxg_so3_knn(oxg_store s, const double* queries, uint64_t Q, uint32_t k, const uint32_t* exclude_idx,
                uint32_t* idx, double* dist) {
  OXG_API_BEGIN
  OXG_TRY(so3_check_store(s));
  if ((Q > 0 && !queries) || (Q > 0 && k > 0 && (!idx || !dist))) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (k > 32) {
    set_error("SO3 k-NN supports k <= 32 (got %u)", k);
    return OXG_E_ARG;
  }
  if (Q > 0x3FFFFFFFull) {
    set_error("too many queries in one call");
    return OXG_E_ARG;
  }
  if (Q == 0 || k == 0) return OXG_OK;
  for (uint64_t i = 0; i < Q * k; ++i) {
    idx[i] = OXG_INVALID_INDEX;
    dist[i] = std::numeric_limits<double>::infinity();
  }
  if (s->size == 0) return OXG_OK;
  OXG_CUDA_TRY(cudaSetDevice(s->device));
  OXG_TRY(so3_upload(s, B_QUERY, queries, Q * 4 * sizeof(double)));
  if (exclude_idx) OXG_TRY(so3_upload(s, B_AUX, exclude_idx, Q * sizeof(uint32_t)));
  So3Params p{};
  const int qn = so3_qn(Q);
  so3_plan(s, Q, p, qn);
  p.queries = s->buf[B_QUERY].as<double>();
  p.exclude = exclude_idx ? s->buf[B_AUX].as<uint32_t>() : nullptr;
  p.k = k;
  // pass 1: A_k = the k-th largest a of every query (-1: fewer than k candidates)
  OXG_TRY(s->buf[B_PDIST].ensure((size_t)Q * p.nsplit * SO3_WARPS * 32 * sizeof(double)));
  OXG_TRY(s->buf[B_ODIST].ensure(Q * sizeof(double)));
  p.part_a = s->buf[B_PDIST].as<double>();
  p.out_a = s->buf[B_ODIST].as<double>();
  const uint64_t qchunks = (Q + qn - 1) / qn;
  dim3 grid((unsigned)qchunks, p.nsplit);
  switch (qn) {
    case 1: so3_kth_kernel<1><<<grid, SO3_THREADS, 0, s->stream>>>(p); break;
    case 2: so3_kth_kernel<2><<<grid, SO3_THREADS, 0, s->stream>>>(p); break;
    case 4: so3_kth_kernel<4><<<grid, SO3_THREADS, 0, s->stream>>>(p); break;
    default: so3_kth_kernel<SO3_QC><<<grid, SO3_THREADS, 0, s->stream>>>(p); break;
  }
  OXG_LAUNCH_CHECK();
  so3_kth_final_kernel<<<(unsigned)((Q + 3) / 4), 128, 0, s->stream>>>(p);
  OXG_LAUNCH_CHECK();
  std::vector<double> thr(Q);
  OXG_CUDA_TRY(cudaMemcpyAsync(thr.data(), p.out_a, Q * sizeof(double), cudaMemcpyDeviceToHost,
                               s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  // Host: the k-th smallest distance is f(A_k); every state at most that far has
  // a >= min{x : f(x) == f(A_k)} (f is non-increasing), and nothing else can be in the list.
  const double clamp_lo = std::nextafter(1.0 - 1e-9, 2.0);  // f == 0 from here up
  for (uint64_t q = 0; q < Q; ++q) {
    const double av = thr[q];
    if (av < 0.0) thr[q] = 0.0;  // fewer than k candidates: all of them (NaN a never passes)
    else if (av > 1.0 - 1e-9) thr[q] = clamp_lo;
    else thr[q] = so3_plateau_low(av);
  }
  // pass 2: gather those states (ascending index) with their a
  OXG_TRY(so3_upload(s, B_AUX2, thr.data(), Q * sizeof(double)));
  p.thr = s->buf[B_AUX2].as<double>();
  std::vector<uint64_t> off(Q + 1);
  uint64_t total = 0;
  OXG_TRY(so3_gather(s, p, qn, true, off.data(), ~0ull, &total));
  if (total == 0) return OXG_OK;
  std::vector<uint32_t> hi(total);
  std::vector<double> ha(total);
  OXG_CUDA_TRY(cudaMemcpyAsync(hi.data(), p.out_idx, total * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                               s->stream));
  OXG_CUDA_TRY(cudaMemcpyAsync(ha.data(), p.out_a, total * sizeof(double), cudaMemcpyDeviceToHost,
                               s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  // Host: distances from libm; stable order by distance keeps the ascending index inside a tie.
  std::vector<uint32_t> order;
  for (uint64_t q = 0; q < Q; ++q) {
    const uint64_t b = off[q], n = off[q + 1] - off[q];
    for (uint64_t e = 0; e < n; ++e) ha[b + e] = so3_f(ha[b + e]);
    order.resize(n);
    for (uint64_t e = 0; e < n; ++e) order[e] = (uint32_t)e;
    std::stable_sort(order.begin(), order.end(),
                     [&](uint32_t x, uint32_t y) { return ha[b + x] < ha[b + y]; });
    for (uint64_t e = 0; e < n && e < k; ++e) {
      idx[q * k + e] = hi[b + order[e]];
      dist[q * k + e] = ha[b + order[e]];
    }
  }
  return OXG_OK;
  OXG_API_END
}

int oxg_so3_radius(oxg_store s, const double* queries, uint64_t Q, double r, uint64_t* offsets,
                   uint32_t* idx, uint64_t cap, double* dist) {
  OXG_API_BEGIN
  OXG_TRY(so3_check_store(s));
  if (!offsets || (Q > 0 && !queries)) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  offsets[0] = 0;
  if (Q == 0) return OXG_OK;
  if (Q > 0x3FFFFFFFull) {
    set_error("too many queries in one call");
    return OXG_E_ARG;
  }
  if (s->size == 0) {
    for (uint64_t q = 0; q < Q; ++q) offsets[q + 1] = 0;
    return OXG_OK;
  }
  OXG_CUDA_TRY(cudaSetDevice(s->device));
  // distance(query, stored) < r  <=>  a >= thr   (NaN r or r <= 0: thr = NaN, nothing matches)
  const double thr = so3_threshold([r](double d) { return d < r; });
  OXG_TRY(so3_upload(s, B_QUERY, queries, Q * 4 * sizeof(double)));
  So3Params p{};
  const int qn = so3_qn(Q);
  so3_plan(s, Q, p, qn);
  p.queries = s->buf[B_QUERY].as<double>();
  p.thr_all = thr;
  uint64_t total = 0;
  OXG_TRY(so3_gather(s, p, qn, dist != nullptr, offsets, idx ? cap : 0, &total));
  if (total == 0) return OXG_OK;
  OXG_CUDA_TRY(cudaMemcpyAsync(idx, p.out_idx, total * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                               s->stream));
  if (dist)
    OXG_CUDA_TRY(cudaMemcpyAsync(dist, p.out_a, total * sizeof(double), cudaMemcpyDeviceToHost,
                                 s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (dist)
    for (uint64_t i = 0; i < total; ++i) dist[i] = so3_f(dist[i]);
  return OXG_OK;
  OXG_API_END
}

int oxg_so3_cones_create(const double* cones, uint64_t n, oxg_so3_cones* out) {
  OXG_API_BEGIN
  if (!out || (n > 0 && !cones)) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  *out = nullptr;
  if (n > 0x7FFFFFFFull) {
    set_error("too many cones");
    return OXG_E_ARG;
  }
  int dev = 0;
  OXG_CUDA_TRY(cudaGetDevice(&dev));
  oxg_so3_cones_s* c = new oxg_so3_cones_s();
  c->device = dev;
  c->n = n;
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
    delete c;
    return OXG_E_CUDA;
  }
  if (n > 0) {
    // valid iff distance(centre, state) > angle  <=>  NOT (f(a) <= angle)  <=>  a < thr
    std::vector<double> h(n * 5);
    for (uint64_t i = 0; i < n; ++i) {
      for (int k = 0; k < 4; ++k) h[i * 5 + k] = cones[i * 5 + k];
      const double angle = cones[i * 5 + 4];
      double thr = so3_threshold([angle](double d) { return d <= angle; });
      // no a reaches the cone (angle < 0 or NaN): "a < +inf" keeps every non-NaN state valid
      // only when the comparison "distance > angle" is true, i.e. angle is not NaN
      if (thr != thr) thr = angle == angle ? std::numeric_limits<double>::infinity() : thr;
      h[i * 5 + 4] = thr;
    }
    e = cudaMalloc(&c->d_cones, h.size() * sizeof(double));
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(c->d_cones, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice,
                          c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
      set_error("cone upload failed: %s", cudaGetErrorString(e));
      oxg_so3_cones_destroy(c);
      return OXG_E_CUDA;
    }
  }
  *out = c;
  return OXG_OK;
  OXG_API_END
}

int oxg_so3_cones_destroy(oxg_so3_cones c) {
  if (!c) return OXG_OK;
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->d_cones) cudaFree(c->d_cones);
  c->states.release();
  c->flags.release();
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return OXG_OK;
}

static int so3_is_valid_impl(oxg_so3_cones c, const double* states, uint64_t Q, uint8_t* out) {
  OXG_CUDA_TRY(cudaSetDevice(c->device));
  OXG_TRY(c->states.ensure(Q * 4 * sizeof(double)));
  OXG_TRY(c->flags.ensure(Q));
  OXG_CUDA_TRY(cudaMemcpyAsync(c->states.p, states, Q * 4 * sizeof(double), cudaMemcpyHostToDevice,
                               c->stream));
  so3_cones_kernel<<<(unsigned)((Q + 255) / 256), 256, 0, c->stream>>>(
      c->d_cones, (uint32_t)c->n, c->states.as<double>(), Q, c->flags.as<uint8_t>());
  OXG_LAUNCH_CHECK();
  OXG_CUDA_TRY(cudaMemcpyAsync(out, c->flags.p, Q, cudaMemcpyDeviceToHost, c->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return OXG_OK;
}

int oxg_so3_is_valid(oxg_so3_cones c, const double* states, uint64_t Q, uint8_t* out) {
  OXG_API_BEGIN
  if (!c || (Q > 0 && (!states || !out))) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (Q == 0) return OXG_OK;
  if (Q > 0x7FFFFFFFull * 256) {
    set_error("too many states in one call");
    return OXG_E_ARG;
  }
  return so3_is_valid_impl(c, states, Q, out);
  OXG_API_END
}

int oxg_so3_check_motion(oxg_so3_cones c, const double* from, const double* to, uint64_t E,
                         double resolution, uint8_t* valid, uint32_t* first_invalid) {
  OXG_API_BEGIN
  if (!c || (E > 0 && (!from || !to || !valid))) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (!(resolution > 0.0)) {
    set_error("resolution must be > 0 (the reference loop would not terminate)");
    return OXG_E_RESOLUTION;
  }
  if (E == 0) return OXG_OK;
  // rrt.rs:90-116: n = ceil(distance / resolution) as usize; n <= 1 tests `to` only, else the
  // states at t = i / n for i = 1..=n, stopping at the first invalid one.
  std::vector<uint64_t> first(E + 1, 0);
  std::vector<double> states;
  for (uint64_t e = 0; e < E; ++e) {
    const double* f = from + e * 4;
    const double* g = to + e * 4;
    const double steps = std::ceil(so3_distance_host(f, g) / resolution);
    uint64_t n = 0;  // Rust `as usize`: NaN -> 0, saturating
    if (steps == steps && steps > 0.0) n = steps >= 4294967296.0 ? 0xFFFFFFFFull : (uint64_t)steps;
    if (n > (1ull << 24)) {
      set_error("edge %llu needs %llu interpolation steps", (unsigned long long)e,
                (unsigned long long)n);
      return OXG_E_ARG;
    }
    if (n <= 1) {
      states.insert(states.end(), g, g + 4);
    } else {
      const size_t at = states.size();
      states.resize(at + n * 4);
      for (uint64_t i = 1; i <= n; ++i)
        so3_interpolate_host(f, g, (double)i / (double)n, states.data() + at + (i - 1) * 4);
    }
    first[e + 1] = states.size() / 4;
  }
  const uint64_t S = states.size() / 4;
  std::vector<uint8_t> ok(S);
  OXG_TRY(so3_is_valid_impl(c, states.data(), S, ok.data()));
  for (uint64_t e = 0; e < E; ++e) {
    uint32_t bad = 0;
    for (uint64_t i = first[e]; i < first[e + 1]; ++i)
      if (!ok[i]) {
        bad = (uint32_t)(i - first[e] + 1);
        break;
      }
    valid[e] = bad == 0;
    if (first_invalid) first_invalid[e] = bad;
  }
  return OXG_OK;
  OXG_API_END
}

// RRT over SO3 (rrt.rs:158-227) on a pre-drawn sample stream.  The loop is sequential -- the
// nearest node of sample j may be a node that sample j-1 added -- but most samples' nearest node is
// an OLD one, so it runs in speculative chunks: one 1-NN launch and one cone launch serve m
// samples against the tree as it stood when the chunk began; the samples are then committed in
// order, and the first one whose nearest node turns out to be a node added earlier in the same
// chunk (host test on <= m new nodes, the same distance function) ends the chunk and is redone
// at the head of the next.  The result is the sequential loop's, iteration for iteration.
int oxg_so3_rrt_solve(oxg_store tree, oxg_so3_cones cones, const oxg_plan_params* params,
                      const double* goal_target, const double* samples, uint64_t n_samples,
                      int64_t* parents, oxg_plan_result* out) {
  OXG_API_BEGIN
  OXG_TRY(so3_check_store(tree));
  if (!cones || !params || !out || !parents || (n_samples > 0 && !samples)) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (!(params->resolution > 0.0)) {
    set_error("resolution must be > 0 (the reference loop would not terminate)");
    return OXG_E_RESOLUTION;
  }
  if (tree->size == 0) {
    set_error("the tree must hold the start state at index 0");
    return OXG_E_EMPTY;
  }
  if (params->max_nodes < tree->size || params->max_nodes > 0x3FFFFFFFull) {
    set_error("max_nodes must lie in [tree size, 2^30)");
    return OXG_E_ARG;
  }
  const auto t_begin = std::chrono::steady_clock::now();
  *out = oxg_plan_result{};
  out->goal_node[0] = out->goal_node[1] = -1;
  const bool has_goal = goal_target && params->goal_radius >= 0.0;
  const double maxd = params->max_distance;
  uint64_t n = tree->size;
  std::vector<double> nodes(n * 4);
  OXG_TRY(oxg_store_read(tree, 0, n, nodes.data()));
  nodes.reserve(std::min<uint64_t>(params->max_nodes, n + n_samples) * 4);
  uint64_t cursor = 0, chunk = 8, chunks = 0;
  std::vector<uint32_t> near;
  std::vector<double> dist, from, to;
  std::vector<uint8_t> valid;
  int status = 1;
  while (cursor < n_samples) {
    if (n >= params->max_nodes) {
      status = 2;
      break;
    }
    if (params->timeout_s > 0.0 &&
        std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count() >
            params->timeout_s) {
      status = 3;
      break;
    }
    const uint64_t m = std::min<uint64_t>({chunk, n_samples - cursor, params->max_nodes - n});
    const double* q = samples + cursor * 4;
    near.resize(m);
    dist.resize(m);
    from.resize(m * 4);
    to.resize(m * 4);
    valid.resize(m);
    OXG_TRY(oxg_so3_nn1(tree, q, m, near.data(), dist.data()));
    for (uint64_t j = 0; j < m; ++j) {  // steering, rrt.rs:200-208
      const double* qn = nodes.data() + (size_t)near[j] * 4;
      std::copy(qn, qn + 4, from.begin() + j * 4);
      if (dist[j] > maxd) so3_interpolate_host(qn, q + j * 4, maxd / dist[j], to.data() + j * 4);
      else std::copy(q + j * 4, q + j * 4 + 4, to.begin() + j * 4);
    }
    OXG_TRY(oxg_so3_check_motion(cones, from.data(), to.data(), m, params->resolution, valid.data(),
                                 nullptr));
    ++chunks;
    const uint64_t n_before = n;
    uint64_t done = 0;
    bool solved = false;
    for (; done < m && !solved; ++done) {
      bool stale = false;  // a node of this chunk is strictly nearer than the old tree's winner
      for (uint64_t p = n_before; p < n && !stale; ++p)
        stale = so3_distance_host(nodes.data() + p * 4, q + done * 4) < dist[done];
      if (stale) break;
      ++out->motion_checks;
      if (!valid[done]) continue;
      nodes.insert(nodes.end(), to.begin() + done * 4, to.begin() + done * 4 + 4);
      parents[n] = near[done];
      ++n;
      if (has_goal && so3_distance_host(to.data() + done * 4, goal_target) <= params->goal_radius) {
        solved = true;
        out->goal_node[0] = (int64_t)n - 1;
      }
    }
    out->motion_checks_executed += m;
    cursor += done;
    if (n > n_before) {
      uint64_t first = 0;
      OXG_TRY(oxg_store_append(tree, nodes.data() + n_before * 4, n - n_before, &first));
    }
    if (solved) {
      status = 0;
      break;
    }
    // a chunk that committed everything earns a longer one; a cut one restarts near its length
    chunk = done == m ? std::min<uint64_t>(chunk * 2, 1024) : std::max<uint64_t>(done, 4);
  }
  if (status == 1 && n >= params->max_nodes && cursor < n_samples) status = 2;
  out->n_nodes[0] = n;
  out->iters = cursor;
  out->solved = status == 0;
  out->status = status;
  out->ctas = (int32_t)std::min<uint64_t>(chunks, 0x7FFFFFFF);
  out->device_ms =
      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  return OXG_OK;
  OXG_API_END
}

// RRT-Connect over SO3 (rrt_connect.rs:227-309, extend :121-159) in speculative chunks, like
// oxg_so3_rrt_solve.  Which tree an iteration grows depends on the sizes the earlier iterations
// left, so both cases are prepared for every sample of a chunk: extend 1 (1-NN of q_rand in tree
// c, steer, motion check) for c = start and c = goal, then extend 2 (1-NN of that q_new in the
// other tree, steer, motion check) for the edges that survived -- six launches per chunk.  The
// commit pass walks the samples in order, picks the case the sizes dictate, and cuts the chunk at
// the first sample one of whose two scans is beaten by a node added inside the chunk.
int oxg_so3_rrt_connect_solve(oxg_store start_tree, oxg_store goal_tree, oxg_so3_cones cones,
                              const oxg_plan_params* params, const double* goal_target,
                              const double* samples, uint64_t n_samples, int64_t* start_parents,
                              int64_t* goal_parents, oxg_plan_result* out) {
  OXG_API_BEGIN
  OXG_TRY(so3_check_store(start_tree));
  OXG_TRY(so3_check_store(goal_tree));
  if (!cones || !params || !out || !start_parents || !goal_parents || (n_samples > 0 && !samples)) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (start_tree == goal_tree) {
    set_error("the two trees must be different stores");
    return OXG_E_ARG;
  }
  if (!(params->resolution > 0.0)) {
    set_error("resolution must be > 0 (the reference loop would not terminate)");
    return OXG_E_RESOLUTION;
  }
  if (start_tree->size == 0 || goal_tree->size == 0) {
    set_error("each tree must hold its root state at index 0");
    return OXG_E_EMPTY;
  }
  const uint64_t cap = params->max_nodes;
  if (cap < start_tree->size || cap < goal_tree->size || cap > 0x3FFFFFFFull) {
    set_error("max_nodes must lie in [tree size, 2^30)");
    return OXG_E_ARG;
  }
  const auto t_begin = std::chrono::steady_clock::now();
  *out = oxg_plan_result{};
  out->goal_node[0] = out->goal_node[1] = -1;
  const bool has_goal = goal_target && params->goal_radius >= 0.0;
  const double maxd = params->max_distance;
  oxg_store tr[2] = {start_tree, goal_tree};
  int64_t* par[2] = {start_parents, goal_parents};
  uint64_t n[2];
  std::vector<double> nodes[2];
  for (int c = 0; c < 2; ++c) {
    n[c] = tr[c]->size;
    nodes[c].resize(n[c] * 4);
    OXG_TRY(oxg_store_read(tr[c], 0, n[c], nodes[c].data()));
  }
  // steering of extend (rrt.rs:200-208 as used by rrt_connect.rs:138-146); returns `reached`
  auto steer = [maxd](const double* q_near, const double* target, double d, double* q_new) {
    if (d > maxd) {
      so3_interpolate_host(q_near, target, maxd / d, q_new);
      return false;
    }
    std::copy(target, target + 4, q_new);
    return true;
  };
  uint64_t cursor = 0, chunk = 8, chunks = 0;
  std::vector<uint32_t> near1[2], near2[2], pos2[2];
  std::vector<double> dist1[2], dist2[2], from, to1[2], q2[2], to2[2];
  std::vector<uint8_t> valid1, valid2, reached2[2];
  std::vector<double> to;
  int status = 1;
  while (cursor < n_samples) {
    if (n[0] >= cap || n[1] >= cap) {
      status = 2;
      break;
    }
    if (params->timeout_s > 0.0 &&
        std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count() >
            params->timeout_s) {
      status = 3;
      break;
    }
    const uint64_t m = std::min<uint64_t>({chunk, n_samples - cursor, cap - std::max(n[0], n[1])});
    const double* q = samples + cursor * 4;
    // extend 1 for both cases: edges [0, m) grow the start tree, [m, 2m) the goal tree
    from.resize(2 * m * 4);
    to.resize(2 * m * 4);
    valid1.resize(2 * m);
    for (int c = 0; c < 2; ++c) {
      near1[c].resize(m);
      dist1[c].resize(m);
      OXG_TRY(oxg_so3_nn1(tr[c], q, m, near1[c].data(), dist1[c].data()));
      for (uint64_t j = 0; j < m; ++j) {
        const double* qn = nodes[c].data() + (size_t)near1[c][j] * 4;
        std::copy(qn, qn + 4, from.begin() + (c * m + j) * 4);
        steer(qn, q + j * 4, dist1[c][j], to.data() + (c * m + j) * 4);
      }
    }
    OXG_TRY(oxg_so3_check_motion(cones, from.data(), to.data(), 2 * m, params->resolution,
                                 valid1.data(), nullptr));
    to1[0].assign(to.begin(), to.begin() + m * 4);
    to1[1].assign(to.begin() + m * 4, to.end());
    // extend 2: the new node of case c is the target of the OTHER tree
    uint64_t e2 = 0, cnt[2];
    for (int c = 0; c < 2; ++c) {
      pos2[c].assign(m, 0xFFFFFFFFu);
      q2[c].clear();
      for (uint64_t j = 0; j < m; ++j)
        if (valid1[c * m + j]) {
          pos2[c][j] = (uint32_t)(q2[c].size() / 4);
          q2[c].insert(q2[c].end(), to1[c].begin() + j * 4, to1[c].begin() + j * 4 + 4);
        }
      cnt[c] = q2[c].size() / 4;
      e2 += cnt[c];
    }
    from.resize(e2 * 4);
    to.resize(e2 * 4);
    valid2.resize(e2);
    for (int c = 0, base = 0; c < 2; base += (int)cnt[c], ++c) {
      near2[c].resize(cnt[c]);
      dist2[c].resize(cnt[c]);
      reached2[c].resize(cnt[c]);
      if (cnt[c] == 0) continue;
      const int o = 1 - c;
      OXG_TRY(oxg_so3_nn1(tr[o], q2[c].data(), cnt[c], near2[c].data(), dist2[c].data()));
      for (uint64_t e = 0; e < cnt[c]; ++e) {
        const double* qn = nodes[o].data() + (size_t)near2[c][e] * 4;
        std::copy(qn, qn + 4, from.begin() + (base + e) * 4);
        reached2[c][e] = steer(qn, q2[c].data() + e * 4, dist2[c][e], to.data() + (base + e) * 4);
      }
    }
    if (e2 > 0)
      OXG_TRY(oxg_so3_check_motion(cones, from.data(), to.data(), e2, params->resolution,
                                   valid2.data(), nullptr));
    ++chunks;
    out->motion_checks_executed += 2 * m + e2;
    // commit in order
    const uint64_t n_before[2] = {n[0], n[1]};
    auto beaten = [&](int t, const double* target, double d) {  // by a node added in this chunk
      for (uint64_t p = n_before[t]; p < n[t]; ++p)
        if (so3_distance_host(nodes[t].data() + p * 4, target) < d) return true;
      return false;
    };
    uint64_t done = 0;
    bool solved = false;
    while (done < m && !solved && n[0] < cap && n[1] < cap) {
      const int c = n[0] <= n[1] ? 0 : 1, o = 1 - c;  // rrt_connect.rs:249-254
      const uint64_t j = done;
      const bool ok1 = valid1[c * m + j] != 0;
      const uint64_t e = ok1 ? pos2[c][j] : 0;
      if (beaten(c, q + j * 4, dist1[c][j])) break;
      if (ok1 && beaten(o, to1[c].data() + j * 4, dist2[c][e])) break;
      ++done;
      ++out->motion_checks;
      if (!ok1) continue;
      const double* q_new = to1[c].data() + j * 4;
      nodes[c].insert(nodes[c].end(), q_new, q_new + 4);
      par[c][n[c]] = near1[c][j];
      const int64_t idx_a = (int64_t)n[c]++;
      if (c == 0 && has_goal && so3_distance_host(q_new, goal_target) <= params->goal_radius) {
        solved = true;  // rrt_connect.rs:271-274
        out->goal_node[0] = idx_a;
        break;
      }
      ++out->motion_checks;
      const uint64_t g = (c == 0 ? 0 : cnt[0]) + e;
      if (!valid2[g]) continue;
      nodes[o].insert(nodes[o].end(), to.begin() + g * 4, to.begin() + g * 4 + 4);
      par[o][n[o]] = near2[c][e];
      const int64_t idx_b = (int64_t)n[o]++;
      if (reached2[c][e]) {  // rrt_connect.rs:280
        solved = true;
        out->goal_node[0] = c == 0 ? idx_a : idx_b;
        out->goal_node[1] = c == 0 ? idx_b : idx_a;
      }
    }
    cursor += done;
    for (int c = 0; c < 2; ++c)
      if (n[c] > n_before[c]) {
        uint64_t first = 0;
        OXG_TRY(oxg_store_append(tr[c], nodes[c].data() + n_before[c] * 4, n[c] - n_before[c], &first));
      }
    if (solved) {
      status = 0;
      break;
    }
    chunk = done == m ? std::min<uint64_t>(chunk * 2, 1024) : std::max<uint64_t>(done, 4);
  }
  if (status == 1 && cursor < n_samples) status = 2;
  out->n_nodes[0] = n[0];
  out->n_nodes[1] = n[1];
  out->iters = cursor;
  out->solved = status == 0;
  out->status = status;
  out->ctas = (int32_t)std::min<uint64_t>(chunks, 0x7FFFFFFF);
  out->device_ms =
      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  return OXG_OK;
  OXG_API_END
}

}  // extern "C"
