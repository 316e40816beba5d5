// oxg_tree.cu — exact k-NN over the store through a 32-ary bounding-box hierarchy.
//
// The batched scans of oxg_nn.cu evaluate every (query, stored state) pair: N pairs per query
// for lists whose members sit in a tiny ball around the query.  For the low-dimensional stores
// the planners build (PRM roadmaps, rrt_star.rs:121-131 / prm.rs:131-138 neighbourhoods) almost
// all of those pairs can be discarded without being evaluated, and the discard can be made
// EXACT: a group of stored states is skipped only when a certified lower bound of every
// reference-order distance in it exceeds the query's current k-th distance.  The answer is the
// same list — the k smallest by (reference distance value, index), bit for bit — because the
// list is ordered by (distance, index) and not by visiting order.
//
// Structure (rebuilt when the store changes; ~1 ms for 1e6 states):
//   * states sorted by a 32-bit space-filling key: the bits of the quantised coordinates
//     interleaved in the order a k-d tree would split (always the dimension whose cell is
//     currently widest), computed on the device from the store's bounding box;
//   * leaves of 32 consecutive sorted states, stored leaf-major SoA ([leaf][d][32] f64) so a warp
//     reads a leaf with D coalesced 256-byte loads;
//   * levels of 32-ary nodes above them, every node the exact (min, max) box of its children,
//     stored [group][2D][32] so a warp tests the 32 children of a node with 2D coalesced loads.
//
// Query (one warp per query, queries themselves processed in key order so neighbouring warps
// touch the same leaves and hit L1): depth-first, nearest child first.  Lane c evaluates the
// box of child c; __ballot / __reduce_min pick the next child; at a leaf lane c evaluates
// stored state c in the reference's operation order (real_vector_state_space.rs:148-153) and
// the survivors of the admission test  d2 <= U  are merged into the warp's sorted list — one
// (distance, index) entry per lane, insertion by ballot + shuffle, lowest index first on ties.
//
// Certified bound.  For a box B and query q let t_d = max(lo_d - q_d, q_d - hi_d, 0).  Every
// state p in B has real squared distance >= sum t_d^2.  The f64 evaluation lb of that sum and
// the reference-order evaluation d2 of the true sum each carry at most D+2 roundings of relative
// size 2^-53, so  d2 >= lb (1 - 1e-13)  whenever lb is not in the underflow range; a node is
// skipped only if  lb (1 - 1e-13) > U  where U >= max{x : sqrt(x) <= s_kth}.  NaN coordinates
// never enter a box (fmin/fmax drop them) and never enter a list (d2 != d2), as in the scans.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "oxg_internal.cuh"
#include "oxg_math.cuh"

namespace oxg {

constexpr int TREE_LEVELS = 5;       // 32^6 = 2^30 states at most (tree_supported checks)
constexpr int TREE_KEY_BITS = 32;
constexpr int TREE_MAX_DIM = 8;      // dimensions the tree kernels are instantiated for
#ifndef OXG_TK_THREADS
#define OXG_TK_THREADS 128
#endif
#ifndef OXG_TK_MIN_CTAS
#define OXG_TK_MIN_CTAS 7
#endif
#ifndef OXG_TK_QPC
#define OXG_TK_QPC 32
#endif
#ifndef OXG_TK_PERSISTENT
#define OXG_TK_PERSISTENT 0
#endif
#ifndef OXG_TK_TWO_LEAVES
#define OXG_TK_TWO_LEAVES 1
#endif
#ifndef OXG_TK_QBOX_SMEM
#define OXG_TK_QBOX_SMEM 1
#endif
#ifndef OXG_TK_MERGE_INLINE
#define OXG_TK_MERGE_INLINE 1
#endif
#ifndef OXG_TK_F32_LEAF
#define OXG_TK_F32_LEAF 1
#endif
#ifndef OXG_TK_MERGE_AT
#define OXG_TK_MERGE_AT 4
#endif
constexpr int TK_THREADS = OXG_TK_THREADS;  // query kernel: warps per CTA x 32
constexpr int TK_QPC = OXG_TK_QPC;          // consecutive (key-ordered) queries per chunk

struct TreePlan {  // device-resident quantisation plan, written by tree_plan_kernel
  double lo[kMaxFastDim];
  double scale[kMaxFastDim];  // cells per unit length
  uint32_t bits[kMaxFastDim];
  uint32_t seq[TREE_KEY_BITS];  // dimension that owns key bit b (most significant first)
  uint32_t nbits;               // key bits in use
  // "box space" of the certified f32 box tests: x' = (x - c_d) * sc with sc a power of two that
  // brings the largest half-extent into [0.5, 1) (clamped to 2^+-100), so the tests are
  // translation- and scale-invariant
  double c[kMaxFastDim];
  double sc, sc2;
  double mabs;  // >= |x'| of every stored coordinate inside the bounding box (box space)
};

struct TreeView {
  const double* pts;     // [n_leaves][D][32]
  const float* pts32;    // the same states in box space, rounded to nearest (filter input only)
  const uint32_t* perm;  // [n_leaves*32] sorted position -> store index (pads: OXG_INVALID_INDEX)
  const double* boxes[TREE_LEVELS];  // level l nodes in groups of 32: [groups][2D][32]
  const float* boxes32[TREE_LEVELS]; // the same boxes in the plan's box space, rounded outward
  uint32_t count[TREE_LEVELS];       // nodes per level (level 0 = leaves)
  int top;                           // highest level; count[top] <= 32
};

struct KdTables;

struct StoreTree {
  uint64_t epoch = ~0ull;
  uint64_t size = 0;
  // measurement (oxg_tuning_set("tree_stats", 1) + oxg_store_tree_stats)
  cudaEvent_t ev_b0 = nullptr, ev_b1 = nullptr, ev_q0 = nullptr, ev_q1 = nullptr;
  bool have_build_ev = false, have_query_ev = false, have_counts = false;
  uint64_t last_Q = 0;
  DevBuf pts32;
  DevBuf pts, perm, keys, keys2, vals, vals2, plan, bounds, qorder, tmp;
  DevBuf boxes[TREE_LEVELS];
  DevBuf boxes32[TREE_LEVELS];
  DevBuf kd_tables, kd_cell_a, kd_cell_b, kd_dim, kd_keys_a, kd_keys_b;
  KdTables* kd_host = nullptr;
  int kind = 0;  // 0 = space-filling key order, 1 = balanced partition order
  TreeView view;
};

// ---- order-preserving u64 encoding of doubles (for atomicMin / atomicMax) ----------------
__device__ __forceinline__ unsigned long long ord_encode(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_decode(unsigned long long e) {
  unsigned long long b = (e & 0x8000000000000000ull) ? (e & 0x7FFFFFFFFFFFFFFFull) : ~e;
  return __longlong_as_double((long long)b);
}

__global__ void tree_bounds_init_kernel(unsigned long long* bounds) {
  const int t = threadIdx.x;
  if (t < kMaxFastDim) {
    bounds[t] = ~0ull;             // min slot
    bounds[kMaxFastDim + t] = 0;   // max slot
  }
}

// Per-dimension min / max over the finite coordinates of the store.
__global__ void tree_bounds_kernel(const double* __restrict__ db, uint64_t pitch, uint32_t n, int dim,
                                   unsigned long long* __restrict__ bounds) {
  for (int d = 0; d < dim; ++d) {
    double lo = __longlong_as_double(0x7FF0000000000000ll), hi = -lo;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const double v = db[(uint64_t)d * pitch + i];
      if (fabs(v) < 1.0e300) {  // finite and not absurd; NaN fails
        lo = fmin(lo, v);
        hi = fmax(hi, v);
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
      atomicMin(&bounds[d], ord_encode(lo));
      atomicMax(&bounds[kMaxFastDim + d], ord_encode(hi));
    }
  }
}

// Allots the key bits: bit after bit goes to the dimension whose cell is currently the widest
// (what a k-d tree splitting the widest side at its midpoint would do on uniform data).
__global__ void tree_plan_kernel(const unsigned long long* __restrict__ bounds, int dim, TreePlan* plan) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double cell[kMaxFastDim];
  bool any = false;
  for (int d = 0; d < dim; ++d) {
    const double lo = ord_decode(bounds[d]), hi = ord_decode(bounds[kMaxFastDim + d]);
    double ext = (lo <= hi) ? hi - lo : 0.0;
    if (!(ext > 0.0) || !(ext < 1.0e300)) ext = 0.0;
    plan->lo[d] = (lo <= hi) ? lo : 0.0;
    plan->bits[d] = 0;
    cell[d] = ext;
    any = any || ext > 0.0;
  }
  uint32_t nb = 0;
  if (any) {
    for (int b = 0; b < TREE_KEY_BITS; ++b) {
      int best = 0;
      for (int d = 1; d < dim; ++d)
        if (cell[d] > cell[best]) best = d;
      plan->seq[b] = (uint32_t)best;
      plan->bits[best] += 1;
      cell[best] *= 0.5;
      ++nb;
    }
  }
  plan->nbits = nb;
  {
    double half = 0.0;
    for (int d = 0; d < dim; ++d) {
      const double lo = ord_decode(bounds[d]), hi = ord_decode(bounds[kMaxFastDim + d]);
      const bool ok = lo <= hi;
      plan->c[d] = ok ? 0.5 * lo + 0.5 * hi : 0.0;
      if (ok) half = fmax(half, fmax(hi - plan->c[d], plan->c[d] - lo));
    }
    int e = 0;
    if (half > 0.0 && half < 1.0e300) {
      int ex = 0;
      frexp(half, &ex);  // half = m 2^ex, m in [0.5, 1)
      e = -ex;
    }
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
    plan->sc = ldexp(1.0, e);
    plan->sc2 = ldexp(1.0, 2 * e);
    plan->mabs = half * plan->sc * 1.000001 + 1e-300;
  }
  for (int d = 0; d < dim; ++d) {
    const double lo = ord_decode(bounds[d]), hi = ord_decode(bounds[kMaxFastDim + d]);
    const double ext = (lo <= hi) ? hi - lo : 0.0;
    double sc = (plan->bits[d] > 0 && ext > 0.0) ? (double)(1ull << plan->bits[d]) / ext : 0.0;
    if (!(sc < 1.0e300)) sc = 0.0;
    plan->scale[d] = sc;
  }
}

// Key of every state of `v` under the plan; vals[i] = i.
__global__ void tree_keys_kernel(StateView v, uint32_t n, int dim, const TreePlan* __restrict__ plan,
                                 uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  __shared__ TreePlan sp;
  for (int t = threadIdx.x; t < (int)(sizeof(TreePlan) / 4); t += blockDim.x)
    reinterpret_cast<uint32_t*>(&sp)[t] = reinterpret_cast<const uint32_t*>(plan)[t];
  __syncthreads();
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t cell[kMaxFastDim];
    uint32_t rem[kMaxFastDim];
    bool bad = false;
    for (int d = 0; d < dim; ++d) {
      const double x = v.p[(uint64_t)i * v.stride_i + (uint64_t)d * v.stride_d];
      const uint32_t nb = sp.bits[d];
      const double top = (double)((1ull << nb) - 1ull);
      double c = floor((x - sp.lo[d]) * sp.scale[d]);
      if (!(c == c)) {
        bad = bad || !(x == x);  // NaN coordinates sort to the end
        c = 0.0;
      }
      c = fmin(fmax(c, 0.0), top);
      cell[d] = (uint32_t)c;
      rem[d] = nb;
    }
    uint32_t key = 0;
    for (uint32_t b = 0; b < sp.nbits; ++b) {
      const uint32_t d = sp.seq[b];
      rem[d] -= 1;
      key = (key << 1) | ((cell[d] >> rem[d]) & 1u);
    }
    keys[i] = bad ? 0xFFFFFFFFu : key;
    vals[i] = i;
  }
}

// One warp per leaf slot: gathers the leaf's states into the leaf-major layout and writes the
// leaf's box.  Slots past the last leaf (padding of the last group) get empty boxes.
template <int D>
__global__ void tree_leaves_kernel(const double* __restrict__ db, uint64_t pitch, uint32_t n,
                                   const uint32_t* __restrict__ sorted, uint32_t n_leaves,
                                   uint32_t n_slots, const TreePlan* __restrict__ plan,
                                   double* __restrict__ pts, float* __restrict__ pts32,
                                   uint32_t* __restrict__ perm, double* __restrict__ boxes0) {
  const int lane = threadIdx.x & 31;
  const uint32_t leaf = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (leaf >= n_slots) return;
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  double lo[D], hi[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    lo[d] = inf;
    hi[d] = -inf;
  }
  if (leaf < n_leaves) {
    const uint32_t pos = leaf * 32u + lane;
    const uint32_t j = pos < n ? sorted[pos] : OXG_INVALID_INDEX;
    perm[pos] = j;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double x = (j != OXG_INVALID_INDEX) ? db[(uint64_t)d * pitch + j]
                                                : __longlong_as_double(0x7FF8000000000000ll);
      pts[((uint64_t)leaf * D + d) * 32 + lane] = x;
      pts32[((uint64_t)leaf * D + d) * 32 + lane] = __double2float_rn((x - plan->c[d]) * plan->sc);
      lo[d] = fmin(lo[d], x);  // fmin / fmax drop NaN
      hi[d] = fmax(hi[d], x);
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d)
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fmin(lo[d], __shfl_xor_sync(0xFFFFFFFFu, lo[d], o));
      hi[d] = fmax(hi[d], __shfl_xor_sync(0xFFFFFFFFu, hi[d], o));
    }
  if (lane == 0) {
    double* g = boxes0 + (uint64_t)(leaf >> 5) * (2 * D * 32) + (leaf & 31u);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      g[d * 32] = lo[d];
      g[(D + d) * 32] = hi[d];
    }
  }
}

// One warp per node slot of level l+1: the union of its 32 children (a group of level l).
template <int D>
__global__ void tree_level_kernel(const double* __restrict__ child_boxes, uint32_t n_children,
                                  uint32_t n_nodes, uint32_t n_slots, double* __restrict__ boxes) {
  const int lane = threadIdx.x & 31;
  const uint32_t node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (node >= n_slots) return;
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  double lo[D], hi[D];
  const bool have = node < n_nodes && node * 32u + lane < n_children;
  const double* c = child_boxes + (uint64_t)node * (2 * D * 32) + lane;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    lo[d] = have ? c[d * 32] : inf;
    hi[d] = have ? c[(D + d) * 32] : -inf;
  }
#pragma unroll
  for (int d = 0; d < D; ++d)
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fmin(lo[d], __shfl_xor_sync(0xFFFFFFFFu, lo[d], o));
      hi[d] = fmax(hi[d], __shfl_xor_sync(0xFFFFFFFFu, hi[d], o));
    }
  if (lane == 0) {
    double* g = boxes + (uint64_t)(node >> 5) * (2 * D * 32) + (node & 31u);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      g[d * 32] = lo[d];
      g[(D + d) * 32] = hi[d];
    }
  }
}

// ---------------------------------------------------------------------------------------
// Balanced partition order ("k-d order").  The space-filling key above is cheap but its leaves
// are arbitrary runs of 32 along the curve: where the curve jumps, a leaf's box spans two distant
// cells (measured on the config-4 store: the leaf boxes add up to 5.6x the volume of the domain
// and a query's k-th-distance ball meets 129 of them; with the partition below 0.67x and 63).
// Here every node of the 32-ary hierarchy is cut into its (up to) 32 children by three rounds of
// splits — 4 ways, 4 ways, 2 ways — each along the widest side of the part's cell, at equal
// counts of children, so children are compact cells and sibling groups never straddle a jump.
// One round = one radix sort of (segment id, quantised coordinate) over all states; the segment
// tables depend on N only and are made on the host.  3 rounds per hierarchy level (9 for 1e6).
// ---------------------------------------------------------------------------------------
struct KdRound {           // device-resident tables of one round (segmentation before the round)
  uint32_t n_seg;          // segments before the round
  uint32_t n_child;        // segments after it
  const uint32_t* begin;   // [n_seg + 1] first position (in points) of every segment
  const uint8_t* parts;    // [n_seg] parts this round cuts the segment into (1 = untouched)
  const uint32_t* c_begin; // [n_child + 1] positions of the segments after the round
  const uint32_t* c_parent;// [n_child]
  const uint8_t* c_part;   // [n_child] which part of its parent
};

// The dimension every splitting segment is cut along: the widest side of its cell.
__global__ void kd_choose_dim_kernel(KdRound r, int dim, const double* __restrict__ cell,
                                     uint8_t* __restrict__ seg_dim) {
  for (uint32_t sidx = blockIdx.x * blockDim.x + threadIdx.x; sidx < r.n_seg; sidx += gridDim.x * blockDim.x) {
    const double* c = cell + (size_t)sidx * 2 * dim;
    int best = 0;
    double w = -1.0;
    for (int d = 0; d < dim; ++d) {
      const double e = c[dim + d] - c[d];
      if (e > w) {  // NaN / empty extents never win
        w = e;
        best = d;
      }
    }
    seg_dim[sidx] = (uint8_t)best;
  }
}

// key = (segment << sub_bits) | order-preserving sub_bits-bit image of the split coordinate
// within the cell (untouched segments keep their order: the position inside the segment, which
// fits because such a segment is at most one child's worth of states).  32-bit keys: a radix sort
// of four 8-bit passes instead of six.
__global__ void kd_keys_kernel(KdRound r, int dim, const double* __restrict__ db, uint64_t pitch, uint32_t n,
                               const uint32_t* __restrict__ order, const double* __restrict__ cell,
                               const uint8_t* __restrict__ seg_dim, int sub_bits, uint32_t* __restrict__ keys) {
  const uint32_t sub_max = (sub_bits >= 32) ? 0xFFFFFFFFu : ((1u << sub_bits) - 1u);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t lo = 0, hi = r.n_seg;  // last segment with begin <= i
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (__ldg(r.begin + mid) <= i) lo = mid; else hi = mid;
    }
    uint32_t sub = min(i - __ldg(r.begin + lo), sub_max);
    if (__ldg(r.parts + lo) > 1) {
      const int d = seg_dim[lo];
      const double c0 = cell[(size_t)lo * 2 * dim + d], c1 = cell[(size_t)lo * 2 * dim + dim + d];
      const double x = db[(uint64_t)d * pitch + order[i]];
      double t = (x - c0) / (c1 - c0);  // NaN (x NaN, or a zero-width cell) sorts last / together
      t = fmin(fmax(t, 0.0), 1.0) * (double)sub_max;
      sub = (t == t) ? (uint32_t)t : sub_max;
      if (!(x == x)) sub = sub_max;
    }
    keys[i] = (sub_bits >= 32) ? sub : ((lo << sub_bits) | sub);
  }
}

// Cells after the round: the parent's cell with the split side narrowed to the part's own
// first / last coordinate (the states are sorted along it).
__global__ void kd_cells_kernel(KdRound r, int dim, const double* __restrict__ db, uint64_t pitch, uint32_t n,
                                const uint32_t* __restrict__ order, const double* __restrict__ cell,
                                const uint8_t* __restrict__ seg_dim, double* __restrict__ cell_out) {
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < r.n_child; c += gridDim.x * blockDim.x) {
    const uint32_t par = r.c_parent[c];
    const double* pc = cell + (size_t)par * 2 * dim;
    double* oc = cell_out + (size_t)c * 2 * dim;
    for (int d = 0; d < 2 * dim; ++d) oc[d] = pc[d];
    const uint32_t parts = r.parts[par];
    if (parts > 1) {
      const int d = seg_dim[par];
      const uint32_t b = r.c_begin[c], e = r.c_begin[c + 1];
      if (e > b) {
        const uint32_t part = r.c_part[c];
        const double first = db[(uint64_t)d * pitch + order[b]];
        const double last = db[(uint64_t)d * pitch + order[e - 1]];
        if (part > 0 && first == first) oc[d] = first;
        if (part + 1 < parts && last == last) oc[dim + d] = last;
      }
    }
  }
}

__global__ void kd_iota_kernel(uint32_t* __restrict__ v, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = i;
}

__global__ void kd_root_cell_kernel(const unsigned long long* __restrict__ bounds, int dim, double* cell) {
  const int d = threadIdx.x;
  if (d < dim) {
    const double lo = ord_decode(bounds[d]), hi = ord_decode(bounds[kMaxFastDim + d]);
    cell[d] = (lo <= hi) ? lo : 0.0;
    cell[dim + d] = (lo <= hi) ? hi : 0.0;
  }
}

// Host tables of all rounds for n states (they depend on n alone).
struct KdTables {
  uint64_t n = 0;
  struct Round {
    std::vector<uint32_t> begin, c_begin, c_parent;
    std::vector<uint8_t> parts, c_part;
  };
  std::vector<Round> rounds;
  std::vector<KdRound> dev_rounds;  // device views of the uploaded tables
  uint32_t max_seg = 1;
  bool uploaded = false;
};

static void kd_make_tables(uint64_t n, KdTables* out) {
  out->n = n;
  out->rounds.clear();
  out->uploaded = false;
  const uint64_t n_leaves = (n + 31) / 32;
  // hierarchy levels above the leaves: a node of level l spans 32^l leaves
  int top = 0;
  for (uint64_t c = n_leaves; c > 32; c = (c + 31) / 32) ++top;
  struct Seg { uint64_t a, b; };  // in leaves
  std::vector<Seg> segs{{0, n_leaves}};
  static const int kArity[3] = {4, 4, 2};
  for (int lvl = top + 1; lvl >= 1; --lvl) {
    uint64_t unit = 1;
    for (int i = 1; i < lvl; ++i) unit *= 32;  // leaves per child of a level-`lvl` node
    for (int sub = 0; sub < 3; ++sub) {
      KdTables::Round rd;
      std::vector<Seg> next;
      bool any = false;
      for (size_t si = 0; si < segs.size(); ++si) {
        const Seg sg = segs[si];
        const uint64_t nch = (sg.b - sg.a + unit - 1) / unit;
        const uint64_t parts = std::min<uint64_t>((uint64_t)kArity[sub], std::max<uint64_t>(nch, 1));
        rd.begin.push_back((uint32_t)std::min<uint64_t>(sg.a * 32, n));
        rd.parts.push_back((uint8_t)parts);
        any = any || parts > 1;
        const uint64_t base = nch / parts, extra = nch % parts;
        uint64_t a = sg.a;
        for (uint64_t pt = 0; pt < parts; ++pt) {
          const uint64_t c = base + (pt < extra ? 1 : 0);
          const uint64_t b = parts == 1 ? sg.b : std::min<uint64_t>(a + c * unit, sg.b);
          rd.c_begin.push_back((uint32_t)std::min<uint64_t>(a * 32, n));
          rd.c_parent.push_back((uint32_t)si);
          rd.c_part.push_back((uint8_t)pt);
          next.push_back({a, b});
          a = b;
        }
      }
      rd.begin.push_back((uint32_t)n);
      rd.c_begin.push_back((uint32_t)n);
      segs.swap(next);
      if (any) out->rounds.push_back(std::move(rd));
    }
  }
}

// f64 boxes of one level -> box space, rounded outward (directed rounding at every step, so
// lo32 <= (lo - c) sc and hi32 >= (hi - c) sc hold exactly).
__global__ void tree_boxes32_kernel(const double* __restrict__ boxes, uint64_t n_vals, int dim,
                                    const TreePlan* __restrict__ plan, float* __restrict__ out) {
  const double sc = plan->sc;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vals;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const int row = (int)((i >> 5) % (uint64_t)(2 * dim));
    const double v = boxes[i];
    if (row < dim)
      out[i] = __double2float_rd(__dmul_rd(__dsub_rd(v, plan->c[row]), sc));
    else
      out[i] = __double2float_ru(__dmul_ru(__dsub_ru(v, plan->c[row - dim]), sc));
  }
}

// ---------------------------------------------------------------------------------------
// query kernel
// ---------------------------------------------------------------------------------------
struct TreeKnnParams {
  TreeView t;
  const TreePlan* plan;
  StateView q;
  uint32_t Q;
  const uint32_t* qorder;  // [Q] query ids in key order
  uint32_t k;
  int exclude_mode;
  const uint32_t* exclude;
  uint32_t self_base;
  uint32_t* out_idx;   // [Q][k]
  double* out_dist;
  uint32_t qpc;               // consecutive key-ordered queries per CTA chunk
  unsigned long long* stats;  // nullable: [0] leaves visited, [1] box groups tested, [2] candidates, [3] merges,
                              // [4] leaves that reached the exact f64 path
};

constexpr int TK_MERGE_AT = OXG_TK_MERGE_AT;  // buffered candidates that trigger a merge into the list

// STATS: the per-query counters of oxg_store_tree_stats (a measurement build of the kernel: the
// counters cost registers the 7-CTA budget does not have).
template <int D, bool STATS = false>
struct WarpQuery {
#if OXG_TK_F32_LEAF
  float q32[D];          // the query in box space, rounded to nearest (filter input only)
  const double* qd;      // shared: the query itself (exact path)
  float T32;             // filter threshold: a32 > T32 proves d2 > U
  double M;              // >= every |coordinate| the filter sees, in box space
#else
  double q[D];
#endif
#if OXG_TK_QBOX_SMEM
  const float* qbox;     // shared: the query in the box space, [0..D) rounded down, [D..2D) rounded up
#else
  float qbox[2 * D];     // the query in the box space, [0..D) rounded down, [D..2D) rounded up
#endif
  double my_s;           // this lane's list entry (lane e holds the e-th smallest)
  uint32_t my_i;
  int have;
  int k;
  unsigned long long ub;  // bits of U (admission: bits(d2) <= ub, unsigned: every NaN pattern is above)
  float Uf;               // U in the box space, rounded up (pruning: lb32 > Uf)
  int bc;                 // buffered candidates
  double* b_d2;           // this warp's candidate buffer (shared memory)
  uint32_t* b_j;
  uint32_t excl;
  int lane;
  unsigned n_leaf, n_box, n_cand, n_merge, n_exact;
};

// Upper bound of max{x : sqrt(x) <= sw}: any upper bound is correct here because admitted
// candidates still go through the exact (sqrt, index) ordering of the merge.
__device__ __forceinline__ double tree_sqrt_le_upper(double sw) {
  const double u = __dmul_ru(sw, sw);
  if (!(u > 1.0e-290)) return sqrt_le_threshold_dev(sw);
  return __dadd_ru(u, __dmul_ru(u, 8.8817841970012523e-16));  // 2^-50
}

// Threshold of the FP32 leaf filter (difference form, the bound of filter_threshold_dev in
// oxg_nn.cu applied in box space).  With u = 2^-24, M >= |q'_d|, |p'_d| and a = 4uM: whenever the
// exact scaled sum S' <= Ub, the FP32 evaluation a32 = sum fma(df, df, .), df = fl32(q32 - p32),
// satisfies  a32 <= [S + 2a sqrt(D S) + 2 D a^2](1 + D 2^-22),  S = Ub (1 + 1e-5).  A pair with
// a32 > T32 therefore has reference-order d2 > U (Ub already carries the 1e-12 of the f64
// roundings); everything else, including NaN and overflow, takes the exact path.
__device__ __forceinline__ float tree_filter_threshold(double Ub, double M, int D) {
  const float inf = __int_as_float(0x7F800000);
  if (!(Ub < 1.0e300) || !(M < 1.0e30)) return inf;
  const double a = M * 2.384185791015625e-07 + 1e-37;  // 2^-22 M
  const double S = Ub * 1.00001;
  double T = S + 2.0 * a * sqrt((double)D * S) + 2.0 * (double)D * a * a + 1e-35;
  T = T * (1.0 + (double)D * 2.384185791015625e-07);
  if (!(T == T)) return inf;
  return __double2float_ru(T);  // rounds up; overflows to +inf
}

template <int D, bool S>
__device__ __forceinline__ void set_bound(const TreeKnnParams& p, WarpQuery<D, S>& w, double sk) {
  const double U = tree_sqrt_le_upper(sk);
  w.ub = (unsigned long long)__double_as_longlong(U);
  // box space: squared lengths scale by sc2 (a power of two); 1e-12 covers the roundings of the
  // reference-order evaluation (d2 >= true (1 - (D+2) 2^-53))
  const double Ub = __dmul_ru(__dmul_ru(U, __ldg(&p.plan->sc2)), 1.0 + 1.0e-12);
  w.Uf = __double2float_ru(Ub);
#if OXG_TK_F32_LEAF
  w.T32 = tree_filter_threshold(Ub, w.M, D);
#endif
}

// (s, j) pairs: lexicographic "a before b".
__device__ __forceinline__ bool pair_before(double sa, uint32_t ja, double sb, uint32_t jb) {
  return (sa < sb) || (sa == sb && ja < jb);
}

// Merges the buffered candidates into the sorted list: bitonic sort of the buffer across the
// lanes, min-exchange against the reversed list, bitonic merge.  Order is the exact
// (sqrt(d2), index) one, so visiting order never shows in the result.
template <int D, bool S>
#if OXG_TK_MERGE_INLINE
__device__ __forceinline__ void merge_buffer(
#else
__device__ __noinline__ void merge_buffer(
#endif
const TreeKnnParams& p, WarpQuery<D, S>& w) {
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  __syncwarp();
  double s = inf;
  uint32_t j = OXG_INVALID_INDEX;
  if (w.lane < w.bc) {
    j = w.b_j[w.lane];
    s = sqrt(w.b_d2[w.lane]);
    if (j == w.excl) {  // self-exclusion by index (oxg_knn contract)
      s = inf;
      j = OXG_INVALID_INDEX;
    }
  }
  __syncwarp();
  const int filled = w.bc;
  w.bc = 0;
  if constexpr (S) ++w.n_merge;
  // lanes >= filled hold +inf, so once k2 / 2 >= filled the lower half is sorted ascending and the
  // upper half is all sentinels: the remaining stages would not move anything
#pragma unroll
  for (int k2 = 2; k2 <= 32; k2 <<= 1) {
    if (k2 > 2 && (k2 >> 1) >= filled) break;
#pragma unroll
    for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
      const double so = __shfl_xor_sync(0xFFFFFFFFu, s, j2);
      const uint32_t jo = __shfl_xor_sync(0xFFFFFFFFu, j, j2);
      const bool asc = (w.lane & k2) == 0;
      const bool lower = (w.lane & j2) == 0;
      const bool take_min = (lower == asc);
      const bool other_first = pair_before(so, jo, s, j);
      if (take_min == other_first) {
        s = so;
        j = jo;
      }
    }
  }
  // the 32 smallest of list + buffer, as a bitonic sequence
  {
    const double so = __shfl_sync(0xFFFFFFFFu, s, 31 - w.lane);
    const uint32_t jo = __shfl_sync(0xFFFFFFFFu, j, 31 - w.lane);
    if (pair_before(so, jo, w.my_s, w.my_i)) {
      w.my_s = so;
      w.my_i = jo;
    }
  }
#pragma unroll
  for (int j2 = 16; j2 > 0; j2 >>= 1) {
    const double so = __shfl_xor_sync(0xFFFFFFFFu, w.my_s, j2);
    const uint32_t jo = __shfl_xor_sync(0xFFFFFFFFu, w.my_i, j2);
    const bool lower = (w.lane & j2) == 0;
    const bool other_first = pair_before(so, jo, w.my_s, w.my_i);
    if (lower == other_first) {
      w.my_s = so;
      w.my_i = jo;
    }
  }
  if (w.lane >= w.k) {  // lanes past the list stay sentinels
    w.my_s = inf;
    w.my_i = OXG_INVALID_INDEX;
  }
  w.have = __popc(__ballot_sync(0xFFFFFFFFu, w.my_i != OXG_INVALID_INDEX));
  if (w.have == w.k) set_bound(p, w, __shfl_sync(0xFFFFFFFFu, w.my_s, w.k - 1));
}

#if OXG_TK_F32_LEAF
// FP32 filter value of state `lane` of a leaf (box space; decides nothing, see tree_filter_threshold).
template <int D, bool S>
__device__ __forceinline__ float leaf_a32(const TreeKnnParams& p, const WarpQuery<D, S>& w, uint32_t leaf) {
  const float* lp = p.t.pts32 + (uint64_t)leaf * (D * 32) + w.lane;
  float pv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) pv[d] = __ldg(lp + d * 32);
  float a = 0.0f;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float df = w.q32[d] - pv[d];
    a = __fmaf_rn(df, df, a);
  }
  return a;
}

// The states the filter could not discard are evaluated in the reference's operation order
// (real_vector_state_space.rs:148-153) and go through the admission test; the survivors go to the
// candidate buffer.
template <int D, bool S>
__device__ __forceinline__ void leaf_commit(const TreeKnnParams& p, WarpQuery<D, S>& w, uint32_t leaf, float a32) {
  if constexpr (S) ++w.n_leaf;
  if (__ballot_sync(0xFFFFFFFFu, !(a32 > w.T32)) == 0u) return;
  if constexpr (S) ++w.n_exact;
  double d2 = __longlong_as_double(0x7FF8000000000000ll);
  if (!(a32 > w.T32)) {
    const double* lp = p.t.pts + (uint64_t)leaf * (D * 32) + w.lane;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double df = __dsub_rn(w.qd[d], __ldg(lp + d * 32));
      const double m = __dmul_rn(df, df);
      d2 = (d == 0) ? m : __dadd_rn(d2, m);
    }
  }
  // d2 >= +0 or NaN; every NaN pattern (pads, filtered lanes) compares above any ub as unsigned
  bool pass = (unsigned long long)__double_as_longlong(d2) <= w.ub;
  uint32_t m = __ballot_sync(0xFFFFFFFFu, pass);
  if (m == 0) return;
  int n = __popc(m);
  if (w.bc + n > 32) {
    merge_buffer(p, w);
    pass = (unsigned long long)__double_as_longlong(d2) <= w.ub;  // the bound moved
    m = __ballot_sync(0xFFFFFFFFu, pass);
    n = __popc(m);
  }
  if (pass) {
    const int slot = w.bc + __popc(m & ((1u << w.lane) - 1u));
    w.b_d2[slot] = d2;
    w.b_j[slot] = __ldg(p.t.perm + (uint64_t)leaf * 32 + w.lane);
  }
  w.bc += n;
  if constexpr (S) w.n_cand += n;
  if (w.bc >= TK_MERGE_AT || w.have < w.k) merge_buffer(p, w);
}
#define OXG_LEAF_EVAL leaf_a32
#else
// Reference-order squared distance of the query to state `lane` of a leaf
// (real_vector_state_space.rs:148-153).
template <int D, bool S>
__device__ __forceinline__ double leaf_d2(const TreeKnnParams& p, const WarpQuery<D, S>& w, uint32_t leaf) {
  const double* lp = p.t.pts + (uint64_t)leaf * (D * 32) + w.lane;
  double pv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) pv[d] = __ldg(lp + d * 32);
  double d2 = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double df = __dsub_rn(w.q[d], pv[d]);
    const double m = __dmul_rn(df, df);
    d2 = (d == 0) ? m : __dadd_rn(d2, m);
  }
  return d2;
}

// Admission test of a leaf's 32 distances; the survivors go to the candidate buffer.
template <int D, bool S>
__device__ __forceinline__ void leaf_commit(const TreeKnnParams& p, WarpQuery<D, S>& w, uint32_t leaf,
                                            double d2) {
  // d2 >= +0 or NaN; every NaN pattern (pads included) compares above any ub as unsigned
  bool pass = (unsigned long long)__double_as_longlong(d2) <= w.ub;
  uint32_t m = __ballot_sync(0xFFFFFFFFu, pass);
  if constexpr (S) ++w.n_leaf;
  if (m == 0) return;
  int n = __popc(m);
  if (w.bc + n > 32) {
    merge_buffer(p, w);
    pass = (unsigned long long)__double_as_longlong(d2) <= w.ub;  // the bound moved
    m = __ballot_sync(0xFFFFFFFFu, pass);
    n = __popc(m);
  }
  if (pass) {
    const int slot = w.bc + __popc(m & ((1u << w.lane) - 1u));
    w.b_d2[slot] = d2;
    w.b_j[slot] = __ldg(p.t.perm + (uint64_t)leaf * 32 + w.lane);
  }
  w.bc += n;
  if constexpr (S) w.n_cand += n;
  if (w.bc >= TK_MERGE_AT || w.have < w.k) merge_buffer(p, w);
}

#define OXG_LEAF_EVAL leaf_d2
#endif

// Tests the 32 children of `group` at level L and descends into the survivors, nearest first.
// Every lane owns one child: `mine` says it is still to be visited, and the lane whose key equals
// the warp minimum is the one being visited (keys carry the lane in their low bits, so they are
// unique) -- no lane index is needed inside the loop.
template <int D, int L>
struct Visit {
  template <bool S>
  static __device__ __forceinline__ void run(const TreeKnnParams& p, WarpQuery<D, S>& w, uint32_t group) {
    const uint32_t child0 = group * 32u;
    const float* bp = p.t.boxes32[L] + (uint64_t)group * (2 * D * 32) + w.lane;
    // certified lower bound in the box space: boxes rounded outward, the query bracketed, every
    // operation rounded down
    float blo[D], bhi[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      blo[d] = __ldg(bp + d * 32);
      bhi[d] = __ldg(bp + (D + d) * 32);
    }
    float lb = 0.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float t = fmaxf(fmaxf(__fsub_rd(blo[d], w.qbox[D + d]), __fsub_rd(w.qbox[d], bhi[d])), 0.0f);
      lb = __fadd_rd(lb, __fmul_rd(t, t));
    }
    if constexpr (S) ++w.n_box;
    const uint32_t key = (__float_as_uint(lb) & 0xFFFFFFE0u) | (uint32_t)w.lane;
    bool mine = child0 + (uint32_t)w.lane < p.t.count[L];
    for (;;) {
      const bool alive = mine && !(lb > w.Uf);
      const uint32_t r = __reduce_min_sync(0xFFFFFFFFu, alive ? key : 0xFFFFFFFFu);
      if (r == 0xFFFFFFFFu) break;
      const uint32_t c = r & 31u;
      if (key == r) mine = false;
      if constexpr (L == 0) {
        // two leaves per round: both loads are in flight before either is consumed
        const uint32_t r1 = OXG_TK_TWO_LEAVES
                                ? __reduce_min_sync(0xFFFFFFFFu, (alive && key != r) ? key : 0xFFFFFFFFu)
                                : 0xFFFFFFFFu;
        if (OXG_TK_TWO_LEAVES && r1 != 0xFFFFFFFFu) {
          const uint32_t c1 = r1 & 31u;
          if (key == r1) mine = false;
          const auto da = OXG_LEAF_EVAL(p, w, child0 + c);
          const auto db = OXG_LEAF_EVAL(p, w, child0 + c1);
          leaf_commit(p, w, child0 + c, da);
          leaf_commit(p, w, child0 + c1, db);
        } else {
          leaf_commit(p, w, child0 + c, OXG_LEAF_EVAL(p, w, child0 + c));
        }
      } else {
        Visit<D, L - 1>::run(p, w, child0 + c);
      }
    }
  }
};

// Persistent CTAs: chunks of TK_QPC key-ordered queries are dealt round-robin to the CTAs, so all
// warps of an SM work on neighbouring queries at the same time and share the leaves through L1.
template <int D, bool STATS>
__global__ void __launch_bounds__(TK_THREADS, OXG_TK_MIN_CTAS) tree_knn_kernel(const TreeKnnParams p) {
#if OXG_TK_F32_LEAF
  __shared__ double s_q[TK_THREADS / 32][D];
#endif
  __shared__ double s_d2[TK_THREADS];
  __shared__ uint32_t s_j[TK_THREADS];
#if OXG_TK_QBOX_SMEM
  __shared__ float s_qbox[TK_THREADS / 32][2 * D];
#endif
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  const double sc = p.plan->sc;
  [[maybe_unused]] unsigned long long st_leaf = 0, st_box = 0, st_cand = 0, st_merge = 0, st_exact = 0;
  const uint32_t n_chunks = (p.Q + p.qpc - 1) / p.qpc;
  for (uint32_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const uint32_t q_begin = chunk * p.qpc;
    const uint32_t q_end = min(p.Q, q_begin + p.qpc);
    for (uint32_t wq = q_begin + wid; wq < q_end; wq += TK_THREADS / 32) {
      const uint32_t qi = p.qorder ? __ldg(p.qorder + wq) : wq;
      WarpQuery<D, STATS> w;
      bool qnan = false;
      __syncwarp();
#if OXG_TK_F32_LEAF
      double mq = 0.0;
#endif
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double qv = p.q.p[(uint64_t)qi * p.q.stride_i + (uint64_t)d * p.q.stride_d];
        qnan = qnan || !(qv == qv);
        const double c = p.plan->c[d];
#if OXG_TK_F32_LEAF
        const double qs = (qv - c) * sc;
        w.q32[d] = __double2float_rn(qs);
        mq = fmax(mq, fabs(qs));
        if (lane == 0) s_q[wid][d] = qv;
#else
        w.q[d] = qv;
#endif
#if OXG_TK_QBOX_SMEM
        if (lane == 0) {
          s_qbox[wid][d] = __double2float_rd(__dmul_rd(__dsub_rd(qv, c), sc));
          s_qbox[wid][D + d] = __double2float_ru(__dmul_ru(__dsub_ru(qv, c), sc));
        }
#else
        w.qbox[d] = __double2float_rd(__dmul_rd(__dsub_rd(qv, c), sc));
        w.qbox[D + d] = __double2float_ru(__dmul_ru(__dsub_ru(qv, c), sc));
#endif
      }
#if OXG_TK_F32_LEAF
      w.qd = s_q[wid];
      w.M = fmax(mq, p.plan->mabs) * 1.000001;  // inf for a query far outside: the filter then passes everything
      w.T32 = __int_as_float(0x7F800000);
#endif
      __syncwarp();
#if OXG_TK_QBOX_SMEM
      w.qbox = s_qbox[wid];
#endif
      w.my_s = inf;
      w.my_i = OXG_INVALID_INDEX;
      w.have = 0;
      w.k = (int)p.k;
      w.ub = 0x7FF0000000000000ull;
      w.Uf = __int_as_float(0x7F800000);
      w.bc = 0;
      w.b_d2 = s_d2 + wid * 32;
      w.b_j = s_j + wid * 32;
      w.lane = lane;
      w.n_leaf = w.n_box = w.n_cand = w.n_merge = w.n_exact = 0;
      w.excl = OXG_INVALID_INDEX;
      if (p.exclude_mode == 1) w.excl = __ldg(p.exclude + qi);
      if (p.exclude_mode == 2) w.excl = p.self_base + qi;
      if (!qnan) {  // a NaN query admits nothing (every distance is NaN)
        switch (p.t.top) {
          case 0: Visit<D, 0>::run(p, w, 0); break;
          case 1: Visit<D, 1>::run(p, w, 0); break;
          case 2: Visit<D, 2>::run(p, w, 0); break;
          case 3: Visit<D, 3>::run(p, w, 0); break;
          default: Visit<D, 4>::run(p, w, 0); break;
        }
        if (w.bc > 0) merge_buffer(p, w);
      }
      if (lane < (int)p.k) {
        const uint64_t o = (uint64_t)qi * p.k + lane;
        p.out_idx[o] = w.my_i;
        p.out_dist[o] = w.my_s;
      }
      if constexpr (STATS) {
        st_leaf += w.n_leaf;
        st_box += w.n_box;
        st_cand += w.n_cand;
        st_merge += w.n_merge;
        st_exact += w.n_exact;
      }
    }
  }
  if (STATS && p.stats && lane == 0) {
    atomicAdd(&p.stats[0], st_leaf);
    atomicAdd(&p.stats[1], st_box);
    atomicAdd(&p.stats[2], st_cand);
    atomicAdd(&p.stats[3], st_merge);
    atomicAdd(&p.stats[4], st_exact);
  }
}

// ---------------------------------------------------------------------------------------
// radius search through the hierarchy: all i with distance(query, stored_i) < r (strict), the
// scans of rrt_star.rs:121-131, prm.rs:131-138 and prm.rs:249-256.  Same certified pruning with a
// fixed bound U = max{x : sqrt(x) < r}; a count pass and a fill pass (the fill emits
// (query << 32 | index) keys, one radix sort puts every row in ascending index order).
// ---------------------------------------------------------------------------------------
struct TreeRadiusParams {
  TreeView t;
  const TreePlan* plan;
  StateView q;
  uint32_t Q;
  const uint32_t* qorder;
  unsigned long long ub;       // bits of U (admission: bits(d2) <= ub)
  double U;
  const uint32_t* idx_limit;   // nullable
  int limit_self;              // limit of query qi = self_base + qi
  uint32_t self_base;
  uint32_t n_db;
  uint64_t* counts;            // count pass: hits per query; fill pass: exclusive offsets
  unsigned long long* keys;    // fill pass
  double* dist;                // fill pass, nullable
  uint32_t qpc;
};

template <int D>
struct RadiusQuery {
  float q32[D];
  const double* qd;
  const float* qbox;
  float T32, Uf;
  unsigned long long ub;
  uint32_t limit, qi;
  uint64_t base;     // fill: first output slot of this query
  uint32_t found;
  int lane;
};

template <int D, bool FILL>
__device__ __forceinline__ void radius_leaf(const TreeRadiusParams& p, RadiusQuery<D>& w, uint32_t leaf) {
  const float* lp32 = p.t.pts32 + (uint64_t)leaf * (D * 32) + w.lane;
  float a = 0.0f;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float df = w.q32[d] - __ldg(lp32 + d * 32);
    a = __fmaf_rn(df, df, a);
  }
  const bool maybe = !(a > w.T32);
  if (__ballot_sync(0xFFFFFFFFu, maybe) == 0u) return;
  bool hit = false;
  double d2 = 0.0;
  uint32_t j = OXG_INVALID_INDEX;
  if (maybe) {
    const double* lp = p.t.pts + (uint64_t)leaf * (D * 32) + w.lane;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double df = __dsub_rn(w.qd[d], __ldg(lp + d * 32));
      const double m = __dmul_rn(df, df);
      d2 = (d == 0) ? m : __dadd_rn(d2, m);
    }
    j = __ldg(p.t.perm + (uint64_t)leaf * 32 + w.lane);
    // NaN patterns compare above any ub as unsigned; pads carry NaN coordinates
    hit = (unsigned long long)__double_as_longlong(d2) <= w.ub && j < w.limit;
  }
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
  if (FILL && hit) {
    const uint64_t pos = w.base + w.found + __popc(m & ((1u << w.lane) - 1u));
    p.keys[pos] = ((unsigned long long)w.qi << 32) | j;
    if (p.dist) p.dist[pos] = sqrt(d2);
  }
  w.found += __popc(m);
}

template <int D, int L, bool FILL>
struct VisitRadius {
  static __device__ __forceinline__ void run(const TreeRadiusParams& p, RadiusQuery<D>& w, uint32_t group) {
    const uint32_t child0 = group * 32u;
    const float* bp = p.t.boxes32[L] + (uint64_t)group * (2 * D * 32) + w.lane;
    float lb = 0.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float lo = __ldg(bp + d * 32), hi = __ldg(bp + (D + d) * 32);
      const float t = fmaxf(fmaxf(__fsub_rd(lo, w.qbox[D + d]), __fsub_rd(w.qbox[d], hi)), 0.0f);
      lb = __fadd_rd(lb, __fmul_rd(t, t));
    }
    uint32_t m = __ballot_sync(0xFFFFFFFFu, child0 + (uint32_t)w.lane < p.t.count[L] && !(lb > w.Uf));
    while (m) {
      const uint32_t c = __ffs(m) - 1;
      m &= m - 1;
      if constexpr (L == 0)
        radius_leaf<D, FILL>(p, w, child0 + c);
      else
        VisitRadius<D, L - 1, FILL>::run(p, w, child0 + c);
    }
  }
};

template <int D, bool FILL>
__global__ void __launch_bounds__(TK_THREADS, OXG_TK_MIN_CTAS) tree_radius_kernel(const TreeRadiusParams p) {
  __shared__ double s_q[TK_THREADS / 32][D];
  __shared__ float s_qbox[TK_THREADS / 32][2 * D];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const double sc = p.plan->sc;
  const uint32_t n_chunks = (p.Q + p.qpc - 1) / p.qpc;
  for (uint32_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const uint32_t q_begin = chunk * p.qpc;
    const uint32_t q_end = min(p.Q, q_begin + p.qpc);
    for (uint32_t wq = q_begin + wid; wq < q_end; wq += TK_THREADS / 32) {
      const uint32_t qi = p.qorder ? __ldg(p.qorder + wq) : wq;
      RadiusQuery<D> w;
      bool qnan = false;
      double mq = 0.0;
      __syncwarp();
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double qv = p.q.p[(uint64_t)qi * p.q.stride_i + (uint64_t)d * p.q.stride_d];
        qnan = qnan || !(qv == qv);
        const double c = p.plan->c[d];
        const double qs = (qv - c) * sc;
        w.q32[d] = __double2float_rn(qs);
        mq = fmax(mq, fabs(qs));
        if (lane == 0) {
          s_q[wid][d] = qv;
          s_qbox[wid][d] = __double2float_rd(__dmul_rd(__dsub_rd(qv, c), sc));
          s_qbox[wid][D + d] = __double2float_ru(__dmul_ru(__dsub_ru(qv, c), sc));
        }
      }
      __syncwarp();
      w.qd = s_q[wid];
      w.qbox = s_qbox[wid];
      w.lane = lane;
      w.qi = qi;
      w.ub = p.ub;
      w.found = 0;
      w.limit = p.limit_self ? p.self_base + qi : (p.idx_limit ? __ldg(p.idx_limit + qi) : p.n_db);
      w.base = FILL ? p.counts[qi] : 0;
      const double Ub = __dmul_ru(__dmul_ru(p.U, __ldg(&p.plan->sc2)), 1.0 + 1.0e-12);
      w.Uf = __double2float_ru(Ub);
      w.T32 = tree_filter_threshold(Ub, fmax(mq, p.plan->mabs) * 1.000001, D);
      if (!qnan && w.limit > 0) {
        switch (p.t.top) {
          case 0: VisitRadius<D, 0, FILL>::run(p, w, 0); break;
          case 1: VisitRadius<D, 1, FILL>::run(p, w, 0); break;
          case 2: VisitRadius<D, 2, FILL>::run(p, w, 0); break;
          case 3: VisitRadius<D, 3, FILL>::run(p, w, 0); break;
          default: VisitRadius<D, 4, FILL>::run(p, w, 0); break;
        }
      }
      if (!FILL && lane == 0) p.counts[qi] = w.found;
    }
  }
}

__global__ void tree_radius_unpack_kernel(const unsigned long long* __restrict__ keys, uint64_t n,
                                          uint32_t* __restrict__ idx) {
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
       t += (uint64_t)gridDim.x * blockDim.x)
    idx[t] = (uint32_t)(keys[t] & 0xFFFFFFFFull);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static unsigned grid_1d(uint64_t n, int threads, unsigned cap = 148u * 16u) {
  return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + threads - 1) / threads, cap));
}

template <int D>
static int tree_build_levels(oxg_store_s* s, StoreTree* t, const uint32_t* sorted) {
  const uint32_t n = (uint32_t)s->size;
  TreeView& v = t->view;
  const uint32_t n_leaves = (n + 31u) / 32u;
  uint32_t cnt = n_leaves;
  int lvl = 0;
  for (;;) {
    v.count[lvl] = cnt;
    const uint32_t groups = (cnt + 31u) / 32u;
    OXG_TRY(t->boxes[lvl].ensure((size_t)groups * 2 * D * 32 * sizeof(double)));
    OXG_TRY(t->boxes32[lvl].ensure((size_t)groups * 2 * D * 32 * sizeof(float)));
    v.boxes[lvl] = t->boxes[lvl].as<double>();
    v.boxes32[lvl] = t->boxes32[lvl].as<float>();
    if (cnt <= 32u || lvl == TREE_LEVELS - 1) break;
    cnt = groups;
    ++lvl;
  }
  v.top = lvl;
  for (int l = lvl + 1; l < TREE_LEVELS; ++l) {
    v.count[l] = 0;
    v.boxes[l] = nullptr;
    v.boxes32[l] = nullptr;
  }
  OXG_TRY(t->pts.ensure((size_t)n_leaves * D * 32 * sizeof(double)));
  OXG_TRY(t->pts32.ensure((size_t)n_leaves * D * 32 * sizeof(float)));
  v.pts32 = t->pts32.as<float>();
  OXG_TRY(t->perm.ensure((size_t)n_leaves * 32 * sizeof(uint32_t)));
  v.pts = t->pts.as<double>();
  v.perm = t->perm.as<uint32_t>();
  {
    const uint32_t slots = ((n_leaves + 31u) / 32u) * 32u;
    tree_leaves_kernel<D><<<(slots * 32u + 255u) / 256u, 256, 0, s->stream>>>(
        s->d_data, s->pitch, n, sorted, n_leaves, slots, t->plan.as<TreePlan>(), t->pts.as<double>(),
        t->pts32.as<float>(), t->perm.as<uint32_t>(), t->boxes[0].as<double>());
    OXG_LAUNCH_CHECK();
  }
  for (int l = 1; l <= lvl; ++l) {
    const uint32_t nodes = v.count[l];
    const uint32_t slots = ((nodes + 31u) / 32u) * 32u;
    tree_level_kernel<D><<<(slots * 32u + 255u) / 256u, 256, 0, s->stream>>>(
        t->boxes[l - 1].as<double>(), v.count[l - 1], nodes, slots, t->boxes[l].as<double>());
    OXG_LAUNCH_CHECK();
  }
  for (int l = 0; l <= lvl; ++l) {
    const uint64_t n_vals = (uint64_t)((v.count[l] + 31u) / 32u) * 2 * D * 32;
    tree_boxes32_kernel<<<grid_1d(n_vals, 256), 256, 0, s->stream>>>(
        t->boxes[l].as<double>(), n_vals, D, t->plan.as<TreePlan>(), t->boxes32[l].as<float>());
    OXG_LAUNCH_CHECK();
  }
  return OXG_OK;
}

static int sort_pairs(oxg_store_s* s, StoreTree* t, const uint32_t* k_in, uint32_t* k_out,
                      const uint32_t* v_in, uint32_t* v_out, uint32_t n) {
  size_t bytes = 0;
  OXG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k_in, k_out, v_in, v_out, (int)n, 0,
                                               TREE_KEY_BITS, s->stream));
  OXG_TRY(t->tmp.ensure(bytes));
  OXG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(t->tmp.p, bytes, k_in, k_out, v_in, v_out, (int)n, 0,
                                               TREE_KEY_BITS, s->stream));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return OXG_OK;
}

static StoreTree* tree_of(oxg_store_s* s) {
  if (!s->tree) s->tree = new StoreTree();
  return static_cast<StoreTree*>(s->tree);
}

void tree_release(oxg_store_s* s) {
  if (!s->tree) return;
  StoreTree* t = static_cast<StoreTree*>(s->tree);
  t->pts.release();
  t->pts32.release();
  t->perm.release();
  t->keys.release();
  t->keys2.release();
  t->vals.release();
  t->vals2.release();
  t->plan.release();
  t->bounds.release();
  t->qorder.release();
  t->tmp.release();
  for (auto& b : t->boxes) b.release();
  for (auto& b : t->boxes32) b.release();
  t->kd_tables.release();
  t->kd_cell_a.release();
  t->kd_cell_b.release();
  t->kd_dim.release();
  t->kd_keys_a.release();
  t->kd_keys_b.release();
  delete t->kd_host;
  if (t->ev_b0) cudaEventDestroy(t->ev_b0);
  if (t->ev_b1) cudaEventDestroy(t->ev_b1);
  if (t->ev_q0) cudaEventDestroy(t->ev_q0);
  if (t->ev_q1) cudaEventDestroy(t->ev_q1);
  delete t;
  s->tree = nullptr;
}

bool tree_supported(const oxg_store_s* s, uint32_t k) {
  return is_pure_rv(s->space) && s->dim >= 1 && s->dim <= (uint32_t)TREE_MAX_DIM && k >= 1 && k <= 32 &&
         s->size >= 1 && s->size <= (1ull << 30);
}

bool tree_is_current(const oxg_store_s* s) {
  const StoreTree* t = static_cast<const StoreTree*>(s->tree);
  return t && g_tune_tree_cache.load() != 0 && t->epoch == s->epoch && t->size == s->size;
}

// 0 = off, 1 = time the build and the query kernel (events only), 2 = also count visits (the
// counters cost the query kernel ~15 %)
static int tree_stats_level() {
  const int v = g_tune_tree_stats.load();
  if (v > 0) return v;
  static const int env_level = [] {
    const char* e = getenv("OXG_TREE_STATS");
    return (e && *e) ? atoi(e) : 0;
  }();
  return env_level;
}
static bool tree_stats_on() { return tree_stats_level() > 0; }

static int tree_events(StoreTree* t) {
  if (t->ev_b0) return OXG_OK;
  OXG_CUDA_TRY(cudaEventCreate(&t->ev_b0));
  OXG_CUDA_TRY(cudaEventCreate(&t->ev_b1));
  OXG_CUDA_TRY(cudaEventCreate(&t->ev_q0));
  OXG_CUDA_TRY(cudaEventCreate(&t->ev_q1));
  return OXG_OK;
}

#define OXG_TREE_DISPATCH_D(D_RT, CALL)              \
  switch (D_RT) {                                    \
    case 1: { constexpr int DD = 1; CALL; } break;   \
    case 2: { constexpr int DD = 2; CALL; } break;   \
    case 3: { constexpr int DD = 3; CALL; } break;   \
    case 4: { constexpr int DD = 4; CALL; } break;   \
    case 5: { constexpr int DD = 5; CALL; } break;   \
    case 6: { constexpr int DD = 6; CALL; } break;   \
    case 7: { constexpr int DD = 7; CALL; } break;   \
    case 8: { constexpr int DD = 8; CALL; } break;   \
    default: break;                                  \
  }

// The balanced partition order of the store's states -> `order` (device, n entries).
static int kd_order(oxg_store_s* s, StoreTree* t, const unsigned long long* bounds, uint32_t** order_out) {
  const uint32_t n = (uint32_t)s->size;
  const int dim = (int)s->dim;
  if (!t->kd_host) t->kd_host = new KdTables();
  KdTables& T = *t->kd_host;
  if (T.n != n || !T.uploaded) {
    // one upload of every round's tables (kept until the store's size changes)
    kd_make_tables(n, &T);
    size_t bytes = 0;
    T.max_seg = 1;
    for (const auto& rd : T.rounds) {
      bytes += (rd.begin.size() + rd.c_begin.size() + rd.c_parent.size()) * sizeof(uint32_t);
      bytes += ((rd.parts.size() + 3) & ~size_t(3)) + ((rd.c_part.size() + 3) & ~size_t(3));
      T.max_seg = std::max<uint32_t>(T.max_seg, (uint32_t)rd.c_parent.size());
    }
    std::vector<unsigned char> blob(bytes);
    OXG_TRY(t->kd_tables.ensure(bytes + 16));
    unsigned char* dbase = t->kd_tables.as<unsigned char>();
    T.dev_rounds.clear();
    size_t off = 0;
    auto put = [&](const void* src, size_t nbytes) {
      memcpy(blob.data() + off, src, nbytes);
      const size_t at = off;
      off += (nbytes + 3) & ~size_t(3);
      return at;
    };
    for (const auto& rd : T.rounds) {
      KdRound r;
      r.n_seg = (uint32_t)rd.parts.size();
      r.n_child = (uint32_t)rd.c_parent.size();
      r.begin = reinterpret_cast<const uint32_t*>(dbase + put(rd.begin.data(), rd.begin.size() * 4));
      r.c_begin = reinterpret_cast<const uint32_t*>(dbase + put(rd.c_begin.data(), rd.c_begin.size() * 4));
      r.c_parent = reinterpret_cast<const uint32_t*>(dbase + put(rd.c_parent.data(), rd.c_parent.size() * 4));
      r.parts = dbase + put(rd.parts.data(), rd.parts.size());
      r.c_part = dbase + put(rd.c_part.data(), rd.c_part.size());
      T.dev_rounds.push_back(r);
    }
    if (bytes) {
      OXG_CUDA_TRY(cudaMemcpyAsync(dbase, blob.data(), bytes, cudaMemcpyHostToDevice, s->stream));
      OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));  // `blob` is pageable host memory
    }
    T.uploaded = true;
  }
  const uint32_t max_seg = T.max_seg;
  const std::vector<KdRound>& dev_rounds = T.dev_rounds;
  OXG_TRY(t->kd_cell_a.ensure((size_t)max_seg * 2 * dim * sizeof(double)));
  OXG_TRY(t->kd_cell_b.ensure((size_t)max_seg * 2 * dim * sizeof(double)));
  OXG_TRY(t->kd_dim.ensure(max_seg));
  OXG_TRY(t->kd_keys_a.ensure((size_t)n * sizeof(uint32_t)));
  OXG_TRY(t->kd_keys_b.ensure((size_t)n * sizeof(uint32_t)));
  double* cell = t->kd_cell_a.as<double>();
  double* cell_next = t->kd_cell_b.as<double>();
  uint32_t* ord = t->vals.as<uint32_t>();
  uint32_t* ord_next = t->vals2.as<uint32_t>();
  uint32_t* keys = t->kd_keys_a.as<uint32_t>();
  uint32_t* keys2 = t->kd_keys_b.as<uint32_t>();
  kd_iota_kernel<<<grid_1d(n, 256), 256, 0, s->stream>>>(ord, n);
  OXG_LAUNCH_CHECK();
  kd_root_cell_kernel<<<1, 32, 0, s->stream>>>(bounds, dim, cell);
  OXG_LAUNCH_CHECK();
  for (const KdRound& r : dev_rounds) {
    kd_choose_dim_kernel<<<grid_1d(r.n_seg, 128), 128, 0, s->stream>>>(r, dim, cell, t->kd_dim.as<uint8_t>());
    OXG_LAUNCH_CHECK();
    int seg_bits = 0;
    while ((1u << seg_bits) < r.n_seg && seg_bits < 31) ++seg_bits;
    // the segment id takes seg_bits, the coordinate image at least 12 bits (the split only has to
    // put equal COUNTS on each side, ties inside a 1/4096 slice of the cell are harmless) and no
    // more than fills whole 8-bit sort passes: 2 passes for the first rounds, 3 for the middle
    // ones, instead of 4 for every round
#ifndef OXG_KD_MIN_SUB
#define OXG_KD_MIN_SUB 12
#endif
    const int key_bits = std::min(32, std::max(16, (seg_bits + OXG_KD_MIN_SUB + 7) / 8 * 8));
    const int sub_bits = std::max(12, key_bits - seg_bits);
    if (seg_bits + sub_bits > 32) {
      set_error("too many segments for the 32-bit partition keys");
      return OXG_E_INTERNAL;
    }
    kd_keys_kernel<<<grid_1d(n, 256), 256, 0, s->stream>>>(r, dim, s->d_data, s->pitch, n, ord, cell,
                                                          t->kd_dim.as<uint8_t>(), sub_bits, keys);
    OXG_LAUNCH_CHECK();
    size_t tmp_bytes = 0;
    OXG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, ord, ord_next, (int)n, 0,
                                                 seg_bits + sub_bits, s->stream));
    OXG_TRY(t->tmp.ensure(tmp_bytes));
    OXG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(t->tmp.p, tmp_bytes, keys, keys2, ord, ord_next, (int)n, 0,
                                                 seg_bits + sub_bits, s->stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    std::swap(ord, ord_next);
    kd_cells_kernel<<<grid_1d(r.n_child, 128), 128, 0, s->stream>>>(r, dim, s->d_data, s->pitch, n, ord, cell,
                                                                   t->kd_dim.as<uint8_t>(), cell_next);
    OXG_LAUNCH_CHECK();
    std::swap(cell, cell_next);
  }
  *order_out = ord;
  return OXG_OK;
}

// (Re)builds the hierarchy of the store's current contents on the store's stream.  want_kd: the
// balanced partition order (a few ms for 1e6 states, about half the leaf visits per query);
// otherwise the space-filling key order (a fraction of a ms).  A current hierarchy of the other
// kind is reused when the cheaper one was asked for.
int tree_build(oxg_store_s* s, bool want_kd) {
  StoreTree* t = tree_of(s);
  if (tree_is_current(s) && (t->kind == 1 || !want_kd)) return OXG_OK;
  const uint32_t n = (uint32_t)s->size;
  OXG_TRY(t->bounds.ensure((2 * kMaxFastDim + 8) * sizeof(unsigned long long)));
  OXG_TRY(t->plan.ensure(sizeof(TreePlan)));
  OXG_TRY(t->keys.ensure((size_t)n * sizeof(uint32_t)));
  OXG_TRY(t->keys2.ensure((size_t)n * sizeof(uint32_t)));
  OXG_TRY(t->vals.ensure((size_t)n * sizeof(uint32_t)));
  OXG_TRY(t->vals2.ensure((size_t)n * sizeof(uint32_t)));
  unsigned long long* bounds = t->bounds.as<unsigned long long>();
  const bool stats = tree_stats_on();
  if (stats) {
    OXG_TRY(tree_events(t));
    OXG_CUDA_TRY(cudaEventRecord(t->ev_b0, s->stream));
  }
  tree_bounds_init_kernel<<<1, 32, 0, s->stream>>>(bounds);
  OXG_LAUNCH_CHECK();
  tree_bounds_kernel<<<grid_1d(n, 256, 148u * 4u), 256, 0, s->stream>>>(s->d_data, s->pitch, n,
                                                                       (int)s->dim, bounds);
  OXG_LAUNCH_CHECK();
  tree_plan_kernel<<<1, 32, 0, s->stream>>>(bounds, (int)s->dim, t->plan.as<TreePlan>());
  OXG_LAUNCH_CHECK();
  if (!want_kd) {
    tree_keys_kernel<<<grid_1d(n, 256), 256, 0, s->stream>>>(store_view(s), n, (int)s->dim,
                                                             t->plan.as<TreePlan>(), t->keys.as<uint32_t>(),
                                                             t->vals.as<uint32_t>());
    OXG_LAUNCH_CHECK();
  }
  uint32_t* order = t->vals2.as<uint32_t>();
  if (want_kd) {
    OXG_TRY(kd_order(s, t, bounds, &order));
  } else {
    OXG_TRY(sort_pairs(s, t, t->keys.as<uint32_t>(), t->keys2.as<uint32_t>(), t->vals.as<uint32_t>(),
                       t->vals2.as<uint32_t>(), n));
  }
  t->kind = want_kd ? 1 : 0;
  int r = OXG_E_INTERNAL;
  OXG_TREE_DISPATCH_D(s->dim, r = (tree_build_levels<DD>(s, t, order)));
  OXG_TRY(r);
  if (stats) {
    OXG_CUDA_TRY(cudaEventRecord(t->ev_b1, s->stream));
    t->have_build_ev = true;
  }
  t->epoch = s->epoch;
  t->size = s->size;
  return OXG_OK;
}

template <int D>
static int launch_tree_knn(TreeKnnParams p, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // A query is one warp's sequential walk (~100 dependent steps): small batches get one query per
  // warp so that they spread over the machine; large ones TK_QPC neighbouring queries per CTA so
  // that its warps share leaves through L1.
  const unsigned warps = TK_THREADS / 32;
  // ... and never fewer than ~16 waves of CTAs, so the last wave is short (measured on a 125 000-query
  // shard of the 1e6-node build: 2.14 ms against 2.30 ms with 32 queries per CTA)
  const unsigned waves = 16;
  const unsigned slots = (unsigned)sms * OXG_TK_MIN_CTAS * waves;
  unsigned qpc = (unsigned)((p.Q + slots - 1) / slots);
  qpc = ((qpc + warps - 1) / warps) * warps;
  p.qpc = std::max(warps, std::min<unsigned>(qpc, TK_QPC));
  const unsigned chunks = (p.Q + p.qpc - 1) / p.qpc;
  const unsigned grid = OXG_TK_PERSISTENT ? std::max(1u, std::min(chunks, (unsigned)sms * OXG_TK_MIN_CTAS)) : chunks;
  if (p.stats)
    tree_knn_kernel<D, true><<<grid, TK_THREADS, 0, st>>>(p);
  else
    tree_knn_kernel<D, false><<<grid, TK_THREADS, 0, st>>>(p);
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

// k-NN of Q device-resident queries through the hierarchy (same contract as nn_knn_device).
int tree_knn_device(oxg_store_s* s, StateView q, uint64_t Q, uint32_t k, int exclude_mode,
                    const uint32_t* d_exclude, uint64_t self_base, uint32_t* d_idx, double* d_dist,
                    unsigned long long* d_stats) {
  int kind = g_tune_tree_kind.load();
  // the balanced partition costs ~0.9 us per 1000 states more to build and saves ~45 ns per query
  if (kind < 0) kind = (Q >= 8192 && Q >= s->size / 64 && s->size >= 65536) ? 1 : 0;
  if (s->size > (1ull << 25)) kind = 0;  // 32-bit partition keys hold at most 2^20 leaf segments
  OXG_TRY(tree_build(s, kind == 1));
  StoreTree* t = tree_of(s);
  // queries in key order: neighbouring warps walk the same part of the hierarchy
  OXG_TRY(t->qorder.ensure((size_t)Q * 4 * sizeof(uint32_t)));
  uint32_t* qk = t->qorder.as<uint32_t>();
  uint32_t* qk2 = qk + Q;
  uint32_t* qv = qk + 2 * Q;
  uint32_t* qv2 = qk + 3 * Q;
  tree_keys_kernel<<<grid_1d(Q, 256), 256, 0, s->stream>>>(q, (uint32_t)Q, (int)s->dim,
                                                           t->plan.as<TreePlan>(), qk, qv);
  OXG_LAUNCH_CHECK();
  OXG_TRY(sort_pairs(s, t, qk, qk2, qv, qv2, (uint32_t)Q));
  TreeKnnParams p;
  p.t = t->view;
  p.plan = t->plan.as<TreePlan>();
  p.q = q;
  p.Q = (uint32_t)Q;
  p.qorder = qv2;
  p.k = k;
  p.exclude_mode = exclude_mode;
  p.exclude = d_exclude;
  p.self_base = (uint32_t)self_base;
  p.out_idx = d_idx;
  p.out_dist = d_dist;
  p.stats = d_stats;
  // counters + kernel time for oxg_store_tree_stats (measurement runs only: the counting costs
  // a few instructions per visited leaf)
  const bool stats = tree_stats_on();
  if (stats && !d_stats) {
    OXG_TRY(tree_events(t));
    if (tree_stats_level() >= 2) {
      p.stats = t->bounds.as<unsigned long long>() + 2 * kMaxFastDim;
      OXG_CUDA_TRY(cudaMemsetAsync(p.stats, 0, 8 * sizeof(unsigned long long), s->stream));
    }
    OXG_CUDA_TRY(cudaEventRecord(t->ev_q0, s->stream));
  }
  int r = OXG_E_INTERNAL;
  OXG_TREE_DISPATCH_D(s->dim, r = (launch_tree_knn<DD>(p, s->stream)));
  if (stats && !d_stats && r == OXG_OK) {
    OXG_CUDA_TRY(cudaEventRecord(t->ev_q1, s->stream));
    t->have_query_ev = true;
    t->have_counts = p.stats != nullptr;
    t->last_Q = Q;
  }
  return r;
}

template <int D>
static int launch_tree_radius(TreeRadiusParams p, bool fill, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned warps = TK_THREADS / 32;
  unsigned qpc = (unsigned)((p.Q + (unsigned)sms * OXG_TK_MIN_CTAS - 1) / ((unsigned)sms * OXG_TK_MIN_CTAS));
  qpc = ((qpc + warps - 1) / warps) * warps;
  p.qpc = std::max(warps, std::min<unsigned>(qpc, TK_QPC));
  const unsigned grid = (p.Q + p.qpc - 1) / p.qpc;
  if (fill)
    tree_radius_kernel<D, true><<<grid, TK_THREADS, 0, st>>>(p);
  else
    tree_radius_kernel<D, false><<<grid, TK_THREADS, 0, st>>>(p);
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

// Radius search of Q device-resident queries through the hierarchy (same contract as
// nn_radius_device with one slice per query): exclusive offsets (Q + 1) in B_OFFSETS, ascending
// indices in B_PIDX, distances in B_PDIST.  `U` = max{x : sqrt(x) < r} (negative: no hits).
int tree_radius_device(oxg_store_s* s, StateView q, uint64_t Q, double U, const uint32_t* d_idx_limit,
                       bool limit_self, uint64_t self_base, bool want_dist, uint64_t* total) {
  *total = 0;
  int kind = g_tune_tree_kind.load();
  if (kind < 0) kind = (Q >= 8192 && Q >= s->size / 64 && s->size >= 65536) ? 1 : 0;
  if (s->size > (1ull << 25)) kind = 0;
  OXG_TRY(tree_build(s, kind == 1));
  StoreTree* t = tree_of(s);
  OXG_TRY(s->buf[B_OFFSETS].ensure((Q + 1) * sizeof(uint64_t)));
  OXG_TRY(s->buf[B_COUNTS].ensure((Q + 1) * sizeof(uint64_t)));
  uint64_t* counts = s->buf[B_COUNTS].as<uint64_t>();
  uint64_t* offsets = s->buf[B_OFFSETS].as<uint64_t>();
  OXG_CUDA_TRY(cudaMemsetAsync(counts, 0, (Q + 1) * sizeof(uint64_t), s->stream));
  if (!(U >= 0.0)) {  // r <= 0 or NaN: nothing is within the radius
    OXG_CUDA_TRY(cudaMemsetAsync(offsets, 0, (Q + 1) * sizeof(uint64_t), s->stream));
    OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    return OXG_OK;
  }
  OXG_TRY(t->qorder.ensure((size_t)Q * 4 * sizeof(uint32_t)));
  uint32_t* qk = t->qorder.as<uint32_t>();
  uint32_t* qk2 = qk + Q;
  uint32_t* qv = qk + 2 * Q;
  uint32_t* qv2 = qk + 3 * Q;
  tree_keys_kernel<<<grid_1d(Q, 256), 256, 0, s->stream>>>(q, (uint32_t)Q, (int)s->dim,
                                                           t->plan.as<TreePlan>(), qk, qv);
  OXG_LAUNCH_CHECK();
  OXG_TRY(sort_pairs(s, t, qk, qk2, qv, qv2, (uint32_t)Q));
  TreeRadiusParams p;
  memset(&p, 0, sizeof(p));
  p.t = t->view;
  p.plan = t->plan.as<TreePlan>();
  p.q = q;
  p.Q = (uint32_t)Q;
  p.qorder = qv2;
  p.U = U;
  memcpy(&p.ub, &U, sizeof(double));
  p.idx_limit = d_idx_limit;
  p.limit_self = limit_self ? 1 : 0;
  p.self_base = (uint32_t)self_base;
  p.n_db = (uint32_t)s->size;
  p.counts = counts;
  int r = OXG_E_INTERNAL;
  OXG_TREE_DISPATCH_D(s->dim, r = (launch_tree_radius<DD>(p, false, s->stream)));
  OXG_TRY(r);
  size_t tmp_bytes = 0;
  OXG_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, offsets, (int)(Q + 1), s->stream));
  OXG_TRY(t->tmp.ensure(tmp_bytes));
  OXG_CUDA_TRY(cub::DeviceScan::ExclusiveSum(t->tmp.p, tmp_bytes, counts, offsets, (int)(Q + 1), s->stream));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  uint64_t tot = 0;
  OXG_CUDA_TRY(cudaMemcpyAsync(&tot, offsets + Q, sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  *total = tot;
  if (tot == 0) return OXG_OK;
  if (tot > 0x7FFFFFF0ull) {
    set_error("radius result of %llu entries is too large for one batch", (unsigned long long)tot);
    return OXG_E_ARG;
  }
  OXG_TRY(s->buf[B_PIDX].ensure(tot * sizeof(uint32_t)));
  OXG_TRY(t->kd_keys_a.ensure(tot * sizeof(unsigned long long)));
  OXG_TRY(t->kd_keys_b.ensure(tot * sizeof(unsigned long long)));
  unsigned long long* keys = t->kd_keys_a.as<unsigned long long>();
  unsigned long long* keys2 = t->kd_keys_b.as<unsigned long long>();
  double* dist_tmp = nullptr;
  if (want_dist) {
    OXG_TRY(s->buf[B_PDIST].ensure(tot * sizeof(double)));
    OXG_TRY(t->keys.ensure(tot * sizeof(double)));
    dist_tmp = t->keys.as<double>();
  }
  p.counts = offsets;
  p.keys = keys;
  p.dist = dist_tmp;
  OXG_TREE_DISPATCH_D(s->dim, r = (launch_tree_radius<DD>(p, true, s->stream)));
  OXG_TRY(r);
  // rows in ascending index order: one sort of (query << 32 | index); the query part already
  // places every key inside its row's slot range
  int qbits = 1;
  while ((1ull << qbits) < Q && qbits < 32) ++qbits;
  tmp_bytes = 0;
  if (want_dist) {
    OXG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, dist_tmp,
                                                 s->buf[B_PDIST].as<double>(), (int)tot, 0, 32 + qbits, s->stream));
    OXG_TRY(t->tmp.ensure(tmp_bytes));
    OXG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(t->tmp.p, tmp_bytes, keys, keys2, dist_tmp,
                                                 s->buf[B_PDIST].as<double>(), (int)tot, 0, 32 + qbits, s->stream));
  } else {
    OXG_CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, keys2, (int)tot, 0, 32 + qbits, s->stream));
    OXG_TRY(t->tmp.ensure(tmp_bytes));
    OXG_CUDA_TRY(cub::DeviceRadixSort::SortKeys(t->tmp.p, tmp_bytes, keys, keys2, (int)tot, 0, 32 + qbits, s->stream));
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  tree_radius_unpack_kernel<<<grid_1d(tot, 256), 256, 0, s->stream>>>(keys2, tot, s->buf[B_PIDX].as<uint32_t>());
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

}  // namespace oxg

extern "C" int oxg_store_tree_stats(oxg_store s, double* out8) {
  using namespace oxg;
  OXG_API_BEGIN
  if (!s || !out8) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  for (int i = 0; i < 8; ++i) out8[i] = 0.0;
  StoreTree* t = static_cast<StoreTree*>(s->tree);
  if (!t || !t->have_query_ev) {
    set_error("no k-NN has run through the hierarchy with \"tree_stats\" switched on");
    return OXG_E_ARG;
  }
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  float ms = 0.f;
  if (t->have_build_ev && cudaEventElapsedTime(&ms, t->ev_b0, t->ev_b1) == cudaSuccess) out8[0] = ms;
  OXG_CUDA_TRY(cudaEventElapsedTime(&ms, t->ev_q0, t->ev_q1));
  out8[1] = ms;
  unsigned long long h[5] = {0, 0, 0, 0, 0};
  if (t->have_counts)
    OXG_CUDA_TRY(cudaMemcpy(h, t->bounds.as<unsigned long long>() + 2 * kMaxFastDim, sizeof(h),
                          cudaMemcpyDeviceToHost));
  for (int i = 0; i < 5; ++i) out8[2 + i] = (double)h[i];
  out8[7] = (double)t->last_Q;
  return OXG_OK;
  OXG_API_END
}
