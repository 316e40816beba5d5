// oxg_internal.cuh — shared declarations for liboxg (not part of the public ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/oxg.h"

namespace oxg {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define OXG_CUDA_TRY(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::oxg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                       \
      return OXG_E_CUDA;                                                                \
    }                                                                                   \
  } while (0)

#define OXG_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != OXG_OK) return _r; \
  } while (0)

// Checks the launch and counts it (bench.py reports the count as gpu_launches).
#define OXG_LAUNCH_CHECK()                                                                     \
  do {                                                                                         \
    ::oxg::g_launches.fetch_add(1, std::memory_order_relaxed);                                 \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) {                                                                   \
      ::oxg::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__,   \
                       __LINE__);                                                              \
      return OXG_E_CUDA;                                                                       \
    }                                                                                          \
  } while (0)

#define OXG_API_BEGIN try {
#define OXG_API_END                                         \
  }                                                         \
  catch (const std::bad_alloc&) {                           \
    ::oxg::set_error("host allocation failed");             \
    return OXG_E_ALLOC;                                     \
  }                                                         \
  catch (const std::exception& e) {                         \
    ::oxg::set_error("internal error: %s", e.what());       \
    return OXG_E_INTERNAL;                                  \
  }                                                         \
  catch (...) {                                             \
    ::oxg::set_error("internal error: unknown exception");  \
    return OXG_E_INTERNAL;                                  \
  }

// ---------------------------------------------------------------- device buffers
// A growable device allocation (never shrinks); contents are not preserved by ensure().
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return OXG_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
      p = nullptr;
      return OXG_E_CUDA;
    }
    cap = want;
    return OXG_OK;
  }
  // Grows like ensure() but keeps the first `keep` bytes (ordered on `st`).
  int ensure_keep(size_t bytes, size_t keep, cudaStream_t st) {
    if (bytes <= cap) return OXG_OK;
    void* np = nullptr;
    size_t want = bytes + bytes / 2 + 256;
    cudaError_t e = cudaMalloc(&np, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
      return OXG_E_CUDA;
    }
    if (p && keep) {
      e = cudaMemcpyAsync(np, p, keep < cap ? keep : cap, cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) {
        cudaFree(np);
        set_error("device copy failed: %s", cudaGetErrorString(e));
        return OXG_E_CUDA;
      }
    }
    if (p) cudaFree(p);
    p = np;
    cap = want;
    return OXG_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

// Named device scratch arenas of a store (each grows on demand, never shrinks).
enum StoreBuf {
  B_STAGE = 0,  // row-major upload / download staging for append & read
  B_QUERY,      // uploaded query rows / edge endpoints
  B_TMP,        // cub temporary storage
  B_PIDX,       // k-NN partial lists (idx) / radius hit indices
  B_PDIST,      // k-NN partial lists (dist) / radius hit distances
  B_COUNTS,     // radius per-(query, slice) counts
  B_OFFSETS,    // radius scanned offsets
  B_OIDX,       // final k-NN indices for host download
  B_ODIST,      // final k-NN distances for host download
  B_AUX,        // exclusion lists / index limits / per-query CSR offsets
  B_AUX2,       // edge lists, flags
  B_AUX3,
  B_PARENT,     // planner trees: parent index of every node (i64, -1 = root)
  B_COST,       // RRT*: cost-to-come of every node (f64)
  B_PLAN,       // planner kernels: control block, barrier scratch, per-CTA neighbour lists
  B_EROW,       // PRM edge phase: owner row of every candidate edge
  B_ROWCNT,     // PRM edge phase: per-row count of obstacle pairs that survive the row cull
  B_ROWLIST,    // PRM edge phase: those pairs
  B_COUNT
};

constexpr int kMaxDim = 32;     // generic (runtime-dimension) kernels keep states in local arrays
constexpr int kMaxFastDim = 16; // RealVector fast paths are instantiated for D = 1..16
constexpr uint32_t kMaxK = 64;

}  // namespace oxg

// ---------------------------------------------------------------- handles
struct oxg_store_s {
  oxg_space_desc space;
  uint32_t dim = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  double* d_data = nullptr;  // SoA: coordinate d of state i at d_data[d * pitch + i]
  float* d_data32 = nullptr; // the same states rounded to f32 (round-to-nearest), same layout,
                             // plus column `dim` = squared norm of each state: input of the
                             // certified FP32 pre-filter, never of a result
  unsigned long long* d_maxabs = nullptr;  // [0] bits of max |coordinate|, [1] bits of max
                                           // squared norm over the store (NaN sorts high)
  uint64_t pitch = 0;        // elements; multiple of 32 (256-byte aligned columns)
  uint64_t size = 0;
  uint64_t n_planned = 0;    // nodes whose B_PARENT / B_COST entries are valid
  oxg::DevBuf buf[oxg::B_COUNT];  // named scratch arenas, see enum StoreBuf
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  uint64_t epoch = 0;        // bumped whenever the contents change (append, clear, planner growth)
  uint32_t last_radius_nsplit = 1;  // slices per query of the most recent nn_radius_device call
  void* tree = nullptr;      // oxg::StoreTree, the bounding-box hierarchy of oxg_tree.cu (lazy)
};

struct oxg_world_s {
  uint32_t dim = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  // Array-of-structs per primitive so a warp-uniform obstacle read is one contiguous burst.
  //   sphere : [c_0..c_{dim-1}, U]                stride sphere_stride (even)
  //   aabb   : [lo_0..lo_{dim-1}, hi_0..hi_{dim-1}] stride 2*dim
  //   capsule: [a_0.., ab_0.., den, U]             stride capsule_stride (even)
  // U = max{x : sqrt(x) <= r}: "sqrt(d2) > r" is evaluated as "d2 > U" (exactly equivalent).
  double* d_spheres = nullptr;
  double* d_aabbs = nullptr;
  double* d_capsules = nullptr;
  // f32 mirrors for the certified pre-filter of the validity kernels (never decide a result):
  //   sphere : [c_0.., U1, sq]            U1 >= U(1+1e-4), sq >= sqrt(dim*U1)     stride sphere32_stride
  //   aabb   : [lo_0.. (rounded down), hi_0.. (rounded up)]                        stride 2*dim
  //   capsule: [a_0.., ab_0.., 1/den, U1, sq]                                      stride capsule32_stride
  float* d_spheres32 = nullptr;
  float* d_aabbs32 = nullptr;
  float* d_capsules32 = nullptr;
  float* d_capsules32p = nullptr;  // capsules in pairs, values interleaved (c, c+1): packed FP32 math
  uint32_t capsule32p_stride = 0;  // floats per pair row (multiple of 4)
  uint32_t sphere32_stride = 0, capsule32_stride = 0;
  double wmax_spheres = 0.0;   // max |centre coordinate|
  double wmax_capsules = 0.0;  // max over |a_d|, |b_d|, |ab_d|
  uint64_t n_spheres = 0, n_aabbs = 0, n_capsules = 0;
  uint32_t sphere_stride = 0, aabb_stride = 0, capsule_stride = 0;
  // SO2 interval obstacles (host copy kept so add() can re-upload)
  uint32_t n_so2 = 0;
  uint32_t h_so2_comp[16];
  double h_so2_lohi[32];
  // Uniform grid over the first min(dim, 3) coordinates (planner kernels): cell -> ids of the
  // obstacles whose bounding box touches it (+-1 cell of margin).  ids: spheres, then aabbs, then
  // capsules.  A conservative candidate list only — every candidate still takes the exact test.
  uint32_t* d_grid_off = nullptr;  // [cells + 1]
  uint32_t* d_grid_ids = nullptr;
  uint32_t grid_dims = 0;
  uint32_t grid_n[3] = {1, 1, 1};
  double grid_lo[3] = {0, 0, 0};
  double grid_inv[3] = {0, 0, 0};  // cells per unit length
  uint32_t* d_so2_comp = nullptr;
  double* d_so2_lohi = nullptr;
  oxg::DevBuf buf[4];
};

namespace oxg {

// View of a world passed by value to kernels.
struct WorldView {
  const double* spheres;
  const double* aabbs;
  const double* capsules;
  const float* spheres32;
  const float* aabbs32;
  const float* capsules32;
  const float* capsules32p;
  uint32_t capsule32p_stride;
  uint32_t sphere32_stride, capsule32_stride;
  double wmax_spheres, wmax_capsules;
  const uint32_t* grid_off;  // nullable
  const uint32_t* grid_ids;
  uint32_t grid_dims;
  uint32_t grid_n[3];
  double grid_lo[3], grid_inv[3];
  const uint32_t* so2_comp;
  const double* so2_lohi;
  uint32_t n_spheres, n_aabbs, n_capsules, n_so2;
  uint32_t sphere_stride, aabb_stride, capsule_stride;
  uint32_t dim;
};
inline WorldView world_view(const oxg_world_s* w) {
  WorldView v;
  v.spheres = w->d_spheres;
  v.aabbs = w->d_aabbs;
  v.capsules = w->d_capsules;
  v.spheres32 = w->d_spheres32;
  v.aabbs32 = w->d_aabbs32;
  v.capsules32 = w->d_capsules32;
  v.capsules32p = w->d_capsules32p;
  v.capsule32p_stride = w->capsule32p_stride;
  v.sphere32_stride = w->sphere32_stride;
  v.capsule32_stride = w->capsule32_stride;
  v.wmax_spheres = w->wmax_spheres;
  v.wmax_capsules = w->wmax_capsules;
  v.grid_off = w->d_grid_off;
  v.grid_ids = w->d_grid_ids;
  v.grid_dims = w->grid_dims;
  for (int i = 0; i < 3; ++i) {
    v.grid_n[i] = w->grid_n[i];
    v.grid_lo[i] = w->grid_lo[i];
    v.grid_inv[i] = w->grid_inv[i];
  }
  v.so2_comp = w->d_so2_comp;
  v.so2_lohi = w->d_so2_lohi;
  v.n_spheres = (uint32_t)w->n_spheres;
  v.n_aabbs = (uint32_t)w->n_aabbs;
  v.n_capsules = (uint32_t)w->n_capsules;
  v.n_so2 = w->n_so2;
  v.sphere_stride = w->sphere_stride;
  v.aabb_stride = w->aabb_stride;
  v.capsule_stride = w->capsule_stride;
  v.dim = w->dim;
  return v;
}

// Where a batch of states lives: coordinate d of state i at p[i * stride_i + d * stride_d].
//   row-major host upload : stride_i = dim, stride_d = 1
//   SoA store             : stride_i = 1,   stride_d = pitch
struct StateView {
  const double* p;
  uint64_t stride_i;
  uint64_t stride_d;
};
inline StateView store_view(const oxg_store_s* s) { return StateView{s->d_data, 1, s->pitch}; }
inline StateView rows_view(const double* p, uint32_t dim) { return StateView{p, dim, 1}; }

// Host-side exact thresholds (std::sqrt is correctly rounded, as is CUDA's sqrt.rn.f64).
//   sqrt_le_threshold(r) = max{x >= 0 : sqrt(x) <= r}   (-1 if none, +inf if r = +inf)
//   sqrt_lt_threshold(r) = max{x >= 0 : sqrt(x) <  r}   (-1 if none)
double sqrt_le_threshold(double r);
double sqrt_lt_threshold(double r);

int validate_space(const oxg_space_desc* sp);
inline bool is_pure_rv(const oxg_space_desc& sp) { return sp.n_so2 == 0; }

// internal entry points shared between translation units
int nn_knn_device(oxg_store_s* s, StateView q, uint64_t Q, uint32_t k, int exclude_mode,
                  const uint32_t* d_exclude, uint64_t self_base, bool stored_first,
                  uint32_t* d_idx, double* d_dist);
// E_dev (nullable): the edge count is read from device memory at kernel start (E is then
// only an upper bound used to size the grid), so a producer kernel can feed the check without
// a host round-trip.
// Row cull (PRM k-NN build): every candidate edge owned by row r lies inside the ball around
// state r whose radius is r's longest candidate edge, so the obstacles that ball cannot touch are
// ruled out once per row instead of once per edge.  e_row[e] = owner row of edge e (relative to
// the first row of the lists), row_cnt[r] = surviving capsule pairs of row r (kRowCullOverflow:
// too many, the edge takes the full pass), row_list[r * kRowCullCap ..] = their pair indices.
constexpr uint32_t kRowCullCap = 64;
constexpr uint16_t kRowCullOverflow = 0xFFFF;
struct RowCull {
  const uint32_t* e_row = nullptr;
  const uint16_t* row_cnt = nullptr;
  const uint16_t* row_list = nullptr;
};
int motion_check_device(const oxg_world_s* w, const oxg_space_desc& sp, cudaStream_t stream,
                        StateView from, StateView to, const uint32_t* d_from_idx,
                        const uint32_t* d_to_idx, uint64_t E, const uint64_t* E_dev,
                        double resolution, uint8_t* d_valid, uint32_t* d_first_invalid,
                        const RowCull* row_cull = nullptr);
// Fills the row lists for rows [row_begin, row_end) of the store.  k-NN mode: the row's candidates
// are d_idx[(row - base_row) * k ..] and the ball's radius is the longest of them; radius mode
// (d_idx == nullptr): every candidate is shorter than fixed_radius.  n_edges = the edge count the
// motion check will be launched with.  *enabled = false when this world / space / size does not
// take the row-cull path (nothing is written then).
int prm_row_cull_device(const oxg_world_s* w, const oxg_space_desc& sp, cudaStream_t stream, StateView st,
                        const uint32_t* d_idx, uint32_t k, uint64_t base_row, uint64_t row_begin,
                        uint64_t row_end, double fixed_radius, uint64_t n_edges, uint16_t* row_cnt,
                        uint16_t* row_list, bool* enabled);
int is_valid_device(const oxg_world_s* w, const oxg_space_desc& sp, cudaStream_t stream,
                    StateView st, uint64_t Q, uint8_t* d_out);
int nn_radius_device(oxg_store_s* s, StateView q, uint64_t Q, double r, const uint32_t* d_idx_limit,
                     bool limit_self, uint64_t self_base, bool want_dist, uint64_t* total);
int nn_radius_nsplit(const oxg_store_s* s, uint64_t Q);
int store_refresh_mirror(oxg_store_s* s, uint64_t first, uint64_t n);
int store_append_staged(oxg_store_s* s, uint64_t n);  // rows already in B_STAGE (device, row-major)
int nn_radius_offsets(oxg_store_s* s, uint64_t Q, uint64_t* d_out);

// oxg_tree.cu: exact k-NN through the store's bounding-box hierarchy
void tree_release(oxg_store_s* s);
bool tree_supported(const oxg_store_s* s, uint32_t k);
bool tree_is_current(const oxg_store_s* s);
int tree_build(oxg_store_s* s, bool want_kd);
int tree_knn_device(oxg_store_s* s, StateView q, uint64_t Q, uint32_t k, int exclude_mode,
                    const uint32_t* d_exclude, uint64_t self_base, uint32_t* d_idx, double* d_dist,
                    unsigned long long* d_stats);
int tree_radius_device(oxg_store_s* s, StateView q, uint64_t Q, double U, const uint32_t* d_idx_limit,
                       bool limit_self, uint64_t self_base, bool want_dist, uint64_t* total);
// runtime tuning switches (oxg_tuning_set): -1 = automatic
extern std::atomic<int> g_tune_knn_tree;
extern std::atomic<int> g_tune_tree_cache;  // 0 = rebuild the hierarchy for every batch
extern std::atomic<int> g_tune_tree_kind;   // -1 auto, 0 space-filling key order, 1 balanced partition
extern std::atomic<int> g_tune_row_cull;    // 0 = the PRM edge phase culls per edge only
extern std::atomic<int> g_tune_tree_stats;  // 1 = count visits and time the query kernel

}  // namespace oxg
