// oxg_dist.cu — the sharded PRM build behind the C ABI (SURVEY §8b "multi-GPU variant", §8e).
//
// One process per GPU.  The store and the world are replicated (every rank appended the same
// samples); rank g owns the query rows [gN/G, (g+1)N/G) and the candidate edges whose owner slot
// lies in those rows (prm_mark_edges_kernel).  Exchanges, all NCCL all-gathers of fixed-width
// per-node rows over NVLink, issued on the store's stream:
//     k-NN rows (idx + dist, one grouped call)  ->  edge phase  ->  validity rows  ->  resolve
// so every rank ends with the complete roadmap, exactly what oxg_prm_build_knn returns on one GPU
// (prm.rs:117-146 for a pre-filled store).  Radius mode (reference-faithful, prm.rs:131-138) cuts
// the rows at N sqrt(g/G) — node i scans j < i, so the work per row grows with i — and exchanges
// the ragged lower-index adjacency with one broadcast per rank inside a group call.
//
// NCCL is resolved at run time (dlopen of the libnccl.so.2 already in the process — the one
// torch.distributed loaded — else the system one): liboxg.so has no link-time dependency on it and
// loads on hosts without NCCL, where these entry points fail with OXG_E_NCCL.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "oxg_internal.cuh"

namespace {

// The slice of nccl.h this file needs (ABI-stable since NCCL 2.0).
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;  // 0 = ncclSuccess
enum { ncclUint8 = 1, ncclUint32 = 3, ncclUint64 = 5, ncclFloat64 = 8 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool ok = false;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);  // reuse the instance already in the process
    if (api.handle) break;
  }
  if (!api.handle)
    for (const char* nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
  if (!api.handle) return nullptr;
#define OXG_NCCL_SYM(field, name)                                        \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name)); \
  if (!api.field) return nullptr;
  OXG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  OXG_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  OXG_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  OXG_NCCL_SYM(AllGather, "ncclAllGather")
  OXG_NCCL_SYM(Broadcast, "ncclBroadcast")
  OXG_NCCL_SYM(GroupStart, "ncclGroupStart")
  OXG_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  OXG_NCCL_SYM(GetErrorString, "ncclGetErrorString")
  OXG_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef OXG_NCCL_SYM
  api.ok = true;
  return &api;
}

}  // namespace

struct oxg_dist_s {
  ncclComm_t comm = nullptr;
  int rank = 0;
  int world = 1;
  int device = 0;
  oxg::DevBuf idx_all, dist_all, valid_all, counters, pad, pad2, off_all, cnt_all;
  cudaStream_t copy_stream = nullptr;  // D2H of the neighbour table while the edge phase runs
  cudaEvent_t ev_knn = nullptr;
};

namespace oxg {

#define OXG_NCCL_TRY(api, expr)                                                          \
  do {                                                                                   \
    ncclResult_t _r = (expr);                                                            \
    if (_r != 0) {                                                                       \
      set_error("%s failed: %s (%s:%d)", #expr, (api)->GetErrorString(_r), __FILE__, __LINE__); \
      return OXG_E_NCCL;                                                                 \
    }                                                                                    \
  } while (0)

// rows [gN/G, (g+1)N/G): k-NN-all work is uniform per row
static void knn_row_range(uint64_t n, int rank, int world, uint64_t* rb, uint64_t* re) {
  *rb = n * (uint64_t)rank / (uint64_t)world;
  *re = n * (uint64_t)(rank + 1) / (uint64_t)world;
}

// rows [N sqrt(g/G), N sqrt((g+1)/G)): node i scans the i earlier nodes (prm.rs:131), so equal
// areas under the triangle balance the radius-mode shards
static uint64_t radius_cut(uint64_t n, int g, int world) {
  if (g <= 0) return 0;
  if (g >= world) return n;
  return std::min<uint64_t>(n, (uint64_t)std::floor((double)n * std::sqrt((double)g / (double)world)));
}

// compacts the padded all-gather layout [world][rows_max][width] into [n][width]
__global__ void unpad_rows_kernel(const unsigned char* __restrict__ pad, unsigned char* __restrict__ out,
                                  uint64_t n, int world, uint64_t rows_max, uint32_t row_bytes) {
  const uint64_t total = n * row_bytes;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t row = t / row_bytes, b = t - row * row_bytes;
    // owner of `row`: the rank g with g n / G <= row < (g+1) n / G
    int g = (int)(((row + 1) * (uint64_t)world - 1) / n);
    while (g > 0 && n * (uint64_t)g / (uint64_t)world > row) --g;
    while (g + 1 < world && n * (uint64_t)(g + 1) / (uint64_t)world <= row) ++g;
    const uint64_t rb = n * (uint64_t)g / (uint64_t)world;
    out[t] = pad[((uint64_t)g * rows_max + (row - rb)) * row_bytes + b];
  }
}

// All-gather of the row shards of several [n][row_bytes] device arrays, in place: every rank
// holds its own rows on entry and all rows on return.  The NCCL calls of all arrays go into one
// group (one launch, one rendezvous of the ranks).  Uneven shards travel through a padded
// [world][rows_max][row_bytes] staging area and are compacted afterwards.
struct GatherItem {
  void* all;
  uint32_t row_bytes;
  DevBuf* pad;
};

static int all_gather_rows(NcclApi* api, oxg_dist_s* d, cudaStream_t st, uint64_t n, GatherItem* items,
                           int n_items) {
  if (d->world == 1) return OXG_OK;
  uint64_t rb, re;
  knn_row_range(n, d->rank, d->world, &rb, &re);
  const bool even = n % (uint64_t)d->world == 0;
  uint64_t rows_max = 0;
  for (int g = 0; g < d->world; ++g) {
    uint64_t a, b;
    knn_row_range(n, g, d->world, &a, &b);
    rows_max = std::max(rows_max, b - a);
  }
  if (!even)
    for (int i = 0; i < n_items; ++i) {
      const uint32_t rbts = items[i].row_bytes;
      OXG_TRY(items[i].pad->ensure((size_t)d->world * rows_max * rbts));
      unsigned char* p = items[i].pad->as<unsigned char>();
      OXG_CUDA_TRY(cudaMemcpyAsync(p + (uint64_t)d->rank * rows_max * rbts,
                                   static_cast<const unsigned char*>(items[i].all) + rb * rbts,
                                   (re - rb) * rbts, cudaMemcpyDeviceToDevice, st));
    }
  if (n_items > 1) OXG_NCCL_TRY(api, api->GroupStart());
  for (int i = 0; i < n_items; ++i) {
    const uint32_t rbts = items[i].row_bytes;
    if (even) {
      OXG_NCCL_TRY(api, api->AllGather(static_cast<const unsigned char*>(items[i].all) + rb * rbts,
                                       items[i].all, (re - rb) * rbts, ncclUint8, d->comm, st));
    } else {
      unsigned char* p = items[i].pad->as<unsigned char>();
      OXG_NCCL_TRY(api, api->AllGather(p + (uint64_t)d->rank * rows_max * rbts, p, rows_max * rbts,
                                       ncclUint8, d->comm, st));
    }
  }
  if (n_items > 1) OXG_NCCL_TRY(api, api->GroupEnd());
  if (!even)
    for (int i = 0; i < n_items; ++i) {
      unpad_rows_kernel<<<148 * 8, 256, 0, st>>>(items[i].pad->as<unsigned char>(),
                                                 static_cast<unsigned char*>(items[i].all), n, d->world,
                                                 rows_max, items[i].row_bytes);
      OXG_LAUNCH_CHECK();
    }
  return OXG_OK;
}

// The device-resident sharded k-NN build: fills d_idx_all / d_dist_all / d_valid_all ([n][k]) on
// every rank; d_counters[0..1] = motion checks executed / valid on THIS rank.
static int prm_build_knn_dist_dev(oxg_dist_s* d, oxg_world_s* w, oxg_store_s* s, uint32_t k,
                                  double resolution, uint32_t* d_idx_all, double* d_dist_all,
                                  uint8_t* d_valid_all, uint64_t* d_counters,
                                  cudaEvent_t after_neighbours = nullptr) {
  NcclApi* api = nccl_api();
  if (d->world > 1 && !api) {
    set_error("NCCL is not available in this process (libnccl.so.2 not found)");
    return OXG_E_NCCL;
  }
  const uint64_t n = s->size;
  uint64_t rb, re;
  knn_row_range(n, d->rank, d->world, &rb, &re);
  OXG_TRY(oxg_knn_rows_dev(s, rb, re, k, d_idx_all + rb * k, d_dist_all + rb * k));
  if (d->world > 1) {
    GatherItem items[2] = {{d_idx_all, k * (uint32_t)sizeof(uint32_t), &d->pad},
                           {d_dist_all, k * (uint32_t)sizeof(double), &d->pad2}};
    OXG_TRY(all_gather_rows(api, d, s->stream, n, items, 2));
  }
  if (after_neighbours) OXG_CUDA_TRY(cudaEventRecord(after_neighbours, s->stream));  // idx / dist are final
  OXG_TRY(oxg_prm_edges_dev(w, s, d_idx_all, k, resolution, rb, re, d_valid_all, d_counters));
  if (d->world > 1) {
    GatherItem item = {d_valid_all, k * (uint32_t)sizeof(uint8_t), &d->pad};
    OXG_TRY(all_gather_rows(api, d, s->stream, n, &item, 1));
  }
  OXG_TRY(oxg_prm_resolve_dev(s, d_idx_all, k, d_valid_all));
  return OXG_OK;
}

}  // namespace oxg

using namespace oxg;

extern "C" {

int oxg_dist_unique_id(void* out_id_128) {
  OXG_API_BEGIN
  if (!out_id_128) {
    set_error("out_id_128 is NULL");
    return OXG_E_NULL;
  }
  NcclApi* api = nccl_api();
  if (!api) {
    set_error("NCCL is not available in this process (libnccl.so.2 not found)");
    return OXG_E_NCCL;
  }
  ncclUniqueId id;
  OXG_NCCL_TRY(api, api->GetUniqueId(&id));
  memcpy(out_id_128, id.internal, sizeof(id.internal));
  return OXG_OK;
  OXG_API_END
}

int oxg_dist_create(const void* unique_id_128, int rank, int world_size, oxg_dist* out) {
  OXG_API_BEGIN
  if (!out) {
    set_error("out is NULL");
    return OXG_E_NULL;
  }
  *out = nullptr;
  if (world_size < 1 || rank < 0 || rank >= world_size) {
    set_error("rank %d is outside a world of %d", rank, world_size);
    return OXG_E_ARG;
  }
  if (world_size > 1 && !unique_id_128) {
    set_error("unique_id_128 is NULL");
    return OXG_E_NULL;
  }
  int dev = 0;
  OXG_CUDA_TRY(cudaGetDevice(&dev));
  oxg_dist_s* d = new oxg_dist_s();
  d->rank = rank;
  d->world = world_size;
  d->device = dev;
  if (world_size > 1) {
    NcclApi* api = nccl_api();
    if (!api) {
      delete d;
      set_error("NCCL is not available in this process (libnccl.so.2 not found)");
      return OXG_E_NCCL;
    }
    ncclUniqueId id;
    memcpy(id.internal, unique_id_128, sizeof(id.internal));
    ncclResult_t r = api->CommInitRank(&d->comm, world_size, id, rank);
    if (r != 0) {
      set_error("ncclCommInitRank failed: %s", api->GetErrorString(r));
      delete d;
      return OXG_E_NCCL;
    }
  }
  *out = d;
  return OXG_OK;
  OXG_API_END
}

int oxg_dist_destroy(oxg_dist d) {
  if (!d) return OXG_OK;
  if (d->comm) {
    NcclApi* api = nccl_api();
    if (api) api->CommDestroy(d->comm);
  }
  if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
  if (d->ev_knn) cudaEventDestroy(d->ev_knn);
  d->idx_all.release();
  d->dist_all.release();
  d->valid_all.release();
  d->counters.release();
  d->pad.release();
  d->pad2.release();
  d->off_all.release();
  d->cnt_all.release();
  delete d;
  return OXG_OK;
}

// Every rank contributes its own slice of a common row set; afterwards every rank's store holds
// all rows, in rank order.  The slices cross PCIe once (each rank uploads only its own) and reach
// the other ranks over NVLink (one broadcast per rank inside a group call).
int oxg_store_append_dist(oxg_dist d, oxg_store s, const double* rows_local, uint64_t n_local,
                          uint64_t* first_index) {
  OXG_API_BEGIN
  if (!d || !s || (!rows_local && n_local > 0)) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (first_index) *first_index = s->size;
  if (d->world == 1) return oxg_store_append(s, rows_local, n_local, first_index);
  NcclApi* api = nccl_api();
  if (!api) {
    set_error("NCCL is not available in this process (libnccl.so.2 not found)");
    return OXG_E_NCCL;
  }
  const int G = d->world;
  // slice sizes of all ranks
  OXG_TRY(d->cnt_all.ensure((size_t)G * sizeof(uint64_t)));
  uint64_t* d_cnt = d->cnt_all.as<uint64_t>();
  OXG_CUDA_TRY(cudaMemcpyAsync(d_cnt + d->rank, &n_local, sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
  OXG_NCCL_TRY(api, api->AllGather(d_cnt + d->rank, d_cnt, 1, ncclUint64, d->comm, s->stream));
  std::vector<uint64_t> cnt(G);
  OXG_CUDA_TRY(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)G * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  std::vector<uint64_t> off(G + 1, 0);
  for (int g = 0; g < G; ++g) off[g + 1] = off[g] + cnt[g];
  const uint64_t total = off[G];
  if (total == 0) return OXG_OK;
  const size_t row_bytes = sizeof(double) * s->dim;
  OXG_TRY(s->buf[B_STAGE].ensure(total * row_bytes));
  unsigned char* stage = s->buf[B_STAGE].as<unsigned char>();
  if (n_local)
    OXG_CUDA_TRY(cudaMemcpyAsync(stage + off[d->rank] * row_bytes, rows_local, n_local * row_bytes,
                                 cudaMemcpyHostToDevice, s->stream));
  OXG_NCCL_TRY(api, api->GroupStart());
  for (int g = 0; g < G; ++g)
    if (cnt[g])
      OXG_NCCL_TRY(api, api->Broadcast(stage + off[g] * row_bytes, stage + off[g] * row_bytes,
                                       cnt[g] * row_bytes, ncclUint8, g, d->comm, s->stream));
  OXG_NCCL_TRY(api, api->GroupEnd());
  return store_append_staged(s, total);
  OXG_API_END
}

int oxg_dist_row_range(uint64_t n, int rank, int world_size, int radius_mode, uint64_t* row_begin,
                       uint64_t* row_end) {
  if (!row_begin || !row_end) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (world_size < 1 || rank < 0 || rank >= world_size) {
    set_error("rank %d is outside a world of %d", rank, world_size);
    return OXG_E_ARG;
  }
  if (radius_mode) {
    *row_begin = radius_cut(n, rank, world_size);
    *row_end = radius_cut(n, rank + 1, world_size);
  } else {
    knn_row_range(n, rank, world_size, row_begin, row_end);
  }
  return OXG_OK;
}

int oxg_prm_build_knn_dist_dev(oxg_dist d, oxg_world w, oxg_store s, uint32_t k, double resolution,
                               uint32_t* d_idx_all, double* d_dist_all, uint8_t* d_valid_all,
                               uint64_t* d_counters) {
  OXG_API_BEGIN
  if (!d || !w || !s || !d_idx_all || !d_dist_all || !d_valid_all || !d_counters) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (!(resolution > 0.0)) {
    set_error("motion resolution must be > 0 (got %g)", resolution);
    return OXG_E_RESOLUTION;
  }
  if (s->size == 0) return OXG_OK;
  return prm_build_knn_dist_dev(d, w, s, k, resolution, d_idx_all, d_dist_all, d_valid_all, d_counters);
  OXG_API_END
}

int oxg_prm_build_knn_dist(oxg_dist d, oxg_world w, oxg_store s, uint32_t k, double resolution,
                           uint32_t* idx, double* dist, uint8_t* valid, oxg_prm_stats* stats) {
  OXG_API_BEGIN
  if (!d || !w || !s) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (!(resolution > 0.0)) {
    set_error("motion resolution must be > 0 (got %g)", resolution);
    return OXG_E_RESOLUTION;
  }
  if (stats) memset(stats, 0, sizeof(*stats));
  const uint64_t n = s->size;
  if (n == 0) return OXG_OK;
  const size_t slots = (size_t)n * k;
  OXG_TRY(d->idx_all.ensure(slots * sizeof(uint32_t)));
  OXG_TRY(d->dist_all.ensure(slots * sizeof(double)));
  OXG_TRY(d->valid_all.ensure(slots));
  OXG_TRY(d->counters.ensure(4 * sizeof(uint64_t)));
  uint32_t* d_idx = d->idx_all.as<uint32_t>();
  double* d_dist = d->dist_all.as<double>();
  uint8_t* d_valid = d->valid_all.as<uint8_t>();
  uint64_t* d_cnt = d->counters.as<uint64_t>();
  // the neighbour table goes to the host on a second stream while the edge phase runs
  if (!d->copy_stream) {
    OXG_CUDA_TRY(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
    OXG_CUDA_TRY(cudaEventCreateWithFlags(&d->ev_knn, cudaEventDisableTiming));
  }
  OXG_CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
  OXG_TRY(prm_build_knn_dist_dev(d, w, s, k, resolution, d_idx, d_dist, d_valid, d_cnt, d->ev_knn));
  OXG_CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
  OXG_CUDA_TRY(cudaStreamWaitEvent(d->copy_stream, d->ev_knn, 0));
  uint64_t h_cnt[2] = {0, 0};
  if (idx)
    OXG_CUDA_TRY(cudaMemcpyAsync(idx, d_idx, slots * sizeof(uint32_t), cudaMemcpyDeviceToHost, d->copy_stream));
  if (dist)
    OXG_CUDA_TRY(cudaMemcpyAsync(dist, d_dist, slots * sizeof(double), cudaMemcpyDeviceToHost, d->copy_stream));
  if (valid) OXG_CUDA_TRY(cudaMemcpyAsync(valid, d_valid, slots, cudaMemcpyDeviceToHost, s->stream));
  OXG_CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(d->copy_stream));
  if (stats) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s->ev0, s->ev1);
    uint64_t rb, re;
    knn_row_range(n, d->rank, d->world, &rb, &re);
    stats->n_rows = re - rb;
    stats->n_candidates = 0;
    stats->n_edges_checked = h_cnt[0];
    stats->n_edges_valid = h_cnt[1];
    stats->knn_ms = ms;  // the whole sharded build on this rank (k-NN + exchanges + edges)
    stats->edge_ms = 0.0;
  }
  return OXG_OK;
  OXG_API_END
}

// Radius-mode (reference-faithful) sharded build: every rank returns the complete lower-index
// adjacency of all N rows in CSR form.
int oxg_prm_build_radius_dist(oxg_dist d, oxg_world w, oxg_store s, double connection_radius,
                              double resolution, uint64_t* offsets, uint32_t* idx, uint64_t cap,
                              oxg_prm_stats* stats) {
  OXG_API_BEGIN
  if (!d || !w || !s || !offsets) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  NcclApi* api = nccl_api();
  if (d->world > 1 && !api) {
    set_error("NCCL is not available in this process (libnccl.so.2 not found)");
    return OXG_E_NCCL;
  }
  const uint64_t n = s->size;
  const int G = d->world;
  const uint64_t rb = radius_cut(n, d->rank, G), re = radius_cut(n, d->rank + 1, G);
  const uint64_t rows = re - rb;
  // own shard through the single-GPU entry point (host CSR of the rank's rows)
  std::vector<uint64_t> my_off(rows + 1, 0);
  std::vector<uint32_t> my_idx;
  oxg_prm_stats st_local;
  memset(&st_local, 0, sizeof(st_local));
  int rc = oxg_prm_build_radius(w, s, connection_radius, resolution, rb, re, my_off.data(), nullptr, 0,
                                &st_local);
  if (rc == OXG_E_CAPACITY) {
    my_idx.resize(my_off[rows]);
    rc = oxg_prm_build_radius(w, s, connection_radius, resolution, rb, re, my_off.data(), my_idx.data(),
                              my_idx.size(), &st_local);
  }
  OXG_TRY(rc);
  if (stats) *stats = st_local;
  if (G == 1) {
    memcpy(offsets, my_off.data(), (rows + 1) * sizeof(uint64_t));
    if (my_off[rows] > cap || (my_off[rows] > 0 && !idx)) {
      set_error("roadmap needs %llu adjacency entries, capacity is %llu", (unsigned long long)my_off[rows],
                (unsigned long long)cap);
      return OXG_E_CAPACITY;
    }
    if (my_off[rows]) memcpy(idx, my_idx.data(), my_off[rows] * sizeof(uint32_t));
    return OXG_OK;
  }
  // ragged exchange: per-row counts with one broadcast per rank (grouped), host prefix, then the
  // index lists with one broadcast per rank (grouped)
  OXG_TRY(d->cnt_all.ensure((n + 1) * sizeof(uint64_t)));
  uint64_t* d_cnt = d->cnt_all.as<uint64_t>();
  std::vector<uint64_t> my_cnt(rows);
  for (uint64_t i = 0; i < rows; ++i) my_cnt[i] = my_off[i + 1] - my_off[i];
  if (rows)
    OXG_CUDA_TRY(cudaMemcpyAsync(d_cnt + rb, my_cnt.data(), rows * sizeof(uint64_t), cudaMemcpyHostToDevice,
                                 s->stream));
  OXG_NCCL_TRY(api, api->GroupStart());
  for (int g = 0; g < G; ++g) {
    const uint64_t a = radius_cut(n, g, G), b = radius_cut(n, g + 1, G);
    if (b > a) OXG_NCCL_TRY(api, api->Broadcast(d_cnt + a, d_cnt + a, b - a, ncclUint64, g, d->comm, s->stream));
  }
  OXG_NCCL_TRY(api, api->GroupEnd());
  std::vector<uint64_t> cnt(n);
  OXG_CUDA_TRY(cudaMemcpyAsync(cnt.data(), d_cnt, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  offsets[0] = 0;
  for (uint64_t i = 0; i < n; ++i) offsets[i + 1] = offsets[i] + cnt[i];
  const uint64_t total = offsets[n];
  if (total > cap || (total > 0 && !idx)) {
    set_error("roadmap needs %llu adjacency entries, capacity is %llu", (unsigned long long)total,
              (unsigned long long)cap);
    return OXG_E_CAPACITY;  // every rank computes the same total, so all ranks return together
  }
  if (total == 0) return OXG_OK;
  OXG_TRY(d->idx_all.ensure(total * sizeof(uint32_t)));
  uint32_t* d_idx = d->idx_all.as<uint32_t>();
  if (!my_idx.empty())
    OXG_CUDA_TRY(cudaMemcpyAsync(d_idx + offsets[rb], my_idx.data(), my_idx.size() * sizeof(uint32_t),
                                 cudaMemcpyHostToDevice, s->stream));
  OXG_NCCL_TRY(api, api->GroupStart());
  for (int g = 0; g < G; ++g) {
    const uint64_t a = offsets[radius_cut(n, g, G)], b = offsets[radius_cut(n, g + 1, G)];
    if (b > a) OXG_NCCL_TRY(api, api->Broadcast(d_idx + a, d_idx + a, b - a, ncclUint32, g, d->comm, s->stream));
  }
  OXG_NCCL_TRY(api, api->GroupEnd());
  OXG_CUDA_TRY(cudaMemcpyAsync(idx, d_idx, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  return OXG_OK;
  OXG_API_END
}

}  // extern "C"
