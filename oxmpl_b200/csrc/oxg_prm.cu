// oxg_prm.cu — batched PRM roadmap construction (replaces the body of
// PRM::construct_roadmap, oxmpl/src/geometric/planners/prm.rs:117-146, for a store pre-filled
// with the valid samples in sample order).
//
//  radius mode : reference-faithful.  Node i sees the nodes j < i (prm.rs:131), connects when
//                distance(q_i, q_j) < connection_radius && check_motion(q_i, q_j) (prm.rs:133-134).
//                The output rows are `new_node.edges` at push time (ascending j).
//  k-NN mode   : count-bounded variant for SURVEY §8 config 4 (the reference has no k-NN).
//                Every candidate {i, j} is checked once, as check_motion(node max(i,j), node
//                min(i,j)) — the direction the incremental reference would have used.
//                A pair that appears in both rows (mutual neighbours) is checked by one
//                "owner" slot, chosen by the parity of i^j so owners spread evenly over rows
//                (this is what balances the shards across GPUs); the other slot copies it.
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "oxg_internal.cuh"
#include "oxg_math.cuh"

namespace oxg {

// Slot t of the idx matrix = (row base_row + t / k, neighbour idx[t]).  Rows of idx available
// for the mutual test: [avail_begin, avail_end).  Slots of rows [row_begin, row_end) are
// classified: 0 = padded; 2 = owned by the mirror slot; else appended to the edge list.
__global__ void prm_mark_edges_kernel(const uint32_t* __restrict__ idx, uint32_t k, uint64_t base_row,
                                      uint64_t avail_begin, uint64_t avail_end, uint64_t row_begin,
                                      uint64_t row_end, uint8_t* __restrict__ valid,
                                      uint32_t* __restrict__ e_from, uint32_t* __restrict__ e_to,
                                      uint32_t* __restrict__ e_slot, uint32_t* __restrict__ e_row,
                                      uint64_t* __restrict__ counters) {
  const uint64_t nslots = (row_end - row_begin) * k;
  for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < nslots;
       u += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = row_begin + u / k;
    const uint64_t t = (i - base_row) * k + (u % k);
    const uint32_t j = idx[t];
    if (j == OXG_INVALID_INDEX) {
      valid[t] = 0;
      continue;
    }
    bool mirror_owned = false;
    if (j >= avail_begin && j < avail_end) {
      const uint32_t* rj = idx + ((uint64_t)j - base_row) * k;
      bool mutual = false;
      for (uint32_t e = 0; e < k; ++e) mutual = mutual || (rj[e] == (uint32_t)i);
      if (mutual) {
        const uint64_t lo = i < j ? i : j, hi = i < j ? j : i;
        const uint64_t owner = ((i ^ (uint64_t)j) & 1ull) ? lo : hi;
        mirror_owned = owner != i;
      }
    }
    if (mirror_owned) {
      valid[t] = 2;
    } else {
      unsigned long long pos = atomicAdd((unsigned long long*)&counters[0], 1ull);
      e_from[pos] = (uint32_t)(i > j ? i : j);
      e_to[pos] = (uint32_t)(i > j ? j : i);
      e_slot[pos] = (uint32_t)t;
      if (e_row) e_row[pos] = (uint32_t)(i - row_begin);  // owner row, for the row cull
    }
  }
}

__global__ void prm_scatter_valid_kernel(const uint8_t* __restrict__ e_valid,
                                         const uint32_t* __restrict__ e_slot,
                                         uint64_t* __restrict__ counters, uint8_t* __restrict__ valid) {
  const uint64_t E = counters[0];
  unsigned long long local = 0;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E;
       e += (uint64_t)gridDim.x * blockDim.x) {
    uint8_t v = e_valid[e];
    valid[e_slot[e]] = v;
    local += v;
  }
  // warp-aggregated count of valid edges
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xFFFFFFFFu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd((unsigned long long*)&counters[1], local);
}

// Replaces every 2 ("owned by the mirror slot") by the mirror slot's value.
__global__ void prm_resolve_kernel(const uint32_t* __restrict__ idx, uint32_t k, uint64_t base_row,
                                   uint64_t row_begin, uint64_t row_end, uint8_t* __restrict__ valid) {
  const uint64_t nslots = (row_end - row_begin) * k;
  for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < nslots;
       u += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = row_begin + u / k;
    const uint64_t t = (i - base_row) * k + (u % k);
    if (valid[t] != 2) continue;
    const uint32_t j = idx[t];
    const uint64_t rj = ((uint64_t)j - base_row) * k;
    uint8_t v = 0;
    for (uint32_t e = 0; e < k; ++e)
      if (idx[rj + e] == (uint32_t)i) v = valid[rj + e];
    valid[t] = v;
  }
}

__global__ void count_candidates_kernel(const uint32_t* __restrict__ idx, uint64_t n,
                                        uint64_t* __restrict__ out) {
  unsigned long long local = 0;
  for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n;
       u += (uint64_t)gridDim.x * blockDim.x)
    local += idx[u] != OXG_INVALID_INDEX;
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xFFFFFFFFu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd((unsigned long long*)out, local);
}

static unsigned grid_for(uint64_t n, int threads) {
  return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + threads - 1) / threads, 148ull * 16));
}

// Edge phase shared by the host and the device-pointer entry points.
static int prm_edges(const oxg_world_s* w, oxg_store_s* s, const uint32_t* d_idx, uint32_t k,
                     double resolution, uint64_t base_row, uint64_t avail_begin, uint64_t avail_end,
                     uint64_t row_begin, uint64_t row_end, uint8_t* d_valid, uint64_t* d_counters) {
  const uint64_t rows = row_end - row_begin;
  const uint64_t nslots = rows * k;
  if (nslots == 0) return OXG_OK;
  if (nslots > 0xFFFFFFF0ull) {
    set_error("too many candidate slots in one call");
    return OXG_E_ARG;
  }
  OXG_TRY(s->buf[B_AUX].ensure(nslots * sizeof(uint32_t)));
  OXG_TRY(s->buf[B_AUX2].ensure(nslots * sizeof(uint32_t)));
  OXG_TRY(s->buf[B_AUX3].ensure(nslots * sizeof(uint32_t)));
  OXG_TRY(s->buf[B_TMP].ensure(nslots));
  uint32_t* e_from = s->buf[B_AUX].as<uint32_t>();
  uint32_t* e_to = s->buf[B_AUX2].as<uint32_t>();
  uint32_t* e_slot = s->buf[B_AUX3].as<uint32_t>();
  uint8_t* e_valid = s->buf[B_TMP].as<uint8_t>();
  OXG_CUDA_TRY(cudaMemsetAsync(d_counters, 0, 2 * sizeof(uint64_t), s->stream));
  // Row cull: the obstacles no candidate edge of a row can touch are ruled out once per row
  // (oxg_internal.cuh); the edge kernel then starts from the row's short list.
  OXG_TRY(s->buf[B_ROWCNT].ensure(rows * sizeof(uint16_t)));
  OXG_TRY(s->buf[B_ROWLIST].ensure(rows * kRowCullCap * sizeof(uint16_t)));
  bool row_cull = false;
  OXG_TRY(prm_row_cull_device(w, s->space, s->stream, store_view(s), d_idx, k, base_row, row_begin, row_end,
                              0.0, nslots, s->buf[B_ROWCNT].as<uint16_t>(),
                              s->buf[B_ROWLIST].as<uint16_t>(), &row_cull));
  uint32_t* e_row = nullptr;
  if (row_cull) {
    OXG_TRY(s->buf[B_EROW].ensure(nslots * sizeof(uint32_t)));
    e_row = s->buf[B_EROW].as<uint32_t>();
  }
  prm_mark_edges_kernel<<<grid_for(nslots, 256), 256, 0, s->stream>>>(
      d_idx, k, base_row, avail_begin, avail_end, row_begin, row_end, d_valid, e_from, e_to, e_slot, e_row,
      d_counters);
  OXG_LAUNCH_CHECK();
  RowCull rc;
  rc.e_row = e_row;
  rc.row_cnt = s->buf[B_ROWCNT].as<uint16_t>();
  rc.row_list = s->buf[B_ROWLIST].as<uint16_t>();
  OXG_TRY(motion_check_device(w, s->space, s->stream, store_view(s), store_view(s), e_from, e_to,
                              nslots, d_counters, resolution, e_valid, nullptr, row_cull ? &rc : nullptr));
  prm_scatter_valid_kernel<<<grid_for(nslots, 256), 256, 0, s->stream>>>(e_valid, e_slot, d_counters,
                                                                        d_valid);
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

// ---- radius mode helpers ------------------------------------------------------------------
// from-index of every candidate hit: the query row that owns it.
__global__ void expand_rows_kernel(const uint64_t* __restrict__ row_off, uint64_t rows,
                                   uint64_t row_begin, uint32_t* __restrict__ e_from,
                                   uint32_t* __restrict__ e_row) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < rows;
       q += (uint64_t)gridDim.x * blockDim.x)
    for (uint64_t t = row_off[q]; t < row_off[q + 1]; ++t) {
      e_from[t] = (uint32_t)(row_begin + q);
      if (e_row) e_row[t] = (uint32_t)q;  // owner row, for the row cull
    }
}

__global__ void count_valid_rows_kernel(const uint64_t* __restrict__ row_off, uint64_t rows,
                                        const uint8_t* __restrict__ e_valid,
                                        uint64_t* __restrict__ counts) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q <= rows;
       q += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t c = 0;
    if (q < rows)
      for (uint64_t t = row_off[q]; t < row_off[q + 1]; ++t) c += e_valid[t];
    counts[q] = c;
  }
}

__global__ void fill_valid_rows_kernel(const uint64_t* __restrict__ row_off, uint64_t rows,
                                       const uint8_t* __restrict__ e_valid,
                                       const uint32_t* __restrict__ hit_idx,
                                       const uint64_t* __restrict__ out_off,
                                       uint32_t* __restrict__ out_idx) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < rows;
       q += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t o = out_off[q];
    for (uint64_t t = row_off[q]; t < row_off[q + 1]; ++t)
      if (e_valid[t]) out_idx[o++] = hit_idx[t];
  }
}

}  // namespace oxg

using namespace oxg;

extern "C" {

int oxg_prm_build_knn(oxg_world w, oxg_store s, uint32_t k, double resolution, uint64_t row_begin,
                      uint64_t row_end, uint32_t* idx, double* dist, uint8_t* valid,
                      oxg_prm_stats* stats) {
  OXG_API_BEGIN
  if (!w || !s || !idx || !dist || !valid) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (row_begin > row_end || row_end > s->size) {
    set_error("row range [%llu, %llu) outside the store (size %llu)", (unsigned long long)row_begin,
              (unsigned long long)row_end, (unsigned long long)s->size);
    return OXG_E_ARG;
  }
  if (!(resolution > 0.0)) {
    set_error("motion resolution must be > 0 (got %g)", resolution);
    return OXG_E_RESOLUTION;
  }
  const uint64_t rows = row_end - row_begin;
  if (stats) memset(stats, 0, sizeof(*stats));
  if (rows == 0) return OXG_OK;
  const size_t n = (size_t)rows * k;
  OXG_TRY(s->buf[B_OIDX].ensure(n * sizeof(uint32_t)));
  OXG_TRY(s->buf[B_ODIST].ensure(n * sizeof(double)));
  OXG_TRY(s->buf[B_STAGE].ensure(n));
  OXG_TRY(s->buf[B_COUNTS].ensure(4 * sizeof(uint64_t)));
  uint32_t* d_idx = s->buf[B_OIDX].as<uint32_t>();
  double* d_dist = s->buf[B_ODIST].as<double>();
  uint8_t* d_valid = s->buf[B_STAGE].as<uint8_t>();
  uint64_t* d_counters = s->buf[B_COUNTS].as<uint64_t>();

  OXG_CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
  OXG_TRY(oxg_knn_rows_dev(s, row_begin, row_end, k, d_idx, d_dist));
  OXG_CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
  OXG_TRY(prm_edges(w, s, d_idx, k, resolution, row_begin, row_begin, row_end, row_begin, row_end,
                    d_valid, d_counters));
  prm_resolve_kernel<<<grid_for(n, 256), 256, 0, s->stream>>>(d_idx, k, row_begin, row_begin, row_end,
                                                              d_valid);
  OXG_LAUNCH_CHECK();
  OXG_CUDA_TRY(cudaMemsetAsync(d_counters + 2, 0, sizeof(uint64_t), s->stream));
  count_candidates_kernel<<<grid_for(n, 256), 256, 0, s->stream>>>(d_idx, n, d_counters + 2);
  OXG_LAUNCH_CHECK();
  OXG_CUDA_TRY(cudaEventRecord(s->ev2, s->stream));
  uint64_t h_counters[3] = {0, 0, 0};
  OXG_CUDA_TRY(cudaMemcpyAsync(idx, d_idx, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  OXG_CUDA_TRY(cudaMemcpyAsync(dist, d_dist, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  OXG_CUDA_TRY(cudaMemcpyAsync(valid, d_valid, n, cudaMemcpyDeviceToHost, s->stream));
  OXG_CUDA_TRY(cudaMemcpyAsync(h_counters, d_counters, sizeof(h_counters), cudaMemcpyDeviceToHost,
                               s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (stats) {
    float ms_knn = 0.f, ms_edge = 0.f;
    cudaEventElapsedTime(&ms_knn, s->ev0, s->ev1);
    cudaEventElapsedTime(&ms_edge, s->ev1, s->ev2);
    stats->n_rows = rows;
    stats->n_candidates = h_counters[2];
    stats->n_edges_checked = h_counters[0];
    stats->n_edges_valid = h_counters[1];
    stats->knn_ms = ms_knn;
    stats->edge_ms = ms_edge;
  }
  return OXG_OK;
  OXG_API_END
}

int oxg_prm_edges_dev(oxg_world w, oxg_store s, const uint32_t* d_idx_all, uint32_t k,
                      double resolution, uint64_t row_begin, uint64_t row_end, uint8_t* d_valid_all,
                      uint64_t* d_counters) {
  OXG_API_BEGIN
  if (!w || !s || !d_idx_all || !d_valid_all || !d_counters) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (row_begin > row_end || row_end > s->size) {
    set_error("row range outside the store");
    return OXG_E_ARG;
  }
  if (!(resolution > 0.0)) {
    set_error("motion resolution must be > 0 (got %g)", resolution);
    return OXG_E_RESOLUTION;
  }
  return prm_edges(w, s, d_idx_all, k, resolution, 0, 0, s->size, row_begin, row_end, d_valid_all,
                   d_counters);
  OXG_API_END
}

int oxg_prm_resolve_dev(oxg_store s, const uint32_t* d_idx_all, uint32_t k, uint8_t* d_valid_all) {
  OXG_API_BEGIN
  if (!s || !d_idx_all || !d_valid_all) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  const uint64_t n = s->size * k;
  if (n == 0) return OXG_OK;
  prm_resolve_kernel<<<grid_for(n, 256), 256, 0, s->stream>>>(d_idx_all, k, 0, 0, s->size,
                                                              d_valid_all);
  OXG_LAUNCH_CHECK();
  return OXG_OK;
  OXG_API_END
}

int oxg_prm_build_radius(oxg_world w, oxg_store s, double connection_radius, double resolution,
                         uint64_t row_begin, uint64_t row_end, uint64_t* offsets, uint32_t* idx,
                         uint64_t cap, oxg_prm_stats* stats) {
  OXG_API_BEGIN
  if (!w || !s || !offsets) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (row_begin > row_end || row_end > s->size) {
    set_error("row range [%llu, %llu) outside the store (size %llu)", (unsigned long long)row_begin,
              (unsigned long long)row_end, (unsigned long long)s->size);
    return OXG_E_ARG;
  }
  if (!(resolution > 0.0)) {
    set_error("motion resolution must be > 0 (got %g)", resolution);
    return OXG_E_RESOLUTION;
  }
  const uint64_t rows = row_end - row_begin;
  if (stats) memset(stats, 0, sizeof(*stats));
  offsets[0] = 0;
  if (rows == 0) return OXG_OK;

  OXG_CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
  // candidates: j < i with distance(q_i, q_j) < r   (prm.rs:131-134)
  StateView q = store_view(s);
  q.p += row_begin;
  uint64_t total = 0;
  OXG_TRY(nn_radius_device(s, q, rows, connection_radius, nullptr, /*limit_self=*/true, row_begin,
                           false, &total));
  OXG_TRY(s->buf[B_AUX].ensure((rows + 1) * sizeof(uint64_t)));   // candidate CSR offsets
  OXG_TRY(s->buf[B_COUNTS].ensure((rows + 2) * sizeof(uint64_t)));  // valid counts
  OXG_TRY(s->buf[B_AUX3].ensure((rows + 2) * sizeof(uint64_t)));  // valid CSR offsets
  uint64_t* d_row_off = s->buf[B_AUX].as<uint64_t>();
  OXG_TRY(nn_radius_offsets(s, rows, d_row_off));
  OXG_CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
  uint64_t n_valid = 0;
  if (total > 0) {
    if (total > 0xFFFFFFF0ull) {
      set_error("too many candidate edges in one call");
      return OXG_E_ARG;
    }
    OXG_TRY(s->buf[B_AUX2].ensure(total * sizeof(uint32_t)));  // from indices
    OXG_TRY(s->buf[B_STAGE].ensure(total));                    // per-candidate validity
    uint32_t* e_from = s->buf[B_AUX2].as<uint32_t>();
    uint32_t* e_to = s->buf[B_PIDX].as<uint32_t>();
    uint8_t* e_valid = s->buf[B_STAGE].as<uint8_t>();
    // row cull (oxg_internal.cuh): every candidate of row i is closer than connection_radius
    OXG_TRY(s->buf[B_ROWCNT].ensure(rows * sizeof(uint16_t)));
    OXG_TRY(s->buf[B_ROWLIST].ensure(rows * kRowCullCap * sizeof(uint16_t)));
    bool row_cull = false;
    OXG_TRY(prm_row_cull_device(w, s->space, s->stream, store_view(s), nullptr, 0, 0, row_begin, row_end,
                                connection_radius, total, s->buf[B_ROWCNT].as<uint16_t>(),
                                s->buf[B_ROWLIST].as<uint16_t>(), &row_cull));
    uint32_t* e_row = nullptr;
    if (row_cull) {
      OXG_TRY(s->buf[B_EROW].ensure(total * sizeof(uint32_t)));
      e_row = s->buf[B_EROW].as<uint32_t>();
    }
    expand_rows_kernel<<<grid_for(rows, 128), 128, 0, s->stream>>>(d_row_off, rows, row_begin, e_from, e_row);
    OXG_LAUNCH_CHECK();
    RowCull rc;
    rc.e_row = e_row;
    rc.row_cnt = s->buf[B_ROWCNT].as<uint16_t>();
    rc.row_list = s->buf[B_ROWLIST].as<uint16_t>();
    OXG_TRY(motion_check_device(w, s->space, s->stream, store_view(s), store_view(s), e_from, e_to,
                                total, nullptr, resolution, e_valid, nullptr, row_cull ? &rc : nullptr));
    uint64_t* d_counts = s->buf[B_COUNTS].as<uint64_t>();
    uint64_t* d_out_off = s->buf[B_AUX3].as<uint64_t>();
    count_valid_rows_kernel<<<grid_for(rows + 1, 128), 128, 0, s->stream>>>(d_row_off, rows, e_valid,
                                                                          d_counts);
    OXG_LAUNCH_CHECK();
    size_t tmp_bytes = 0;
    OXG_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_counts, d_out_off,
                                               (int)(rows + 1), s->stream));
    OXG_TRY(s->buf[B_TMP].ensure(tmp_bytes));
    OXG_CUDA_TRY(cub::DeviceScan::ExclusiveSum(s->buf[B_TMP].p, tmp_bytes, d_counts, d_out_off,
                                               (int)(rows + 1), s->stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    OXG_TRY(s->buf[B_PDIST].ensure(total * sizeof(uint32_t)));  // compacted neighbour indices
    uint32_t* d_out_idx = s->buf[B_PDIST].as<uint32_t>();
    fill_valid_rows_kernel<<<grid_for(rows, 128), 128, 0, s->stream>>>(d_row_off, rows, e_valid, e_to,
                                                                      d_out_off, d_out_idx);
    OXG_LAUNCH_CHECK();
    OXG_CUDA_TRY(cudaEventRecord(s->ev2, s->stream));
    OXG_CUDA_TRY(cudaMemcpyAsync(offsets, d_out_off, (rows + 1) * sizeof(uint64_t),
                                 cudaMemcpyDeviceToHost, s->stream));
    OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    n_valid = offsets[rows];
    if (n_valid > cap || (n_valid > 0 && !idx)) {
      set_error("roadmap needs %llu adjacency entries, capacity is %llu", (unsigned long long)n_valid,
                (unsigned long long)cap);
      return OXG_E_CAPACITY;
    }
    if (n_valid > 0) {
      OXG_CUDA_TRY(cudaMemcpyAsync(idx, d_out_idx, n_valid * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                   s->stream));
      OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    }
  } else {
    OXG_CUDA_TRY(cudaEventRecord(s->ev2, s->stream));
    OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    for (uint64_t q2 = 0; q2 <= rows; ++q2) offsets[q2] = 0;
  }
  if (stats) {
    float ms_a = 0.f, ms_b = 0.f;
    cudaEventElapsedTime(&ms_a, s->ev0, s->ev1);
    cudaEventElapsedTime(&ms_b, s->ev1, s->ev2);
    stats->n_rows = rows;
    stats->n_candidates = total;
    stats->n_edges_checked = total;
    stats->n_edges_valid = n_valid;
    stats->knn_ms = ms_a;
    stats->edge_ms = ms_b;
  }
  return OXG_OK;
  OXG_API_END
}

}  // extern "C"
