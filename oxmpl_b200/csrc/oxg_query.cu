// oxg_query.cu — the query phase of PRM on a GPU-built roadmap: PRM::solve,
// oxmpl/src/geometric/planners/prm.rs:227-307.
//
//   start validity (:244)  ->  start connection: radius scan + check_motion(start, roadmap[i])
//   (:249-256)  ->  goal scan: goal.is_satisfied over all nodes (:259-264, the ball goal of the
//   reference's tests)  ->  breadth-first search over `roadmap[i].edges` (:271-301)  ->  path.
//
// The roadmap arrives as the library built it: the lower-index adjacency of
// oxg_prm_build_radius (`new_node.edges` at push time, prm.rs:135) or the k-NN table of
// oxg_prm_build_knn with its validity bytes.  It is symmetrised on the device (prm.rs:143-145):
// every edge in both directions, one radix sort by (node, neighbour), so each row is ascending —
// which is the reference's order: node i first collects its earlier neighbours in ascending
// order, then later nodes append themselves as they arrive.
//
// The search reproduces the reference's BFS exactly, not just some shortest path: the queue
// order decides which parent a node gets.  Level by level, every frontier node u (queue rank
// r_u) offers each unvisited neighbour v the key (r_u, position of v in u's row); the smallest
// key wins (atomicMin) — that is the u whose turn comes first in the sequential loop — and the
// next level's queue is the winners sorted by key, which is the order the sequential loop pushes
// them in.  The answer is the first goal node in queue order and its parent chain.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include <algorithm>
#include <vector>

#include "oxg_internal.cuh"
#include "oxg_math.cuh"

namespace oxg {

static unsigned grid_q(uint64_t n, int threads) {
  return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + threads - 1) / threads, 148ull * 16));
}

// both directions of every lower-index edge as (node << 32 | neighbour)
__global__ void query_keys_from_lower_kernel(const uint64_t* __restrict__ off, const uint32_t* __restrict__ idx,
                                             uint64_t n, unsigned long long* __restrict__ keys) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    for (uint64_t t = off[i]; t < off[i + 1]; ++t) {
      const unsigned long long j = idx[t];
      keys[2 * t] = (i << 32) | j;
      keys[2 * t + 1] = (j << 32) | i;
    }
}

// k-NN table: slot (i, s) with valid = 1 gives the edge {i, idx[i][s]}; a mutual pair appears in
// both rows and is emitted once per direction from each — duplicates are dropped after the sort.
__global__ void query_keys_from_knn_kernel(const uint32_t* __restrict__ idx, const uint8_t* __restrict__ valid,
                                           uint64_t n, uint32_t k, unsigned long long* __restrict__ keys) {
  const uint64_t total = n * k;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long i = t / k, j = idx[t];
    const bool ok = valid[t] == 1 && j != OXG_INVALID_INDEX;
    keys[2 * t] = ok ? ((i << 32) | j) : ~0ull;
    keys[2 * t + 1] = ok ? ((j << 32) | i) : ~0ull;
  }
}

__global__ void query_flag_unique_kernel(const unsigned long long* __restrict__ keys, uint64_t m,
                                         uint8_t* __restrict__ flag) {
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < m;
       t += (uint64_t)gridDim.x * blockDim.x)
    flag[t] = keys[t] != ~0ull && (t == 0 || keys[t] != keys[t - 1]);
}

// CSR offsets of the sorted (node << 32 | neighbour) keys: off[v] = first key with node >= v
__global__ void query_row_offsets_kernel(const unsigned long long* __restrict__ keys, uint64_t m, uint64_t n,
                                         uint64_t* __restrict__ off, uint32_t* __restrict__ adj) {
  for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v <= n;
       v += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t lo = 0, hi = m;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if ((keys[mid] >> 32) < v) lo = mid + 1; else hi = mid;
    }
    off[v] = lo;
  }
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < m;
       t += (uint64_t)gridDim.x * blockDim.x)
    adj[t] = (uint32_t)(keys[t] & 0xFFFFFFFFull);
}

// goal.is_satisfied(state) = distance(state, target) <= radius
// (CircularGoalRegion, oxmpl/tests/rrt_rvss_tests.rs:46-50)
__global__ void query_goal_kernel(StateView st, uint64_t n, uint32_t n_real, uint32_t n_so2, double w_real,
                                  double w_so2, const double* __restrict__ target, double radius,
                                  uint8_t* __restrict__ is_goal) {
  const uint32_t dim = n_real + n_so2;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    double x[kMaxDim], t[kMaxDim];
    for (uint32_t d = 0; d < dim; ++d) {
      x[d] = st.p[i * st.stride_i + (uint64_t)d * st.stride_d];
      t[d] = target[d];
    }
    is_goal[i] = space_distance_dev(n_real, n_so2, w_real, w_so2, x, t) <= radius ? 1 : 0;
  }
}

struct BfsCtl {
  unsigned long long goal_rank;  // smallest queue rank of a goal node in the current level (~0 = none)
  unsigned int n_next;
  unsigned int pad;
};

// level 0: the start connections in ascending index order
__global__ void bfs_seed_kernel(const uint32_t* __restrict__ conn, const uint8_t* __restrict__ ok, uint32_t m,
                                uint32_t* __restrict__ queue, uint32_t* __restrict__ rank,
                                uint32_t* __restrict__ parent, unsigned int* __restrict__ n_out) {
  // single thread block, ordered compaction (m is the size of one radius neighbourhood)
  __shared__ unsigned int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (uint32_t t0 = 0; t0 < m; t0 += blockDim.x) {
    const uint32_t t = t0 + threadIdx.x;
    const bool keep = t < m && ok[t];
    // block-wide exclusive count of `keep` in thread order
    __shared__ unsigned int warp_cnt[32];
    const unsigned b = __ballot_sync(0xFFFFFFFFu, keep);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[wid] = __popc(b);
    __syncthreads();
    unsigned int before = 0;
    for (int w2 = 0; w2 < wid; ++w2) before += warp_cnt[w2];
    unsigned int total = 0;
    for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) total += warp_cnt[w2];
    if (keep) {
      const unsigned int r = base + before + __popc(b & ((1u << lane) - 1u));
      const uint32_t v = conn[t];
      queue[r] = v;
      rank[v] = r;
      parent[v] = OXG_INVALID_INDEX;  // parent_map.insert(idx, None)  (prm.rs:276)
    }
    __syncthreads();
    if (threadIdx.x == 0) base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_out = base;
}

// `if goal_indices.contains(&current_idx)` at pop time (prm.rs:287): the first goal node in queue order
__global__ void bfs_goal_kernel(const uint32_t* __restrict__ queue, uint32_t level_begin, uint32_t level_end,
                                const uint8_t* __restrict__ is_goal, BfsCtl* ctl) {
  for (uint32_t r = level_begin + blockIdx.x * blockDim.x + threadIdx.x; r < level_end;
       r += gridDim.x * blockDim.x)
    if (is_goal[queue[r]]) atomicMin(&ctl->goal_rank, (unsigned long long)r);
}

// every frontier node offers its unvisited neighbours the key (its rank, position in its row)
__global__ void bfs_expand_kernel(const uint32_t* __restrict__ queue, uint32_t level_begin, uint32_t level_end,
                                  const uint64_t* __restrict__ off, const uint32_t* __restrict__ adj,
                                  const uint32_t* __restrict__ rank, unsigned long long* __restrict__ key) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = level_begin + warp; r < level_end; r += nwarps) {
    const uint32_t u = queue[r];
    const uint64_t a = off[u], b = off[u + 1];
    for (uint64_t t = a + lane; t < b; t += 32) {
      const uint32_t v = adj[t];
      if (rank[v] == OXG_INVALID_INDEX)
        atomicMin(&key[v], ((unsigned long long)r << 32) | (unsigned long long)(t - a));
    }
  }
}

// candidates of the next level: nodes that received a key in this level
__global__ void bfs_collect_kernel(const uint32_t* __restrict__ queue, uint32_t level_begin, uint32_t level_end,
                                   const uint64_t* __restrict__ off, const uint32_t* __restrict__ adj,
                                   const uint32_t* __restrict__ rank, const unsigned long long* __restrict__ key,
                                   unsigned long long* __restrict__ cand_key, uint32_t* __restrict__ cand_node,
                                   BfsCtl* ctl) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = level_begin + warp; r < level_end; r += nwarps) {
    const uint32_t u = queue[r];
    const uint64_t a = off[u], b = off[u + 1];
    for (uint64_t t = a + lane; t < b; t += 32) {
      const uint32_t v = adj[t];
      // exactly one (u, position) pair holds v's winning key: that one emits v
      if (rank[v] == OXG_INVALID_INDEX &&
          key[v] == (((unsigned long long)r << 32) | (unsigned long long)(t - a))) {
        const unsigned int pos = atomicAdd(&ctl->n_next, 1u);
        cand_key[pos] = key[v];
        cand_node[pos] = v;
      }
    }
  }
}

// the next level in queue order: ranks continue after the current queue
__global__ void bfs_commit_kernel(const unsigned long long* __restrict__ sorted_key,
                                  const uint32_t* __restrict__ sorted_node, uint32_t m, uint32_t queue_end,
                                  uint32_t* __restrict__ queue, uint32_t* __restrict__ rank,
                                  uint32_t* __restrict__ parent) {
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x) {
    const uint32_t v = sorted_node[t];
    queue[queue_end + t] = v;
    rank[v] = queue_end + t;
    parent[v] = queue[(uint32_t)(sorted_key[t] >> 32)];  // parent_map.insert(neighbor, Some(current))
  }
}

__global__ void fill_u32_kernel(uint32_t* p, uint64_t n, uint32_t v) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    p[i] = v;
}
__global__ void fill_u64_kernel(unsigned long long* p, uint64_t n, unsigned long long v) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

struct QueryBufs {  // device scratch of one query (freed on return)
  DevBuf keys, keys2, flag, ukeys, nsel, off, adj, goal, rank, parent, queue, key, ckey, cnode, ckey2, cnode2,
      ctl, tmp, conn_ok, target;
  ~QueryBufs() {
    for (DevBuf* b : {&keys, &keys2, &flag, &ukeys, &nsel, &off, &adj, &goal, &rank, &parent, &queue, &key, &ckey,
                      &cnode, &ckey2, &cnode2, &ctl, &tmp, &conn_ok, &target})
      b->release();
  }
};

// The query on a device edge-key array (2 keys per undirected edge occurrence, ~0 = none).
static int prm_solve_impl(oxg_world_s* w, oxg_store_s* s, QueryBufs& B, uint64_t n_keys, double connection_radius,
                          double resolution, const double* start, const double* goal_target, double goal_radius,
                          uint32_t* path, uint32_t path_cap, oxg_prm_query_result* out) {
  cudaStream_t st = s->stream;
  const uint64_t n = s->size;
  const uint32_t dim = s->dim;
  memset(out, 0, sizeof(*out));
  out->goal_node = OXG_INVALID_INDEX;
  OXG_CUDA_TRY(cudaEventRecord(s->ev0, st));

  // ---- is_valid(start)  (prm.rs:244)
  OXG_TRY(s->buf[B_QUERY].ensure(sizeof(double) * dim));
  OXG_TRY(B.target.ensure(sizeof(double) * dim));
  OXG_TRY(B.conn_ok.ensure(std::max<uint64_t>(n, 16)));
  OXG_CUDA_TRY(cudaMemcpyAsync(s->buf[B_QUERY].p, start, sizeof(double) * dim, cudaMemcpyHostToDevice, st));
  OXG_CUDA_TRY(cudaMemcpyAsync(B.target.p, goal_target, sizeof(double) * dim, cudaMemcpyHostToDevice, st));
  const StateView q = rows_view(s->buf[B_QUERY].as<double>(), dim);
  OXG_TRY(is_valid_device(w, s->space, st, q, 1, B.conn_ok.as<uint8_t>()));
  uint8_t start_ok = 0;
  OXG_CUDA_TRY(cudaMemcpyAsync(&start_ok, B.conn_ok.p, 1, cudaMemcpyDeviceToHost, st));
  OXG_CUDA_TRY(cudaStreamSynchronize(st));
  if (!start_ok) {
    out->status = 2;  // PlanningError::InvalidStartState
    return OXG_OK;
  }

  // ---- symmetric adjacency, rows ascending  (prm.rs:135, 143-145)
  uint64_t m = 0;
  if (n_keys > 0) {
    OXG_TRY(B.keys2.ensure(n_keys * sizeof(unsigned long long)));
    size_t tb = 0;
    OXG_CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, tb, B.keys.as<unsigned long long>(),
                                                B.keys2.as<unsigned long long>(), (int)n_keys, 0, 64, st));
    OXG_TRY(B.tmp.ensure(tb));
    OXG_CUDA_TRY(cub::DeviceRadixSort::SortKeys(B.tmp.p, tb, B.keys.as<unsigned long long>(),
                                                B.keys2.as<unsigned long long>(), (int)n_keys, 0, 64, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    OXG_TRY(B.flag.ensure(n_keys));
    OXG_TRY(B.ukeys.ensure(n_keys * sizeof(unsigned long long)));
    OXG_TRY(B.nsel.ensure(sizeof(unsigned long long)));
    query_flag_unique_kernel<<<grid_q(n_keys, 256), 256, 0, st>>>(B.keys2.as<unsigned long long>(), n_keys,
                                                                 B.flag.as<uint8_t>());
    OXG_LAUNCH_CHECK();
    tb = 0;
    OXG_CUDA_TRY(cub::DeviceSelect::Flagged(nullptr, tb, B.keys2.as<unsigned long long>(), B.flag.as<uint8_t>(),
                                            B.ukeys.as<unsigned long long>(), B.nsel.as<unsigned long long>(),
                                            (int)n_keys, st));
    OXG_TRY(B.tmp.ensure(tb));
    OXG_CUDA_TRY(cub::DeviceSelect::Flagged(B.tmp.p, tb, B.keys2.as<unsigned long long>(), B.flag.as<uint8_t>(),
                                            B.ukeys.as<unsigned long long>(), B.nsel.as<unsigned long long>(),
                                            (int)n_keys, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    unsigned long long h_m = 0;
    OXG_CUDA_TRY(cudaMemcpyAsync(&h_m, B.nsel.p, sizeof(h_m), cudaMemcpyDeviceToHost, st));
    OXG_CUDA_TRY(cudaStreamSynchronize(st));
    m = h_m;
  }
  OXG_TRY(B.off.ensure((n + 1) * sizeof(uint64_t)));
  OXG_TRY(B.adj.ensure(std::max<uint64_t>(m, 1) * sizeof(uint32_t)));
  query_row_offsets_kernel<<<grid_q(std::max<uint64_t>(m, n + 1), 256), 256, 0, st>>>(
      B.ukeys.as<unsigned long long>(), m, n, B.off.as<uint64_t>(), B.adj.as<uint32_t>());
  OXG_LAUNCH_CHECK();
  out->n_adjacency = m;

  // ---- start connection  (prm.rs:249-256): distance(start, roadmap[i]) < r && check_motion(start, roadmap[i])
  uint64_t total = 0;
  OXG_TRY(nn_radius_device(s, q, 1, connection_radius, nullptr, false, 0, false, &total));
  uint32_t n_conn = 0;
  OXG_TRY(B.rank.ensure(n * sizeof(uint32_t)));
  OXG_TRY(B.parent.ensure(n * sizeof(uint32_t)));
  OXG_TRY(B.queue.ensure(n * sizeof(uint32_t)));
  OXG_TRY(B.key.ensure(n * sizeof(unsigned long long)));
  OXG_TRY(B.ctl.ensure(sizeof(BfsCtl) + 16));
  fill_u32_kernel<<<grid_q(n, 256), 256, 0, st>>>(B.rank.as<uint32_t>(), n, OXG_INVALID_INDEX);
  OXG_LAUNCH_CHECK();
  fill_u64_kernel<<<grid_q(n, 256), 256, 0, st>>>(B.key.as<unsigned long long>(), n, ~0ull);
  OXG_LAUNCH_CHECK();
  if (total > 0) {
    const uint32_t* d_conn = s->buf[B_PIDX].as<uint32_t>();  // ascending indices of the radius scan
    // from = the start state for every candidate (stride 0), to = store rows by index
    StateView from{s->buf[B_QUERY].as<double>(), 0, 1};
    OXG_TRY(motion_check_device(w, s->space, st, from, store_view(s), nullptr, d_conn, total, nullptr, resolution,
                                B.conn_ok.as<uint8_t>(), nullptr));
    unsigned int* d_n = reinterpret_cast<unsigned int*>(B.ctl.as<unsigned char>() + sizeof(BfsCtl));
    bfs_seed_kernel<<<1, 1024, 0, st>>>(d_conn, B.conn_ok.as<uint8_t>(), (uint32_t)total, B.queue.as<uint32_t>(),
                                        B.rank.as<uint32_t>(), B.parent.as<uint32_t>(), d_n);
    OXG_LAUNCH_CHECK();
    OXG_CUDA_TRY(cudaMemcpyAsync(&n_conn, d_n, sizeof(n_conn), cudaMemcpyDeviceToHost, st));
  }

  // ---- goal scan  (prm.rs:259-264)
  OXG_TRY(B.goal.ensure(std::max<uint64_t>(n, 1)));
  query_goal_kernel<<<grid_q(n, 128), 128, 0, st>>>(store_view(s), n, s->space.n_real, s->space.n_so2,
                                                   s->space.w_real, s->space.w_so2, B.target.as<double>(),
                                                   goal_radius, B.goal.as<uint8_t>());
  OXG_LAUNCH_CHECK();
  std::vector<uint8_t> h_goal(n);
  OXG_CUDA_TRY(cudaMemcpyAsync(h_goal.data(), B.goal.p, n, cudaMemcpyDeviceToHost, st));
  OXG_CUDA_TRY(cudaStreamSynchronize(st));
  uint32_t n_goal = 0;
  for (uint8_t g : h_goal) n_goal += g;
  out->n_start_connections = n_conn;
  out->n_goal_nodes = n_goal;
  if (n_conn == 0 || n_goal == 0) {
    out->status = 1;  // PlanningError::NoSolutionFound (prm.rs:266-268)
    return OXG_OK;
  }

  // ---- breadth-first search in the reference's queue order  (prm.rs:271-301)
  OXG_TRY(B.ckey.ensure(n * sizeof(unsigned long long)));
  OXG_TRY(B.ckey2.ensure(n * sizeof(unsigned long long)));
  OXG_TRY(B.cnode.ensure(n * sizeof(uint32_t)));
  OXG_TRY(B.cnode2.ensure(n * sizeof(uint32_t)));
  BfsCtl* d_ctl = B.ctl.as<BfsCtl>();
  uint32_t level_begin = 0, level_end = n_conn, levels = 0;
  uint32_t goal_rank = OXG_INVALID_INDEX;
  for (;;) {
    BfsCtl h{~0ull, 0u, 0u};
    OXG_CUDA_TRY(cudaMemcpyAsync(d_ctl, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    const uint32_t width = level_end - level_begin;
    bfs_goal_kernel<<<grid_q(width, 256), 256, 0, st>>>(B.queue.as<uint32_t>(), level_begin, level_end,
                                                       B.goal.as<uint8_t>(), d_ctl);
    OXG_LAUNCH_CHECK();
    bfs_expand_kernel<<<grid_q((uint64_t)width * 32, 256), 256, 0, st>>>(
        B.queue.as<uint32_t>(), level_begin, level_end, B.off.as<uint64_t>(), B.adj.as<uint32_t>(),
        B.rank.as<uint32_t>(), B.key.as<unsigned long long>());
    OXG_LAUNCH_CHECK();
    bfs_collect_kernel<<<grid_q((uint64_t)width * 32, 256), 256, 0, st>>>(
        B.queue.as<uint32_t>(), level_begin, level_end, B.off.as<uint64_t>(), B.adj.as<uint32_t>(),
        B.rank.as<uint32_t>(), B.key.as<unsigned long long>(), B.ckey.as<unsigned long long>(),
        B.cnode.as<uint32_t>(), d_ctl);
    OXG_LAUNCH_CHECK();
    OXG_CUDA_TRY(cudaMemcpyAsync(&h, d_ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
    OXG_CUDA_TRY(cudaStreamSynchronize(st));
    ++levels;
    if (h.goal_rank != ~0ull) {
      goal_rank = (uint32_t)h.goal_rank;
      break;
    }
    if (h.n_next == 0) break;
    size_t tb = 0;
    OXG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, B.ckey.as<unsigned long long>(),
                                                 B.ckey2.as<unsigned long long>(), B.cnode.as<uint32_t>(),
                                                 B.cnode2.as<uint32_t>(), (int)h.n_next, 0, 64, st));
    OXG_TRY(B.tmp.ensure(tb));
    OXG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(B.tmp.p, tb, B.ckey.as<unsigned long long>(),
                                                 B.ckey2.as<unsigned long long>(), B.cnode.as<uint32_t>(),
                                                 B.cnode2.as<uint32_t>(), (int)h.n_next, 0, 64, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    bfs_commit_kernel<<<grid_q(h.n_next, 256), 256, 0, st>>>(B.ckey2.as<unsigned long long>(),
                                                            B.cnode2.as<uint32_t>(), h.n_next, level_end,
                                                            B.queue.as<uint32_t>(), B.rank.as<uint32_t>(),
                                                            B.parent.as<uint32_t>());
    OXG_LAUNCH_CHECK();
    level_begin = level_end;
    level_end += h.n_next;
  }
  out->bfs_levels = levels;
  out->n_visited = level_end;
  if (goal_rank == OXG_INVALID_INDEX) {
    out->status = 1;  // `goal_reached.ok_or(PlanningError::NoSolutionFound)` (prm.rs:304)
    return OXG_OK;
  }
  // ---- reconstruct_path (prm.rs:189-208): the parent chain of the goal node, start side first
  std::vector<uint32_t> h_queue(level_end), h_parent(n);
  OXG_CUDA_TRY(cudaMemcpyAsync(h_queue.data(), B.queue.p, (size_t)level_end * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, st));
  OXG_CUDA_TRY(cudaMemcpyAsync(h_parent.data(), B.parent.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  OXG_CUDA_TRY(cudaEventRecord(s->ev1, st));
  OXG_CUDA_TRY(cudaStreamSynchronize(st));
  std::vector<uint32_t> chain;
  for (uint32_t v = h_queue[goal_rank]; v != OXG_INVALID_INDEX; v = h_parent[v]) chain.push_back(v);
  std::reverse(chain.begin(), chain.end());
  out->goal_node = h_queue[goal_rank];
  out->path_len = (uint32_t)chain.size();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, s->ev0, s->ev1);
  out->device_ms = ms;
  if (chain.size() > path_cap || !path) {
    set_error("the path has %zu roadmap nodes, capacity is %u", chain.size(), path_cap);
    return OXG_E_CAPACITY;
  }
  std::copy(chain.begin(), chain.end(), path);
  out->status = 0;
  return OXG_OK;
}

static int check_query_args(oxg_world w, oxg_store s, const double* start, const double* goal_target,
                            double resolution, oxg_prm_query_result* out) {
  if (!w || !s || !start || !goal_target || !out) {
    set_error("NULL argument");
    return OXG_E_NULL;
  }
  if (!(resolution > 0.0)) {
    set_error("motion resolution must be > 0 (got %g)", resolution);
    return OXG_E_RESOLUTION;
  }
  if (s->size > 0xFFFFFFF0ull) {
    set_error("roadmap too large");
    return OXG_E_ARG;
  }
  return OXG_OK;
}

}  // namespace oxg

using namespace oxg;

extern "C" {

int oxg_prm_solve(oxg_world w, oxg_store s, const uint64_t* lower_offsets, const uint32_t* lower_idx,
                  double connection_radius, double resolution, const double* start, const double* goal_target,
                  double goal_radius, uint32_t* path, uint32_t path_cap, oxg_prm_query_result* out) {
  OXG_API_BEGIN
  OXG_TRY(check_query_args(w, s, start, goal_target, resolution, out));
  memset(out, 0, sizeof(*out));
  if (s->size == 0) {
    out->status = 3;  // PlanningError::UnsampledStateSpace (prm.rs:239-241)
    return OXG_OK;
  }
  if (!lower_offsets) {
    set_error("lower_offsets is NULL");
    return OXG_E_NULL;
  }
  const uint64_t n = s->size, e = lower_offsets[n];
  if (e > 0 && !lower_idx) {
    set_error("lower_idx is NULL");
    return OXG_E_NULL;
  }
  QueryBufs B;
  if (e > 0) {
    DevBuf d_off, d_idx;
    OXG_TRY(d_off.ensure((n + 1) * sizeof(uint64_t)));
    int rc = d_idx.ensure(e * sizeof(uint32_t));
    if (rc == OXG_OK) rc = B.keys.ensure(2 * e * sizeof(unsigned long long));
    if (rc == OXG_OK) {
      cudaMemcpyAsync(d_off.p, lower_offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream);
      cudaMemcpyAsync(d_idx.p, lower_idx, e * sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream);
      query_keys_from_lower_kernel<<<grid_q(n, 128), 128, 0, s->stream>>>(
          d_off.as<uint64_t>(), d_idx.as<uint32_t>(), n, B.keys.as<unsigned long long>());
      g_launches.fetch_add(1, std::memory_order_relaxed);
      if (cudaStreamSynchronize(s->stream) != cudaSuccess) {
        set_error("adjacency upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = OXG_E_CUDA;
      }
    }
    d_off.release();
    d_idx.release();
    OXG_TRY(rc);
  }
  return prm_solve_impl(w, s, B, 2 * e, connection_radius, resolution, start, goal_target, goal_radius, path,
                        path_cap, out);
  OXG_API_END
}

int oxg_prm_solve_knn(oxg_world w, oxg_store s, const uint32_t* idx, const uint8_t* valid, uint32_t k,
                      double connection_radius, double resolution, const double* start,
                      const double* goal_target, double goal_radius, uint32_t* path, uint32_t path_cap,
                      oxg_prm_query_result* out) {
  OXG_API_BEGIN
  OXG_TRY(check_query_args(w, s, start, goal_target, resolution, out));
  memset(out, 0, sizeof(*out));
  if (s->size == 0) {
    out->status = 3;
    return OXG_OK;
  }
  if (!idx || !valid || k == 0) {
    set_error("NULL roadmap table");
    return OXG_E_NULL;
  }
  const uint64_t n = s->size, slots = n * k;
  QueryBufs B;
  {
    DevBuf d_idx, d_valid;
    int rc = d_idx.ensure(slots * sizeof(uint32_t));
    if (rc == OXG_OK) rc = d_valid.ensure(slots);
    if (rc == OXG_OK) rc = B.keys.ensure(2 * slots * sizeof(unsigned long long));
    if (rc == OXG_OK) {
      cudaMemcpyAsync(d_idx.p, idx, slots * sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream);
      cudaMemcpyAsync(d_valid.p, valid, slots, cudaMemcpyHostToDevice, s->stream);
      query_keys_from_knn_kernel<<<grid_q(slots, 256), 256, 0, s->stream>>>(
          d_idx.as<uint32_t>(), d_valid.as<uint8_t>(), n, k, B.keys.as<unsigned long long>());
      g_launches.fetch_add(1, std::memory_order_relaxed);
      if (cudaStreamSynchronize(s->stream) != cudaSuccess) {
        set_error("roadmap upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = OXG_E_CUDA;
      }
    }
    d_idx.release();
    d_valid.release();
    OXG_TRY(rc);
  }
  return prm_solve_impl(w, s, B, 2 * slots, connection_radius, resolution, start, goal_target, goal_radius, path,
                        path_cap, out);
  OXG_API_END
}

}  // extern "C"
