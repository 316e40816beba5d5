// oxg_api.cu — library plumbing: errors, device selection, the SoA state store and the
// declarative world.  See include/oxg.h for the contract of each entry point.
#include <cstdarg>
#include <cstdlib>
#include <vector>

#include "oxg_internal.cuh"
#include "oxg_math.cuh"

namespace oxg {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_tune_knn_tree{-1};
std::atomic<int> g_tune_tree_cache{1};
std::atomic<int> g_tune_tree_stats{0};
std::atomic<int> g_tune_tree_kind{-1};
std::atomic<int> g_tune_row_cull{1};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int validate_space(const oxg_space_desc* sp) {
  if (!sp) {
    set_error("space descriptor is NULL");
    return OXG_E_NULL;
  }
  uint32_t dim = sp->n_real + sp->n_so2;
  if (dim == 0 || dim > (uint32_t)kMaxDim || sp->n_so2 > 16) {
    set_error("unsupported space: n_real=%u n_so2=%u (need 1 <= dim <= %d, n_so2 <= 16)", sp->n_real,
              sp->n_so2, kMaxDim);
    return OXG_E_DIM;
  }
  return OXG_OK;
}

// max{x >= 0 : sqrt(x) <= r}.  Host mirror of sqrt_le_threshold_dev.
double sqrt_le_threshold(double r) {
  if (!(r == r)) return INFINITY;  // "sqrt(d2) > NaN" is always false -> every state hits
  if (r < 0.0) return -1.0;
  if (std::isinf(r)) return INFINITY;
  double x = r * r;
  if (std::isinf(x)) x = 1.7976931348623157e308;
  while (std::sqrt(x) > r && x > 0.0) x = std::nextafter(x, 0.0);
  for (;;) {
    double nx = std::nextafter(x, INFINITY);
    if (!std::isinf(nx) && std::sqrt(nx) <= r)
      x = nx;
    else
      break;
  }
  return x;
}

// max{x >= 0 : sqrt(x) < r}; -1 if there is none (r <= 0 or NaN).
double sqrt_lt_threshold(double r) {
  if (!(r == r) || r <= 0.0) return -1.0;
  if (std::isinf(r)) return 1.7976931348623157e308;
  double x = sqrt_le_threshold(r);  // sqrt(x) <= r
  while (x >= 0.0 && std::sqrt(x) >= r) {
    if (x == 0.0) return -1.0;
    x = std::nextafter(x, 0.0);
  }
  return x;
}

// Row-major [n][dim] -> SoA columns, appended at `first`.  Also maintains the f32 mirror used
// by the certified pre-filter and the running max |coordinate| (as f64 bits; |x| >= +0 so the
// bit patterns order like the values, and NaN sorts above +inf, which disables the filter).
__global__ void scatter_rows_kernel(const double* __restrict__ rows, uint64_t n, uint32_t dim,
                                    double* __restrict__ data, float* __restrict__ data32,
                                    uint64_t pitch, uint64_t first,
                                    unsigned long long* __restrict__ maxabs) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t total = n * dim;
  unsigned long long m = 0ull;
  for (; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t i = t / dim;
    uint32_t d = (uint32_t)(t - i * dim);
    double v = rows[t];
    data[(uint64_t)d * pitch + first + i] = v;
    data32[(uint64_t)d * pitch + first + i] = __double2float_rn(v);
    unsigned long long b = (unsigned long long)__double_as_longlong(fabs(v));
    m = b > m ? b : m;
  }
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long x = __shfl_down_sync(0xFFFFFFFFu, m, o);
    m = x > m ? x : m;
  }
  if ((threadIdx.x & 31) == 0 && m) atomicMax(maxabs, m);
}

__global__ void gather_rows_kernel(const double* __restrict__ data, uint64_t pitch, uint64_t first,
                                   uint64_t n, uint32_t dim, double* __restrict__ rows) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t total = n * dim;
  for (; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t i = t / dim;
    uint32_t d = (uint32_t)(t - i * dim);
    rows[t] = data[(uint64_t)d * pitch + first + i];
  }
}

// Squared norm of states [first, first+n): column `dim` of the f32 mirror (input of the
// dot-product form of the certified pre-filter) and the running maximum (maxabs[1], f64 bits).
__global__ void norms_kernel(const double* __restrict__ data, float* __restrict__ data32,
                             uint64_t pitch, uint64_t first, uint64_t n, uint32_t dim,
                             unsigned long long* __restrict__ maxabs) {
  unsigned long long m = 0ull;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = first + t;
    double s = 0.0;
    for (uint32_t d = 0; d < dim; ++d) {
      const double v = data[(uint64_t)d * pitch + i];
      s += v * v;
    }
    data32[(uint64_t)dim * pitch + i] = __double2float_rn(s);
    unsigned long long b = (unsigned long long)__double_as_longlong(s);
    m = b > m ? b : m;
  }
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long x = __shfl_down_sync(0xFFFFFFFFu, m, o);
    m = x > m ? x : m;
  }
  if ((threadIdx.x & 31) == 0 && m) atomicMax(maxabs + 1, m);
}

static int launch_norms(oxg_store_s* s, uint64_t first, uint64_t n) {
  if (n == 0) return OXG_OK;
  int threads = 256;
  int blocks = (int)std::min<uint64_t>((n + threads - 1) / threads, 148 * 16);
  norms_kernel<<<blocks, threads, 0, s->stream>>>(s->d_data, s->d_data32, s->pitch, first, n, s->dim,
                                                  s->d_maxabs);
  OXG_LAUNCH_CHECK();
  return OXG_OK;
}

// f32 mirror + max |coordinate| for states [first, first+n) written directly into the SoA
// columns by a device kernel (the fused planners append on the device).
__global__ void mirror_rows_kernel(const double* __restrict__ data, float* __restrict__ data32,
                                   uint64_t pitch, uint64_t first, uint64_t n, uint32_t dim,
                                   unsigned long long* __restrict__ maxabs) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t total = n * dim;
  unsigned long long m = 0ull;
  for (; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t d = (uint32_t)(t / n);
    uint64_t i = first + (t - (uint64_t)d * n);
    double v = data[(uint64_t)d * pitch + i];
    data32[(uint64_t)d * pitch + i] = __double2float_rn(v);
    unsigned long long b = (unsigned long long)__double_as_longlong(fabs(v));
    m = b > m ? b : m;
  }
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long x = __shfl_down_sync(0xFFFFFFFFu, m, o);
    m = x > m ? x : m;
  }
  if ((threadIdx.x & 31) == 0 && m) atomicMax(maxabs, m);
}

int store_refresh_mirror(oxg_store_s* s, uint64_t first, uint64_t n) {
  if (n == 0) return OXG_OK;
  uint64_t total = n * s->dim;
  int threads = 256;
  int blocks = (int)std::min<uint64_t>((total + threads - 1) / threads, 148 * 16);
  mirror_rows_kernel<<<blocks, threads, 0, s->stream>>>(s->d_data, s->d_data32, s->pitch, first, n,
                                                        s->dim, s->d_maxabs);
  OXG_LAUNCH_CHECK();
  return launch_norms(s, first, n);
}

static uint64_t round_up(uint64_t v, uint64_t m) { return (v + m - 1) / m * m; }

static int store_grow(oxg_store_s* s, uint64_t capacity) {
  if (capacity <= s->pitch) return OXG_OK;
  uint64_t new_pitch = round_up(std::max<uint64_t>(capacity, 1024), 32);
  double* nd = nullptr;
  cudaError_t e = cudaMalloc(&nd, sizeof(double) * new_pitch * s->dim);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaMalloc of %llu x %u f64 store failed: %s", (unsigned long long)new_pitch, s->dim,
              cudaGetErrorString(e));
    return OXG_E_ALLOC;
  }
  float* nd32 = nullptr;
  e = cudaMalloc(&nd32, sizeof(float) * new_pitch * (s->dim + 1));  // + the squared-norm column
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaFree(nd);
    set_error("cudaMalloc of %llu x %u f32 store mirror failed: %s", (unsigned long long)new_pitch,
              s->dim, cudaGetErrorString(e));
    return OXG_E_ALLOC;
  }
  if (s->d_data && s->size > 0) {
    for (uint32_t d = 0; d < s->dim; ++d)
      OXG_CUDA_TRY(cudaMemcpyAsync(nd + (uint64_t)d * new_pitch, s->d_data + (uint64_t)d * s->pitch,
                                   sizeof(double) * s->size, cudaMemcpyDeviceToDevice, s->stream));
    for (uint32_t d = 0; d <= s->dim; ++d)
      OXG_CUDA_TRY(cudaMemcpyAsync(nd32 + (uint64_t)d * new_pitch,
                                   s->d_data32 + (uint64_t)d * s->pitch, sizeof(float) * s->size,
                                   cudaMemcpyDeviceToDevice, s->stream));
    OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  }
  if (s->d_data) cudaFree(s->d_data);
  if (s->d_data32) cudaFree(s->d_data32);
  s->d_data = nd;
  s->d_data32 = nd32;
  s->pitch = new_pitch;
  return OXG_OK;
}

// Appends the n row-major states waiting in B_STAGE (device) to the SoA store and its f32 mirror.
int store_append_staged(oxg_store_s* s, uint64_t n) {
  if (n == 0) return OXG_OK;
  if (s->size + n > 0xFFFFFFF0ull) {
    set_error("store would exceed the u32 index range");
    return OXG_E_CAPACITY;
  }
  if (s->size + n > s->pitch) OXG_TRY(store_grow(s, std::max(s->size + n, s->pitch * 2)));
  uint64_t total = n * s->dim;
  int threads = 256;
  int blocks = (int)std::min<uint64_t>((total + threads - 1) / threads, 148 * 16);
  scatter_rows_kernel<<<blocks, threads, 0, s->stream>>>(s->buf[B_STAGE].as<double>(), n, s->dim,
                                                         s->d_data, s->d_data32, s->pitch, s->size,
                                                         s->d_maxabs);
  OXG_LAUNCH_CHECK();
  OXG_TRY(launch_norms(s, s->size, n));
  // the caller may reuse its host rows immediately
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  s->size += n;
  s->epoch += 1;
  return OXG_OK;
}

}  // namespace oxg

using namespace oxg;

extern "C" {

int oxg_version(void) { return OXG_VERSION; }

const char* oxg_last_error(void) { return g_err; }

int oxg_device_count(int* out_count) {
  if (!out_count) {
    set_error("out_count is NULL");
    return OXG_E_NULL;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *out_count = 0;
    set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    return OXG_E_CUDA;
  }
  *out_count = n;
  return OXG_OK;
}

int oxg_set_device(int device) {
  OXG_CUDA_TRY(cudaSetDevice(device));
  return OXG_OK;
}

uint64_t oxg_kernel_launches(int reset) {
  if (reset) return g_launches.exchange(0);
  return g_launches.load();
}

int oxg_tuning_set(const char* key, int64_t value) {
  if (!key) {
    set_error("key is NULL");
    return OXG_E_NULL;
  }
  if (strcmp(key, "knn_tree") == 0) {
    g_tune_knn_tree.store(value < 0 ? -1 : (value ? 1 : 0));
    return OXG_OK;
  }
  if (strcmp(key, "tree_cache") == 0) {
    g_tune_tree_cache.store(value ? 1 : 0);
    return OXG_OK;
  }
  if (strcmp(key, "tree_kind") == 0) {
    g_tune_tree_kind.store(value < 0 ? -1 : (value ? 1 : 0));
    return OXG_OK;
  }
  if (strcmp(key, "tree_stats") == 0) {
    g_tune_tree_stats.store(value <= 0 ? 0 : (value >= 2 ? 2 : 1));
    return OXG_OK;
  }
  if (strcmp(key, "row_cull") == 0) {
    g_tune_row_cull.store(value ? 1 : 0);
    return OXG_OK;
  }
  set_error("unknown tuning key '%s'", key);
  return OXG_E_ARG;
}

// ------------------------------------------------------------------ store
int oxg_store_create(const oxg_space_desc* space, uint64_t capacity, oxg_store* out) {
  OXG_API_BEGIN
  if (!out) {
    set_error("out is NULL");
    return OXG_E_NULL;
  }
  *out = nullptr;
  OXG_TRY(validate_space(space));
  if (capacity > 0xFFFFFFF0ull) {
    set_error("capacity %llu exceeds the u32 index range", (unsigned long long)capacity);
    return OXG_E_ARG;
  }
  int dev = 0;
  OXG_CUDA_TRY(cudaGetDevice(&dev));
  oxg_store_s* s = new oxg_store_s();
  s->space = *space;
  s->dim = space->n_real + space->n_so2;
  s->device = dev;
  cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&s->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&s->ev1);
  if (e == cudaSuccess) e = cudaEventCreate(&s->ev2);
  if (e != cudaSuccess) {
    set_error("stream/event creation failed: %s", cudaGetErrorString(e));
    delete s;
    return OXG_E_CUDA;
  }
  if (cudaMalloc((void**)&s->d_maxabs, 2 * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMemsetAsync(s->d_maxabs, 0, 2 * sizeof(unsigned long long), s->stream) != cudaSuccess) {
    cudaGetLastError();
    set_error("store allocation failed");
    oxg_store_destroy(s);
    return OXG_E_ALLOC;
  }
  int r = store_grow(s, capacity);
  if (r != OXG_OK) {
    oxg_store_destroy(s);
    return r;
  }
  *out = s;
  return OXG_OK;
  OXG_API_END
}

int oxg_store_destroy(oxg_store s) {
  if (!s) return OXG_OK;
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->d_data) cudaFree(s->d_data);
  if (s->d_data32) cudaFree(s->d_data32);
  if (s->d_maxabs) cudaFree(s->d_maxabs);
  for (auto& b : s->buf) b.release();
  tree_release(s);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->ev2) cudaEventDestroy(s->ev2);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return OXG_OK;
}

int oxg_store_clear(oxg_store s) {
  if (!s) {
    set_error("store is NULL");
    return OXG_E_NULL;
  }
  s->size = 0;
  s->n_planned = 0;
  s->epoch += 1;
  OXG_CUDA_TRY(cudaMemsetAsync(s->d_maxabs, 0, 2 * sizeof(unsigned long long), s->stream));
  return OXG_OK;
}

int oxg_store_reserve(oxg_store s, uint64_t capacity) {
  OXG_API_BEGIN
  if (!s) {
    set_error("store is NULL");
    return OXG_E_NULL;
  }
  if (capacity > 0xFFFFFFF0ull) {
    set_error("capacity %llu exceeds the u32 index range", (unsigned long long)capacity);
    return OXG_E_ARG;
  }
  return store_grow(s, capacity);
  OXG_API_END
}

int oxg_store_append(oxg_store s, const double* rows, uint64_t n, uint64_t* first_index) {
  OXG_API_BEGIN
  if (!s || (!rows && n > 0)) {
    set_error("store or rows is NULL");
    return OXG_E_NULL;
  }
  if (first_index) *first_index = s->size;
  if (n == 0) return OXG_OK;
  if (s->size + n > 0xFFFFFFF0ull) {
    set_error("store would exceed the u32 index range");
    return OXG_E_CAPACITY;
  }
  size_t bytes = sizeof(double) * n * s->dim;
  OXG_TRY(s->buf[B_STAGE].ensure(bytes));
  OXG_CUDA_TRY(cudaMemcpyAsync(s->buf[B_STAGE].p, rows, bytes, cudaMemcpyHostToDevice, s->stream));
  return store_append_staged(s, n);
  OXG_API_END
}

int oxg_store_size(oxg_store s, uint64_t* out_n) {
  if (!s || !out_n) {
    set_error("store or out_n is NULL");
    return OXG_E_NULL;
  }
  *out_n = s->size;
  return OXG_OK;
}

int oxg_store_read(oxg_store s, uint64_t first, uint64_t n, double* rows_out) {
  OXG_API_BEGIN
  if (!s || (!rows_out && n > 0)) {
    set_error("store or rows_out is NULL");
    return OXG_E_NULL;
  }
  if (first + n > s->size) {
    set_error("read range [%llu, %llu) exceeds store size %llu", (unsigned long long)first,
              (unsigned long long)(first + n), (unsigned long long)s->size);
    return OXG_E_ARG;
  }
  if (n == 0) return OXG_OK;
  size_t bytes = sizeof(double) * n * s->dim;
  OXG_TRY(s->buf[B_STAGE].ensure(bytes));
  uint64_t total = n * s->dim;
  int threads = 256;
  int blocks = (int)std::min<uint64_t>((total + threads - 1) / threads, 148 * 16);
  gather_rows_kernel<<<blocks, threads, 0, s->stream>>>(s->d_data, s->pitch, first, n, s->dim,
                                                        s->buf[B_STAGE].as<double>());
  OXG_LAUNCH_CHECK();
  OXG_CUDA_TRY(cudaMemcpyAsync(rows_out, s->buf[B_STAGE].p, bytes, cudaMemcpyDeviceToHost, s->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  return OXG_OK;
  OXG_API_END
}

int oxg_store_sync(oxg_store s) {
  if (!s) {
    set_error("store is NULL");
    return OXG_E_NULL;
  }
  OXG_CUDA_TRY(cudaStreamSynchronize(s->stream));
  return OXG_OK;
}

int oxg_store_stream(oxg_store s, void** out_stream) {
  if (!s || !out_stream) {
    set_error("store or out_stream is NULL");
    return OXG_E_NULL;
  }
  *out_stream = (void*)s->stream;
  return OXG_OK;
}

// ------------------------------------------------------------------ world
static int upload32(float** dst, const std::vector<float>& h, cudaStream_t st) {
  *dst = nullptr;
  if (h.empty()) return OXG_OK;
  OXG_CUDA_TRY(cudaMalloc((void**)dst, sizeof(float) * h.size()));
  OXG_CUDA_TRY(cudaMemcpyAsync(*dst, h.data(), sizeof(float) * h.size(), cudaMemcpyHostToDevice, st));
  return OXG_OK;
}

// float >= v (round up) / float <= v (round down)
static float f32_up(double v) {
  float f = (float)v;
  if ((double)f < v) f = std::nextafterf(f, INFINITY);
  return f;
}
static float f32_down(double v) {
  float f = (float)v;
  if ((double)f > v) f = std::nextafterf(f, -INFINITY);
  return f;
}

// OXG_WORLD_GRID=0 switches the planners' obstacle grid off (A/B measurements; same results).
static bool env_grid_enabled() {
  const char* v = getenv("OXG_WORLD_GRID");
  return !(v && v[0] == '0');
}

static int upload_u32(uint32_t** dst, const std::vector<uint32_t>& h, cudaStream_t st) {
  *dst = nullptr;
  if (h.empty()) return OXG_OK;
  OXG_CUDA_TRY(cudaMalloc((void**)dst, sizeof(uint32_t) * h.size()));
  OXG_CUDA_TRY(cudaMemcpyAsync(*dst, h.data(), sizeof(uint32_t) * h.size(), cudaMemcpyHostToDevice, st));
  return OXG_OK;
}

static int upload(double** dst, const std::vector<double>& h, cudaStream_t st) {
  *dst = nullptr;
  if (h.empty()) return OXG_OK;
  OXG_CUDA_TRY(cudaMalloc((void**)dst, sizeof(double) * h.size()));
  OXG_CUDA_TRY(cudaMemcpyAsync(*dst, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, st));
  return OXG_OK;
}

int oxg_world_create(uint32_t dim, const double* spheres, uint64_t n_spheres, const double* aabbs,
                     uint64_t n_aabbs, const double* capsules, uint64_t n_capsules, oxg_world* out) {
  OXG_API_BEGIN
  if (!out) {
    set_error("out is NULL");
    return OXG_E_NULL;
  }
  *out = nullptr;
  if ((n_spheres && !spheres) || (n_aabbs && !aabbs) || (n_capsules && !capsules)) {
    set_error("obstacle array is NULL with a non-zero count");
    return OXG_E_NULL;
  }
  if (dim > (uint32_t)kMaxDim || (dim == 0 && (n_spheres || n_aabbs || n_capsules))) {
    set_error("unsupported world dimension %u", dim);
    return OXG_E_DIM;
  }
  if (n_spheres > 0x7FFFFFFFull || n_aabbs > 0x7FFFFFFFull || n_capsules > 0x7FFFFFFFull) {
    set_error("too many obstacles");
    return OXG_E_ARG;
  }
  int dev = 0;
  OXG_CUDA_TRY(cudaGetDevice(&dev));
  oxg_world_s* w = new oxg_world_s();
  w->dim = dim;
  w->device = dev;
  w->n_spheres = n_spheres;
  w->n_aabbs = n_aabbs;
  w->n_capsules = n_capsules;
  w->sphere_stride = (dim + 1 + 1) & ~1u;
  w->aabb_stride = 2 * dim;
  w->capsule_stride = (2 * dim + 2 + 1) & ~1u;
  cudaError_t e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
    delete w;
    return OXG_E_CUDA;
  }
  std::vector<double> hs(n_spheres * w->sphere_stride, 0.0);
  for (uint64_t i = 0; i < n_spheres; ++i) {
    const double* src = spheres + i * (dim + 1);
    double* dst = hs.data() + i * w->sphere_stride;
    for (uint32_t d = 0; d < dim; ++d) dst[d] = src[d];
    dst[dim] = sqrt_le_threshold(src[dim]);
  }
  std::vector<double> ha(aabbs, aabbs + n_aabbs * 2 * dim);
  std::vector<double> hc(n_capsules * w->capsule_stride, 0.0);
  for (uint64_t i = 0; i < n_capsules; ++i) {
    const double* a = capsules + i * (2 * dim + 1);
    const double* b = a + dim;
    double* dst = hc.data() + i * w->capsule_stride;
    double den = 0.0;
    for (uint32_t d = 0; d < dim; ++d) {
      double ab = b[d] - a[d];
      double m = ab * ab;
      den = (d == 0) ? m : den + m;
      dst[d] = a[d];
      dst[dim + d] = ab;
    }
    dst[2 * dim] = den;
    dst[2 * dim + 1] = sqrt_le_threshold(a[2 * dim]);
  }
  // ---- f32 mirrors for the certified pre-filter (see oxg_math.cuh, *_maybe_hit32)
  w->sphere32_stride = (dim + 2 + 3) & ~3u;
  w->capsule32_stride = (2 * dim + 3 + 3) & ~3u;
  auto u1_of = [](double U) { return U >= 0.0 ? f32_up(U * 1.0001) : 0.0f; };
  auto sq_of = [&](double U) {
    return U >= 0.0 ? f32_up(std::sqrt((double)dim * U * 1.0001) * 1.0001) : 0.0f;
  };
  std::vector<float> hs32(n_spheres * w->sphere32_stride, 0.0f);
  for (uint64_t i = 0; i < n_spheres; ++i) {
    const double* src = hs.data() + i * w->sphere_stride;
    float* dst = hs32.data() + i * w->sphere32_stride;
    for (uint32_t d = 0; d < dim; ++d) {
      dst[d] = (float)src[d];
      w->wmax_spheres = std::fmax(w->wmax_spheres, std::fabs(src[d]));
      if (!(src[d] == src[d])) w->wmax_spheres = INFINITY;
    }
    dst[dim] = u1_of(src[dim]);
    dst[dim + 1] = sq_of(src[dim]);
  }
  std::vector<float> ha32(ha.size());
  for (uint64_t i = 0; i < n_aabbs; ++i)
    for (uint32_t d = 0; d < dim; ++d) {
      ha32[i * 2 * dim + d] = f32_down(ha[i * 2 * dim + d]);
      ha32[i * 2 * dim + dim + d] = f32_up(ha[i * 2 * dim + dim + d]);
    }
  std::vector<float> hc32(n_capsules * w->capsule32_stride, 0.0f);
  for (uint64_t i = 0; i < n_capsules; ++i) {
    const double* src = hc.data() + i * w->capsule_stride;
    float* dst = hc32.data() + i * w->capsule32_stride;
    for (uint32_t d = 0; d < dim; ++d) {
      dst[d] = (float)src[d];
      dst[dim + d] = (float)src[dim + d];
      const double bd = capsules[i * (2 * dim + 1) + dim + d];
      const double m = std::fmax(std::fmax(std::fabs(src[d]), std::fabs(src[dim + d])), std::fabs(bd));
      w->wmax_capsules = std::fmax(w->wmax_capsules, m);
      if (!(m == m)) w->wmax_capsules = INFINITY;
    }
    const double den = src[2 * dim];
    dst[2 * dim] = den > 0.0 ? (float)(1.0 / den) : 0.0f;
    dst[2 * dim + 1] = u1_of(src[2 * dim + 1]);
    dst[2 * dim + 2] = sq_of(src[2 * dim + 1]);
  }
  // capsule pairs (c, c+1) with interleaved values; an odd tail pairs with a dummy whose
  // threshold is -inf (never "maybe hit")
  w->capsule32p_stride = (2 * (2 * dim + 3) + 3) & ~3u;
  const uint64_t n_pairs = (n_capsules + 1) / 2;
  std::vector<float> hc32p(n_pairs * w->capsule32p_stride, 0.0f);
  for (uint64_t pr = 0; pr < n_pairs; ++pr) {
    float* dst = hc32p.data() + pr * w->capsule32p_stride;
    for (int h = 0; h < 2; ++h) {
      const uint64_t c = 2 * pr + h;
      for (uint32_t v = 0; v < 2 * dim + 3; ++v) {
        float val = 0.0f;
        if (c < n_capsules) val = hc32[c * w->capsule32_stride + v];
        else if (v == 2 * dim + 1) val = -INFINITY;
        dst[2 * v + h] = val;
      }
    }
  }
  // ---- uniform grid of the obstacles' bounding boxes over the first min(dim, 3) coordinates
  // (used by the fused planner kernels: a state only meets the obstacles listed in its cell)
  std::vector<uint32_t> g_off, g_ids;
  {
    const uint64_t n_obst = n_spheres + n_aabbs + n_capsules;
    const uint32_t gd = std::min<uint32_t>(dim, 3);
    bool ok = gd >= 1 && n_obst >= 64 && n_obst < 0x7FFFFFFFull && env_grid_enabled();
    std::vector<double> bb(n_obst * 2 * gd);  // [lo.., hi..] per obstacle
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint64_t o = 0; o < n_obst && ok; ++o) {
      for (uint32_t d = 0; d < gd; ++d) {
        double a, b;
        if (o < n_spheres) {
          const double* s0 = spheres + o * (dim + 1);
          const double r = std::fabs(s0[dim]) * (1.0 + 1e-9);
          a = s0[d] - r;
          b = s0[d] + r;
        } else if (o < n_spheres + n_aabbs) {
          const double* b0 = aabbs + (o - n_spheres) * 2 * dim;
          a = b0[d];
          b = b0[dim + d];
        } else {
          const double* c0 = capsules + (o - n_spheres - n_aabbs) * (2 * dim + 1);
          const double r = std::fabs(c0[2 * dim]) * (1.0 + 1e-9);
          a = std::fmin(c0[d], c0[dim + d]) - r;
          b = std::fmax(c0[d], c0[dim + d]) + r;
        }
        if (!std::isfinite(a) || !std::isfinite(b)) ok = false;
        if (a > b) std::swap(a, b);  // an inverted box hits nothing; any cell range is fine
        bb[o * 2 * gd + d] = a;
        bb[o * 2 * gd + gd + d] = b;
        lo[d] = std::fmin(lo[d], a);
        hi[d] = std::fmax(hi[d], b);
      }
    }
    if (ok) {
      const double per_dim = std::pow((double)n_obst, 1.0 / gd) * 1.5;
      const uint32_t cap = gd == 1 ? 4096 : (gd == 2 ? 96 : 24);
      uint64_t cells = 1;
      for (uint32_t d = 0; d < gd; ++d) {
        const double ext = hi[d] - lo[d];
        uint32_t nc = (uint32_t)std::min<double>(cap, std::max(2.0, std::ceil(per_dim)));
        if (!(ext > 0.0)) nc = 1;
        w->grid_n[d] = nc;
        w->grid_lo[d] = lo[d];
        w->grid_inv[d] = ext > 0.0 ? (double)nc / ext : 0.0;
        if (!std::isfinite(w->grid_inv[d])) ok = false;
        cells *= nc;
      }
      if (ok) {
        auto cell_of = [&](double x, uint32_t d) -> int64_t {  // the device uses the same formula
          double c = std::floor((x - w->grid_lo[d]) * w->grid_inv[d]);
          if (!(c == c)) c = 0.0;
          c = std::fmin(std::fmax(c, 0.0), (double)(w->grid_n[d] - 1));
          return (int64_t)c;
        };
        std::vector<uint32_t> count(cells + 1, 0);
        auto for_cells = [&](uint64_t o, auto&& fn) {
          int64_t a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
          for (uint32_t d = 0; d < gd; ++d) {
            a[d] = std::max<int64_t>(0, cell_of(bb[o * 2 * gd + d], d) - 1);
            b[d] = std::min<int64_t>((int64_t)w->grid_n[d] - 1, cell_of(bb[o * 2 * gd + gd + d], d) + 1);
          }
          for (int64_t z = a[2]; z <= b[2]; ++z)
            for (int64_t y = a[1]; y <= b[1]; ++y)
              for (int64_t x = a[0]; x <= b[0]; ++x)
                fn((uint64_t)((z * w->grid_n[1] + y) * w->grid_n[0] + x));
        };
        for (uint64_t o = 0; o < n_obst; ++o) for_cells(o, [&](uint64_t c) { count[c + 1]++; });
        uint64_t total = 0;
        for (uint64_t c = 0; c < cells; ++c) total += count[c + 1];
        // worth it only when a cell's list fits one warp round: long lists (many obstacles in
        // more dimensions than the grid has) are served better by the chunk sweep, which keeps 32
        // obstacles per warp resident and streams the states past them
        if (total < 0x7FFFFFFFull && total <= 24 * cells) {
          g_off.assign(cells + 1, 0);
          for (uint64_t c = 0; c < cells; ++c) g_off[c + 1] = g_off[c] + count[c + 1];
          g_ids.assign(std::max<uint64_t>(total, 1), 0);
          std::vector<uint32_t> fill(g_off.begin(), g_off.end() - 1);
          for (uint64_t o = 0; o < n_obst; ++o) for_cells(o, [&](uint64_t c) { g_ids[fill[c]++] = (uint32_t)o; });
          w->grid_dims = gd;
        }
      }
    }
  }
  int r = upload(&w->d_spheres, hs, w->stream);
  if (r == OXG_OK) r = upload(&w->d_aabbs, ha, w->stream);
  if (r == OXG_OK) r = upload(&w->d_capsules, hc, w->stream);
  if (r == OXG_OK) r = upload32(&w->d_spheres32, hs32, w->stream);
  if (r == OXG_OK) r = upload32(&w->d_aabbs32, ha32, w->stream);
  if (r == OXG_OK) r = upload32(&w->d_capsules32, hc32, w->stream);
  if (r == OXG_OK) r = upload32(&w->d_capsules32p, hc32p, w->stream);
  if (r == OXG_OK && w->grid_dims > 0) {
    r = upload_u32(&w->d_grid_off, g_off, w->stream);
    if (r == OXG_OK) r = upload_u32(&w->d_grid_ids, g_ids, w->stream);
  }
  if (r == OXG_OK && cudaStreamSynchronize(w->stream) != cudaSuccess) {
    set_error("world upload failed");
    r = OXG_E_CUDA;
  }
  if (r != OXG_OK) {
    oxg_world_destroy(w);
    return r;
  }
  *out = w;
  return OXG_OK;
  OXG_API_END
}

int oxg_world_add_so2_interval(oxg_world w, uint32_t so2_component, double lo, double hi) {
  OXG_API_BEGIN
  if (!w) {
    set_error("world is NULL");
    return OXG_E_NULL;
  }
  if (w->n_so2 >= 16 || so2_component >= 16) {
    set_error("at most 16 SO2 intervals / components are supported");
    return OXG_E_ARG;
  }
  w->h_so2_comp[w->n_so2] = so2_component;
  w->h_so2_lohi[2 * w->n_so2] = lo;
  w->h_so2_lohi[2 * w->n_so2 + 1] = hi;
  w->n_so2 += 1;
  if (!w->d_so2_comp) {
    OXG_CUDA_TRY(cudaMalloc((void**)&w->d_so2_comp, sizeof(uint32_t) * 16));
    OXG_CUDA_TRY(cudaMalloc((void**)&w->d_so2_lohi, sizeof(double) * 32));
  }
  OXG_CUDA_TRY(cudaMemcpyAsync(w->d_so2_comp, w->h_so2_comp, sizeof(uint32_t) * 16,
                               cudaMemcpyHostToDevice, w->stream));
  OXG_CUDA_TRY(cudaMemcpyAsync(w->d_so2_lohi, w->h_so2_lohi, sizeof(double) * 32,
                               cudaMemcpyHostToDevice, w->stream));
  OXG_CUDA_TRY(cudaStreamSynchronize(w->stream));
  return OXG_OK;
  OXG_API_END
}

int oxg_world_destroy(oxg_world w) {
  if (!w) return OXG_OK;
  if (w->stream) cudaStreamSynchronize(w->stream);
  if (w->d_spheres) cudaFree(w->d_spheres);
  if (w->d_aabbs) cudaFree(w->d_aabbs);
  if (w->d_capsules) cudaFree(w->d_capsules);
  if (w->d_spheres32) cudaFree(w->d_spheres32);
  if (w->d_aabbs32) cudaFree(w->d_aabbs32);
  if (w->d_capsules32) cudaFree(w->d_capsules32);
  if (w->d_capsules32p) cudaFree(w->d_capsules32p);
  if (w->d_so2_comp) cudaFree(w->d_so2_comp);
  if (w->d_so2_lohi) cudaFree(w->d_so2_lohi);
  if (w->d_grid_off) cudaFree(w->d_grid_off);
  if (w->d_grid_ids) cudaFree(w->d_grid_ids);
  for (auto& b : w->buf) b.release();
  if (w->stream) cudaStreamDestroy(w->stream);
  delete w;
  return OXG_OK;
}

int oxg_world_sync(oxg_world w) {
  if (!w) {
    set_error("world is NULL");
    return OXG_E_NULL;
  }
  OXG_CUDA_TRY(cudaStreamSynchronize(w->stream));
  return OXG_OK;
}

}  // extern "C"
