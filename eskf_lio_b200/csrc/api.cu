// api.cu — context, device clouds and the registration entry points of
// include/eskf_gpu.h.  No CPU fallback: without a CUDA device every call
// fails loudly with ESKF_ERR_NO_DEVICE.
#include <cstdarg>

#include <cstdlib>

#include "internal.h"

namespace eskf {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

int voxelize_max_blocks(int sm_count, int* out);
int voxelize_cluster_max(int* out);
int align_max_blocks(int sm_count, int* out);
int align_sharded(eskf_ctx* ctx, const AlignArgs& a, eskf_allreduce_fn allreduce, void* user,
                  double T_out[16], eskf_align_info* info);

int ctx_pinned(eskf_ctx* ctx, size_t bytes, void** out) {
  if (bytes > ctx->pinned_bytes) {
    if (ctx->pinned) {
      cudaStreamSynchronize(ctx->stream);
      cudaFreeHost(ctx->pinned);
    }
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
    const size_t want = bytes + bytes / 2 + 4096;
    ESKF_CUDA(cudaHostAlloc(&ctx->pinned, want, cudaHostAllocDefault));
    ctx->pinned_bytes = want;
  }
  *out = ctx->pinned;
  return ESKF_OK;
}

int wait_mail(eskf_ctx* ctx, const volatile unsigned* word, unsigned seq) {
  // (acquire loads: the payload words the device wrote before the sequence word must not be read ahead
  // of it on weakly ordered hosts -- Grace / aarch64; a plain load on x86)
  for (unsigned spins = 1;; ++spins) {
    if (__atomic_load_n(const_cast<const unsigned*>(word), __ATOMIC_ACQUIRE) == seq) return ESKF_OK;
    if ((spins & 2047u) == 0u) {  // every few microseconds: is the stream still busy?
      const cudaError_t q = cudaStreamQuery(ctx->stream);
      if (q == cudaSuccess) return __atomic_load_n(const_cast<const unsigned*>(word), __ATOMIC_ACQUIRE) == seq ? ESKF_OK : 1;
      if (q != cudaErrorNotReady) {
        set_error("stream error while waiting for a kernel result: %s", cudaGetErrorString(q));
        return ESKF_ERR_CUDA;
      }
    }
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
}

namespace {

// AoS (host layout) <-> SoA (device layout)
__global__ void aos_to_soa_kernel(const double* __restrict__ xyz, const double* __restrict__ cov,
                                  unsigned n, size_t pitch, double* x, double* y, double* z,
                                  double* c, float4* c4, float2* c2) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  x[i] = xyz[3 * static_cast<size_t>(i)];
  y[i] = xyz[3 * static_cast<size_t>(i) + 1];
  z[i] = xyz[3 * static_cast<size_t>(i) + 2];
  if (cov) {
    double v[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      v[k] = cov[9 * static_cast<size_t>(i) + k];
      c[k * pitch + i] = v[k];
    }
    c4[i] = make_float4(static_cast<float>(v[0]), static_cast<float>(v[1]), static_cast<float>(v[2]),
                        static_cast<float>(v[4]));
    c2[i] = make_float2(static_cast<float>(v[5]), static_cast<float>(v[8]));
  }
}

__global__ void f32_to_soa_kernel(const float* __restrict__ xyz, unsigned n, double* x, double* y,
                                  double* z) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  x[i] = static_cast<double>(xyz[3 * static_cast<size_t>(i)]);
  y[i] = static_cast<double>(xyz[3 * static_cast<size_t>(i) + 1]);
  z[i] = static_cast<double>(xyz[3 * static_cast<size_t>(i) + 2]);
}

__global__ void soa_to_aos_kernel(const double* x, const double* y, const double* z,
                                  const double* c, size_t pitch, unsigned n, double* xyz,
                                  double* cov) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (xyz) {
    xyz[3 * static_cast<size_t>(i)] = x[i];
    xyz[3 * static_cast<size_t>(i) + 1] = y[i];
    xyz[3 * static_cast<size_t>(i) + 2] = z[i];
  }
  if (cov)
#pragma unroll
    for (int k = 0; k < 9; ++k) cov[9 * static_cast<size_t>(i) + k] = c[k * pitch + i];
}

__global__ void build_c32_kernel(const double* c, size_t pitch, unsigned n, float4* c4, float2* c2) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  c4[i] = make_float4(static_cast<float>(c[i]), static_cast<float>(c[pitch + i]),
                      static_cast<float>(c[2 * pitch + i]), static_cast<float>(c[4 * pitch + i]));
  c2[i] = make_float2(static_cast<float>(c[5 * pitch + i]), static_cast<float>(c[8 * pitch + i]));
}

__global__ void transform_cloud_kernel(double* x, double* y, double* z, double* c, size_t pitch,
                                       unsigned n, const double* Tin) {
  __shared__ double T[12];
  if (threadIdx.x < 12) T[threadIdx.x] = Tin[threadIdx.x];
  __syncthreads();
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double px = x[i], py = y[i], pz = z[i];
  transform_point_rn(T, px, py, pz);
  x[i] = px;
  y[i] = py;
  z[i] = pz;
  if (c) {
    double C[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) C[k] = c[k * pitch + i];
    rotate_cov_rn(T, C);
#pragma unroll
    for (int k = 0; k < 9; ++k) c[k * pitch + i] = C[k];
  }
}

void free_cloud_buffers(eskf_cloud* c) {
  if (c->xyz) cudaFree(c->xyz);
  if (c->cov) cudaFree(c->cov);
  if (c->c4) cudaFree(c->c4);
  if (c->c2) cudaFree(c->c2);
  if (c->src) cudaFree(c->src);
  c->xyz = c->cov = nullptr;
  c->c4 = nullptr;
  c->c2 = nullptr;
  c->src = nullptr;
  c->cap = 0;
}

}  // namespace

// (re)allocate for `cap` points; contents are NOT preserved
int cloud_reserve(eskf_cloud* c, size_t cap, bool with_cov) {
  const bool have_cov = c->cov != nullptr;
  if (cap <= c->cap && (!with_cov || have_cov)) return ESKF_OK;
  size_t want = cap > c->cap ? cap + cap / 8 + 64 : c->cap;
  want = (want + 63) / 64 * 64;
  cudaStreamSynchronize(c->ctx->stream);
  free_cloud_buffers(c);
  ESKF_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->xyz), want * 3 * sizeof(double)));
  ESKF_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->src), want * sizeof(uint32_t)));
  if (with_cov || have_cov) {
    ESKF_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->cov), want * 9 * sizeof(double)));
    ESKF_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->c4), want * sizeof(float4)));
    ESKF_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->c2), want * sizeof(float2)));
  }
  c->cap = want;
  c->n = 0;
  c->has_cov = c->has_c32 = c->has_src = false;
  return ESKF_OK;
}

int cloud_build_c32(eskf_cloud* c) {
  if (c->has_c32 || c->n == 0) return ESKF_OK;
  ESKF_REQUIRE(c->has_cov, "cloud has no covariances");
  const unsigned n = static_cast<unsigned>(c->n);
  build_c32_kernel<<<(n + 255) / 256, 256, 0, c->ctx->stream>>>(c->cov, c->cap, n, c->c4, c->c2);
  ESKF_CUDA(cudaGetLastError());
  count_launch(c->ctx);
  c->has_c32 = true;
  return ESKF_OK;
}

}  // namespace eskf

using namespace eskf;

extern "C" {

int eskf_abi_version(void) { return ESKF_GPU_ABI_VERSION; }

const char* eskf_last_error(void) { return g_last_error.c_str(); }

int eskf_device_count(int* n) {
  ESKF_REQUIRE(n, "null n");
  *n = 0;
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess || *n <= 0) {
    *n = 0;
    set_error("no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    return ESKF_ERR_NO_DEVICE;
  }
  return ESKF_OK;
}

int eskf_host_alloc(size_t bytes, void** out) {
  ESKF_REQUIRE(out, "null out");
  ESKF_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return ESKF_OK;
}

int eskf_host_free(void* p) {
  if (p) ESKF_CUDA(cudaFreeHost(p));
  return ESKF_OK;
}

int eskf_ctx_create(int device, void* cuda_stream, eskf_ctx** out) {
  ESKF_REQUIRE(out, "null out");
  *out = nullptr;
  int n = 0;
  ESKF_TRY(eskf_device_count(&n));
  ESKF_REQUIRE(device >= 0 && device < n, "device index out of range");
  ESKF_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  ESKF_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
              prop.major, prop.minor);
    return ESKF_ERR_NO_DEVICE;
  }
  eskf_ctx* ctx = new eskf_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  // option defaults may come from the environment (A/B runs without touching the caller)
  if (const char* e = getenv("ESKF_ALIGN_DYNAMIC")) ctx->opt_align_dynamic = atoi(e) != 0;
  if (const char* e = getenv("ESKF_L2_PERSIST")) ctx->opt_l2_persist = atoi(e) != 0;
  if (const char* e = getenv("ESKF_TRACE")) ctx->opt_trace = atoi(e) != 0;
  if (const char* e = getenv("ESKF_ALIGN_AUTOTUNE")) ctx->opt_align_autotune = atoi(e) != 0;
  if (const char* e = getenv("ESKF_VOX_CLUSTER")) {
    const int v = atoi(e);
    if (v == 0 || v == 1 || v == 8 || v == 16) ctx->opt_vox_cluster = v;
  }
  if (const char* e = getenv("ESKF_MAPPED_RESULTS")) ctx->opt_mapped_results = atoi(e) != 0;
  if (const char* e = getenv("ESKF_ALIGN_BLOCK")) {
    const int v = atoi(e);
    if (v == 0 || v == 256 || v == 257 || v == 384 || v == 448 || v == 512 || v == 640 || v == 768 || v == 769) ctx->opt_align_block = v;
  }
  if (const char* e = getenv("ESKF_ALIGN_DEPTH")) {
    const int v = atoi(e);
    if (v == 0 || (v >= 3 && v <= 11)) ctx->opt_align_depth = v;
  }
  if (const char* e = getenv("ESKF_ALIGN_RESIDENT")) ctx->opt_align_resident = atoi(e);
  if (const char* e = getenv("ESKF_ALIGN_FAT_POINTS")) ctx->opt_align_fat_points = atoll(e);
  if (const char* e = getenv("ESKF_ALIGN_LL")) ctx->opt_align_ll = atoi(e) != 0;
  if (const char* e = getenv("ESKF_ALIGN_XCHG_LL")) ctx->opt_align_xchg_ll = atoi(e) != 0;
  if (const char* e = getenv("ESKF_ALIGN_FLAGS")) ctx->opt_align_flags = atoi(e);
  if (const char* e = getenv("ESKF_ALIGN_FILTER")) ctx->opt_align_filter = atoi(e) != 0;
  if (const char* e = getenv("ESKF_ALIGN_STAMPS")) ctx->opt_align_stamps = atoi(e) != 0;
  if (const char* e = getenv("ESKF_L2_CARVEOUT")) ctx->opt_l2_carveout = atoi(e) != 0;
  if (const char* e = getenv("ESKF_ALIGN_CHUNK")) {
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4) ctx->opt_align_chunk = v;
  }
  if (cuda_stream) {
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    ctx->own_stream = false;
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      set_error("cudaStreamCreate: %s", cudaGetErrorString(e));
      delete ctx;
      return ESKF_ERR_CUDA;
    }
    ctx->own_stream = true;
  }
  // persisting-L2 carve-out for the map's tag array (registration.cu)
  if (!ctx->opt_l2_carveout) {
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
  } else if (prop.persistingL2CacheMaxSize > 0 &&
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, prop.persistingL2CacheMaxSize) == cudaSuccess) {
    ctx->l2_persist_bytes = static_cast<size_t>(prop.persistingL2CacheMaxSize);
    ctx->l2_window_max = static_cast<size_t>(prop.accessPolicyMaxWindowSize);
  }
  cudaGetLastError();
  int st = voxelize_max_blocks(ctx->sm_count, &ctx->max_blocks_voxelize);
  if (st == ESKF_OK) st = voxelize_cluster_max(&ctx->vox_cluster_max);
  if (st == ESKF_OK) st = align_max_blocks(ctx->sm_count, &ctx->max_blocks_align);
  if (st == ESKF_OK) {
    void* hp = nullptr;
    void* dp = nullptr;
    if (cudaHostAlloc(&hp, sizeof(HostMail), cudaHostAllocMapped) == cudaSuccess &&
        cudaHostGetDevicePointer(&dp, hp, 0) == cudaSuccess) {
      std::memset(hp, 0, sizeof(HostMail));
      ctx->mail_h = static_cast<HostMail*>(hp);
      ctx->mail_d = static_cast<HostMail*>(dp);
    } else {
      if (hp) cudaFreeHost(hp);
      cudaGetLastError();
      ctx->opt_mapped_results = 0;  // no zero-copy memory: results come back by copy
    }
  }
  if (st == ESKF_OK && cudaEventCreate(&ctx->ev0) != cudaSuccess) st = ESKF_ERR_CUDA;
  if (st == ESKF_OK && cudaEventCreate(&ctx->ev1) != cudaSuccess) st = ESKF_ERR_CUDA;
  if (st != ESKF_OK) {
    eskf_ctx_destroy(ctx);
    return st;
  }
  *out = ctx;
  return ESKF_OK;
}

int eskf_ctx_destroy(eskf_ctx* ctx) {
  if (!ctx) return ESKF_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto& c : ctx->tmp_cloud) {
    if (c) eskf_cloud_destroy(c);
    c = nullptr;
  }
  if (ctx->crop_cloud) eskf_cloud_destroy(ctx->crop_cloud);
  ctx->crop_cloud = nullptr;
  eskf::DevBuf* bufs[] = {&ctx->stage, &ctx->sortbuf, &ctx->hist, &ctx->hdr, &ctx->runs,
                          &ctx->sorted_xyz, &ctx->segs, &ctx->work, &ctx->spill, &ctx->partials, &ctx->astate,
                          &ctx->misc, &ctx->knn_levels, &ctx->knn_nbr, &ctx->link, &ctx->crop_orig, &ctx->crop_cnt, &ctx->vox_stamps, &ctx->xform};
  for (auto* b : bufs) b->release();
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->mail_h) cudaFreeHost(ctx->mail_h);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return ESKF_OK;
}

int eskf_ctx_sync(eskf_ctx* ctx) {
  ESKF_REQUIRE(ctx, "null ctx");
  ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
  return ESKF_OK;
}

int eskf_ctx_stream(eskf_ctx* ctx, void** cuda_stream) {
  ESKF_REQUIRE(ctx && cuda_stream, "null argument");
  *cuda_stream = ctx->stream;
  return ESKF_OK;
}

int eskf_ctx_launch_count(eskf_ctx* ctx, uint64_t* n) {
  ESKF_REQUIRE(ctx && n, "null argument");
  *n = ctx->launches;
  return ESKF_OK;
}

int eskf_ctx_get_option(eskf_ctx* ctx, const char* name, int64_t* value) {
  ESKF_REQUIRE(ctx && name && value, "null argument");
  const std::string n(name);
  if (n == "align_tuned_block") *value = ctx->tuned_block;          // 0 until a large cloud was registered
  else if (n == "align_autotune") *value = ctx->opt_align_autotune;
  else if (n == "align_block") *value = ctx->opt_align_block;
  else if (n == "align_depth") *value = ctx->opt_align_depth;
  else if (n == "vox_cluster") *value = ctx->opt_vox_cluster;
  else if (n == "vox_cluster_max") *value = ctx->vox_cluster_max;   // CTAs of the largest placeable cluster (0: none)
  else if (n == "stamps_sorted") *value = ctx->opt_stamps_sorted;
  else {
    set_error("unknown option '%s'", name);
    return ESKF_ERR_INVALID;
  }
  return ESKF_OK;
}

int eskf_ctx_set_option(eskf_ctx* ctx, const char* name, int64_t value) {
  ESKF_REQUIRE(ctx && name, "null argument");
  const std::string n(name);
  if (n == "align_dynamic_tiles") {
    ctx->opt_align_dynamic = value != 0;
  } else if (n == "l2_persist") {
    ctx->opt_l2_persist = value != 0;
  } else if (n == "mapped_results") {
    ctx->opt_mapped_results = (value != 0 && ctx->mail_h != nullptr) ? 1 : 0;
  } else if (n == "map_insert_sorted") {
    ctx->opt_insert_sorted = value != 0;
  } else if (n == "align_block") {
    ESKF_REQUIRE(value == 0 || value == 256 || value == 257 || value == 384 || value == 448 || value == 512 ||
                     value == 640 || value == 768 || value == 769,
                 "align_block must be 0 (by cloud size), 256, 257 (256 threads, 3-stage loop on the probe filter), 384, 448, 512, 640, 768 or 769 (768 threads, 3-stage loop)");
    ctx->opt_align_block = static_cast<int>(value);
  } else if (n == "align_depth") {
    ESKF_REQUIRE(value == 0 || (value >= 3 && value <= 11), "align_depth must be 0 (default) or 3 .. 11");
    ctx->opt_align_depth = static_cast<int>(value);
  } else if (n == "align_resident") {
    ESKF_REQUIRE(value >= -1 && value <= 1024, "align_resident must be -1 (auto) or a tile count");
    ctx->opt_align_resident = static_cast<int>(value);
  } else if (n == "align_fat_points") {
    ESKF_REQUIRE(value >= 0, "align_fat_points must be non-negative");
    ctx->opt_align_fat_points = value;
  } else if (n == "align_filter") {
    ctx->opt_align_filter = value != 0;
  } else if (n == "align_cons") {
    ESKF_REQUIRE(value >= 0 && value <= 20, "align_cons must be 0 (adaptive) or a warp count");
    ctx->opt_align_cons = static_cast<int>(value);
  } else if (n == "align_flags") {
    ESKF_REQUIRE(value >= 0 && value < 16384, "align_flags is a bit mask below 16384");
    ctx->opt_align_flags = static_cast<int>(value);
  } else if (n == "align_xchg_ll") {
    ctx->opt_align_xchg_ll = value != 0;
  } else if (n == "align_ll") {
    ctx->opt_align_ll = value != 0;
  } else if (n == "align_autotune") {
    ctx->opt_align_autotune = value != 0;
    ctx->tuned_block = 0;
  } else if (n == "vox_cluster") {
    ESKF_REQUIRE(value == 0 || value == 1 || value == 8 || value == 16, "vox_cluster must be 0, 1, 8 or 16");
    ctx->opt_vox_cluster = static_cast<int>(value);
  } else if (n == "stamps_sorted") {
    ESKF_REQUIRE(value >= -1 && value <= 1, "stamps_sorted must be -1 (check), 0 or 1");
    ctx->opt_stamps_sorted = static_cast<int>(value);
  } else if (n == "align_dyn16") {
    ESKF_REQUIRE(value >= 1 && value <= 12, "align_dyn16 must be in [1, 12]");
    ctx->opt_align_dyn16 = static_cast<int>(value);
  } else if (n == "align_ticket_chunk") {
    ESKF_REQUIRE(value == 1 || value == 2 || value == 4, "align_ticket_chunk must be 1, 2 or 4");
    ctx->opt_align_chunk = static_cast<int>(value);
  } else if (n == "knn_buffer") {
    ESKF_REQUIRE(value >= 1 && value <= 128, "knn_buffer must be in [1, 128]");
    ctx->opt_knn_buffer = static_cast<int>(value);
  } else {
    set_error("unknown option '%s'", name);
    return ESKF_ERR_INVALID;
  }
  return ESKF_OK;
}

int eskf_ctx_timer_start(eskf_ctx* ctx) {
  ESKF_REQUIRE(ctx, "null ctx");
  ESKF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  return ESKF_OK;
}

int eskf_ctx_timer_stop(eskf_ctx* ctx, float* elapsed_ms) {
  ESKF_REQUIRE(ctx && elapsed_ms, "null argument");
  ESKF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  ESKF_CUDA(cudaEventSynchronize(ctx->ev1));
  ESKF_CUDA(cudaEventElapsedTime(elapsed_ms, ctx->ev0, ctx->ev1));
  return ESKF_OK;
}

// ------------------------------------------------------------------ cloud
int eskf_cloud_create(eskf_ctx* ctx, size_t capacity, eskf_cloud** out) {
  ESKF_REQUIRE(ctx && out, "null argument");
  ESKF_CUDA(cudaSetDevice(ctx->device));
  eskf_cloud* c = new eskf_cloud();
  c->ctx = ctx;
  int st = cloud_reserve(c, capacity ? capacity : 64, false);
  if (st != ESKF_OK) {
    free_cloud_buffers(c);
    delete c;
    return st;
  }
  *out = c;
  return ESKF_OK;
}

int eskf_cloud_destroy(eskf_cloud* c) {
  if (!c) return ESKF_OK;
  cudaSetDevice(c->ctx->device);
  cudaStreamSynchronize(c->ctx->stream);
  free_cloud_buffers(c);
  delete c;
  return ESKF_OK;
}

int eskf_cloud_upload(eskf_cloud* c, const double* xyz, const double* cov, size_t n) {
  ESKF_REQUIRE(c, "null cloud");
  ESKF_REQUIRE(n == 0 || xyz, "null xyz");
  ESKF_REQUIRE(n < (1ull << 31), "cloud too large");
  eskf_ctx* ctx = c->ctx;
  ESKF_CUDA(cudaSetDevice(ctx->device));
  ESKF_TRY(cloud_reserve(c, n, cov != nullptr));
  c->n = n;
  c->has_cov = cov != nullptr;
  c->has_c32 = cov != nullptr;
  c->has_src = false;
  if (n == 0) return ESKF_OK;
  const size_t b_xyz = n * 24, b_cov = cov ? n * 72 : 0;
  ESKF_TRY(ctx->stage.ensure(b_xyz + b_cov));
  char* d = ctx->stage.as<char>();
  ESKF_CUDA(cudaMemcpyAsync(d, xyz, b_xyz, cudaMemcpyHostToDevice, ctx->stream));
  if (cov) ESKF_CUDA(cudaMemcpyAsync(d + b_xyz, cov, b_cov, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned nn = static_cast<unsigned>(n);
  aos_to_soa_kernel<<<(nn + 255) / 256, 256, 0, ctx->stream>>>(
      reinterpret_cast<const double*>(d), cov ? reinterpret_cast<const double*>(d + b_xyz) : nullptr,
      nn, c->cap, c->x(), c->y(), c->z(), c->cov, c->c4, c->c2);
  ESKF_CUDA(cudaGetLastError());
  count_launch(ctx);
  return ESKF_OK;
}

int eskf_cloud_upload_f32(eskf_cloud* c, const float* xyz, size_t n) {
  ESKF_REQUIRE(c, "null cloud");
  ESKF_REQUIRE(n == 0 || xyz, "null xyz");
  ESKF_REQUIRE(n < (1ull << 31), "cloud too large");
  eskf_ctx* ctx = c->ctx;
  ESKF_CUDA(cudaSetDevice(ctx->device));
  ESKF_TRY(cloud_reserve(c, n, false));
  c->n = n;
  c->has_cov = c->has_c32 = c->has_src = false;
  if (n == 0) return ESKF_OK;
  ESKF_TRY(ctx->stage.ensure(n * 12));
  ESKF_CUDA(cudaMemcpyAsync(ctx->stage.p, xyz, n * 12, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned nn = static_cast<unsigned>(n);
  f32_to_soa_kernel<<<(nn + 255) / 256, 256, 0, ctx->stream>>>(ctx->stage.as<float>(), nn, c->x(),
                                                               c->y(), c->z());
  ESKF_CUDA(cudaGetLastError());
  count_launch(ctx);
  return ESKF_OK;
}

int eskf_cloud_download(eskf_cloud* c, double* xyz, double* cov, uint32_t* src_index,
                        size_t capacity, size_t* n) {
  ESKF_REQUIRE(c, "null cloud");
  eskf_ctx* ctx = c->ctx;
  ESKF_CUDA(cudaSetDevice(ctx->device));
  if (n) *n = c->n;
  if (c->n == 0) return ESKF_OK;
  if (capacity < c->n) {
    set_error("download capacity %zu < %zu points", capacity, c->n);
    return ESKF_ERR_CAPACITY;
  }
  ESKF_REQUIRE(!cov || c->has_cov, "cloud has no covariances");
  ESKF_REQUIRE(!src_index || c->has_src, "cloud has no source indices");
  const size_t b_xyz = xyz ? c->n * 24 : 0, b_cov = cov ? c->n * 72 : 0;
  if (b_xyz + b_cov) {
    ESKF_TRY(ctx->stage.ensure(b_xyz + b_cov));
    char* d = ctx->stage.as<char>();
    const unsigned nn = static_cast<unsigned>(c->n);
    soa_to_aos_kernel<<<(nn + 255) / 256, 256, 0, ctx->stream>>>(
        c->x(), c->y(), c->z(), c->cov, c->cap, nn, xyz ? reinterpret_cast<double*>(d) : nullptr,
        cov ? reinterpret_cast<double*>(d + b_xyz) : nullptr);
    ESKF_CUDA(cudaGetLastError());
    count_launch(ctx);
    if (xyz) ESKF_CUDA(cudaMemcpyAsync(xyz, d, b_xyz, cudaMemcpyDeviceToHost, ctx->stream));
    if (cov) ESKF_CUDA(cudaMemcpyAsync(cov, d + b_xyz, b_cov, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (src_index)
    ESKF_CUDA(cudaMemcpyAsync(src_index, c->src, c->n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
  return ESKF_OK;
}

int eskf_cloud_size(eskf_cloud* c, size_t* n) {
  ESKF_REQUIRE(c && n, "null argument");
  *n = c->n;
  return ESKF_OK;
}

int eskf_cloud_transform(eskf_cloud* c, const double T[16]) {
  ESKF_REQUIRE(c && T, "null argument");
  eskf_ctx* ctx = c->ctx;
  ESKF_CUDA(cudaSetDevice(ctx->device));
  if (c->n == 0) return ESKF_OK;
  double* hT = nullptr;
  ESKF_TRY(ctx_pinned(ctx, 12 * sizeof(double), reinterpret_cast<void**>(&hT)));
  // the pinned scratch is shared: make sure earlier async users are done
  ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) hT[3 * i + j] = T[4 * i + j];
    hT[9 + i] = T[4 * i + 3];
  }
  ESKF_TRY(ctx->misc.ensure(256));
  ESKF_CUDA(cudaMemcpyAsync(ctx->misc.p, hT, 12 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  const unsigned nn = static_cast<unsigned>(c->n);
  transform_cloud_kernel<<<(nn + 255) / 256, 256, 0, ctx->stream>>>(
      c->x(), c->y(), c->z(), c->has_cov ? c->cov : nullptr, c->cap, nn, ctx->misc.as<double>());
  ESKF_CUDA(cudaGetLastError());
  count_launch(ctx);
  ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
  c->has_c32 = false;
  return ESKF_OK;
}

int eskf_cloud_copy(eskf_cloud* dst, const eskf_cloud* src) {
  ESKF_REQUIRE(dst && src && dst != src, "bad clouds");
  ESKF_REQUIRE(dst->ctx == src->ctx, "clouds belong to different contexts");
  eskf_ctx* ctx = dst->ctx;
  ESKF_CUDA(cudaSetDevice(ctx->device));
  ESKF_TRY(cloud_reserve(dst, src->n, src->has_cov));
  dst->n = src->n;
  dst->has_cov = src->has_cov;
  dst->has_c32 = src->has_cov && src->has_c32;
  dst->has_src = src->has_src;
  if (src->n == 0) return ESKF_OK;
  const size_t n = src->n;
  for (int k = 0; k < 3; ++k)
    ESKF_CUDA(cudaMemcpyAsync(dst->xyz + k * dst->cap, src->xyz + k * src->cap, n * 8,
                              cudaMemcpyDeviceToDevice, ctx->stream));
  if (src->has_cov) {
    for (int k = 0; k < 9; ++k)
      ESKF_CUDA(cudaMemcpyAsync(dst->cov + k * dst->cap, src->cov + k * src->cap, n * 8,
                                cudaMemcpyDeviceToDevice, ctx->stream));
    if (src->has_c32) {
      ESKF_CUDA(cudaMemcpyAsync(dst->c4, src->c4, n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
      ESKF_CUDA(cudaMemcpyAsync(dst->c2, src->c2, n * sizeof(float2), cudaMemcpyDeviceToDevice, ctx->stream));
    }
  }
  if (src->has_src)
    ESKF_CUDA(cudaMemcpyAsync(dst->src, src->src, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  return ESKF_OK;
}

// ----------------------------------------------------------- registration
static int fill_align_args(const eskf_map* map, const eskf_cloud* cloud, const double guess[16],
                           const eskf_icp_params* prm, AlignArgs* a) {
  ESKF_REQUIRE(map && cloud && guess && prm, "null argument");
  std::memset(a, 0, sizeof *a);
  a->map = map;
  a->cloud = cloud;
  std::memcpy(a->guess, guess, 16 * sizeof(double));
  a->max_iteration = prm->max_iteration;
  a->neighbor_mode = prm->neighbor_mode == 7 ? 7 : 1;
  a->trans_sq_thr = prm->translation_sq_threshold;
  a->cos_thr = prm->cosine_threshold;
  return ESKF_OK;
}

int eskf_align_cloud(eskf_ctx* ctx, const eskf_map* map, const eskf_cloud* cloud,
                     const double guess[16], const eskf_icp_params* params, double T_out[16],
                     eskf_align_info* info) {
  ESKF_REQUIRE(ctx && T_out, "null argument");
  AlignArgs a;
  ESKF_TRY(fill_align_args(map, cloud, guess, params, &a));
  ESKF_TRY(cloud_build_c32(const_cast<eskf_cloud*>(cloud)));
  return align_device(ctx, a, T_out, info);
}

int eskf_align_cloud_begin(eskf_ctx* ctx, const eskf_map* map, const eskf_cloud* cloud,
                           const double guess[16], const eskf_icp_params* params,
                           const eskf_align_info* info) {
  ESKF_REQUIRE(ctx, "null argument");
  AlignArgs a;
  ESKF_TRY(fill_align_args(map, cloud, guess, params, &a));
  ESKF_TRY(cloud_build_c32(const_cast<eskf_cloud*>(cloud)));
  return align_begin(ctx, a, info);
}

int eskf_align_end(eskf_ctx* ctx, double T_out[16], eskf_align_info* info) {
  ESKF_REQUIRE(ctx && T_out, "null argument");
  return align_end(ctx, T_out, info);
}

int eskf_align_batch(eskf_ctx* const* ctxs, int n_ctx, const eskf_map* const* maps,
                     const eskf_cloud* const* clouds, const double* guesses, size_t n,
                     const eskf_icp_params* params, double* T_out, eskf_align_info* infos) {
  ESKF_REQUIRE(ctxs && n_ctx > 0 && n_ctx <= 256, "eskf_align_batch needs 1..256 contexts");
  for (int s = 0; s < n_ctx; ++s) {
    ESKF_REQUIRE(ctxs[s], "null context in the batch");
    for (int r = 0; r < s; ++r) ESKF_REQUIRE(ctxs[r] != ctxs[s], "the contexts of a batch must be distinct");
  }
  ESKF_REQUIRE(params, "null params");
  if (n == 0) return ESKF_OK;
  ESKF_REQUIRE(maps && clouds && guesses && T_out, "null argument");
  std::vector<long long> pending(static_cast<size_t>(n_ctx), -1);  // job in flight on each context
  auto collect = [&](int s) -> int {
    const long long j = pending[static_cast<size_t>(s)];
    pending[static_cast<size_t>(s)] = -1;
    return eskf_align_end(ctxs[s], T_out + 16 * j, infos ? &infos[j] : nullptr);
  };
  int rc = ESKF_OK;
  std::string first_error;
  for (size_t i = 0; i < n && rc == ESKF_OK; ++i) {
    const int s = static_cast<int>(i % static_cast<size_t>(n_ctx));
    if (pending[static_cast<size_t>(s)] >= 0) rc = collect(s);
    if (rc == ESKF_OK) {
      rc = eskf_align_cloud_begin(ctxs[s], maps[i], clouds[i], guesses + 16 * i, params,
                                  infos ? &infos[i] : nullptr);
      if (rc == ESKF_OK) pending[static_cast<size_t>(s)] = static_cast<long long>(i);
    }
  }
  if (rc != ESKF_OK) first_error = eskf_last_error();
  for (int s = 0; s < n_ctx; ++s) {
    if (pending[static_cast<size_t>(s)] < 0) continue;
    const int r2 = collect(s);
    if (rc == ESKF_OK && r2 != ESKF_OK) {
      rc = r2;
      first_error = eskf_last_error();
    }
  }
  if (rc != ESKF_OK) set_error("%s", first_error.c_str());
  return rc;
}

int eskf_align(eskf_ctx* ctx, const eskf_map* map, const double* xyz, const double* cov, size_t n,
               const double guess[16], const eskf_icp_params* params, double T_out[16],
               eskf_align_info* info) {
  ESKF_REQUIRE(ctx && map, "null argument");
  ESKF_REQUIRE(n == 0 || (xyz && cov), "null xyz/cov");
  ESKF_REQUIRE(guess && params && T_out, "null argument");
  if (n == 0) {
    // zero correspondences: zero step, "converged" after one iteration (the
    // reference's LDLT of a zero system solves to zero)
    std::memcpy(T_out, guess, 16 * sizeof(double));
    if (info) {
      info->iterations = 1;
      info->converged = 1;
      info->n_corr_last = 0;
    }
    return ESKF_OK;
  }
  if (!ctx->tmp_cloud[0]) ESKF_TRY(eskf_cloud_create(ctx, n, &ctx->tmp_cloud[0]));
  ESKF_TRY(eskf_cloud_upload(ctx->tmp_cloud[0], xyz, cov, n));
  return eskf_align_cloud(ctx, map, ctx->tmp_cloud[0], guess, params, T_out, info);
}

int eskf_align_cloud_fixed(eskf_ctx* ctx, const eskf_map* map, const eskf_cloud* cloud,
                           const double guess[16], int iterations, int neighbor_mode,
                           double T_out[16], eskf_align_info* info) {
  ESKF_REQUIRE(ctx && T_out, "null argument");
  ESKF_REQUIRE(iterations > 0, "iterations must be positive");
  eskf_icp_params prm = {iterations, neighbor_mode, 0.0, 2.0};
  AlignArgs a;
  ESKF_TRY(fill_align_args(map, cloud, guess, &prm, &a));
  a.fixed_iterations = iterations;
  ESKF_TRY(cloud_build_c32(const_cast<eskf_cloud*>(cloud)));
  return align_device(ctx, a, T_out, info);
}

int eskf_linearize(eskf_ctx* ctx, const eskf_map* map, const double* xyz, const double* cov,
                   size_t n, const double T[16], int neighbor_mode, int fp64_math, double H36[36],
                   double b6[6], uint8_t* hit, uint64_t* n_corr) {
  ESKF_REQUIRE(ctx && map && T && H36 && b6, "null argument");
  ESKF_REQUIRE(n > 0 && xyz && cov, "empty input");
  if (!ctx->tmp_cloud[0]) ESKF_TRY(eskf_cloud_create(ctx, n, &ctx->tmp_cloud[0]));
  ESKF_TRY(eskf_cloud_upload(ctx->tmp_cloud[0], xyz, cov, n));
  eskf_icp_params prm = {1, neighbor_mode, 0.0, 2.0};
  AlignArgs a;
  ESKF_TRY(fill_align_args(map, ctx->tmp_cloud[0], T, &prm, &a));
  a.fixed_iterations = 1;
  a.fp64_math = fp64_math;
  const size_t nn = neighbor_mode == 7 ? 7 : 1;
  if (hit) {
    ESKF_TRY(ctx->misc.ensure(n * nn + 256));
    a.d_hit = ctx->misc.as<uint8_t>() + 256;
  }
  double Tout[16];
  uint64_t nc = 0;
  eskf_align_info info;
  std::memset(&info, 0, sizeof info);
  info.trace_H = H36;
  info.trace_b = b6;
  info.trace_ncorr = &nc;
  ESKF_TRY(align_device(ctx, a, Tout, &info));
  if (n_corr) *n_corr = nc;
  if (hit) {
    ESKF_CUDA(cudaMemcpyAsync(hit, a.d_hit, n * nn, cudaMemcpyDeviceToHost, ctx->stream));
    ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return ESKF_OK;
}

int eskf_align_cloud_sharded(eskf_ctx* ctx, const eskf_map* map, const eskf_cloud* cloud,
                             const double guess[16], const eskf_icp_params* params,
                             eskf_allreduce_fn allreduce, void* user, int fixed_iterations,
                             double T_out[16], eskf_align_info* info) {
  ESKF_REQUIRE(ctx && T_out, "null argument");
  AlignArgs a;
  ESKF_TRY(fill_align_args(map, cloud, guess, params, &a));
  a.fixed_iterations = fixed_iterations > 0 ? fixed_iterations : 0;
  ESKF_TRY(cloud_build_c32(const_cast<eskf_cloud*>(cloud)));
  return align_sharded(ctx, a, allreduce, user, T_out, info);
}

// ---------------------------------------------------------- multi-GPU comm
int eskf_comm_create(eskf_ctx* ctx, int rank, int world, eskf_comm** out) {
  ESKF_REQUIRE(ctx && out, "null argument");
  ESKF_REQUIRE(world >= 1 && world <= ESKF_MAX_WORLD, "world must be in [1, 16]");
  ESKF_REQUIRE(rank >= 0 && rank < world, "rank out of range");
  ESKF_CUDA(cudaSetDevice(ctx->device));
  eskf_comm* c = new eskf_comm();
  c->ctx = ctx;
  c->rank = rank;
  c->world = world;
  const size_t bytes = static_cast<size_t>(4) * world * 64 * sizeof(double);  // [call parity][iteration parity][rank][64]
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->local), bytes);
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, bytes);
  if (e != cudaSuccess) {
    set_error("mailbox allocation: %s", cudaGetErrorString(e));
    delete c;
    return ESKF_ERR_CUDA;
  }
  c->peers[rank] = c->local;
  c->connected = world == 1;
  *out = c;
  return ESKF_OK;
}

int eskf_comm_destroy(eskf_comm* c) {
  if (!c) return ESKF_OK;
  cudaSetDevice(c->ctx->device);
  cudaStreamSynchronize(c->ctx->stream);
  for (int r = 0; r < c->world; ++r)
    if (c->opened[r]) cudaIpcCloseMemHandle(c->peers[r]);
  cudaFree(c->local);
  delete c;
  return ESKF_OK;
}

int eskf_comm_local_handle(eskf_comm* c, void* handle64) {
  ESKF_REQUIRE(c && handle64, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == ESKF_COMM_HANDLE_BYTES, "IPC handle size");
  ESKF_CUDA(cudaSetDevice(c->ctx->device));
  cudaIpcMemHandle_t h;
  ESKF_CUDA(cudaIpcGetMemHandle(&h, c->local));
  std::memcpy(handle64, &h, sizeof h);
  return ESKF_OK;
}

int eskf_comm_connect(eskf_comm* c, const void* handles) {
  ESKF_REQUIRE(c && handles, "null argument");
  ESKF_CUDA(cudaSetDevice(c->ctx->device));
  const char* hp = static_cast<const char*>(handles);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank || c->opened[r]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, hp + static_cast<size_t>(r) * sizeof h, sizeof h);
    void* ptr = nullptr;
    ESKF_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->peers[r] = static_cast<double*>(ptr);
    c->opened[r] = true;
  }
  c->connected = true;
  return ESKF_OK;
}

int eskf_align_cloud_p2p(eskf_ctx* ctx, const eskf_map* map, const eskf_cloud* cloud,
                         const double guess[16], const eskf_icp_params* params, eskf_comm* comm,
                         int fixed_iterations, double T_out[16], eskf_align_info* info) {
  ESKF_REQUIRE(ctx && T_out && comm, "null argument");
  AlignArgs a;
  ESKF_TRY(fill_align_args(map, cloud, guess, params, &a));
  a.fixed_iterations = fixed_iterations > 0 ? fixed_iterations : 0;
  a.comm = comm;
  if (cloud->n > 0) ESKF_TRY(cloud_build_c32(const_cast<eskf_cloud*>(cloud)));
  return align_device(ctx, a, T_out, info);
}

}  // extern "C"
