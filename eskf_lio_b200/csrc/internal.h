// internal.h — handle layouts and host helpers behind include/eskf_gpu.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/eskf_gpu.h"
#include "common.cuh"

namespace eskf {

// ------------------------------------------------------------------ errors
void set_error(const char* fmt, ...);

#define ESKF_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      eskf::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return ESKF_ERR_CUDA;                                                              \
    }                                                                                    \
  } while (0)

#define ESKF_TRY(call)            \
  do {                            \
    int s__ = (call);             \
    if (s__ != ESKF_OK) return s__; \
  } while (0)

#define ESKF_REQUIRE(cond, msg)                   \
  do {                                            \
    if (!(cond)) {                                \
      eskf::set_error("invalid argument: %s", msg); \
      return ESKF_ERR_INVALID;                    \
    }                                             \
  } while (0)

// grow-only device buffer
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return ESKF_OK;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    size_t want = need + need / 4 + 256;
    ESKF_CUDA(cudaMalloc(&p, want));
    bytes = want;
    return ESKF_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <typename T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

// ---------------------------------------------------------------- layouts
// One voxel record of the open-addressing table, 64 B = one HBM access.
// Sector 0: key (verifies a tag match), count, mean; sector 1: covariance.
// The mean is stored relative to the voxel centre so fp32 keeps ~3e-8 m.
struct __align__(16) VoxelSlot {
  uint64_t key;    // pack_key(), kEmptyKey when free
  uint32_t count;  // numPoints (capped)
  uint32_t pad0;
  float mx, my, mz;  // mean - (k + 0.5) * voxel
  float pad1;
  float c00, c01, c02, c11;
  float c12, c22, pad2, pad3;
};
static_assert(sizeof(VoxelSlot) == 64, "VoxelSlot must be 64 bytes");

constexpr int kMasterStride = 12;  // fp64 master: mean[3] + cov[9] per slot

struct DeskewSeg {
  uint32_t begin, end;
  double T[12];
};

// written by the voxelize kernel, read by its consumers
struct VoxelHeader {
  int mn[3];      // min voxel coordinate per axis   (memset 0x7F before launch)
  int nmx[3];     // min of the NEGATED coordinate   (same)
  int pad[2];
  // --- zeroed before launch
  unsigned error;    // bit0: coordinate out of range, bit1: barrier timeout
  unsigned sel;      // which ping-pong buffer holds the sorted (key, idx)
  unsigned n_out;    // runs (mode 0) or kept points (mode 1)
  unsigned bits;     // bits per axis of the rebased coordinates
  unsigned knn_next; // next kept point to hand to a k-NN search warp (dynamic scheduling)
  unsigned pad2[3];
  GridBarrier gb;
};

// Host-mapped (zero-copy) result words: a kernel stores its few result bytes here, fences at
// system scope and bumps the sequence word; the host polls it instead of paying a D2H copy + a
// stream synchronisation per call (and, for the preprocessor, returns as soon as the voxelize
// kernel has published the kept-point count, while the k-NN kernels are still running).
struct HostMail {
  unsigned vox_seq;       // = the call's sequence number once the fields below are valid
  unsigned vox_n_out;
  unsigned vox_error;     // VoxelHeader::error
  unsigned vox_gb_error;  // grid barrier timeout
  unsigned align_seq;
  unsigned align_error;
  int align_iter;
  int align_converged;
  unsigned long long align_ncorr;
  double align_T[12];
};

}  // namespace eskf

struct eskf_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 0;
  uint64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // scratch
  eskf::DevBuf stage;      // AoS upload / download staging
  eskf::DevBuf sortbuf;    // keyA keyB idxA idxB
  eskf::DevBuf hist;       // 2 x [G][256] + blk counts
  eskf::DevBuf hdr;        // VoxelHeader
  eskf::DevBuf runs;       // run_start / keep_flag / kept_src / kept_pos
  eskf::DevBuf sorted_xyz; // positions gathered in sorted order (kNN)
  eskf::DevBuf knn_levels; // block-range tables of the k-NN search (kKnnHashLevels x stride uint4)
  eskf::DevBuf knn_nbr;    // neighbour lists handed from the search to the finish kernel
  eskf::DevBuf link;       // next[] / slot_of[] of the sort-free map insert
  eskf::DevBuf segs;       // deskew segments
  eskf::DevBuf work;       // align working positions (SoA)
  eskf::DevBuf spill;      // align depth 7: hit-list entries beyond shared memory
  // range crop of the preprocessor (eskf_ctx_set_range_crop; off by default: the reference has none)
  bool crop = false;
  double crop_min2 = 0.0, crop_max2 = 0.0;  // squared bounds, LiDAR frame
  eskf::DevBuf crop_orig;   // [n'] index of every surviving point in the uncropped sweep
  eskf::DevBuf crop_cnt;    // per-block survivor counts + the total
  eskf_cloud* crop_cloud = nullptr;  // the surviving points (what the rest of the preprocessor runs on)
  eskf::DevBuf partials;   // align per-block partial sums
  eskf::DevBuf astate;     // align state + traces
  eskf::DevBuf misc;       // small outputs (query / export counters)
  void* pinned = nullptr;  // pinned host scratch for small read-backs
  size_t pinned_bytes = 0;
  eskf_cloud* tmp_cloud[3] = {nullptr, nullptr, nullptr};  // host-buffer entry points
  int max_blocks_voxelize = 0;
  eskf::DevBuf xform;                 // transformed sweep when the preprocessor must leave its input as delivered
  eskf::DevBuf vox_stamps;            // ESKF_TRACE: phase stamps of the one-cluster voxelize kernel
  bool knn_levels_clean = false;  // every entry of knn_levels is empty (kept so by knn_finish_kernel; false after (re)allocation or a failed call)
  int tuned_block = 0;          // large clouds: CTA size the autotune chose (0: not tuned)
  size_t tuned_n = 0;           // ... on a cloud of this many points
  int vox_cluster_max = 0;      // CTAs of the largest cluster the one-cluster voxelize kernel can be placed in (0: off)
  int max_blocks_align = 0;
  int opt_align_dynamic = 1;    // eskf_ctx_set_option knobs (initialised from the environment)
  int opt_l2_persist = 1;
  eskf::HostMail* mail_h = nullptr;  // cudaHostAllocMapped
  eskf::HostMail* mail_d = nullptr;  // the same memory as the device sees it
  unsigned vox_seq = 0, align_seq = 0;
  int opt_mapped_results = 1;   // 0: always cudaMemcpyAsync + cudaStreamSynchronize
  std::shared_ptr<void> pending_align;  // registration.cu: state between align_begin / align_end
  int opt_trace = 0;            // ESKF_TRACE=1: per-kernel CUDA-event timings on stderr (debug aid)
  std::vector<std::pair<const char*, cudaEvent_t>> trace_marks;
  int opt_knn_buffer = 128;
  int opt_insert_sorted = 0;    // 1: always take the radix-sort insert path
  int opt_align_block = 0;      // CTA size of the 1-neighbour fp32 align kernel: 0 = by cloud size, 256 | 384 | 512 | 640 | 768
  int opt_align_autotune = 1;   // time the two large-cloud loop shapes once per context and keep the faster one
  int opt_vox_cluster = 1;      // 0 = sweeps go through the grid-wide voxelize kernel, 1 = through one cluster when they fit, 8 = cluster capped at 8 CTAs
  int opt_stamps_sorted = -1;   // deskew: -1 = check the stamps (one pass over them), 1 = caller vouches they are non-decreasing, 0 = they are not
  int opt_align_dyn16 = 3;      // sixteenths of an align pass dealt by tickets (the rest is a fixed stride per warp)
  int opt_align_chunk = 2;      // tiles taken per ticket in the dynamic tail of an align pass (1 | 2 | 4)
  int opt_align_depth = 0;      // large-cloud align kernel: 0 = default (6), 3 | 4 = round-1 register pipelines, 5 = SM-resident positions + probe filter + bulk-copy ring, 6 = producer / consumer warps around a shared-memory hit queue
  int opt_align_resident = -1;  // depth 5: resident warp tiles per warp (-1 = as many as shared memory holds)
  int64_t opt_align_fat_points = 1 << 17;  // clouds with at least this many points take the large-cloud kernel
  int opt_align_stamps = 0;     // ESKF_ALIGN_STAMPS=1: globaltimer stamps of the iteration hand-off on stderr (traced calls)
  int opt_align_filter = 1;     // depth 4 (fat CTAs): probe the 8-bit L2-resident filter instead of the 16-bit tags
  int opt_align_cons = 0;       // depth 6: consumer warps per CTA (0 = follow the hit rate)
  int opt_align_flags = 16;     // L2 policy of the large-cloud kernels (registration.cu, kFlag*): 16 = filter windows evict_last
  int opt_l2_carveout = 1;      // 0: no persisting-L2 set-aside (ESKF_L2_CARVEOUT=0; decided at context creation)
  int opt_align_xchg_ll = 1;    // multi-GPU H/b exchange: flagged words (1) or data words + a release flag (0)
  int opt_align_ll = 1;         // depth 5: flagged-word (LL) pose broadcast instead of epoch word + second round trip
  size_t l2_persist_bytes = 0;  // persisting-L2 carve-out (0 = unavailable)
  size_t l2_window_max = 0;     // max access-policy window
  const void* l2_win_ptr = nullptr;  // window currently set on the stream
  size_t l2_win_bytes = 0;
};

struct eskf_cloud {
  eskf_ctx* ctx = nullptr;
  size_t n = 0, cap = 0;
  double* xyz = nullptr;  // SoA: x[cap] y[cap] z[cap]
  double* cov = nullptr;  // SoA: 9 x [cap], row-major entry index outermost
  float4* c4 = nullptr;   // fp32 mirror of the covariance: (c00 c01 c02 c11)
  float2* c2 = nullptr;   //                                (c12 c22)
  uint32_t* src = nullptr;
  bool has_cov = false, has_c32 = false, has_src = false;
  double* x() const { return xyz; }
  double* y() const { return xyz + cap; }
  double* z() const { return xyz + 2 * cap; }
};

struct eskf_map {
  eskf_ctx* ctx = nullptr;
  double voxel = 0.0;
  uint32_t cap_pts = 0;
  uint64_t n_slots = 0;                // < 2^32, any size
  eskf::tag_t* tags = nullptr;         // [n_slots] 0 = empty; probed instead of the records
  eskf::VoxelSlot* slots = nullptr;    // [n_slots] 64 B records
  double* master = nullptr;            // [n_slots][12]
  unsigned long long* d_count = nullptr;  // occupied voxels (+1 word: table-full error)
  uint64_t count_upper = 0;            // host-side upper bound of *d_count
  // the table the last same-size rebuild (eviction sweep) moved out of: the
  // next sweep moves back into it, so steady-state eviction never calls
  // cudaMalloc / cudaFree
  uint32_t* head = nullptr;            // [head_n] pending-list heads of the sort-free insert (all ~0 between batches)
  uint64_t head_n = 0;
  eskf::tag_t* spare_tags = nullptr;
  eskf::VoxelSlot* spare_slots = nullptr;
  double* spare_master = nullptr;
  uint64_t spare_n = 0;
  // 8-bit probe filter of the table (local_map.cu, map_probe_filter): derived from `tags` on demand
  // by the dense registration, valid while filt_version == version
  uint64_t version = 1;                // bumped by every call that changes the table
  mutable uint8_t* filt = nullptr;     // [filt_slots + kFilterPad]
  mutable uint64_t filt_slots = 0;
  mutable uint64_t filt_version = 0;
};

#define ESKF_MAX_WORLD 16

// peer-mapped H/b mailboxes of a sharded registration (registration.cu, exchange_sums)
struct eskf_comm {
  eskf_ctx* ctx = nullptr;
  int rank = 0, world = 1;
  double* local = nullptr;                  // [2 call parities][2 iteration parities][world][64] 8-byte words, cudaMalloc'ed
  double* peers[ESKF_MAX_WORLD] = {};       // peers[rank] == local
  bool opened[ESKF_MAX_WORLD] = {};         // mapped with cudaIpcOpenMemHandle
  bool connected = false;
  unsigned seq = 0;
};

namespace eskf {

// pinned scratch of at least `bytes`
int ctx_pinned(eskf_ctx* ctx, size_t bytes, void** out);
int cloud_reserve(eskf_cloud* c, size_t cap, bool with_cov);
int cloud_build_c32(eskf_cloud* c);

// voxelize.cu ------------------------------------------------------------
struct VoxelizeArgs {
  const double* in_x;
  const double* in_y;
  const double* in_z;
  int in_stride;
  double* out_x;
  double* out_y;
  double* out_z;
  double* cov;       // SoA 9 x pitch, rotated in place by T1 (nullable)
  size_t cov_pitch;
  unsigned n;
  double voxel;
  int has_T1;
  double T1[12];
  const DeskewSeg* segs;
  int n_segs;
  const uint32_t* orig;  // nullable: point i is point orig[i] of the sweep the deskew segments index (range crop)
  int mode;  // 0: runs in sorted order (map insert), 1: kept points in source order (preprocess)
  HostMail* mail;     // mode 1, nullable: publish n_out / error words here when known
  unsigned mail_seq;
};
// After it returns (asynchronously): ctx->hdr holds the VoxelHeader; the
// sorted (key, idx) are in sort buffer hdr.sel; mode 0: runs[0..n_out] =
// run starts (+ sentinel n); mode 1: kept_src / kept_pos lists and the
// sorted-order position arrays.
int voxelize(eskf_ctx* ctx, const VoxelizeArgs& a);

struct SortView {
  uint64_t* key[2];
  uint32_t* idx[2];
  uint32_t* run_start;  // mode 0
  uint32_t* keep_flag;  // mode 1
  uint32_t* kept_src;
  uint32_t* kept_pos;
  double* sx;
  double* sy;
  double* sz;
  VoxelHeader* hdr;
  uint4* levels;          // mode 1: k-NN block-range tables
  unsigned level_stride;  // entries between consecutive level tables
};
SortView sort_view(eskf_ctx* ctx, unsigned n);

// local_map.cu ------------------------------------------------------------
int map_reserve(eskf_map* m, uint64_t incoming_points);
// The 8-bit probe filter of the map's table: filt[i] = 0 for an empty slot, else filter_tag(tags[i])
// in 1..255; the first kFilterPad entries are repeated after the end so that a 16-entry window never
// wraps.  1 B/slot: the filter of the 36 M-slot dense map is 36 MB and stays L2-resident next to the
// position / covariance streams, where the 72 MB tag array did not (78 % of the tag probes went to HBM,
// profiles/r1_prof_align_ncu.md).  (Re)built only when the table changed since the last build.
constexpr uint32_t kFilterPad = 16;
__host__ __device__ __forceinline__ uint32_t filter_tag(uint32_t tag16) {
  const uint32_t t = tag16 >> 8;
  return t != 0u ? t : 1u;
}
int map_probe_filter(const eskf_map* m, const uint8_t** filt);

// preprocess.cu -----------------------------------------------------------
// spin on a host-mapped sequence word; ESKF_OK when it reached `seq`, 1 when the stream went idle
// without it (caller falls back to a copy), ESKF_ERR_CUDA on a stream error
int wait_mail(eskf_ctx* ctx, const volatile unsigned* word, unsigned seq);

int compute_deskew_segments(const double* point_time, size_t n, const eskf_state* states,
                            size_t n_states, std::vector<DeskewSeg>* out, int sorted_hint = -1);

// registration.cu ---------------------------------------------------------
struct AlignArgs {
  const eskf_map* map;
  const eskf_cloud* cloud;
  double guess[16];
  int max_iteration;
  int neighbor_mode;
  double trans_sq_thr;
  double cos_thr;
  int fixed_iterations;  // > 0: run exactly this many, ignore convergence
  int fp64_math;
  uint8_t* d_hit;        // device, optional: hit mask of iteration 0
  eskf_comm* comm;       // non-null + world > 1: fused NVLink exchange of the sums
};
int align_device(eskf_ctx* ctx, const AlignArgs& a, double T_out[16], eskf_align_info* info);
int align_begin(eskf_ctx* ctx, const AlignArgs& a, const eskf_align_info* info);
int align_end(eskf_ctx* ctx, double T_out[16], eskf_align_info* info);

inline void count_launch(eskf_ctx* ctx, int n = 1) { ctx->launches += n; }

// ESKF_TRACE=1: event marks between kernels; trace_flush synchronises and prints the gaps
inline void trace_mark(eskf_ctx* ctx, const char* label) {
  if (!ctx->opt_trace) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, ctx->stream);
  ctx->trace_marks.emplace_back(label, e);
}
inline void trace_flush(eskf_ctx* ctx, const char* what) {
  if (!ctx->opt_trace || ctx->trace_marks.empty()) return;
  cudaStreamSynchronize(ctx->stream);
  std::fprintf(stderr, "[eskf trace] %s:", what);
  for (size_t i = 1; i < ctx->trace_marks.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->trace_marks[i - 1].second, ctx->trace_marks[i].second);
    std::fprintf(stderr, " %s %.1f", ctx->trace_marks[i].first, ms * 1e3f);
  }
  std::fprintf(stderr, " (us)\n");
  for (auto& m : ctx->trace_marks) cudaEventDestroy(m.second);
  ctx->trace_marks.clear();
}

}  // namespace eskf
