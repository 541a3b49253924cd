/*
 * eskf_gpu.h — C ABI of the B200 (sm_100a) hot path of ESKF_LIO.
 *
 * The reference (LimHaeryong/ESKF_LIO) has no plugin / FFI layer; its
 * boundary is three C++ classes.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference tree).  The host
 * classes in eskf_lio_b200/host/ESKF_LIO/ keep the reference signatures and
 * call only these symbols; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns an eskf_status (0 = OK) and never throws;
 *     eskf_last_error() gives the message of the calling thread's last error.
 *   - host buffers are caller-owned; 4x4 transforms are row-major double[16];
 *     covariances are row-major double[9] per point; xyz is double[3] per point
 *     (layout of std::vector<Eigen::Vector3d> / <Eigen::Matrix3d> in the
 *     reference, include/ESKF_LIO/LocalMap.hpp:21-22).
 *   - opaque handles own all device memory.  One eskf_ctx = one device + one
 *     CUDA stream = one caller thread (the reference calls the three classes
 *     from its single main thread, src/main.cpp:70).
 *   - there is NO CPU fallback: without a CUDA device every call fails with
 *     ESKF_ERR_NO_DEVICE.
 *   - voxel coordinates must satisfy |floor(p / voxel_size)| < 2^20 per axis
 *     (21-bit packed keys); violations return ESKF_ERR_RANGE.
 */
#ifndef ESKF_GPU_H_
#define ESKF_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESKF_GPU_ABI_VERSION 1

typedef enum {
  ESKF_OK = 0,
  ESKF_ERR_CUDA = 1,       /* a CUDA runtime call failed */
  ESKF_ERR_INVALID = 2,    /* bad argument */
  ESKF_ERR_NO_DEVICE = 3,  /* no usable CUDA device (no CPU fallback exists) */
  ESKF_ERR_CAPACITY = 4,   /* output buffer too small / table cannot grow */
  ESKF_ERR_RANGE = 5,      /* voxel coordinate outside the 21-bit key range */
  ESKF_ERR_INTERNAL = 6    /* device-side invariant violated (e.g. barrier timeout) */
} eskf_status;

typedef struct eskf_ctx eskf_ctx;
typedef struct eskf_map eskf_map;
typedef struct eskf_cloud eskf_cloud;
typedef struct eskf_comm eskf_comm;
#define ESKF_COMM_HANDLE_BYTES 64 /* sizeof(cudaIpcMemHandle_t) */

/* registration.* keys of config/hilti_config.yaml:50-53, read by
 * include/ESKF_LIO/Registration.hpp:23-27.  neighbor_mode 1 = the reference's
 * single-voxel lookup (src/LocalMap.cpp:93-100); 7 = DIRECT7 extension
 * (6 face neighbours too; BASELINE.json configs[3], not in the reference). */
typedef struct {
  int32_t max_iteration;            /* default 100  */
  int32_t neighbor_mode;            /* 1 or 7       */
  double translation_sq_threshold;  /* default 1e-6 */
  double cosine_threshold;          /* default 0.9999 */
} eskf_icp_params;

/* Result of eskf_align*.  The trace pointers are optional caller-owned host
 * buffers, each sized for params.max_iteration entries (NULL = not wanted). */
typedef struct {
  int32_t iterations;    /* computeTransform calls performed (Registration.cpp:15-28) */
  int32_t converged;     /* convergenceCheck result of the last step (:22-25) */
  uint64_t n_corr_last;  /* correspondences of the last iteration */
  double* trace_H;       /* [it][36] row-major J^T W J  */
  double* trace_b;       /* [it][6]  J^T W r            */
  uint64_t* trace_ncorr; /* [it]                        */
  double* trace_step;    /* [it][16] per-iteration SE(3) step */
} eskf_align_info;

/* The fields of ESKF_LIO::State (include/ESKF_LIO/Types.hpp:31-52) that the
 * hot path reads (CloudPreprocessor::deskew only). */
typedef struct {
  double timestamp;
  double position[3];
  double attitude_xyzw[4]; /* Eigen::Quaterniond coefficient order x,y,z,w */
} eskf_state;

/* Multi-GPU sharding of ONE registration: the source cloud is
 * split by point range over `world` ranks, the map is replicated, and the 27
 * H/b partial sums are exchanged every Gauss-Newton iteration.  The exchange
 * is a caller-provided callback (e.g. torch.distributed all_reduce on the
 * device buffer), invoked with the context's stream. */
typedef int (*eskf_allreduce_fn)(void* user, void* device_buf_f64, size_t count,
                                 void* cuda_stream);

/* ------------------------------------------------------------------ misc */
int eskf_abi_version(void);
const char* eskf_last_error(void);
int eskf_device_count(int* n);

/* pinned host memory for asynchronous H2D/D2H of caller buffers */
int eskf_host_alloc(size_t bytes, void** out);
int eskf_host_free(void* p);

/* ------------------------------------------------------------------- ctx */
/* cuda_stream: a cudaStream_t to run on (e.g. torch's current stream), or
 * NULL to let the context create its own non-blocking stream. */
int eskf_ctx_create(int device, void* cuda_stream, eskf_ctx** out);
int eskf_ctx_destroy(eskf_ctx* ctx);
int eskf_ctx_sync(eskf_ctx* ctx);
int eskf_ctx_stream(eskf_ctx* ctx, void** cuda_stream);
/* number of kernels launched by this context so far (bench "gpu_launches") */
int eskf_ctx_launch_count(eskf_ctx* ctx, uint64_t* n);
/* Tuning / test knobs (all have sensible defaults; unknown names return ESKF_ERR_INVALID):
 *   "align_dynamic_tiles"  0 = fully static tile schedule in the registration kernel (bit-
 *                          reproducible summation order), 1 = static + dynamic tail (default)
 *   "l2_persist"           0 = no persisting-L2 window on the map's tag array (default 1)
 *   "mapped_results"       0 = results always come back by cudaMemcpyAsync + stream sync (default 1:
 *                          kernels publish them in host-mapped memory and the host polls)
 *   "map_insert_sorted"    1 = every map insert goes through the radix-sort path (default 0:
 *                          batches up to 131072 points take the sort-free list path; both give
 *                          bit-identical maps)
 *   "knn_buffer"           candidates the 30-NN selection keeps in shared memory before it falls
 *                          back to serial insertion (1..128, default 128; tests force the fallback)
 *   "align_block"          CTA size of the 1-neighbour fp32 registration kernel: 0 = chosen by cloud
 *                          size (default: 256 threads x 3 CTAs per SM, or one 512-thread CTA per SM
 *                          from "align_fat_points" points on); 256, 384, 448, 512, 640 or 768; 257 = 256
 *                          threads with the 3-stage loop on the probe filter, 769 = 768 threads with the
 *                          3-stage loop (what the autotune calls its candidates)
 *   "align_fat_points"     cloud size from which the one-CTA-per-SM shape is used (default 131072)
 *   "align_depth"          arrangement of a pass over a large cloud: 0 = default (4), 3 = loads issued
 *                          and consumed in the same trip, 4 = consumed one trip later (register
 *                          rotation), 5..11 = the experimental arrangements of DESIGN.md section 8
 *                          (shared-memory staging by cp.async.bulk, role-specialised warps, phase
 *                          split, parked candidates, three launches per iteration): same results
 *   "align_filter"         1 = large clouds probe the 8-bit filter derived from the tags (default), 0 =
 *                          the 16-bit tags
 *   "align_flags"          L2 eviction-policy bits of the pass (default 16 = filter windows evict_last)
 *   "align_ll"             1 = the step of an iteration reaches the CTAs as flagged 8-byte words
 *                          (default), 0 = data + release flag
 *   "align_xchg_ll"        the same choice for the exchange of the sums between GPUs (default 1)
 *   "align_ticket_chunk"   warp tiles taken per ticket in the load-balanced tail of a pass over a
 *                          large cloud (1, 2 or 4; default 2)
 *   "align_dyn16"          sixteenths of a pass dealt by tickets (1..12, default 3)
 *   "align_autotune"       1 = the first registration of a large cloud on a context times the two loop
 *                          shapes (4-deep 512 threads; 3-stage 3 x 256 or 1 x 768 threads, on the tags or the
 *                          filter) on its own device and data and keeps the fastest (default; the pool's GPUs differ); 0 = always the 4-deep one.
 *                          Ignored while "align_block" or "align_depth" is set explicitly
 *   "vox_cluster"          1 = a batch of up to 65536 points is voxelised and sorted by ONE thread-block
 *                          cluster (hardware cluster barriers between the phases; default), 0 = always the
 *                          grid-wide kernel with software barriers, 8 = clusters capped at 8 CTAs
 *   "stamps_sorted"        deskew: -1 = find out on every call, overlapped with the kernels (default), 1 / 0 =
 *                          the caller states they are / are not non-decreasing (see eskf_stamps_sorted) */
int eskf_ctx_set_option(eskf_ctx* ctx, const char* name, int64_t value);
/* Read-back of a few of them and of what the context found out about its device: "align_tuned_block"
 * (CTA size the large-cloud autotune settled on, 0 = not run yet), "align_autotune", "align_block",
 * "align_depth", "vox_cluster", "vox_cluster_max" (CTAs of the largest cluster the device places for the
 * one-cluster voxelize kernel, 0 = none), "stamps_sorted". */
int eskf_ctx_get_option(eskf_ctx* ctx, const char* name, int64_t* value);
/* CUDA-event timing on the context's own stream (bench.py's roofline leg) */
int eskf_ctx_timer_start(eskf_ctx* ctx);
int eskf_ctx_timer_stop(eskf_ctx* ctx, float* elapsed_ms);

/* ----------------------------------------------------------------- cloud */
/* Device-resident PointCloud (include/ESKF_LIO/Types.hpp:11-12: points_ +
 * covariances_), stored SoA in HBM. */
int eskf_cloud_create(eskf_ctx* ctx, size_t capacity, eskf_cloud** out);
int eskf_cloud_destroy(eskf_cloud* c);
/* cov may be NULL (raw scan).  Grows the cloud if n > capacity. */
int eskf_cloud_upload(eskf_cloud* c, const double* xyz, const double* cov, size_t n);
/* float32 xyz (the PointCloud2 wire format, include/ESKF_LIO/Subscriber.hpp:89-95),
 * widened to fp64 on the device exactly like the reference's cast. */
int eskf_cloud_upload_f32(eskf_cloud* c, const float* xyz, size_t n);
/* any of xyz / cov / src_index may be NULL; capacity in points */
int eskf_cloud_download(eskf_cloud* c, double* xyz, double* cov, uint32_t* src_index,
                        size_t capacity, size_t* n);
int eskf_cloud_size(eskf_cloud* c, size_t* n);
/* Open3D PointCloud::Transform (call sites src/Registration.cpp:13,27,
 * src/LocalMap.cpp:15, src/CloudPreprocessor.cpp:16): p <- T p ; C <- R C R^T */
int eskf_cloud_transform(eskf_cloud* c, const double T[16]);
/* dst <- src (device to device) */
int eskf_cloud_copy(eskf_cloud* dst, const eskf_cloud* src);

/* ------------------------------------------------------------------- map */
/* LocalMap(double voxelSize, size_t maxNumPointsPerVoxel)
 * (include/ESKF_LIO/LocalMap.hpp:54-61).  capacity_hint = expected number of
 * occupied voxels (the table grows by itself when it fills up). */
int eskf_map_create(eskf_ctx* ctx, double voxel_size, uint32_t max_points_per_voxel,
                    uint64_t capacity_hint, eskf_map** out);
int eskf_map_destroy(eskf_map* m);
/* the transform + insert loop of LocalMap::updateLocalMap
 * (src/LocalMap.cpp:15,47-58 + Voxel::addPoint LocalMap.hpp:72-87).
 * Host-buffer form: xyz/cov are body-frame inputs and are NOT modified. */
int eskf_map_insert(eskf_map* m, const double* xyz, const double* cov, size_t n,
                    const double T[16]);
/* device form: the cloud is transformed in place to the world frame, exactly
 * like the reference mutates the caller's cloud (src/LocalMap.cpp:15) */
int eskf_map_insert_cloud(eskf_map* m, eskf_cloud* cloud, const double T[16]);
/* the eviction sweep of updateLocalMap (src/LocalMap.cpp:62-69) with
 * needsPointRemoval (:149-154): erase voxels whose centre is farther than
 * dist_thresh (strict >) from pos. */
int eskf_map_evict(eskf_map* m, const double pos[3], double dist_thresh, uint64_t* removed);
int eskf_map_size(eskf_map* m, uint64_t* n_voxels);
int eskf_map_capacity(eskf_map* m, uint64_t* n_slots);
/* Re-hash the map into a table of exactly 2 x (occupied voxels) slots.  Bulk inserts grow the
 * table for the worst case (every incoming point a new voxel); after a map build this brings the
 * load factor back to 1/2, i.e. the smallest (most L2-resident) tag array.  Contents unchanged. */
int eskf_map_compact(eskf_map* m);
/* parity / debug: per-point lookup (getVoxelIndex + find, src/LocalMap.cpp:93-100,
 * 114-118).  Any output may be NULL. */
int eskf_map_query(eskf_map* m, const double* xyz, size_t n, int32_t* key_xyz, uint8_t* hit,
                   uint32_t* count, double* mean, double* cov);
/* parity / debug: dump all voxels sorted by (kx,ky,kz).  capacity in voxels. */
int eskf_map_export(eskf_map* m, size_t capacity, size_t* n, int32_t* key_xyz, uint32_t* count,
                    double* mean, double* cov);

/* ------------------------------------------------------------ preprocess */
/* CloudPreprocessor::process (src/CloudPreprocessor.cpp:10-23): T_il
 * transform, deskew against the state history (skipped when n_states == 0),
 * first-point-per-voxel downsample, 30-NN covariance + regularisation.
 * Outputs (capacity n points each, any may be NULL) are in ascending source
 * index order; *n_out = kept points.  The input buffer is NOT modified. */
/* Range crop ahead of the downsample (BASELINE.json north_star "voxel-grid downsample and range crop";
 * the reference's preprocessor has none, seam: src/CloudPreprocessor.cpp:16-20): a point of the sweep
 * survives iff min_range^2 <= x^2 + y^2 + z^2 <= max_range^2 in the LiDAR frame (as delivered, before
 * T_il); max_range 0 = unbounded; (0, 0) = off, the default.  T_il and the deskew act on every point
 * as before (segments and the sweep's end time come from the full stamp array); cropped points are
 * erased ahead of voxelDownsampleAndEstimateCovariances.  Source indices reported by the preprocessor
 * stay indices into the uncropped sweep.  Applies to every later eskf_preprocess* call on ctx. */
int eskf_ctx_set_range_crop(eskf_ctx* ctx, double min_range, double max_range);

/* 1 when the per-point stamps are non-decreasing (the reference's own precondition for deskew,
 * src/CloudPreprocessor.cpp:33), else 0.  The preprocessor needs to know: on sorted stamps the segment
 * ends of the reference's forward scan (:54-61) are binary searches, otherwise the scan itself is taken.
 * By default eskf_preprocess* assumes sorted stamps, launches its kernels and runs this check on the host
 * while the GPU works; if the stamps are not sorted it repeats the call with the scan's segments (the
 * first attempt leaves the raw cloud untouched).  A caller that knows states it with
 * eskf_ctx_set_option(ctx, "stamps_sorted", 0 | 1) and saves the pass (-1 = find out, the default); a
 * wrong 1 is not detected.  Host-only, needs no GPU. */
int eskf_stamps_sorted(const double* point_time, size_t n);
int eskf_preprocess(eskf_ctx* ctx, const double* xyz, const double* point_time, size_t n,
                    const double T_il[16], const eskf_state* states, size_t n_states,
                    double voxel_size, size_t* n_out, double* xyz_out, double* cov_out,
                    uint32_t* src_index_out);
/* device form: raw (device; may be clobbered, like the reference mutates
 * lidarMeas->cloud) -> out (device, with covariances). point_time is a host
 * array of raw->size doubles or NULL when n_states == 0. */
int eskf_preprocess_cloud(eskf_ctx* ctx, eskf_cloud* raw, const double* point_time,
                          const double T_il[16], const eskf_state* states, size_t n_states,
                          double voxel_size, eskf_cloud* out);
/* CloudPreprocessor::voxelDownsampleAndEstimateCovariances
 * (src/CloudPreprocessor.cpp:76-127) alone */
int eskf_downsample_cov(eskf_ctx* ctx, const double* xyz, size_t n, double voxel_size,
                        size_t* n_out, double* xyz_out, double* cov_out,
                        uint32_t* src_index_out);

/* ----------------------------------------------------------- registration */
/* ICP::align (src/Registration.cpp:7-35): the whole Gauss-Newton loop runs
 * on the device in one persistent kernel.  Host-buffer form. */
int eskf_align(eskf_ctx* ctx, const eskf_map* map, const double* xyz, const double* cov, size_t n,
               const double guess[16], const eskf_icp_params* params, double T_out[16],
               eskf_align_info* info);
/* device form: cloud must carry covariances; it is not modified */
int eskf_align_cloud(eskf_ctx* ctx, const eskf_map* map, const eskf_cloud* cloud,
                     const double guess[16], const eskf_icp_params* params, double T_out[16],
                     eskf_align_info* info);
/* the same in two halves, so that host work which does not depend on the result (the Kalman
 * gain of src/ErrorStateKF.cpp:136) overlaps the Gauss-Newton kernel: begin launches and
 * returns, end collects the pose.  One registration in flight per context; `info` at begin
 * only tells whether traces are wanted (may be NULL). */
int eskf_align_cloud_begin(eskf_ctx* ctx, const eskf_map* map, const eskf_cloud* cloud,
                           const double guess[16], const eskf_icp_params* params,
                           const eskf_align_info* info);
int eskf_align_end(eskf_ctx* ctx, double T_out[16], eskf_align_info* info);
/* n independent registrations (BASELINE.json configs[4]; SURVEY.md 8b eskf_align_batch): job i runs on
 * ctxs[i % n_ctx], to which maps[i] and clouds[i] must belong; the n_ctx contexts must be distinct.
 * One host thread keeps one registration in flight per context with the begin / end pair above, so
 * the launches of the other contexts overlap the kernel each end waits for (a 15k-point
 * registration occupies well under half of the SMs).  guesses, T_out: n x 16; infos: n entries or
 * NULL.  On an error no further job is started, the ones in flight are still collected (the
 * contexts stay usable) and the first error is returned. */
int eskf_align_batch(eskf_ctx* const* ctxs, int n_ctx, const eskf_map* const* maps,
                     const eskf_cloud* const* clouds, const double* guesses, size_t n,
                     const eskf_icp_params* params, double* T_out, eskf_align_info* infos);
/* parity / debug: one linearisation of the cloud posed at T
 * (LocalMap::correspondenceMatching src/LocalMap.cpp:78-112 + the accumulation
 * of ICP::computeTransform src/Registration.cpp:56-76).  hit: n bytes
 * (mode 1) or 7n bytes (mode 7), may be NULL.  fp64_math != 0 runs the
 * per-point algebra in fp64 instead of fp32. */
int eskf_linearize(eskf_ctx* ctx, const eskf_map* map, const double* xyz, const double* cov,
                   size_t n, const double T[16], int neighbor_mode, int fp64_math, double H36[36],
                   double b6[6], uint8_t* hit, uint64_t* n_corr);
/* a fixed number of Gauss-Newton iterations with no convergence exit (bench:
 * Mpts/s per GN iteration and the HBM roofline of the linearise kernel) */
int eskf_align_cloud_fixed(eskf_ctx* ctx, const eskf_map* map, const eskf_cloud* cloud,
                           const double guess[16], int iterations, int neighbor_mode,
                           double T_out[16], eskf_align_info* info);

/* sharded registration: this rank holds cloud = its point range of the source;
 * every iteration's 27 partial sums go through `allreduce` (called with the
 * context's stream) before the solve, so all ranks apply the identical step. */
int eskf_align_cloud_sharded(eskf_ctx* ctx, const eskf_map* map, const eskf_cloud* cloud,
                             const double guess[16], const eskf_icp_params* params,
                             eskf_allreduce_fn allreduce, void* user, int fixed_iterations,
                             double T_out[16], eskf_align_info* info);

/* ------------------------------------------------ multi-GPU, fused exchange
 * One process per GPU.  Each rank owns a small mailbox in its HBM that every
 * peer maps through CUDA IPC (NVLink / NVSwitch peer access).  Inside the
 * persistent Gauss-Newton kernel each rank stores its 28 partial sums straight
 * into every peer's mailbox and sums the `world` contributions in rank order,
 * so all ranks apply the bit-identical step with no collective call and no
 * host round trip per iteration.  Setup: create on every rank,
 * exchange the ESKF_COMM_HANDLE_BYTES-byte handles by any means (e.g.
 * torch.distributed.all_gather), connect.  world <= 16. */
int eskf_comm_create(eskf_ctx* ctx, int rank, int world, eskf_comm** out);
int eskf_comm_destroy(eskf_comm* c);
int eskf_comm_local_handle(eskf_comm* c, void* handle64);
/* handles: world x ESKF_COMM_HANDLE_BYTES bytes in rank order (own entry ignored) */
int eskf_comm_connect(eskf_comm* c, const void* handles);
/* ICP::align with the source cloud sharded by point range: `cloud` is this
 * rank's range (may be empty), the map is replicated.  COLLECTIVE: every rank
 * of the communicator must call it with the same guess / params; all ranks
 * return the same pose.  fixed_iterations > 0 ignores the convergence test. */
int eskf_align_cloud_p2p(eskf_ctx* ctx, const eskf_map* map, const eskf_cloud* cloud,
                         const double guess[16], const eskf_icp_params* params, eskf_comm* comm,
                         int fixed_iterations, double T_out[16], eskf_align_info* info);

#ifdef __cplusplus
}
#endif
#endif /* ESKF_GPU_H_ */
