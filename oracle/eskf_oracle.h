/*
 * eskf_oracle.h — C ABI of the CPU ORACLE for the ESKF_LIO hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under eskf_lio_b200/ (the product) may
 * include, link or dlopen this.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker or as
 * the CPU baseline — never as the thing shipped.
 *
 * PARITY: the reference (LimHaeryong/ESKF_LIO) ships no tests, golden vectors
 * or fixtures, and cannot be built as shipped in this environment (Eigen,
 * Open3D, yaml-cpp and rclcpp are all absent, no network).  This file is a
 * dependency-free restatement of the reference algorithm; it is pinned by
 *   (0) the reference's OWN sources (Registration, LocalMap, CloudPreprocessor,
 *       Utils, ErrorStateKF, Odometry .cpp), compiled where they lie
 *       against from-scratch shims of the Eigen / Open3D / yaml-cpp API subset
 *       they use (oracle/refshim, `make ref`): bit-identical voxel keys, kept
 *       sets, map statistics, correspondence sets and per-point J^T W J terms,
 *       poses / filter states to 1e-12, the reference's Odometry::run loop
 *       reproducing the golden trajectory (tests/test_reference_shim.py).  That
 *       pins the CONTROL FLOW AND FORMULAS to the reference's text; the
 *       third-party arithmetic underneath (Eigen's LDLT / inverse / SVD /
 *       quaternions, Open3D's Transform / k-NN / ComputeCovariance) is restated
 *       in the shims as it is here, so in that respect parity stays UNPINNED;
 *   (1) an independent NumPy/SciPy second oracle (oracle/np_oracle.py),
 *   (2) analytic known-answer tests (tests/test_oracle_kat.py),
 *   (3) committed fixtures generated from it (tests/golden/).
 *
 * Arithmetic contract (the "bit-exact" sets are defined against THIS):
 *   - all math fp64, compiled -O3 -ffp-contract=off (the reference builds with
 *     no -march / no fast-math, CMakeLists.txt:6-9 => no FMA contraction);
 *   - dot products of length 3 are evaluated ((a0*b0 + a1*b1) + a2*b2);
 *   - point transform: ((R0*x + R1*y) + R2*z) + t  (Open3D PointCloud::Transform
 *     = 4x4 homogeneous product then division by w == 1.0, which is exact);
 *   - covariance transform: (R*C)*R^T, two 3x3 products;
 *   - voxel index: (int) floor(p / voxel_size) with a true fp64 division
 *     (src/LocalMap.cpp:114-118, src/CloudPreprocessor.cpp:129-133).
 * Third-party arithmetic that is NOT under /root/reference (Open3D, Eigen,
 * both unpinned versions) is restated from its published algorithm; see the
 * comments at each function.
 *
 * Output-order conventions where the reference's order is unspecified
 * (unordered_map iteration order, OpenMP critical order): ascending source
 * index for the downsample output and for correspondences; ascending packed
 * voxel key (kx, ky, kz) for the map export.
 */
#ifndef ESKF_ORACLE_H_
#define ESKF_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_map orc_map;

/* registration.{max_iteration, translation_sq_threshold, cosine_threshold}
 * (config/hilti_config.yaml:50-53, include/ESKF_LIO/Registration.hpp:23-27);
 * neighbor_mode 1 = reference (single voxel), 7 = DIRECT7 extension
 * (not in the reference). */
typedef struct {
  int32_t max_iteration;
  int32_t neighbor_mode;
  double translation_sq_threshold;
  double cosine_threshold;
} orc_icp_params;

typedef struct {
  int32_t iterations; /* number of computeTransform calls performed */
  int32_t converged;
} orc_align_info;

/* The fields of ESKF_LIO::State (include/ESKF_LIO/Types.hpp:31-52) that the
 * hot path reads (deskew only): timestamp, position, attitude (x,y,z,w). */
typedef struct {
  double timestamp;
  double position[3];
  double attitude_xyzw[4];
} orc_state;

int orc_num_threads(void);
void orc_set_num_threads(int n);

/* ---- Utils (src/Utils.cpp) ------------------------------------------- */
void orc_skew(const double v[3], double out9[9]);                 /* Utils.cpp:5-11  */
void orc_compute_J(const double r[3], double out9[9]);            /* Utils.cpp:40-54 */
void orc_rotvec_to_matrix(const double r[3], double out9[9]);     /* Utils.cpp:28-32 */
void orc_se3_to_SE3(const double se3[6], double T16[16]);         /* Utils.cpp:56-63 */
void orc_quat_to_matrix(const double q_xyzw[4], double out9[9]);  /* Eigen toRotationMatrix */
void orc_interpolate_SE3(const orc_state* s1, const orc_state* s2, double t,
                         double T16[16]);                          /* Utils.cpp:65-75 */

/* ---- Open3D PointCloud::Transform (call sites Registration.cpp:13,27,
 *      LocalMap.cpp:15, CloudPreprocessor.cpp:16) ---------------------- */
void orc_transform_cloud(double* xyz, double* cov /* nullable */, size_t n,
                         const double T16[16]);

/* ---- voxel index (LocalMap.cpp:114-118 / CloudPreprocessor.cpp:129-133) */
void orc_voxel_index(const double* xyz, size_t n, double voxel_size, int32_t* out_xyz);

/* ---- Registration (src/Registration.cpp) ----------------------------- */
void orc_jtj_jtr(const double p[3], const double mu[3], const double C9[9],
                 double H36[36], double b6[6]);                    /* :83-102 */
void orc_ldlt_solve6(const double H36[36], const double b6[6], double x6[6]); /* Eigen LDLT */
int orc_convergence_check(const double T16[16], double trans_sq_thr, double cos_thr); /* :37-50 */

/* one linearisation at the cloud's CURRENT pose (no transform applied):
 * correspondenceMatching (LocalMap.cpp:78-112) + the accumulation half of
 * computeTransform (Registration.cpp:56-76).  hit (nullable): n bytes (mode 1)
 * or 7n bytes (mode 7, offset-major per point).  Returns the number of
 * correspondences. */
size_t orc_linearize(const orc_map* map, const double* xyz, const double* cov, size_t n,
                     int neighbor_mode, double H36[36], double b6[6], uint8_t* hit);

/* ICP::align (Registration.cpp:7-35).  Optional traces (nullable), each
 * max_iteration long: trace_H [it*36], trace_b [it*6], trace_ncorr [it],
 * trace_step [it*16]. */
int orc_align(const orc_map* map, const double* xyz, const double* cov, size_t n,
              const double guess16[16], const orc_icp_params* prm, double T_out16[16],
              orc_align_info* info, double* trace_H, double* trace_b,
              uint64_t* trace_ncorr, double* trace_step);

/* ---- LocalMap (src/LocalMap.cpp, include/ESKF_LIO/LocalMap.hpp) ------ */
orc_map* orc_map_create(double voxel_size, uint64_t max_points_per_voxel);
void orc_map_destroy(orc_map* m);
/* update.{translation_sq_threshold,cosine_threshold}, remove_distant_points.* */
void orc_map_set_update_params(orc_map* m, double trans_sq_thr, double cos_thr,
                               int remove_enabled, double distance_thr, double remove_period);
/* updateLocalMap (LocalMap.cpp:10-76).  xyz/cov are transformed IN PLACE to
 * the world frame (the reference mutates the caller's cloud, :15).  `now` is
 * the clock the eviction period is tested against (the reference reads
 * omp_get_wtime(), :60,70; injected here so runs are reproducible).
 * Returns 1 if points were inserted, 0 if gated out by needsMapUpdate.
 * removed (nullable) receives the number of evicted voxels or 0. */
int orc_map_update(orc_map* m, double* xyz, double* cov, size_t n, const double T16[16],
                   int initialize, double now, uint64_t* removed);
/* the insert loop alone (LocalMap.cpp:47-58), points already in world frame */
void orc_map_insert(orc_map* m, const double* xyz, const double* cov, size_t n);
uint64_t orc_map_evict(orc_map* m, const double pos[3], double distance_thr); /* :62-69,149-154 */
int orc_needs_map_update(const double prev16[16], const double cur16[16],
                         double trans_sq_thr, double cos_thr);      /* :132-147 */
uint64_t orc_map_size(const orc_map* m);
/* dump sorted by packed key (kx, ky, kz lexicographic) */
void orc_map_export(const orc_map* m, int32_t* keys_xyz, uint64_t* count, double* mean,
                    double* cov);
void orc_map_query(const orc_map* m, const double* xyz, size_t n, int32_t* keys_xyz,
                   uint8_t* hit, uint64_t* count, double* mean, double* cov);

/* ---- CloudPreprocessor (src/CloudPreprocessor.cpp) -------------------- */
/* exact k-NN with a KD-tree (Open3D KDTreeFlann / nanoflann semantics: exact,
 * results ascending by distance).  out_idx/out_d2: nq*k, rows padded with -1. */
void orc_knn(const double* xyz, size_t n, const double* queries, size_t nq, int k,
             int32_t* out_idx, double* out_d2);
void orc_knn_bruteforce(const double* xyz, size_t n, const double* queries, size_t nq, int k,
                        int32_t* out_idx, double* out_d2);
/* Open3D utility::ComputeCovariance + the SVD regularisation
 * (CloudPreprocessor.cpp:115-123) */
void orc_cov_from_indices(const double* xyz, const int32_t* idx, int k, double out9[9]);
void orc_regularize_cov(const double C9[9], double out9[9]);
/* voxelDownsampleAndEstimateCovariances (:76-127).  Outputs sized >= n.
 * Returns the number of kept points (ascending source index order). */
size_t orc_downsample_cov(const double* xyz, size_t n, double voxel_size, double* out_xyz,
                          double* out_cov, uint32_t* out_src_index);
/* deskew (:25-74), in place.  Returns 0, or -1 when the reference would
 * dereference past the state deque (no state <= lidarEndTime). */
int orc_deskew(double* xyz, const double* point_time, size_t n, const orc_state* states,
               size_t n_states);
/* process (:10-23): T_il, deskew (if n_states>0), downsample+cov.  xyz is
 * clobbered (transformed in place, like the reference). */
long orc_preprocess(double* xyz, const double* point_time, size_t n, const double T_il16[16],
                    const orc_state* states, size_t n_states, double voxel_size,
                    double* out_xyz, double* out_cov, uint32_t* out_src_index);

/* ---- callers either side of the hot path (oracle/odom_oracle.cpp) -------
 * ErrorStateKF (src/ErrorStateKF.cpp) + the call order of Odometry::run
 * (src/Odometry.cpp:16-98), ROS-free.  Keys/defaults of config/hilti_config.yaml. */
typedef struct orc_odom orc_odom;
typedef struct {
  double imu_update_rate;          /* sensors.imu.update_rate            */
  double bias_a[3], bias_g[3], gravity[3];
  double accel_noise_density[3];
  double accel_zero_g_offset, gyro_noise_density, gyro_zero_rate_offset;
  double translation_noise, rotation_noise;      /* kalman_filter.update */
  double lidar_quaternion_xyzw[4], lidar_translation[3]; /* sensors.lidar.extrinsics */
  double map_voxel_size;           /* local_map.*                        */
  uint64_t max_points_per_voxel;
  double update_translation_sq_threshold, update_cosine_threshold;
  int32_t remove_enabled;
  double remove_distance_threshold, remove_period;
  double preprocess_voxel_size;    /* cloud_preprocessor.voxel_size      */
  int32_t max_iteration, neighbor_mode;          /* registration.*       */
  double icp_translation_sq_threshold, icp_cosine_threshold;
} orc_odom_config;
typedef struct {
  uint64_t frames, n_states, map_voxels, last_kept, last_removed;
  int32_t last_iterations, last_inserted;
  double stage_avg_ms[3], stage_max_ms[3]; /* preprocess, filter update, map update (Odometry.cpp:99-109) */
} orc_odom_info_t;
void orc_odom_default_config(orc_odom_config* c);
orc_odom* orc_odom_create(const orc_odom_config* cfg);
void orc_odom_destroy(orc_odom* o);
void orc_odom_feed_imu(orc_odom* o, double t, const double gyro[3], const double acc[3]);
void orc_odom_feed_lidar(orc_odom* o, const double* xyz, const double* point_time, size_t n,
                         double start_time, double end_time);
/* one trip of the loop of Odometry::run: 1 = a LiDAR frame was consumed */
int orc_odom_spin_once(orc_odom* o);
void orc_odom_last_pose(const orc_odom* o, double T16[16]);
void orc_odom_info(const orc_odom* o, orc_odom_info_t* out);
/* newest state: t, p, v, q(xyzw), ba, bg, g = 20 doubles; P324 nullable */
void orc_odom_last_state(const orc_odom* o, double out20[20], double* P324);
const orc_map* orc_odom_map(const orc_odom* o);
void orc_kf_process(orc_odom* o, double t, const double gyro[3], const double acc[3]); /* ErrorStateKF.cpp:76-113 */
void orc_rotation_matrix_to_vector(const double R9[9], double r3[3]);                  /* Utils.cpp:22-26 */
/* ErrorStateKF::update (ErrorStateKF.cpp:115-162) with the ICP pose given by the caller */
void orc_kf_update_with_observation(orc_odom* o, double lidar_end, const double obs16[16],
                                    double guess16_out[16], double T_out16[16]);

#ifdef __cplusplus
}
#endif
#endif /* ESKF_ORACLE_H_ */
