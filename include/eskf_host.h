/*
 * eskf_host.h — C ABI of the ROS-free host driver (libeskf_host.so): the
 * reference's Odometry + ErrorStateKF (include/ESKF_LIO/Odometry.hpp,
 * src/Odometry.cpp:9-110, src/ErrorStateKF.cpp) as C++ host classes
 * (the headers under eskf_lio_b200/host/ESKF_LIO) whose three hot-path calls go to the
 * B200 through include/eskf_gpu.h.  This layer exists so that a non-C++
 * harness (bench.py, the tests) can replay a sensor log through exactly the
 * classes a C++ user of the reference would instantiate; a C++ user includes
 * the headers directly.
 *
 * Conventions as in eskf_gpu.h: extern "C", plain pointers, int status
 * (0 = OK; the message of the last failure is eskf_host_last_error()).
 * There is no CPU fallback: creation fails without a CUDA device.
 */
#ifndef ESKF_HOST_H_
#define ESKF_HOST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct eskf_odom eskf_odom;

/* The keys of config/hilti_config.yaml the path reads (same defaults:
 * eskf_odom_default_config). */
typedef struct {
  double imu_update_rate;                 /* sensors.imu.update_rate */
  double bias_a[3], bias_g[3], gravity[3];
  double accel_noise_density[3];
  double accel_zero_g_offset, gyro_noise_density, gyro_zero_rate_offset;
  double translation_noise, rotation_noise;               /* kalman_filter.update */
  double lidar_quaternion_xyzw[4], lidar_translation[3];  /* sensors.lidar.extrinsics */
  double map_voxel_size;                  /* local_map.* */
  uint64_t max_points_per_voxel;
  double update_translation_sq_threshold, update_cosine_threshold;
  int32_t remove_enabled;
  double remove_distance_threshold, remove_period;
  double preprocess_voxel_size;           /* cloud_preprocessor.voxel_size */
  int32_t max_iteration, neighbor_mode;   /* registration.* (neighbor_mode 7 = DIRECT7 extension) */
  double icp_translation_sq_threshold, icp_cosine_threshold;
  int32_t device_resident;                /* 1: frames stay in HBM between the three calls */
  uint64_t map_capacity_hint;             /* not in the reference: voxels the HBM table is sized for up
                                           * front (it doubles by itself when it fills up) */
} eskf_odom_config;

typedef struct {
  uint64_t frames;        /* frames through the steady-state branch (src/Odometry.cpp:89) */
  uint64_t n_states;      /* ErrorStateKF::getStates().size() */
  uint64_t map_voxels;
  uint64_t last_removed;  /* voxels evicted by the last sweep */
  int32_t last_iterations; /* Gauss-Newton iterations of the last ICP::align */
  int32_t last_inserted;   /* 1 when the last frame passed the keyframe gate */
  double stage_avg_ms[3], stage_max_ms[3]; /* preprocess, filter update, map update (src/Odometry.cpp:99-109) */
  double stage_sum_ms[3];     /* the same three accumulators, not averaged */
  double device_frame_ms_sum; /* the three stages bracketed by CUDA events on the context's stream */
  double device_frame_ms_last;
} eskf_odom_info;

const char* eskf_host_last_error(void);
void eskf_odom_default_config(eskf_odom_config* cfg);
/* Odometry::Odometry (include/ESKF_LIO/Odometry.hpp:22-34) on CUDA device `device`
 * (the process-wide context of the host classes; first call wins). */
int eskf_odom_create(const eskf_odom_config* cfg, int device, eskf_odom** out);
int eskf_odom_destroy(eskf_odom* o);
/* ImuSubscriber::imuCallback (include/ESKF_LIO/Subscriber.hpp:38-52) */
int eskf_odom_feed_imu(eskf_odom* o, double t, const double gyro[3], const double acc[3]);
/* LidarSubscriber::cloudCallback (Subscriber.hpp:80-103): float32 x,y,z per
 * point + one double time stamp per point (ascending).  With device_resident
 * the sweep is copied to HBM asynchronously and the times are read in place:
 * keep xyz (pinned memory, eskf_host_alloc) and point_time valid until the frame
 * has been consumed by eskf_odom_spin_once. */
int eskf_odom_feed_lidar(eskf_odom* o, const float* xyz, const double* point_time, size_t n);
/* a sweep already resident in HBM: raw_cloud is an eskf_cloud* (include/eskf_gpu.h,
 * xyz only) created on eskf_odom_context(); it is clobbered when the frame is
 * consumed and stays owned by the caller (device_resident mode only) */
int eskf_odom_feed_lidar_cloud(eskf_odom* o, void* raw_cloud, const double* point_time, size_t n);
/* the eskf_ctx* (include/eskf_gpu.h) all host classes of this process share */
int eskf_odom_context(eskf_odom* o, void** ctx);
/* one trip of the loop of Odometry::run (src/Odometry.cpp:17-97);
 * *consumed = 1 when a LiDAR frame went through */
int eskf_odom_spin_once(eskf_odom* o, int* consumed);
/* the transform the last frame was inserted with (row-major 4x4) */
int eskf_odom_last_pose(eskf_odom* o, double T[16]);
/* newest filter state: t, p, v, q(xyzw), ba, bg, g = 20 doubles; P324 nullable */
int eskf_odom_last_state(eskf_odom* o, double out20[20], double* P324);
int eskf_odom_info_get(eskf_odom* o, eskf_odom_info* out);
/* the map handle (include/eskf_gpu.h) for parity checks; owned by the odometry */
int eskf_odom_map(eskf_odom* o, void** map_handle);
/* kernels launched so far by the host classes' context (bench "gpu_launches") */
int eskf_odom_launch_count(eskf_odom* o, uint64_t* n);

/* ---- flat binary sensor log (eskf_lio_b200/host/ESKF_LIO/SensorLog.hpp): the wire format of
 * the reference's two subscribers without ROS (include/ESKF_LIO/Subscriber.hpp:38-52 sensor_msgs/Imu,
 * :80-103 sensor_msgs/PointCloud2 with float32 x, y, z + float64 "timestamp" per point), records in
 * callback order.  "ESKFLOG1", u32 version 1, u32 0; then {u32 type, u32 count, payload}:
 * type 1 = IMU {f64 stamp, f64 gyro[3], f64 acc[3]}, type 2 = sweep of count x {f32 x, y, z, f64 stamp}. */
/* Replays a log through the odometry: every record is delivered the way its callback would
 * (feed_imu / feed_lidar) followed by trips of Odometry::run's loop.  poses (nullable): row-major
 * 4x4 of every frame that went through, at most `capacity` of them; *n_frames = all of them. */
int eskf_odom_replay_log(eskf_odom* o, const char* path, double* poses, size_t capacity, size_t* n_frames);
/* counts = {IMU records, sweeps, points}, stamps = {earliest, latest}; needs no GPU */
int eskf_log_summary(const char* path, uint64_t counts[3], double stamps[2]);
typedef struct eskf_log_writer eskf_log_writer;
int eskf_log_writer_open(const char* path, eskf_log_writer** out);
int eskf_log_writer_imu(eskf_log_writer* w, double stamp, const double gyro[3], const double acc[3]);
int eskf_log_writer_lidar(eskf_log_writer* w, const float* xyz, const double* point_time, size_t n);
int eskf_log_writer_close(eskf_log_writer* w);

#ifdef __cplusplus
}
#endif
#endif /* ESKF_HOST_H_ */
