// Types.hpp — boundary types of the hot path, restated without Eigen / Open3D
// (neither exists in this environment).  Mirrors include/ESKF_LIO/Types.hpp:11-52
// of the reference: PointCloud (= open3d::geometry::PointCloud: points_ +
// covariances_ + Transform()), ImuMeasurement, LidarMeasurement, State, and
// just enough of Eigen::Vector3d / Matrix3d / Isometry3d / Quaterniond for the
// three class interfaces.  Memory layout of the vectors equals the reference's
// std::vector<Eigen::Vector3d> / <Eigen::Matrix3d> (3 / 9 contiguous doubles;
// Matrix3d is stored ROW-major here, the C ABI's convention).
#ifndef ESKF_LIO_B200_TYPES_HPP_
#define ESKF_LIO_B200_TYPES_HPP_

#include <array>
#include <cmath>
#include <cstddef>
#include <deque>
#include <memory>
#include <vector>

struct eskf_cloud;  // include/eskf_gpu.h: device-resident cloud handle

namespace ESKF_LIO
{
struct Vector3d
{
  double v[3] = {0.0, 0.0, 0.0};
  Vector3d() = default;
  Vector3d(double x, double y, double z) : v{x, y, z} {}
  double & operator()(int i) {return v[i];}
  double operator()(int i) const {return v[i];}
  double & x() {return v[0];}
  double & y() {return v[1];}
  double & z() {return v[2];}
  double squaredNorm() const {return (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];}
  static Vector3d Zero() {return Vector3d();}
};

struct Matrix3d
{
  double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // row-major
  double & operator()(int r, int c) {return m[3 * r + c];}
  double operator()(int r, int c) const {return m[3 * r + c];}
  double trace() const {return (m[0] + m[4]) + m[8];}
  static Matrix3d Identity()
  {
    Matrix3d I;
    I.m[0] = I.m[4] = I.m[8] = 1.0;
    return I;
  }
};

struct Quaterniond
{
  double x = 0.0, y = 0.0, z = 0.0, w = 1.0;  // Eigen coefficient order x,y,z,w
  static Quaterniond Identity() {return Quaterniond();}
  Matrix3d toRotationMatrix() const
  {
    Matrix3d R;
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R.m[0] = 1.0 - (tyy + tzz); R.m[1] = txy - twz; R.m[2] = txz + twy;
    R.m[3] = txy + twz; R.m[4] = 1.0 - (txx + tzz); R.m[5] = tyz - twx;
    R.m[6] = txz - twy; R.m[7] = tyz + twx; R.m[8] = 1.0 - (txx + tyy);
    return R;
  }
};

// rigid transform; matrix() gives the row-major 4x4 the C ABI takes
struct Isometry3d
{
  Matrix3d R = Matrix3d::Identity();
  Vector3d t;
  static Isometry3d Identity() {return Isometry3d();}
  Matrix3d & linear() {return R;}
  const Matrix3d & linear() const {return R;}
  Vector3d & translation() {return t;}
  const Vector3d & translation() const {return t;}
  std::array<double, 16> matrix() const
  {
    return {R.m[0], R.m[1], R.m[2], t.v[0], R.m[3], R.m[4], R.m[5], t.v[1],
      R.m[6], R.m[7], R.m[8], t.v[2], 0.0, 0.0, 0.0, 1.0};
  }
  static Isometry3d fromMatrix(const double * T)
  {
    Isometry3d a;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {a.R.m[3 * i + j] = T[4 * i + j];}
      a.t.v[i] = T[4 * i + 3];
    }
    return a;
  }
  // Eigen Isometry3d * Isometry3d: linear = La*Lb ; translation = La*tb + ta
  Isometry3d operator*(const Isometry3d & b) const
  {
    Isometry3d r;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {
        r.R.m[3 * i + j] = (R.m[3 * i] * b.R.m[j] + R.m[3 * i + 1] * b.R.m[3 + j]) +
          R.m[3 * i + 2] * b.R.m[6 + j];
      }
      r.t.v[i] = ((R.m[3 * i] * b.t.v[0] + R.m[3 * i + 1] * b.t.v[1]) + R.m[3 * i + 2] * b.t.v[2]) +
        t.v[i];
    }
    return r;
  }
  Isometry3d inverse() const
  {
    Isometry3d r;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {r.R.m[3 * i + j] = R.m[3 * j + i];}
    }
    for (int i = 0; i < 3; ++i) {
      r.t.v[i] = -((r.R.m[3 * i] * t.v[0] + r.R.m[3 * i + 1] * t.v[1]) + r.R.m[3 * i + 2] * t.v[2]);
    }
    return r;
  }
};

// open3d::geometry::PointCloud as far as the hot path uses it
struct PointCloud
{
  std::vector<Vector3d> points_;
  std::vector<Matrix3d> covariances_;
  // Not in the reference: HBM mirror of this cloud.  With Config::device_resident
  // the three classes hand the frame to each other through this handle and the
  // host vectors stay empty (no PCIe round trip between process / align /
  // updateLocalMap); otherwise it is unused and the host vectors are the cloud.
  std::shared_ptr<eskf_cloud> device_;
  bool HasCovariances() const {return !points_.empty() && covariances_.size() == points_.size();}
  // Open3D PointCloud::Transform: p <- T p ; C <- R C R^T  (host-side, same
  // evaluation order as the device kernels: ((a0 b0 + a1 b1) + a2 b2) + t)
  PointCloud & Transform(const Isometry3d & T)
  {
    const double * R = T.R.m;
    for (auto & p : points_) {
      const double x = p.v[0], y = p.v[1], z = p.v[2];
      p.v[0] = ((R[0] * x + R[1] * y) + R[2] * z) + T.t.v[0];
      p.v[1] = ((R[3] * x + R[4] * y) + R[5] * z) + T.t.v[1];
      p.v[2] = ((R[6] * x + R[7] * y) + R[8] * z) + T.t.v[2];
    }
    for (auto & C : covariances_) {
      double A[9];
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
          A[3 * i + j] = (R[3 * i] * C.m[j] + R[3 * i + 1] * C.m[3 + j]) + R[3 * i + 2] * C.m[6 + j];
        }
      }
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
          C.m[3 * i + j] = (A[3 * i] * R[3 * j] + A[3 * i + 1] * R[3 * j + 1]) +
            A[3 * i + 2] * R[3 * j + 2];
        }
      }
    }
    return *this;
  }
};
using PointCloudPtr = std::shared_ptr<PointCloud>;

struct ImuMeasurement
{
  double timestamp;
  Vector3d angularVelocity;
  Vector3d acceleration;
};
using ImuMeasurementPtr = std::shared_ptr<ImuMeasurement>;

struct LidarMeasurement
{
  PointCloudPtr cloud;
  std::vector<double> pointTime;
  double startTime;
  double endTime;
  // Not in the reference (Config::device_resident): per-point times left in the caller's
  // buffer instead of being copied into pointTime (512 KB per sweep); valid until the
  // frame has been consumed.  Null => pointTime is used.
  const double * pointTimeView = nullptr;
  std::size_t pointTimeCount = 0;
  int stampsSorted = -1;  // 1 / 0: the driver knows the stamps are (not) non-decreasing, -1: unknown (the preprocessor finds out while the GPU works)
};
using LidarMeasurementPtr = std::shared_ptr<LidarMeasurement>;

// ESKF_LIO::State (Types.hpp:29-40): nominal state + the 18x18 error-state
// covariance P (row-major here), initialised to 1e-3 * Identity
struct State
{
  double timestamp = 0.0;
  Vector3d position;
  Vector3d velocity;
  Quaterniond attitude;
  Vector3d biasAccel;
  Vector3d biasGyro;
  Vector3d gravity;
  std::array<double, 18 * 18> P = identityP();

  static std::array<double, 18 * 18> identityP()
  {
    std::array<double, 18 * 18> p{};
    for (int i = 0; i < 18; ++i) {p[19 * i] = 1e-3;}
    return p;
  }
};

// plain-struct stand-in for the YAML::Node the reference's constructors take
// (config/hilti_config.yaml; yaml-cpp is not available here).  Same keys, same
// defaults.
struct Config
{
  struct {
    int max_iteration = 100;
    double translation_sq_threshold = 1.0e-6;
    double cosine_threshold = 0.9999;
    int neighbor_mode = 1;  // extension: 7 = DIRECT7
  } registration;
  struct {
    double voxel_size = 0.3;
    std::size_t max_num_points_per_voxel = 1000;
    struct {double translation_sq_threshold = 1.0e-2; double cosine_threshold = 0.985;} update;
    struct {bool enabled = true; double distance_threshold = 100.0; double removing_period = 10.0;}
    remove_distant_points;
    // Not in the reference: voxels the HBM table is sized for up front (2 slots per voxel,
    // 162 B per slot => 340 MB); it doubles by itself when it fills up, but a rebuild
    // allocates device memory and costs milliseconds, so the default covers a 100 m window.
    std::size_t capacity_hint = std::size_t(1) << 20;
  } local_map;
  struct {
    double voxel_size = 0.3;
    // range crop in the LiDAR frame (BASELINE.json north_star; not in the reference: off by default;
    // max_range 0 = unbounded) -- eskf_ctx_set_range_crop
    double min_range = 0.0;
    double max_range = 0.0;
  } cloud_preprocessor;
  struct {
    double quaternion[4] = {0.7071068, -0.7071068, 0.0, 0.0};  // x,y,z,w
    double translation[3] = {-0.001, -0.00855, 0.055};
  } lidar_extrinsics;
  struct {  // sensors.imu (hilti_config.yaml:2-17)
    double update_rate = 400.0;
    double bias_a[3] = {0.06080652138668933, 0.08353074835853214, 0.057072968234636895};
    double bias_g[3] = {-0.0015351229643790084, -0.0013449146576507546, 0.00030127855524786183};
    double gravity[3] = {0.01165152782783894, -0.008749296634685332, 9.804989173462031};
    double accel_noise_density[3] = {105.0, 105.0, 135.0};
    double accel_zero_g_offset = 20.0;
    double gyro_noise_density = 0.014;
    double gyro_zero_rate_offset = 1.0;
  } imu;
  struct {double translation_noise = 1.0e-6; double rotation_noise = 1.0e-6;} kalman_filter;
  // Not in the reference: keep each frame in HBM between process / align /
  // updateLocalMap (PointCloud::device_) instead of round-tripping host vectors.
  bool device_resident = false;
};

}  // namespace ESKF_LIO

#endif  // ESKF_LIO_B200_TYPES_HPP_
