// ref_capi.cpp — C ABI over the REFERENCE'S OWN classes, compiled from the sources where they
// lie under /root/reference against the API shims in oracle/refshim/include (Eigen, Open3D and
// yaml-cpp are absent from this environment).  TEST INFRASTRUCTURE: used only by tests/ to pin
// the oracle's restatement of the reference's control flow and formulas (oracle/Makefile target
// _ref/libref_shim.so; outputs stay in oracle/_ref/, no reference source is copied).
//
// Matrices cross this boundary row-major (numpy), 4x4 poses as 16 doubles, covariances as 9.
#include <cstdint>
#include <cstring>
#include <deque>
#include <chrono>
#include <memory>
#include <thread>
#include <vector>

#include <omp.h>

// the wrapper reads a few private members (voxelGrid_, T_il_, computeTransform, ...): same
// translation-unit trick a white-box test would use; class layouts are unaffected
#define private public
#include "ESKF_LIO/CloudPreprocessor.hpp"
#include "ESKF_LIO/ErrorStateKF.hpp"
#include "ESKF_LIO/LocalMap.hpp"
#include "ESKF_LIO/Odometry.hpp"
#include "ESKF_LIO/Registration.hpp"
#include "ESKF_LIO/Utils.hpp"
#undef private

using namespace ESKF_LIO;

namespace
{
Eigen::Matrix4d m4(const double * T)
{
  Eigen::Matrix4d m;
  for (int i = 0; i < 4; ++i) {for (int j = 0; j < 4; ++j) {m(i, j) = T[4 * i + j];}}
  return m;
}
Eigen::Isometry3d iso(const double * T)
{
  Eigen::Isometry3d a;
  a.matrix() = m4(T);
  return a;
}
void out16(const Eigen::Isometry3d & a, double * T)
{
  for (int i = 0; i < 4; ++i) {for (int j = 0; j < 4; ++j) {T[4 * i + j] = a.matrix()(i, j);}}
}
Eigen::Matrix3d m3(const double * c)
{
  Eigen::Matrix3d m;
  for (int i = 0; i < 3; ++i) {for (int j = 0; j < 3; ++j) {m(i, j) = c[3 * i + j];}}
  return m;
}
void out9(const Eigen::Matrix3d & m, double * c)
{
  for (int i = 0; i < 3; ++i) {for (int j = 0; j < 3; ++j) {c[3 * i + j] = m(i, j);}}
}
PointCloudPtr make_cloud(const double * xyz, const double * cov, size_t n)
{
  auto c = std::make_shared<PointCloud>();
  c->points_.resize(n);
  for (size_t i = 0; i < n; ++i) {c->points_[i] = Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);}
  if (cov) {
    c->covariances_.resize(n);
    for (size_t i = 0; i < n; ++i) {c->covariances_[i] = m3(cov + 9 * i);}
  }
  return c;
}

struct RefConfig  // the keys of config/hilti_config.yaml the five classes read
{
  double voxel_map, voxel_pre;
  uint64_t max_points_per_voxel;
  double update_tsq, update_cos;
  int32_t remove_enabled;
  double remove_distance, remove_period;
  int32_t max_iteration;
  double icp_tsq, icp_cos;
  double lidar_quat_xyzw[4], lidar_trans[3];
  double imu_rate, bias_a[3], bias_g[3], gravity[3], accel_noise_density[3];
  double accel_zero_g_offset, gyro_noise_density, gyro_zero_rate_offset;
  double translation_noise, rotation_noise;
};

YAML::Node to_yaml(const RefConfig & c)
{
  YAML::Node root;
  auto vec = [](const double * p, int n) {return std::vector<double>(p, p + n);};
  YAML::Node & lm = root.child("local_map");
  lm.child("voxel_size").set(c.voxel_map);
  lm.child("max_num_points_per_voxel").set(static_cast<double>(c.max_points_per_voxel));
  lm.child("update").child("translation_sq_threshold").set(c.update_tsq);
  lm.child("update").child("cosine_threshold").set(c.update_cos);
  YAML::Node & rm = lm.child("remove_distant_points");
  rm.child("enabled").set(static_cast<double>(c.remove_enabled));
  rm.child("distance_threshold").set(c.remove_distance);
  rm.child("removing_period").set(c.remove_period);
  root.child("cloud_preprocessor").child("voxel_size").set(c.voxel_pre);
  YAML::Node & reg = root.child("registration");
  reg.child("max_iteration").set(static_cast<double>(c.max_iteration));
  reg.child("translation_sq_threshold").set(c.icp_tsq);
  reg.child("cosine_threshold").set(c.icp_cos);
  YAML::Node & sensors = root.child("sensors");
  YAML::Node & ext = sensors.child("lidar").child("extrinsics");
  ext.child("quaternion").set(vec(c.lidar_quat_xyzw, 4));
  ext.child("translation").set(vec(c.lidar_trans, 3));
  YAML::Node & imu = sensors.child("imu");
  imu.child("update_rate").set(c.imu_rate);
  YAML::Node & prm = imu.child("intrinsics").child("parameters");
  prm.child("bias_a").set(vec(c.bias_a, 3));
  prm.child("bias_g").set(vec(c.bias_g, 3));
  prm.child("gravity").set(vec(c.gravity, 3));
  prm.child("accel_noise_density").set(vec(c.accel_noise_density, 3));
  prm.child("accel_zero_g_offset").set(c.accel_zero_g_offset);
  prm.child("gyro_noise_density").set(c.gyro_noise_density);
  prm.child("gyro_zero_rate_offset").set(c.gyro_zero_rate_offset);
  YAML::Node & kf = root.child("kalman_filter").child("update");
  kf.child("translation_noise").set(c.translation_noise);
  kf.child("rotation_noise").set(c.rotation_noise);
  return root;
}

struct RefState  // what tests exchange for a ESKF_LIO::State
{
  double timestamp, position[3], velocity[3], attitude_xyzw[4], bias_a[3], bias_g[3], gravity[3];
};
State to_state(const RefState & s)
{
  State o;
  o.timestamp = s.timestamp;
  o.position = Eigen::Vector3d(s.position[0], s.position[1], s.position[2]);
  o.velocity = Eigen::Vector3d(s.velocity[0], s.velocity[1], s.velocity[2]);
  o.attitude = Eigen::Quaterniond(s.attitude_xyzw[3], s.attitude_xyzw[0], s.attitude_xyzw[1], s.attitude_xyzw[2]);
  o.biasAccel = Eigen::Vector3d(s.bias_a[0], s.bias_a[1], s.bias_a[2]);
  o.biasGyro = Eigen::Vector3d(s.bias_g[0], s.bias_g[1], s.bias_g[2]);
  o.gravity = Eigen::Vector3d(s.gravity[0], s.gravity[1], s.gravity[2]);
  return o;
}
void from_state(const State & s, RefState * o, double * P)
{
  o->timestamp = s.timestamp;
  for (int i = 0; i < 3; ++i) {
    o->position[i] = s.position(i);
    o->velocity[i] = s.velocity(i);
    o->bias_a[i] = s.biasAccel(i);
    o->bias_g[i] = s.biasGyro(i);
    o->gravity[i] = s.gravity(i);
  }
  o->attitude_xyzw[0] = s.attitude.x();
  o->attitude_xyzw[1] = s.attitude.y();
  o->attitude_xyzw[2] = s.attitude.z();
  o->attitude_xyzw[3] = s.attitude.w();
  if (P) {for (int i = 0; i < 18; ++i) {for (int j = 0; j < 18; ++j) {P[18 * i + j] = s.P(i, j);}}}
}
}  // namespace

extern "C" {

int ref_tree_redux(void) {return ESHIM_TREE_REDUX;}
void ref_set_num_threads(int n) {omp_set_num_threads(n);}

// ------------------------------------------------------------------ Utils
void ref_skew(const double v[3], double o[9]) {out9(Utils::skewSymmetric(Eigen::Vector3d(v[0], v[1], v[2])), o);}
void ref_rotvec_to_matrix(const double r[3], double R[9])
{
  out9(Utils::rotationVectorToMatrix(Eigen::Vector3d(r[0], r[1], r[2])), R);
}
void ref_rotation_matrix_to_vector(const double R[9], double r[3])
{
  const Eigen::Vector3d v = Utils::rotationMatrixToVector(m3(R));
  for (int i = 0; i < 3; ++i) {r[i] = v(i);}
}
void ref_se3_to_SE3(const double se3[6], double T[16])
{
  Eigen::Vector<double, 6> s;
  for (int i = 0; i < 6; ++i) {s(i) = se3[i];}
  out16(Utils::se3ToSE3(s), T);
}
void ref_interpolate_SE3(const RefState * s1, const RefState * s2, double t, double T[16])
{
  out16(Utils::interpolateSE3(to_state(*s1), to_state(*s2), t), T);
}

// ------------------------------------------------------------ PointCloud
void ref_transform_cloud(double * xyz, double * cov, size_t n, const double T[16])
{
  auto c = make_cloud(xyz, cov, n);
  c->Transform(m4(T));
  for (size_t i = 0; i < n; ++i) {
    for (int d = 0; d < 3; ++d) {xyz[3 * i + d] = c->points_[i](d);}
    if (cov) {out9(c->covariances_[i], cov + 9 * i);}
  }
}

// ---------------------------------------------------- CloudPreprocessor
// process() (src/CloudPreprocessor.cpp:10-23) on one sweep.  T_il overrides the extrinsics the
// constructor derives from the quaternion, so that both sides get the identical matrix.  The output
// order is the reference's own (unordered_map iteration order).  Returns the number of kept points.
long long ref_preprocess(
  const RefConfig * cfg, const double * xyz, const double * point_time, size_t n, const double T_il[16],
  const RefState * states, size_t n_states, double * out_xyz, double * out_cov)
{
  CloudPreprocessor pre(to_yaml(*cfg));
  pre.T_il_.matrix() = m4(T_il);
  auto meas = std::make_shared<LidarMeasurement>();
  meas->cloud = make_cloud(xyz, nullptr, n);
  meas->pointTime.assign(point_time, point_time + n);
  meas->startTime = n ? point_time[0] : 0.0;
  meas->endTime = n ? point_time[n - 1] : 0.0;
  std::deque<State> st;
  for (size_t i = 0; i < n_states; ++i) {st.push_back(to_state(states[i]));}
  pre.process(st, meas);
  const size_t m = meas->cloud->points_.size();
  for (size_t i = 0; i < m; ++i) {
    for (int d = 0; d < 3; ++d) {out_xyz[3 * i + d] = meas->cloud->points_[i](d);}
    out9(meas->cloud->covariances_[i], out_cov + 9 * i);
  }
  return static_cast<long long>(m);
}

void ref_voxel_index(const RefConfig * cfg, const double * xyz, size_t n, int32_t * out)
{
  LocalMap map(to_yaml(*cfg), open3d::camera::PinholeCameraParameters(), false);
  for (size_t i = 0; i < n; ++i) {
    const Eigen::Vector3i k = map.getVoxelIndex(Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    for (int d = 0; d < 3; ++d) {out[3 * i + d] = k(d);}
  }
}

// -------------------------------------------------------------- LocalMap
struct ref_map
{
  std::unique_ptr<LocalMap> map;
  std::unique_ptr<ICP> icp;
};

ref_map * ref_map_create(const RefConfig * cfg)
{
  auto * m = new ref_map();
  const YAML::Node y = to_yaml(*cfg);
  m->map = std::make_unique<LocalMap>(y, open3d::camera::PinholeCameraParameters(), false);
  m->icp = std::make_unique<ICP>(y);
  return m;
}
void ref_map_destroy(ref_map * m) {delete m;}

// updateLocalMap (src/LocalMap.cpp:10-76); xyz / cov come back transformed like the caller's cloud
void ref_map_update(ref_map * m, double * xyz, double * cov, size_t n, const double T[16], int initialize)
{
  auto c = make_cloud(xyz, cov, n);
  m->map->updateLocalMap(c, iso(T), initialize != 0);
  for (size_t i = 0; i < n; ++i) {
    for (int d = 0; d < 3; ++d) {xyz[3 * i + d] = c->points_[i](d);}
    out9(c->covariances_[i], cov + 9 * i);
  }
}
uint64_t ref_map_size(const ref_map * m) {return m->map->voxelGrid_.size();}
// voxels in the unordered_map's order: key, numPoints, mean, covariance
void ref_map_export(const ref_map * m, int32_t * keys, uint64_t * count, double * mean, double * cov)
{
  size_t i = 0;
  for (const auto & kv : m->map->voxelGrid_) {
    for (int d = 0; d < 3; ++d) {keys[3 * i + d] = kv.first(d); mean[3 * i + d] = kv.second.mean(d);}
    count[i] = kv.second.numPoints;
    out9(kv.second.covariance, cov + 9 * i);
    ++i;
  }
}
int ref_needs_map_update(ref_map * m, const double prev[16], const double cur[16])
{
  m->map->prevTransform_ = iso(prev);
  return m->map->needsMapUpdate(iso(cur)) ? 1 : 0;
}
// correspondenceMatching (src/LocalMap.cpp:78-112), single-threaded so that the output keeps the
// input order; returns the number of correspondences
uint64_t ref_correspondences(
  const ref_map * m, const double * xyz, const double * cov, size_t n, double * src_xyz, double * src_cov,
  double * map_mean, double * map_cov)
{
  auto c = make_cloud(xyz, cov, n);
  const int nt = omp_get_max_threads();
  omp_set_num_threads(1);
  auto corr = m->map->correspondenceMatching(c->points_, c->covariances_);
  omp_set_num_threads(nt);
  auto & [sp, sc, mp, mc] = corr;
  for (size_t i = 0; i < sp.size(); ++i) {
    for (int d = 0; d < 3; ++d) {src_xyz[3 * i + d] = sp[i](d); map_mean[3 * i + d] = mp[i](d);}
    out9(sc[i], src_cov + 9 * i);
    out9(mc[i], map_cov + 9 * i);
  }
  return sp.size();
}

// -------------------------------------------------------------------- ICP
void ref_jtj_jtr(const ref_map * m, const double p[3], const double mu[3], const double C[9], double H[36], double b[6])
{
  auto r = m->icp->computeJTJAndJTr(Eigen::Vector3d(p[0], p[1], p[2]), Eigen::Vector3d(mu[0], mu[1], mu[2]), m3(C));
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < 6; ++j) {H[6 * i + j] = r.first(i, j);}
    b[i] = r.second(i);
  }
}
// one Gauss-Newton step at the cloud's current position: correspondenceMatching + computeTransform
// (src/Registration.cpp:16-19, 52-81); single-threaded => input-order summation
uint64_t ref_gn_step(const ref_map * m, const double * xyz, const double * cov, size_t n, double T_step[16])
{
  auto c = make_cloud(xyz, cov, n);
  const int nt = omp_get_max_threads();
  omp_set_num_threads(1);
  auto corr = m->map->correspondenceMatching(c->points_, c->covariances_);
  const uint64_t nc = std::get<0>(corr).size();
  out16(m->icp->computeTransform(corr), T_step);
  omp_set_num_threads(nt);
  return nc;
}
int ref_convergence_check(const ref_map * m, const double T[16]) {return m->icp->convergenceCheck(iso(T)) ? 1 : 0;}
// ICP::align (src/Registration.cpp:7-35); returns the sticky converged_ flag (reset before the call)
int ref_align(ref_map * m, const double * xyz, const double * cov, size_t n, const double guess[16], double T_out[16])
{
  auto c = make_cloud(xyz, cov, n);
  m->icp->converged_ = false;
  out16(m->icp->align(*c, *m->map, iso(guess)), T_out);
  return m->icp->converged_ ? 1 : 0;
}

// ----------------------------------------------------------- ErrorStateKF
struct ref_eskf
{
  std::unique_ptr<ErrorStateKF> kf;
};
ref_eskf * ref_eskf_create(const RefConfig * cfg)
{
  auto * e = new ref_eskf();
  e->kf = std::make_unique<ErrorStateKF>(to_yaml(*cfg));
  return e;
}
void ref_eskf_destroy(ref_eskf * e) {delete e;}
void ref_eskf_feed_imu(ref_eskf * e, double t, const double gyro[3], const double acc[3])
{
  auto imu = std::make_shared<ImuMeasurement>();
  imu->timestamp = t;
  imu->angularVelocity = Eigen::Vector3d(gyro[0], gyro[1], gyro[2]);
  imu->acceleration = Eigen::Vector3d(acc[0], acc[1], acc[2]);
  e->kf->feedImu(imu);
}
void ref_eskf_initialize(ref_eskf * e, double lidar_end_time) {e->kf->initialize(lidar_end_time);}
// process() of one sample without queueing it (src/ErrorStateKF.cpp:73-113)
void ref_eskf_process(ref_eskf * e, double t, const double gyro[3], const double acc[3])
{
  auto imu = std::make_shared<ImuMeasurement>();
  imu->timestamp = t;
  imu->angularVelocity = Eigen::Vector3d(gyro[0], gyro[1], gyro[2]);
  imu->acceleration = Eigen::Vector3d(acc[0], acc[1], acc[2]);
  e->kf->process(imu);
}
// update() (src/ErrorStateKF.cpp:115-162) with a preprocessed cloud against `map`
void ref_eskf_update(
  ref_eskf * e, ref_map * map, const double * xyz, const double * cov, size_t n, double lidar_end_time,
  double T_out[16])
{
  LidarMeasurement meas;
  meas.cloud = make_cloud(xyz, cov, n);
  meas.startTime = lidar_end_time;
  meas.endTime = lidar_end_time;
  out16(e->kf->update(meas, *map->map), T_out);
}
uint64_t ref_eskf_num_states(const ref_eskf * e) {return e->kf->getStates().size();}
void ref_eskf_state(const ref_eskf * e, long long index, RefState * out, double * P /* 18x18 or NULL */)
{
  const auto & st = e->kf->getStates();
  const size_t i = index < 0 ? st.size() + index : static_cast<size_t>(index);
  from_state(st[i], out, P);
}

// ---------------------------------------------------------------- Odometry
// Odometry::run (src/Odometry.cpp:9-98) over a whole log.  Both queues are filled before run()
// starts, so the loop's outcome does not depend on thread timing: the first trip queues every IMU
// sample and initialises on sweep 0 (which processes the whole IMU log), every later trip finds
// the filter ahead of the sweep, rolls back to its end time and replays the rest — the states at
// and before each sweep's end are those of the online interleaving.  run() is stopped with
// setExit() once the last sweep has reached updateLocalMap (the trip in flight completes).
// Outputs: the pose of every updateLocalMap call (LocalMap::trajectory_, one per sweep), the
// filter's last state at or before `state_time`, the map's voxel count.
long long ref_odom_run(
  const RefConfig * cfg, const double * imu /* n_imu x (t, gyro xyz, acc xyz) */, size_t n_imu, const double * xyz,
  const double * point_time, const uint64_t * n_per_scan, size_t n_scans, double * poses /* n_scans x 16 */,
  double state_time, RefState * state, double * P, uint64_t * map_voxels)
{
  auto imuBuf = std::make_shared<SynchronizedQueue<ImuMeasurementPtr>>();
  auto cloudBuf = std::make_shared<SynchronizedQueue<LidarMeasurementPtr>>();
  for (size_t i = 0; i < n_imu; ++i) {
    auto m = std::make_shared<ImuMeasurement>();
    m->timestamp = imu[7 * i];
    m->angularVelocity = Eigen::Vector3d(imu[7 * i + 1], imu[7 * i + 2], imu[7 * i + 3]);
    m->acceleration = Eigen::Vector3d(imu[7 * i + 4], imu[7 * i + 5], imu[7 * i + 6]);
    imuBuf->push(m);
  }
  size_t off = 0;
  for (size_t k = 0; k < n_scans; ++k) {
    const size_t n = n_per_scan[k];
    auto meas = std::make_shared<LidarMeasurement>();
    meas->cloud = make_cloud(xyz + 3 * off, nullptr, n);
    meas->pointTime.assign(point_time + off, point_time + off + n);
    meas->startTime = point_time[off];
    meas->endTime = point_time[off + n - 1];
    cloudBuf->push(meas);
    off += n;
  }
  Odometry odom(to_yaml(*cfg), imuBuf, cloudBuf, open3d::camera::PinholeCameraParameters(), false);
  std::thread th([&odom]() {odom.run();});
  const auto t0 = std::chrono::steady_clock::now();
  bool timed_out = false;
  for (;;) {
    // (a racy read of a vector's size from the polling thread: benign here — it only decides when
    //  to raise the exit flag; everything read after join() is synchronised by it)
    if (odom.localMap_->trajectory_.parameters_.size() >= n_scans) {break;}
    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(600)) {timed_out = true; break;}
    std::this_thread::sleep_for(std::chrono::milliseconds(2));
  }
  odom.setExit();
  th.join();
  if (timed_out) {return -1;}
  const auto & traj = odom.localMap_->trajectory_.parameters_;
  for (size_t k = 0; k < n_scans; ++k) {
    for (int i = 0; i < 4; ++i) {for (int j = 0; j < 4; ++j) {poses[16 * k + 4 * i + j] = traj[k].extrinsic_(i, j);}}
  }
  const auto & st = odom.kalmanFilter_->getStates();
  size_t pick = 0;
  for (size_t i = 0; i < st.size(); ++i) {if (st[i].timestamp <= state_time) {pick = i;}}
  from_state(st[pick], state, P);
  *map_voxels = odom.localMap_->voxelGrid_.size();
  return static_cast<long long>(st.size());
}

}  // extern "C"
