// CloudPreprocessor.hpp — drop-in for include/ESKF_LIO/CloudPreprocessor.hpp +
// src/CloudPreprocessor.cpp of the reference.
#ifndef ESKF_LIO_B200_CLOUD_PREPROCESSOR_HPP_
#define ESKF_LIO_B200_CLOUD_PREPROCESSOR_HPP_

#include "ESKF_LIO/GpuContext.hpp"
#include "ESKF_LIO/Types.hpp"

namespace ESKF_LIO
{
class CloudPreprocessor
{
public:
  explicit CloudPreprocessor(const Config & config)
  : voxelSize_(config.cloud_preprocessor.voxel_size)
  {
    Quaterniond q;
    q.x = config.lidar_extrinsics.quaternion[0];
    q.y = config.lidar_extrinsics.quaternion[1];
    q.z = config.lidar_extrinsics.quaternion[2];
    q.w = config.lidar_extrinsics.quaternion[3];
    T_il_.R = q.toRotationMatrix();
    T_il_.t = Vector3d(
      config.lidar_extrinsics.translation[0], config.lidar_extrinsics.translation[1],
      config.lidar_extrinsics.translation[2]);
  }

  // CloudPreprocessor::process (src/CloudPreprocessor.cpp:10-23): mutates
  // lidarMeas->cloud in place and frees pointTime, like the reference
  void process(const std::deque<State> & states, LidarMeasurementPtr lidarMeas) const
  {
    std::vector<eskf_state> st(states.size());
    for (std::size_t i = 0; i < states.size(); ++i) {
      st[i].timestamp = states[i].timestamp;
      for (int k = 0; k < 3; ++k) {st[i].position[k] = states[i].position.v[k];}
      st[i].attitude_xyzw[0] = states[i].attitude.x;
      st[i].attitude_xyzw[1] = states[i].attitude.y;
      st[i].attitude_xyzw[2] = states[i].attitude.z;
      st[i].attitude_xyzw[3] = states[i].attitude.w;
    }
    const auto T = T_il_.matrix();
    run(*lidarMeas->cloud, lidarMeas->pointTime.data(), T.data(), st.data(), st.size());
    lidarMeas->pointTime.clear();
    lidarMeas->pointTime.shrink_to_fit();
  }

  // CloudPreprocessor::voxelDownsampleAndEstimateCovariances (:76-127)
  void voxelDownsampleAndEstimateCovariances(PointCloud & cloud) const
  {
    run(cloud, nullptr, nullptr, nullptr, 0);
  }

private:
  CloudPreprocessor() = delete;

  void run(
    PointCloud & cloud, const double * pointTime, const double * T_il, const eskf_state * states,
    std::size_t nStates) const
  {
    const std::size_t n = cloud.points_.size();
    std::vector<Vector3d> pointsDown(n);
    std::vector<Matrix3d> covs(n);
    std::size_t m = 0;
    gpuCheck(
      eskf_preprocess(
        GpuContext::get(), reinterpret_cast<const double *>(cloud.points_.data()), pointTime, n,
        T_il, states, nStates, voxelSize_, &m, reinterpret_cast<double *>(pointsDown.data()),
        reinterpret_cast<double *>(covs.data()), nullptr), "eskf_preprocess");
    pointsDown.resize(m);
    covs.resize(m);
    std::swap(cloud.points_, pointsDown);  // :126
    std::swap(cloud.covariances_, covs);
  }

  double voxelSize_;
  Isometry3d T_il_;
};
}  // namespace ESKF_LIO

#endif  // ESKF_LIO_B200_CLOUD_PREPROCESSOR_HPP_
