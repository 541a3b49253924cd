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
  : voxelSize_(config.cloud_preprocessor.voxel_size), minRange_(config.cloud_preprocessor.min_range),
    maxRange_(config.cloud_preprocessor.max_range), deviceResident_(config.device_resident)
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
    // deskew (:25-74) walks the state deque from its begin, but a state whose
    // stamp is <= the first point time consumes no point (:54-65): only the
    // states from the last such one onwards are handed to the device path, so
    // the per-frame cost does not grow with the (never trimmed) history.
    const double * pointTime =
      lidarMeas->pointTimeView ? lidarMeas->pointTimeView : lidarMeas->pointTime.data();
    const std::size_t nTimes =
      lidarMeas->pointTimeView ? lidarMeas->pointTimeCount : lidarMeas->pointTime.size();
    std::size_t first = 0;
    if (!states.empty() && nTimes > 0) {
      const double t0 = pointTime[0];
      std::size_t lo = 0, hi = states.size();  // first state with timestamp > t0
      while (lo < hi) {
        const std::size_t mid = (lo + hi) / 2;
        if (states[mid].timestamp <= t0) {lo = mid + 1;} else {hi = mid;}
      }
      first = lo > 0 ? lo - 1 : 0;
    }
    std::vector<eskf_state> st(states.size() - first);
    for (std::size_t i = first; i < states.size(); ++i) {
      eskf_state & o = st[i - first];
      o.timestamp = states[i].timestamp;
      for (int k = 0; k < 3; ++k) {o.position[k] = states[i].position.v[k];}
      o.attitude_xyzw[0] = states[i].attitude.x;
      o.attitude_xyzw[1] = states[i].attitude.y;
      o.attitude_xyzw[2] = states[i].attitude.z;
      o.attitude_xyzw[3] = states[i].attitude.w;
    }
    const auto T = T_il_.matrix();
    gpuCheck(
      eskf_ctx_set_option(GpuContext::get(), "stamps_sorted", lidarMeas->stampsSorted),
      "eskf_ctx_set_option(stamps_sorted)");
    run(*lidarMeas->cloud, pointTime, T.data(), st.data(), st.size());
    gpuCheck(eskf_ctx_set_option(GpuContext::get(), "stamps_sorted", -1), "eskf_ctx_set_option(stamps_sorted)");
    lidarMeas->pointTime.clear();
    lidarMeas->pointTime.shrink_to_fit();
    lidarMeas->pointTimeView = nullptr;
    lidarMeas->pointTimeCount = 0;
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
    // (the context is shared by the host classes: the crop is this preprocessor's setting, stated per call)
    gpuCheck(eskf_ctx_set_range_crop(GpuContext::get(), minRange_, maxRange_), "eskf_ctx_set_range_crop");
    if (deviceResident_) {
      // raw scan: already in HBM (uploaded on arrival) or uploaded now
      std::shared_ptr<eskf_cloud> raw = cloud.device_;
      if (!raw) {
        raw = rawPool_.acquire(cloud.points_.size());
        gpuCheck(
          eskf_cloud_upload(
            raw.get(), reinterpret_cast<const double *>(cloud.points_.data()), nullptr,
            cloud.points_.size()), "eskf_cloud_upload");
      }
      std::shared_ptr<eskf_cloud> out = outPool_.acquire(1u << 16);
      gpuCheck(
        eskf_preprocess_cloud(
          GpuContext::get(), raw.get(), pointTime, T_il, states, nStates, voxelSize_, out.get()),
        "eskf_preprocess_cloud");
      cloud.device_ = std::move(out);
      cloud.points_.clear();
      cloud.covariances_.clear();
      return;
    }
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
  double minRange_, maxRange_;
  bool deviceResident_;
  Isometry3d T_il_;
  mutable DeviceCloudPool rawPool_, outPool_;
};
}  // namespace ESKF_LIO

#endif  // ESKF_LIO_B200_CLOUD_PREPROCESSOR_HPP_
