// Odometry.hpp — ROS-free drop-in for include/ESKF_LIO/Odometry.hpp +
// src/Odometry.cpp of the reference: the same members, the same per-frame
// call order  process -> ErrorStateKF::update{ICP::align} -> updateLocalMap
// (src/Odometry.cpp:73-87), with the three hot-path calls on the B200.
//
// Differences, all forced by the environment or by undefined behaviour in the
// reference (SURVEY.md sections 5 and 7):
//   - no visualiser (Open3D GUI); saveMapcloud exports voxel means + the trajectory;
//   - spinOnce() is one trip of run()'s busy loop, so a test or bench can
//     drive it deterministically; run() loops on it until setExit();
//   - the first scan is inserted with initialize = true (the reference tests
//     the keyframe gate against an uninitialised prevTransform_,
//     src/Odometry.cpp:61 + include/ESKF_LIO/LocalMap.hpp:113);
//   - the eviction period is tested against the LiDAR clock instead of
//     omp_get_wtime() (src/LocalMap.cpp:60), so runs are reproducible;
//   - with Config::device_resident a sweep can be fed as the float32 wire
//     format (feedLidar) and is uploaded to HBM when it arrives, overlapping
//     the copy with the host's IMU propagation.
#ifndef ESKF_LIO_B200_ODOMETRY_HPP_
#define ESKF_LIO_B200_ODOMETRY_HPP_

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <memory>
#include <vector>

#include "ESKF_LIO/CloudPreprocessor.hpp"
#include "ESKF_LIO/ErrorStateKF.hpp"
#include "ESKF_LIO/LocalMap.hpp"
#include "ESKF_LIO/SynchronizedQueue.hpp"

namespace ESKF_LIO
{
class Odometry
{
public:
  using ImuBuffer = typename std::shared_ptr<SynchronizedQueue<ImuMeasurementPtr>>;
  using CloudBuffer = typename std::shared_ptr<SynchronizedQueue<LidarMeasurementPtr>>;

  struct StageTimes  // the six accumulators of src/Odometry.cpp:11-15, seconds
  {
    double cloudPreprocessor = 0.0, filterUpdate = 0.0, mapUpdate = 0.0;
    double cloudPreprocessorMax = 0.0, filterUpdateMax = 0.0, mapUpdateMax = 0.0;
    int numFrames = 0;
    // not in the reference: the same three stages bracketed by CUDA events on the
    // context's stream (host gaps inside a frame included), milliseconds
    double deviceFrameMs = 0.0, deviceFrameMsLast = 0.0;
  };

  Odometry(const Config & config, ImuBuffer imuBuffer, CloudBuffer cloudBuffer)
  : imuBuffer_(std::move(imuBuffer))
    , cloudBuffer_(std::move(cloudBuffer))
    , kalmanFilter_(std::make_shared<ErrorStateKF>(config))
    , localMap_(std::make_shared<LocalMap>(config))
    , cloudPreprocessor_(std::make_shared<CloudPreprocessor>(config))
    , deviceResident_(config.device_resident)
  {
    localMap_->setClock([this] {return lidarClock_;});
    localMap_->setVerbose(false);
    warmUp(config);
  }

  // Not in the reference: one synthetic 64k-point frame through the three device calls on a scratch
  // map, so that scratch-buffer allocations, lazily loaded kernels and the first cooperative launch
  // are paid at construction instead of by the first real frames (the 1000-frame run of round 1 had one
  // frame at 3.9 ms; round 2 measured 2.8 ms on the first registration and <= 0.58 ms on every other frame).
  static void warmUp(const Config & config)
  {
    const auto guard = GpuContext::lock();
    eskf_ctx * ctx = GpuContext::get();
    constexpr std::size_t kSide = 256, kN = kSide * kSide;  // one sweep's worth of points
    std::vector<float> xyz(3 * kN);
    std::uint32_t lcg = 12345u;
    auto jitter = [&lcg] {
        lcg = lcg * 1664525u + 1013904223u;
        return (static_cast<float>(lcg >> 8) / 16777216.0f - 0.5f) * 0.04f;
      };
    for (std::size_t i = 0; i < kSide; ++i) {
      for (std::size_t j = 0; j < kSide; ++j) {  // a floor and, folded up at one edge, a wall
        float * p = &xyz[3 * (i * kSide + j)];
        const float u = 0.2f * static_cast<float>(i) - 25.0f, v = 0.2f * static_cast<float>(j) - 25.0f;
        p[0] = u + jitter();
        p[1] = (j < 200 ? v : 15.0f) + jitter();
        p[2] = (j < 200 ? -1.5f : -1.5f + 0.2f * static_cast<float>(j - 200)) + jitter();
      }
    }
    eskf_cloud * raw = nullptr, * ds = nullptr;
    eskf_map * map = nullptr;
    const double I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    gpuCheck(eskf_cloud_create(ctx, kN, &raw), "eskf_cloud_create");
    gpuCheck(eskf_cloud_create(ctx, kN, &ds), "eskf_cloud_create");
    gpuCheck(
      eskf_map_create(
        ctx, config.local_map.voxel_size, static_cast<uint32_t>(config.local_map.max_num_points_per_voxel),
        1u << 16, &map), "eskf_map_create");
    const eskf_icp_params prm = {config.registration.max_iteration, config.registration.neighbor_mode,
      config.registration.translation_sq_threshold, config.registration.cosine_threshold};
    double T[16];
    // (the second pass deskews against two identity states, so that the segment table and the scratch
    // buffer of the deskewing call exist before the first real sweep needs them)
    std::vector<double> stamps(kN);
    for (std::size_t i = 0; i < kN; ++i) {stamps[i] = 1e-6 * static_cast<double>(i);}
    eskf_state st[2] = {};
    st[0].timestamp = -1.0;
    st[1].timestamp = 1.0;
    st[0].attitude_xyzw[3] = st[1].attitude_xyzw[3] = 1.0;
    for (int pass = 0; pass < 2; ++pass) {  // (the second insert takes the "voxel exists" paths)
      gpuCheck(eskf_cloud_upload_f32(raw, xyz.data(), kN), "eskf_cloud_upload_f32");
      gpuCheck(eskf_ctx_set_range_crop(ctx, 0.0, 0.0), "eskf_ctx_set_range_crop");
      gpuCheck(
        eskf_preprocess_cloud(
          ctx, raw, pass == 1 ? stamps.data() : nullptr, I, pass == 1 ? st : nullptr, pass == 1 ? 2 : 0,
          config.cloud_preprocessor.voxel_size, ds),
        "eskf_preprocess_cloud");
      if (pass == 1) {gpuCheck(eskf_align_cloud(ctx, map, ds, I, &prm, T, nullptr), "eskf_align_cloud");}
      gpuCheck(eskf_map_insert_cloud(map, ds, I), "eskf_map_insert_cloud");
    }
    gpuCheck(eskf_ctx_sync(ctx), "eskf_ctx_sync");
    eskf_map_destroy(map);
    eskf_cloud_destroy(ds);
    eskf_cloud_destroy(raw);
  }

  // src/Odometry.cpp:9-110
  void run()
  {
    while (!exitFlag_) {spinOnce();}
  }

  void setExit() {exitFlag_ = true;}

  // include/ESKF_LIO/Odometry.hpp:37-40 (see LocalMap::save for what the cloud holds)
  void saveMapcloud(const std::string & cloud_path, const std::string & trajectory_path) const
  {
    localMap_->save(cloud_path, trajectory_path);
  }

  // One trip of the loop body (src/Odometry.cpp:17-97).  True when a LiDAR
  // frame was consumed (initialisation frame included).
  bool spinOnce()
  {
    auto imuMeas = imuBuffer_->popAll();  // :24
    while (!imuMeas.empty()) {            // :27-41
      if (initialized_) {kalmanFilter_->process(imuMeas.front());}
      kalmanFilter_->feedImu(imuMeas.front());
      imuMeas.pop_front();
    }
    if (lidarMeas_ == nullptr) {          // :44-49
      auto lidarMeas = cloudBuffer_->pop();
      if (lidarMeas.has_value()) {lidarMeas_ = lidarMeas.value();}
    }
    if (lidarMeas_ == nullptr) {return false;}
    const auto guard = GpuContext::lock();  // (feedLidar may be uploading from the subscriber's thread)

    const double lidarEndTime = lidarMeas_->endTime;
    lidarClock_ = lidarEndTime;
    if (!initialized_) {                  // :55-63
      initialized_ = true;
      kalmanFilter_->initialize(lidarEndTime);
      auto lidarMeasCopy = lidarMeas_;
      lidarMeas_ = nullptr;
      cloudPreprocessor_->process({}, lidarMeasCopy);
      localMap_->updateLocalMap(std::move(lidarMeasCopy->cloud), Isometry3d::Identity(), true);
      lastTransform_ = Isometry3d::Identity();
      return true;
    }
    if (kalmanFilter_->getLastStateTime() < lidarEndTime) {return false;}  // :65-69

    const auto & states = kalmanFilter_->getStates();
    gpuCheck(eskf_ctx_timer_start(GpuContext::get()), "eskf_ctx_timer_start");
    const double t0 = now();
    cloudPreprocessor_->process(states, lidarMeas_);                    // :74
    const double t1 = now();
    lastTransform_ = kalmanFilter_->update(*lidarMeas_, *localMap_);    // :79
    const double t2 = now();
    auto lidarMeasCopy = lidarMeas_;
    lidarMeas_ = nullptr;
    localMap_->updateLocalMap(std::move(lidarMeasCopy->cloud), lastTransform_);  // :86
    const double t3 = now();
    float deviceMs = 0.0f;
    gpuCheck(eskf_ctx_timer_stop(GpuContext::get(), &deviceMs), "eskf_ctx_timer_stop");
    times_.deviceFrameMs += deviceMs;
    times_.deviceFrameMsLast = deviceMs;

    ++times_.numFrames;                                                  // :89-96
    times_.cloudPreprocessor += t1 - t0;
    times_.filterUpdate += t2 - t1;
    times_.mapUpdate += t3 - t2;
    times_.cloudPreprocessorMax = std::max(times_.cloudPreprocessorMax, t1 - t0);
    times_.filterUpdateMax = std::max(times_.filterUpdateMax, t2 - t1);
    times_.mapUpdateMax = std::max(times_.mapUpdateMax, t3 - t2);
    return true;
  }

  // ---- not in the reference
  // The LidarSubscriber callback (include/ESKF_LIO/Subscriber.hpp:80-103)
  // without ROS: float32 x,y,z + double per-point time.  In device-resident
  // mode the sweep goes to HBM right away (pinned source => asynchronous).
  void feedLidar(const float * xyz, const double * pointTime, std::size_t n)
  {
    auto measurement = std::make_shared<LidarMeasurement>();
    auto cloud = std::make_shared<PointCloud>();
    if (deviceResident_) {
      const auto guard = GpuContext::lock();  // (the subscriber thread's entry into the shared context)
      cloud->device_ = rawPool_.acquire(n);
      gpuCheck(eskf_cloud_upload_f32(cloud->device_.get(), xyz, n), "eskf_cloud_upload_f32");
      measurement->pointTimeView = pointTime;  // the caller keeps both buffers until the frame is consumed
      measurement->pointTimeCount = n;
    } else {
      cloud->points_.resize(n);
      for (std::size_t i = 0; i < n; ++i) {
        cloud->points_[i] = Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
      }
      measurement->pointTime.assign(pointTime, pointTime + n);
    }
    measurement->startTime = pointTime[0];
    measurement->endTime = pointTime[n - 1];
    measurement->cloud = std::move(cloud);
    cloudBuffer_->push(std::move(measurement));
  }

  // A sweep that already lives in HBM (caller-owned eskf_cloud on the host
  // classes' context, xyz only); it is clobbered by process() like the
  // reference clobbers lidarMeas->cloud.
  void feedLidarDevice(eskf_cloud * raw, const double * pointTime, std::size_t n)
  {
    auto measurement = std::make_shared<LidarMeasurement>();
    auto cloud = std::make_shared<PointCloud>();
    cloud->device_ = std::shared_ptr<eskf_cloud>(raw, [](eskf_cloud *) {});  // borrowed
    measurement->pointTimeView = pointTime;
    measurement->pointTimeCount = n;
    measurement->startTime = pointTime[0];
    measurement->endTime = pointTime[n - 1];
    measurement->cloud = std::move(cloud);
    cloudBuffer_->push(std::move(measurement));
  }

  const Isometry3d & lastTransform() const {return lastTransform_;}
  const StageTimes & stageTimes() const {return times_;}
  const ErrorStateKF & kalmanFilter() const {return *kalmanFilter_;}
  const LocalMap & localMap() const {return *localMap_;}

private:
  Odometry() = delete;
  static double now()
  {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }

  bool initialized_ = false;
  std::atomic<bool> exitFlag_{false};  // setExit() comes from another thread (src/main.cpp:63-68)
  ImuBuffer imuBuffer_;
  CloudBuffer cloudBuffer_;
  std::shared_ptr<ErrorStateKF> kalmanFilter_;
  std::shared_ptr<LocalMap> localMap_;
  std::shared_ptr<CloudPreprocessor> cloudPreprocessor_;
  LidarMeasurementPtr lidarMeas_ = nullptr;
  bool deviceResident_;
  double lidarClock_ = 0.0;
  Isometry3d lastTransform_;
  StageTimes times_;
  DeviceCloudPool rawPool_;
};
}  // namespace ESKF_LIO

#endif  // ESKF_LIO_B200_ODOMETRY_HPP_
