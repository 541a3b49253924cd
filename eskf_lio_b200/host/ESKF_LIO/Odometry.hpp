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
#include <memory>

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
