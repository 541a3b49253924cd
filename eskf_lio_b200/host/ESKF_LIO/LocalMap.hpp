// LocalMap.hpp — drop-in for include/ESKF_LIO/LocalMap.hpp + src/LocalMap.cpp
// of the reference: same public names and argument meaning, the voxel map
// itself lives in HBM behind the C ABI (include/eskf_gpu.h).
// Out of scope: the Open3D visualiser and the raw per-voxel
// point list; save() therefore exports the voxel MEANS instead of the raw points.
#ifndef ESKF_LIO_B200_LOCAL_MAP_HPP_
#define ESKF_LIO_B200_LOCAL_MAP_HPP_

#include <chrono>
#include <cstdio>
#include <fstream>
#include <functional>
#include <iostream>
#include <string>
#include <limits>
#include <tuple>

#include "ESKF_LIO/GpuContext.hpp"
#include "ESKF_LIO/Types.hpp"

namespace ESKF_LIO
{
class LocalMap
{
public:
  using PointVector = typename std::vector<Vector3d>;
  using CovarianceVector = typename std::vector<Matrix3d>;
  using Correspondence = typename std::tuple<PointVector, CovarianceVector, PointVector,
      CovarianceVector>;

  // LocalMap(const YAML::Node &, ...) (LocalMap.hpp:28-52)
  explicit LocalMap(const Config & config, bool visualize = false)
  : voxelSize_(config.local_map.voxel_size)
    , maxNumPointsPerVoxel_(config.local_map.max_num_points_per_voxel)
    , translationSquaredThreshold_(config.local_map.update.translation_sq_threshold)
    , cosineThreshold_(config.local_map.update.cosine_threshold)
    , removeDistantPoints_(config.local_map.remove_distant_points.enabled)
    , distanceThreshold_(config.local_map.remove_distant_points.distance_threshold)
    , removePeriod_(config.local_map.remove_distant_points.removing_period)
    , capacityHint_(config.local_map.capacity_hint)
  {
    (void)visualize;
    create();
  }

  // LocalMap(double voxelSize, size_t maxNumPointsPerVoxel, bool visualize = false)
  // (LocalMap.hpp:54-61; the reference leaves the update / removal members
  // uninitialised here — they get the YAML defaults instead)
  LocalMap(double voxelSize, std::size_t maxNumPointsPerVoxel, bool visualize = false)
  : voxelSize_(voxelSize), maxNumPointsPerVoxel_(maxNumPointsPerVoxel)
  {
    (void)visualize;
    create();
  }

  ~LocalMap()
  {
    const auto guard = GpuContext::lock();
    eskf_map_destroy(map_);
  }
  LocalMap(const LocalMap &) = delete;
  LocalMap & operator=(const LocalMap &) = delete;

  // LocalMap::updateLocalMap (src/LocalMap.cpp:10-76)
  void updateLocalMap(PointCloudPtr cloud, const Isometry3d & transform, bool initialize = false)
  {
    const auto T = transform.matrix();
    trajectory_.push_back(transform);  // :16-18 (PinholeCameraTrajectory of the reference)
    lastInserted_ = false;
    if (initialize == false && needsMapUpdate(transform) == false) {
      // :15 — the caller's cloud always ends up in the world frame
      if (cloud->device_) {
        gpuCheck(eskf_cloud_transform(cloud->device_.get(), T.data()), "eskf_cloud_transform");
      } else {
        cloud->Transform(transform);
      }
      prevTransform_ = transform;   // :40
      return;
    }
    lastInserted_ = true;
    if (cloud->device_) {  // frame already in HBM: transformed in place on the device
      gpuCheck(eskf_map_insert_cloud(map_, cloud->device_.get(), T.data()), "eskf_map_insert_cloud");
    } else {
      gpuCheck(
        eskf_map_insert(
          map_, reinterpret_cast<const double *>(cloud->points_.data()),
          reinterpret_cast<const double *>(cloud->covariances_.data()), cloud->points_.size(),
          T.data()), "eskf_map_insert");
      cloud->Transform(transform);
    }
    const double now = clock_();
    if (removeDistantPoints_ && now - currentRemoveTime_ > removePeriod_) {  // :60
      uint64_t removed = 0;
      gpuCheck(eskf_map_evict(map_, transform.t.v, distanceThreshold_, &removed), "eskf_map_evict");
      currentRemoveTime_ = now;
      lastRemoved_ = removed;
      if (verbose_) {std::cout << "removed " << removed << " voxels\n";}  // :71
    }
    prevTransform_ = transform;  // :74
  }

  // LocalMap::correspondenceMatching (src/LocalMap.cpp:78-112); output in
  // ascending source index (the reference's order is OpenMP arrival order)
  Correspondence correspondenceMatching(
    const PointVector & points, const CovarianceVector & covariances) const
  {
    Correspondence correspondence;
    auto & [srcPoints, srcCovs, mapPoints, mapCovs] = correspondence;
    const std::size_t n = points.size();
    std::vector<uint8_t> hit(n);
    std::vector<double> mean(3 * n), cov(9 * n);
    gpuCheck(
      eskf_map_query(
        map_, reinterpret_cast<const double *>(points.data()), n, nullptr, hit.data(), nullptr,
        mean.data(), cov.data()), "eskf_map_query");
    for (std::size_t i = 0; i < n; ++i) {
      if (!hit[i]) {continue;}
      srcPoints.push_back(points[i]);
      srcCovs.push_back(covariances[i]);
      mapPoints.emplace_back(mean[3 * i], mean[3 * i + 1], mean[3 * i + 2]);
      Matrix3d C;
      for (int k = 0; k < 9; ++k) {C.m[k] = cov[9 * i + k];}
      mapCovs.push_back(C);
    }
    return correspondence;
  }

  // LocalMap::save (src/LocalMap.cpp:156-167).  The reference writes every raw point it kept
  // per voxel with Open3D; the HBM table keeps per-voxel statistics only, so the cloud written
  // here holds one point per voxel (its mean), as an ASCII PCD, and the trajectory is written
  // in the layout of Open3D's PinholeCameraTrajectory JSON (extrinsic = the pose, column-major).
  void save(const std::string & cloud_path, const std::string & trajectory_path) const
  {
    const std::size_t n = size();
    std::vector<int32_t> keys(3 * n);
    std::vector<uint32_t> count(n);
    std::vector<double> mean(3 * n), cov(9 * n);
    std::size_t m = 0;
    gpuCheck(
      eskf_map_export(map_, n, &m, keys.data(), count.data(), mean.data(), cov.data()),
      "eskf_map_export");
    std::ofstream pcd(cloud_path);
    pcd << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 8 8 8\n"
        << "TYPE F F F\nCOUNT 1 1 1\nWIDTH " << m << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS "
        << m << "\nDATA ascii\n";
    char line[128];
    for (std::size_t i = 0; i < m; ++i) {
      std::snprintf(line, sizeof line, "%.9g %.9g %.9g\n", mean[3 * i], mean[3 * i + 1], mean[3 * i + 2]);
      pcd << line;
    }
    std::ofstream js(trajectory_path);
    js << "{\n\t\"class_name\" : \"PinholeCameraTrajectory\",\n\t\"parameters\" :\n\t[";
    for (std::size_t k = 0; k < trajectory_.size(); ++k) {
      const auto M = trajectory_[k].matrix();  // row-major
      js << (k ? "," : "") << "\n\t\t{\n\t\t\t\"class_name\" : \"PinholeCameraParameters\",\n\t\t\t\"extrinsic\" : [";
      for (int c = 0; c < 4; ++c) {
        for (int r = 0; r < 4; ++r) {
          std::snprintf(line, sizeof line, "%s%.17g", (c || r) ? ", " : " ", M[4 * r + c]);
          js << line;
        }
      }
      js << " ],\n\t\t\t\"intrinsic\" : { \"height\" : -1, \"width\" : -1, \"intrinsic_matrix\" : "
         << "[ 0, 0, 0, 0, 0, 0, 0, 0, 0 ] },\n\t\t\t\"version_major\" : 1,\n\t\t\t\"version_minor\" : 0\n\t\t}";
    }
    js << "\n\t],\n\t\"version_major\" : 1,\n\t\"version_minor\" : 0\n}\n";
  }

  const std::vector<Isometry3d> & trajectory() const {return trajectory_;}

  // ---- not in the reference: handles / test seams
  eskf_map * handle() const {return map_;}
  std::size_t size() const
  {
    uint64_t n = 0;
    gpuCheck(eskf_map_size(map_, &n), "eskf_map_size");
    return n;
  }
  // the eviction period is tested against omp_get_wtime() in the reference
  // (:60,70); inject a clock to make runs reproducible
  void setClock(std::function<double()> clock) {clock_ = std::move(clock);}
  void setVerbose(bool v) {verbose_ = v;}
  bool lastInserted() const {return lastInserted_;}
  uint64_t lastRemoved() const {return lastRemoved_;}

private:
  void create()
  {
    context_ = GpuContext::share();  // the map's context must outlive the map (static / global odometries)
    gpuCheck(
      eskf_map_create(
        GpuContext::get(), voxelSize_, static_cast<uint32_t>(maxNumPointsPerVoxel_), capacityHint_,
        &map_), "eskf_map_create");
    clock_ = [] {
        return std::chrono::duration<double>(
          std::chrono::steady_clock::now().time_since_epoch()).count();
      };
  }

  // LocalMap::needsMapUpdate (src/LocalMap.cpp:132-147), vs the previous FRAME
  bool needsMapUpdate(const Isometry3d & transform) const
  {
    const Isometry3d moved = prevTransform_.inverse() * transform;
    const double cosine = 0.5 * (moved.R.trace() - 1.0);
    if (cosine < cosineThreshold_) {return true;}
    if (moved.t.squaredNorm() > translationSquaredThreshold_) {return true;}
    return false;
  }

  double voxelSize_;
  std::size_t maxNumPointsPerVoxel_;
  double translationSquaredThreshold_ = 1.0e-2;
  double cosineThreshold_ = 0.985;
  bool removeDistantPoints_ = true;
  double distanceThreshold_ = 100.0;
  double removePeriod_ = 10.0;
  std::size_t capacityHint_ = std::size_t(1) << 16;
  double currentRemoveTime_ = std::numeric_limits<double>::lowest();  // LocalMap.hpp:40
  Isometry3d prevTransform_;  // uninitialised in the reference; identity here
  std::vector<Isometry3d> trajectory_;
  std::function<double()> clock_;
  bool verbose_ = true;
  bool lastInserted_ = false;
  uint64_t lastRemoved_ = 0;
  std::shared_ptr<GpuContext> context_;  // declared before map_: destroyed after it
  eskf_map * map_ = nullptr;
};
}  // namespace ESKF_LIO

#endif  // ESKF_LIO_B200_LOCAL_MAP_HPP_
