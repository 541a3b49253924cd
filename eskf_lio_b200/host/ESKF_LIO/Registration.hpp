// Registration.hpp — drop-in for include/ESKF_LIO/Registration.hpp +
// src/Registration.cpp of the reference (class ICP).  The Gauss-Newton loop
// runs on the device (eskf_align); the host keeps the sticky converged_ flag
// and the "ICP not converged!" message (src/Registration.cpp:23,30-32).
#ifndef ESKF_LIO_B200_REGISTRATION_HPP_
#define ESKF_LIO_B200_REGISTRATION_HPP_

#include <iostream>

#include "ESKF_LIO/LocalMap.hpp"

namespace ESKF_LIO
{
class ICP
{
public:
  explicit ICP(const Config & config)
  : maxIteration_(config.registration.max_iteration)
    , translationSquaredThreshold_(config.registration.translation_sq_threshold)
    , cosineThreshold_(config.registration.cosine_threshold)
    , neighborMode_(config.registration.neighbor_mode)
  {
  }

  // ICP::align (src/Registration.cpp:7-35)
  Isometry3d align(const PointCloud & cloud, const LocalMap & localMap, const Isometry3d & guess)
  {
    eskf_icp_params prm = {maxIteration_, neighborMode_, translationSquaredThreshold_,
      cosineThreshold_};
    eskf_align_info info = {};
    double T[16];
    const auto G = guess.matrix();
    if (cloud.device_) {  // frame already in HBM (Config::device_resident)
      gpuCheck(
        eskf_align_cloud(
          GpuContext::get(), localMap.handle(), cloud.device_.get(), G.data(), &prm, T, &info),
        "eskf_align_cloud");
    } else {
      gpuCheck(
        eskf_align(
        GpuContext::get(), localMap.handle(), reinterpret_cast<const double *>(cloud.points_.data()),
        reinterpret_cast<const double *>(cloud.covariances_.data()), cloud.points_.size(), G.data(),
          &prm, T, &info), "eskf_align");
    }
    lastIterations_ = info.iterations;
    if (info.converged) {converged_ = true;}
    if (!converged_) {std::cout << "ICP not converged!\n";}
    return Isometry3d::fromMatrix(T);
  }

  // ---- not in the reference: align() in two halves, so the caller's host work that does not
  // depend on the result overlaps the Gauss-Newton kernel (device-resident clouds; a host
  // cloud is simply registered synchronously in alignBegin)
  void alignBegin(const PointCloud & cloud, const LocalMap & localMap, const Isometry3d & guess)
  {
    if (!cloud.device_) {
      pendingHost_ = true;
      hostResult_ = align(cloud, localMap, guess);
      return;
    }
    pendingHost_ = false;
    eskf_icp_params prm = {maxIteration_, neighborMode_, translationSquaredThreshold_,
      cosineThreshold_};
    const auto G = guess.matrix();
    gpuCheck(
      eskf_align_cloud_begin(
        GpuContext::get(), localMap.handle(), cloud.device_.get(), G.data(), &prm, nullptr),
      "eskf_align_cloud_begin");
  }

  Isometry3d alignEnd()
  {
    if (pendingHost_) {return hostResult_;}
    eskf_align_info info = {};
    double T[16];
    gpuCheck(eskf_align_end(GpuContext::get(), T, &info), "eskf_align_end");
    lastIterations_ = info.iterations;
    if (info.converged) {converged_ = true;}
    if (!converged_) {std::cout << "ICP not converged!\n";}
    return Isometry3d::fromMatrix(T);
  }

  int lastIterations() const {return lastIterations_;}  // not in the reference

private:
  ICP() = delete;
  int maxIteration_;
  double translationSquaredThreshold_;
  double cosineThreshold_;
  int neighborMode_;
  bool converged_ = false;  // sticky, like the reference (Registration.hpp:50)
  int lastIterations_ = 0;
  bool pendingHost_ = false;
  Isometry3d hostResult_;
};
}  // namespace ESKF_LIO

#endif  // ESKF_LIO_B200_REGISTRATION_HPP_
