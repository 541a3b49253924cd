// GpuContext.hpp — process-wide eskf_ctx shared by the three host classes
// (the reference drives them from one thread, src/main.cpp:70).
#ifndef ESKF_LIO_B200_GPU_CONTEXT_HPP_
#define ESKF_LIO_B200_GPU_CONTEXT_HPP_

#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "eskf_gpu.h"

namespace ESKF_LIO
{
inline void gpuCheck(int status, const char * what)
{
  if (status != ESKF_OK) {
    // the reference never reports errors from these calls; a missing GPU is
    // not something to limp through, so this throws (there is no CPU fallback)
    throw std::runtime_error(std::string(what) + ": " + eskf_last_error());
  }
}

// One eskf_ctx = one stream = one caller at a time (include/eskf_gpu.h).  The reference drives the
// three classes from its main thread only (src/main.cpp:70) while its subscriber threads just queue
// measurements (Subscriber.hpp:51,102); here Odometry::feedLidar uploads a sweep from whichever thread
// delivers it, so every entry into the context from the host classes holds lock().  Maps, clouds and
// pools keep the context alive through share(): a static or global Odometry may outlive main().
class GpuContext
{
public:
  static eskf_ctx * get(int device = 0) {return share(device)->ctx_;}
  static std::shared_ptr<GpuContext> share(int device = 0)
  {
    static std::shared_ptr<GpuContext> instance(new GpuContext(device));
    return instance;
  }
  static std::unique_lock<std::recursive_mutex> lock() {return std::unique_lock<std::recursive_mutex>(share()->mutex_);}
  ~GpuContext() {eskf_ctx_destroy(ctx_);}

private:
  explicit GpuContext(int device) {gpuCheck(eskf_ctx_create(device, nullptr, &ctx_), "eskf_ctx_create");}
  eskf_ctx * ctx_ = nullptr;
  std::recursive_mutex mutex_;
};

// Reusable device clouds (Config::device_resident): a cloud goes back into
// circulation once nobody but the pool holds it, so steady-state frames never
// call cudaMalloc.
class DeviceCloudPool
{
public:
  std::shared_ptr<eskf_cloud> acquire(std::size_t capacity)
  {
    for (auto & c : pool_) {
      if (c.use_count() == 1) {return c;}
    }
    eskf_cloud * raw = nullptr;
    gpuCheck(eskf_cloud_create(GpuContext::get(), capacity, &raw), "eskf_cloud_create");
    std::shared_ptr<GpuContext> keep = GpuContext::share();  // the cloud's context outlives the cloud
    pool_.emplace_back(raw, [keep](eskf_cloud * c) {eskf_cloud_destroy(c);});
    return pool_.back();
  }

private:
  std::vector<std::shared_ptr<eskf_cloud>> pool_;
};
}  // namespace ESKF_LIO

#endif  // ESKF_LIO_B200_GPU_CONTEXT_HPP_
