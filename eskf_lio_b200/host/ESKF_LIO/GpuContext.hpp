// GpuContext.hpp — process-wide eskf_ctx shared by the three host classes
// (the reference drives them from one thread, src/main.cpp:70).
#ifndef ESKF_LIO_B200_GPU_CONTEXT_HPP_
#define ESKF_LIO_B200_GPU_CONTEXT_HPP_

#include <memory>
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

class GpuContext
{
public:
  static eskf_ctx * get(int device = 0)
  {
    static GpuContext instance(device);
    return instance.ctx_;
  }

private:
  explicit GpuContext(int device) {gpuCheck(eskf_ctx_create(device, nullptr, &ctx_), "eskf_ctx_create");}
  ~GpuContext() {eskf_ctx_destroy(ctx_);}
  eskf_ctx * ctx_ = nullptr;
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
    pool_.emplace_back(raw, [](eskf_cloud * c) {eskf_cloud_destroy(c);});
    return pool_.back();
  }

private:
  std::vector<std::shared_ptr<eskf_cloud>> pool_;
};
}  // namespace ESKF_LIO

#endif  // ESKF_LIO_B200_GPU_CONTEXT_HPP_
