// odometry_capi.cpp — include/eskf_host.h over the header-only host classes.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>

#include "ESKF_LIO/Odometry.hpp"
#include "ESKF_LIO/SensorLog.hpp"
#include "eskf_host.h"

using namespace ESKF_LIO;

struct eskf_odom
{
  Odometry::ImuBuffer imu = std::make_shared<SynchronizedQueue<ImuMeasurementPtr>>();
  Odometry::CloudBuffer cloud = std::make_shared<SynchronizedQueue<LidarMeasurementPtr>>();
  std::unique_ptr<Odometry> odom;
  bool device_resident = false;
};

namespace
{
thread_local std::string g_error;

int fail(const char * what)
{
  g_error = what;
  return 1;
}

template<typename F>
int guarded(F && f)
{
  try {
    f();
    return 0;
  } catch (const std::exception & e) {
    return fail(e.what());
  } catch (...) {
    return fail("unknown exception");
  }
}

Config toConfig(const eskf_odom_config & c)
{
  Config k;
  k.imu.update_rate = c.imu_update_rate;
  for (int i = 0; i < 3; ++i) {
    k.imu.bias_a[i] = c.bias_a[i];
    k.imu.bias_g[i] = c.bias_g[i];
    k.imu.gravity[i] = c.gravity[i];
    k.imu.accel_noise_density[i] = c.accel_noise_density[i];
    k.lidar_extrinsics.translation[i] = c.lidar_translation[i];
  }
  for (int i = 0; i < 4; ++i) {k.lidar_extrinsics.quaternion[i] = c.lidar_quaternion_xyzw[i];}
  k.imu.accel_zero_g_offset = c.accel_zero_g_offset;
  k.imu.gyro_noise_density = c.gyro_noise_density;
  k.imu.gyro_zero_rate_offset = c.gyro_zero_rate_offset;
  k.kalman_filter.translation_noise = c.translation_noise;
  k.kalman_filter.rotation_noise = c.rotation_noise;
  k.local_map.voxel_size = c.map_voxel_size;
  k.local_map.max_num_points_per_voxel = c.max_points_per_voxel;
  k.local_map.update.translation_sq_threshold = c.update_translation_sq_threshold;
  k.local_map.update.cosine_threshold = c.update_cosine_threshold;
  k.local_map.remove_distant_points.enabled = c.remove_enabled != 0;
  k.local_map.remove_distant_points.distance_threshold = c.remove_distance_threshold;
  k.local_map.remove_distant_points.removing_period = c.remove_period;
  k.cloud_preprocessor.voxel_size = c.preprocess_voxel_size;
  k.registration.max_iteration = c.max_iteration;
  k.registration.neighbor_mode = c.neighbor_mode;
  k.registration.translation_sq_threshold = c.icp_translation_sq_threshold;
  k.registration.cosine_threshold = c.icp_cosine_threshold;
  k.device_resident = c.device_resident != 0;
  if (c.map_capacity_hint != 0) {k.local_map.capacity_hint = static_cast<std::size_t>(c.map_capacity_hint);}
  return k;
}
}  // namespace

extern "C" {

const char * eskf_host_last_error(void) {return g_error.c_str();}

void eskf_odom_default_config(eskf_odom_config * c)
{
  const Config k;  // the struct's defaults are config/hilti_config.yaml
  c->imu_update_rate = k.imu.update_rate;
  for (int i = 0; i < 3; ++i) {
    c->bias_a[i] = k.imu.bias_a[i];
    c->bias_g[i] = k.imu.bias_g[i];
    c->gravity[i] = k.imu.gravity[i];
    c->accel_noise_density[i] = k.imu.accel_noise_density[i];
    c->lidar_translation[i] = k.lidar_extrinsics.translation[i];
  }
  for (int i = 0; i < 4; ++i) {c->lidar_quaternion_xyzw[i] = k.lidar_extrinsics.quaternion[i];}
  c->accel_zero_g_offset = k.imu.accel_zero_g_offset;
  c->gyro_noise_density = k.imu.gyro_noise_density;
  c->gyro_zero_rate_offset = k.imu.gyro_zero_rate_offset;
  c->translation_noise = k.kalman_filter.translation_noise;
  c->rotation_noise = k.kalman_filter.rotation_noise;
  c->map_voxel_size = k.local_map.voxel_size;
  c->max_points_per_voxel = k.local_map.max_num_points_per_voxel;
  c->update_translation_sq_threshold = k.local_map.update.translation_sq_threshold;
  c->update_cosine_threshold = k.local_map.update.cosine_threshold;
  c->remove_enabled = k.local_map.remove_distant_points.enabled ? 1 : 0;
  c->remove_distance_threshold = k.local_map.remove_distant_points.distance_threshold;
  c->remove_period = k.local_map.remove_distant_points.removing_period;
  c->preprocess_voxel_size = k.cloud_preprocessor.voxel_size;
  c->max_iteration = k.registration.max_iteration;
  c->neighbor_mode = k.registration.neighbor_mode;
  c->icp_translation_sq_threshold = k.registration.translation_sq_threshold;
  c->icp_cosine_threshold = k.registration.cosine_threshold;
  c->device_resident = 1;
  c->map_capacity_hint = k.local_map.capacity_hint;
}

int eskf_odom_create(const eskf_odom_config * cfg, int device, eskf_odom ** out)
{
  if (!cfg || !out) {return fail("null argument");}
  *out = nullptr;
  return guarded(
    [&] {
      GpuContext::get(device);  // fails here, loudly, when there is no CUDA device
      auto o = std::make_unique<eskf_odom>();
      o->odom = std::make_unique<Odometry>(toConfig(*cfg), o->imu, o->cloud);
      o->device_resident = cfg->device_resident != 0;
      *out = o.release();
    });
}

int eskf_odom_destroy(eskf_odom * o)
{
  delete o;
  return 0;
}

int eskf_odom_feed_imu(eskf_odom * o, double t, const double gyro[3], const double acc[3])
{
  if (!o || !gyro || !acc) {return fail("null argument");}
  return guarded(
    [&] {
      auto m = std::make_shared<ImuMeasurement>();
      m->timestamp = t;
      m->angularVelocity = Vector3d(gyro[0], gyro[1], gyro[2]);
      m->acceleration = Vector3d(acc[0], acc[1], acc[2]);
      o->imu->push(std::move(m));
    });
}

int eskf_odom_feed_lidar(eskf_odom * o, const float * xyz, const double * point_time, size_t n)
{
  if (!o || !xyz || !point_time || n == 0) {return fail("null or empty sweep");}
  return guarded([&] {o->odom->feedLidar(xyz, point_time, n);});
}

int eskf_odom_feed_lidar_cloud(eskf_odom * o, void * raw_cloud, const double * point_time, size_t n)
{
  if (!o || !raw_cloud || !point_time || n == 0) {return fail("null or empty sweep");}
  if (!o->device_resident) {return fail("eskf_odom_feed_lidar_cloud needs device_resident = 1");}
  return guarded(
    [&] {o->odom->feedLidarDevice(static_cast<eskf_cloud *>(raw_cloud), point_time, n);});
}

int eskf_odom_context(eskf_odom * o, void ** ctx)
{
  if (!o || !ctx) {return fail("null argument");}
  return guarded([&] {*ctx = GpuContext::get();});
}

int eskf_odom_spin_once(eskf_odom * o, int * consumed)
{
  if (!o) {return fail("null argument");}
  return guarded(
    [&] {
      const bool c = o->odom->spinOnce();
      if (consumed) {*consumed = c ? 1 : 0;}
    });
}

int eskf_odom_last_pose(eskf_odom * o, double T[16])
{
  if (!o || !T) {return fail("null argument");}
  const auto M = o->odom->lastTransform().matrix();
  std::memcpy(T, M.data(), sizeof(double) * 16);
  return 0;
}

int eskf_odom_last_state(eskf_odom * o, double s[20], double * P324)
{
  if (!o || !s) {return fail("null argument");}
  const State & st = o->odom->kalmanFilter().getStates().back();
  s[0] = st.timestamp;
  for (int i = 0; i < 3; ++i) {
    s[1 + i] = st.position(i);
    s[4 + i] = st.velocity(i);
    s[11 + i] = st.biasAccel(i);
    s[14 + i] = st.biasGyro(i);
    s[17 + i] = st.gravity(i);
  }
  s[7] = st.attitude.x; s[8] = st.attitude.y; s[9] = st.attitude.z; s[10] = st.attitude.w;
  if (P324) {std::memcpy(P324, st.P.data(), sizeof(double) * 324);}
  return 0;
}

int eskf_odom_info_get(eskf_odom * o, eskf_odom_info * out)
{
  if (!o || !out) {return fail("null argument");}
  return guarded(
    [&] {
      const auto & t = o->odom->stageTimes();
      const double f = t.numFrames > 0 ? 1e3 / t.numFrames : 0.0;
      out->frames = static_cast<uint64_t>(t.numFrames);
      out->n_states = o->odom->kalmanFilter().getStates().size();
      out->map_voxels = o->odom->localMap().size();
      out->last_removed = o->odom->localMap().lastRemoved();
      out->last_iterations = o->odom->kalmanFilter().lastIcpIterations();
      out->last_inserted = o->odom->localMap().lastInserted() ? 1 : 0;
      out->stage_avg_ms[0] = t.cloudPreprocessor * f;
      out->stage_avg_ms[1] = t.filterUpdate * f;
      out->stage_avg_ms[2] = t.mapUpdate * f;
      out->stage_max_ms[0] = t.cloudPreprocessorMax * 1e3;
      out->stage_max_ms[1] = t.filterUpdateMax * 1e3;
      out->stage_max_ms[2] = t.mapUpdateMax * 1e3;
      out->stage_sum_ms[0] = t.cloudPreprocessor * 1e3;
      out->stage_sum_ms[1] = t.filterUpdate * 1e3;
      out->stage_sum_ms[2] = t.mapUpdate * 1e3;
      out->device_frame_ms_sum = t.deviceFrameMs;
      out->device_frame_ms_last = t.deviceFrameMsLast;
    });
}

int eskf_odom_map(eskf_odom * o, void ** map_handle)
{
  if (!o || !map_handle) {return fail("null argument");}
  *map_handle = o->odom->localMap().handle();
  return 0;
}

int eskf_odom_launch_count(eskf_odom * o, uint64_t * n)
{
  if (!o || !n) {return fail("null argument");}
  return guarded([&] {gpuCheck(eskf_ctx_launch_count(GpuContext::get(), n), "eskf_ctx_launch_count");});
}

// ---- flat binary sensor log (ESKF_LIO/SensorLog.hpp) -------------------------
int eskf_odom_replay_log(eskf_odom * o, const char * path, double * poses, size_t capacity, size_t * n_frames)
{
  if (!o || !path || !n_frames) {return fail("null argument");}
  *n_frames = 0;
  return guarded(
    [&] {
      SensorLogReader rd(path);
      // sweeps handed to the odometry but not consumed yet: in device-resident mode the per-point
      // stamps are read in place when the frame is processed, so their buffers stay alive here
      std::deque<std::pair<std::vector<float>, std::vector<double>>> pending;
      for (;;) {
        const SensorLogReader::Type t = rd.next();
        if (t == SensorLogReader::kEnd) {break;}
        if (t == SensorLogReader::kImu) {
          auto m = std::make_shared<ImuMeasurement>();  // ImuSubscriber::imuCallback (Subscriber.hpp:38-52)
          m->timestamp = rd.stamp;
          m->angularVelocity = Vector3d(rd.gyro[0], rd.gyro[1], rd.gyro[2]);
          m->acceleration = Vector3d(rd.acc[0], rd.acc[1], rd.acc[2]);
          o->imu->push(std::move(m));
        } else {
          if (rd.pointTime.empty()) {continue;}
          pending.emplace_back(std::move(rd.xyz), std::move(rd.pointTime));  // cloudCallback (:80-103)
          o->odom->feedLidar(pending.back().first.data(), pending.back().second.data(), pending.back().second.size());
        }
        // one trip of Odometry::run's loop per delivery, more while frames keep going through
        while (o->odom->spinOnce()) {
          if (poses && *n_frames < capacity) {
            const auto M = o->odom->lastTransform().matrix();
            std::memcpy(poses + 16 * *n_frames, M.data(), sizeof(double) * 16);
          }
          ++*n_frames;
          if (!pending.empty()) {pending.pop_front();}
        }
      }
    });
}

int eskf_log_summary(const char * path, uint64_t counts[3], double stamps[2])
{
  if (!path || !counts || !stamps) {return fail("null argument");}
  return guarded(
    [&] {
      SensorLogReader rd(path);
      counts[0] = counts[1] = counts[2] = 0;
      stamps[0] = stamps[1] = 0.0;
      bool first = true;
      for (;;) {
        const SensorLogReader::Type t = rd.next();
        if (t == SensorLogReader::kEnd) {break;}
        double a, b;
        if (t == SensorLogReader::kImu) {
          ++counts[0];
          a = b = rd.stamp;
        } else {
          ++counts[1];
          counts[2] += rd.pointTime.size();
          if (rd.pointTime.empty()) {continue;}
          a = rd.pointTime.front();
          b = rd.pointTime.back();
        }
        if (first || a < stamps[0]) {stamps[0] = a;}
        if (first || b > stamps[1]) {stamps[1] = b;}
        first = false;
      }
    });
}

struct eskf_log_writer
{
  std::unique_ptr<SensorLogWriter> w;
};

int eskf_log_writer_open(const char * path, eskf_log_writer ** out)
{
  if (!path || !out) {return fail("null argument");}
  *out = nullptr;
  return guarded(
    [&] {
      auto h = std::make_unique<eskf_log_writer>();
      h->w = std::make_unique<SensorLogWriter>(path);
      *out = h.release();
    });
}

int eskf_log_writer_imu(eskf_log_writer * w, double stamp, const double gyro[3], const double acc[3])
{
  if (!w || !gyro || !acc) {return fail("null argument");}
  return guarded([&] {w->w->writeImu(stamp, gyro, acc);});
}

int eskf_log_writer_lidar(eskf_log_writer * w, const float * xyz, const double * point_time, size_t n)
{
  if (!w || (n && (!xyz || !point_time))) {return fail("null argument");}
  return guarded([&] {w->w->writeLidar(xyz, point_time, n);});
}

int eskf_log_writer_close(eskf_log_writer * w)
{
  delete w;
  return 0;
}

}  // extern "C"
