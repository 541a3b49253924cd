// SensorLog.hpp -- flat binary sensor log: the wire format of the two ROS subscribers
// (include/ESKF_LIO/Subscriber.hpp:38-52 sensor_msgs/Imu, :80-103 sensor_msgs/PointCloud2 with
// float32 x, y, z and a float64 "timestamp" per point) without ROS, so that recorded data (e.g. a
// converted Hilti bag) can be replayed through Odometry::feedImu / feedLidar.
//
//   file   = header record*
//   header = "ESKFLOG1" (8 bytes)  u32 version (1)  u32 reserved (0)                  little endian
//   record = u32 type  u32 count  payload
//     type 1 (IMU, count 1):      f64 stamp, f64 angular_velocity[3], f64 linear_acceleration[3]
//     type 2 (LiDAR sweep):       count x { f32 x, f32 y, f32 z, f64 timestamp }      (20 B, packed)
// Records are stored in the order the callbacks fired.
#ifndef ESKF_LIO_B200_SENSOR_LOG_HPP_
#define ESKF_LIO_B200_SENSOR_LOG_HPP_

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace ESKF_LIO
{
class SensorLogWriter
{
public:
  explicit SensorLogWriter(const std::string & path)
  : f_(std::fopen(path.c_str(), "wb"))
  {
    if (!f_) {throw std::runtime_error("SensorLogWriter: cannot open " + path);}
    const char magic[8] = {'E', 'S', 'K', 'F', 'L', 'O', 'G', '1'};
    const std::uint32_t hdr[2] = {1u, 0u};
    put(magic, 8);
    put(hdr, 8);
  }
  ~SensorLogWriter() {if (f_) {std::fclose(f_);}}
  SensorLogWriter(const SensorLogWriter &) = delete;
  SensorLogWriter & operator=(const SensorLogWriter &) = delete;

  void writeImu(double stamp, const double gyro[3], const double acc[3])
  {
    const std::uint32_t h[2] = {1u, 1u};
    put(h, 8);
    put(&stamp, 8);
    put(gyro, 24);
    put(acc, 24);
  }

  void writeLidar(const float * xyz, const double * pointTime, std::size_t n)
  {
    const std::uint32_t h[2] = {2u, static_cast<std::uint32_t>(n)};
    put(h, 8);
    std::vector<char> buf(n * 20);
    for (std::size_t i = 0; i < n; ++i) {
      std::memcpy(buf.data() + 20 * i, xyz + 3 * i, 12);
      std::memcpy(buf.data() + 20 * i + 12, pointTime + i, 8);
    }
    put(buf.data(), buf.size());
  }

private:
  void put(const void * p, std::size_t bytes)
  {
    if (bytes && std::fwrite(p, 1, bytes, f_) != bytes) {throw std::runtime_error("SensorLogWriter: short write");}
  }
  std::FILE * f_;
};

class SensorLogReader
{
public:
  enum Type : std::uint32_t {kEnd = 0, kImu = 1, kLidar = 2};

  explicit SensorLogReader(const std::string & path)
  : f_(std::fopen(path.c_str(), "rb"))
  {
    if (!f_) {throw std::runtime_error("SensorLogReader: cannot open " + path);}
    char magic[8];
    std::uint32_t hdr[2];
    if (!get(magic, 8) || !get(hdr, 8) || std::memcmp(magic, "ESKFLOG1", 8) != 0) {
      std::fclose(f_);
      f_ = nullptr;
      throw std::runtime_error("SensorLogReader: " + path + " is not an ESKFLOG1 file");
    }
    if (hdr[0] != 1u) {
      std::fclose(f_);
      f_ = nullptr;
      throw std::runtime_error("SensorLogReader: unsupported log version");
    }
  }
  ~SensorLogReader() {if (f_) {std::fclose(f_);}}
  SensorLogReader(const SensorLogReader &) = delete;
  SensorLogReader & operator=(const SensorLogReader &) = delete;

  // Reads the next record.  IMU: stamp / gyro / acc are filled.  LiDAR: xyz (3 floats per point) and
  // pointTime are filled.  Returns kEnd at a clean end of file; throws on a truncated record.
  Type next()
  {
    std::uint32_t h[2];
    const std::size_t got = std::fread(h, 1, 8, f_);
    if (got == 0) {return kEnd;}
    if (got != 8) {throw std::runtime_error("SensorLogReader: truncated record header");}
    if (h[0] == kImu) {
      if (h[1] != 1u) {throw std::runtime_error("SensorLogReader: IMU record with count != 1");}
      double v[7];
      if (!get(v, 56)) {throw std::runtime_error("SensorLogReader: truncated IMU record");}
      stamp = v[0];
      for (int k = 0; k < 3; ++k) {
        gyro[k] = v[1 + k];
        acc[k] = v[4 + k];
      }
      return kImu;
    }
    if (h[0] == kLidar) {
      const std::size_t n = h[1];
      buf_.resize(n * 20);
      if (!get(buf_.data(), buf_.size())) {throw std::runtime_error("SensorLogReader: truncated sweep");}
      xyz.resize(3 * n);
      pointTime.resize(n);
      for (std::size_t i = 0; i < n; ++i) {
        std::memcpy(xyz.data() + 3 * i, buf_.data() + 20 * i, 12);
        std::memcpy(pointTime.data() + i, buf_.data() + 20 * i + 12, 8);
      }
      return kLidar;
    }
    throw std::runtime_error("SensorLogReader: unknown record type");
  }

  double stamp = 0.0;
  double gyro[3] = {0, 0, 0}, acc[3] = {0, 0, 0};
  std::vector<float> xyz;
  std::vector<double> pointTime;

private:
  bool get(void * p, std::size_t bytes) {return bytes == 0 || std::fread(p, 1, bytes, f_) == bytes;}
  std::FILE * f_;
  std::vector<char> buf_;
};
}  // namespace ESKF_LIO

#endif  // ESKF_LIO_B200_SENSOR_LOG_HPP_
