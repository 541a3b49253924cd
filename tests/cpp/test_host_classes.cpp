// Drives the C++ drop-in classes (eskf_lio_b200/host/ESKF_LIO/*.hpp) the way
// Odometry::run does (src/Odometry.cpp:55-87): process -> align -> updateLocalMap.
// Inputs come from a binary file written by the pytest; results go to stdout
// as "key value..." lines that the pytest compares with the CPU oracle.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "ESKF_LIO/CloudPreprocessor.hpp"
#include "ESKF_LIO/LocalMap.hpp"
#include "ESKF_LIO/Registration.hpp"

using namespace ESKF_LIO;

static std::vector<double> readDoubles(std::ifstream & f, std::size_t n)
{
  std::vector<double> v(n);
  f.read(reinterpret_cast<char *>(v.data()), static_cast<std::streamsize>(n * sizeof(double)));
  return v;
}

int main(int argc, char ** argv)
{
  if (argc < 2) {return 2;}
  std::ifstream f(argv[1], std::ios::binary);
  uint64_t nScans = 0;
  f.read(reinterpret_cast<char *>(&nScans), 8);
  Config config;
  config.local_map.voxel_size = 0.5;
  config.cloud_preprocessor.voxel_size = 0.5;
  config.local_map.remove_distant_points.enabled = false;
  CloudPreprocessor pre(config);
  LocalMap map(config);
  ICP icp(config);
  for (uint64_t s = 0; s < nScans; ++s) {
    uint64_t n = 0;
    f.read(reinterpret_cast<char *>(&n), 8);
    auto xyz = readDoubles(f, 3 * n);
    auto times = readDoubles(f, n);
    auto pose = readDoubles(f, 16);   // map-building pose or align guess
    auto meas = std::make_shared<LidarMeasurement>();
    meas->cloud = std::make_shared<PointCloud>();
    meas->cloud->points_.resize(n);
    for (uint64_t i = 0; i < n; ++i) {meas->cloud->points_[i] = Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);}
    meas->pointTime = times;
    pre.process({}, meas);
    std::printf("scan %llu kept %zu\n", static_cast<unsigned long long>(s), meas->cloud->points_.size());
    Isometry3d T = Isometry3d::fromMatrix(pose.data());
    if (s + 1 == nScans) {
      const Isometry3d out = icp.align(*meas->cloud, map, T);
      const auto M = out.matrix();
      std::printf("align iterations %d pose", icp.lastIterations());
      for (double v : M) {std::printf(" %.17g", v);}
      std::printf("\n");
      auto corr = map.correspondenceMatching(meas->cloud->points_, meas->cloud->covariances_);
      std::printf("corr_body %zu\n", std::get<0>(corr).size());
      T = out;
    }
    map.updateLocalMap(meas->cloud, T, true);
    std::printf("map %zu first_world %.17g %.17g %.17g\n", map.size(), meas->cloud->points_[0].v[0],
      meas->cloud->points_[0].v[1], meas->cloud->points_[0].v[2]);
  }
  if (argc > 3) {
    map.save(argv[2], argv[3]);
    std::printf("saved %zu poses\n", map.trajectory().size());
  }
  return 0;
}
