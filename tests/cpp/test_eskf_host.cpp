// Host-only check of the product's ErrorStateKF::process (no GPU call is made:
// ICP is constructed but never aligned).  Reads IMU rows (t, gyro, acc) from a
// binary file, propagates, prints the newest state and P for the pytest to
// compare with the CPU oracle.
#include <cstdio>
#include <fstream>
#include <vector>

#include "ESKF_LIO/ErrorStateKF.hpp"

using namespace ESKF_LIO;

int main(int argc, char ** argv)
{
  if (argc < 2) {return 2;}
  std::ifstream f(argv[1], std::ios::binary);
  uint64_t n = 0;
  f.read(reinterpret_cast<char *>(&n), 8);
  std::vector<double> rows(7 * n);
  f.read(reinterpret_cast<char *>(rows.data()), static_cast<std::streamsize>(rows.size() * 8));
  Config config;
  ErrorStateKF kf(config);
  for (uint64_t k = 0; k < n; ++k) {
    auto m = std::make_shared<ImuMeasurement>();
    m->timestamp = rows[7 * k];
    m->angularVelocity = Vector3d(rows[7 * k + 1], rows[7 * k + 2], rows[7 * k + 3]);
    m->acceleration = Vector3d(rows[7 * k + 4], rows[7 * k + 5], rows[7 * k + 6]);
    kf.process(m);
  }
  const State & s = kf.getStates().back();
  std::printf("states %zu\n", kf.getStates().size());
  std::printf("state %.17g", s.timestamp);
  for (int i = 0; i < 3; ++i) {std::printf(" %.17g", s.position(i));}
  for (int i = 0; i < 3; ++i) {std::printf(" %.17g", s.velocity(i));}
  std::printf(" %.17g %.17g %.17g %.17g", s.attitude.x, s.attitude.y, s.attitude.z, s.attitude.w);
  std::printf("\nP");
  for (double v : s.P) {std::printf(" %.17g", v);}
  std::printf("\n");
  // Utils::rotationMatrixToVector on a fixed rotation (both quaternion branches)
  for (double ang : {0.3, 2.9}) {
    Quaterniond q = Utils::angleAxisToQuaternion(ang, Vector3d(0.2, -0.5, 0.84));
    const Vector3d r = Utils::rotationMatrixToVector(q.toRotationMatrix());
    std::printf("rotvec %.17g %.17g %.17g\n", r(0), r(1), r(2));
  }
  return 0;
}
