// ErrorStateKF.hpp — host-side 18-state error-state Kalman filter, drop-in for
// include/ESKF_LIO/ErrorStateKF.hpp + src/ErrorStateKF.cpp of the reference.
// The filter stays on the host (BASELINE.json north_star); its one expensive
// call, ICP::align (src/ErrorStateKF.cpp:130), goes to the B200.
//
// No Eigen here: Mat<R,C> below is a minimal row-major fixed-size matrix.
// Error-state order (ErrorStateKF.cpp:164-172): dp(0) dv(3) dtheta(6) dba(9)
// dbg(12) dg(15).
//
// The reference never zero-initialises Q_ (ErrorStateKF.hpp:43; only the four
// diagonal blocks are assigned, ErrorStateKF.cpp:37-40); the off-diagonal
// blocks are zero here.
#ifndef ESKF_LIO_B200_ERROR_STATE_KALMAN_FILTER_HPP_
#define ESKF_LIO_B200_ERROR_STATE_KALMAN_FILTER_HPP_

#include <cmath>
#include <deque>
#include <memory>
#include <utility>

#include "ESKF_LIO/LocalMap.hpp"
#include "ESKF_LIO/Registration.hpp"
#include "ESKF_LIO/Types.hpp"

namespace ESKF_LIO
{
template<int R, int C>
struct Mat
{
  double a[R * C];
  Mat() {for (double & x : a) {x = 0.0;}}
  double & operator()(int r, int c) {return a[r * C + c];}
  double operator()(int r, int c) const {return a[r * C + c];}
  static Mat Identity()
  {
    Mat m;
    for (int i = 0; i < (R < C ? R : C); ++i) {m(i, i) = 1.0;}
    return m;
  }
  template<int K>
  Mat<R, K> operator*(const Mat<C, K> & b) const
  {
    Mat<R, K> out;
    for (int i = 0; i < R; ++i) {
      for (int l = 0; l < C; ++l) {
        const double v = (*this)(i, l);
        if (v == 0.0) {continue;}
        for (int j = 0; j < K; ++j) {out(i, j) += v * b(l, j);}
      }
    }
    return out;
  }
  Mat<C, R> transpose() const
  {
    Mat<C, R> t;
    for (int i = 0; i < R; ++i) {
      for (int j = 0; j < C; ++j) {t(j, i) = (*this)(i, j);}
    }
    return t;
  }
  template<int BR, int BC>
  void setBlock(int r0, int c0, const Mat<BR, BC> & b)
  {
    for (int i = 0; i < BR; ++i) {
      for (int j = 0; j < BC; ++j) {(*this)(r0 + i, c0 + j) = b(i, j);}
    }
  }
};

namespace Utils
{
// Utils::skewSymmetric (src/Utils.cpp:5-11)
inline Mat<3, 3> skewSymmetric(const Vector3d & v)
{
  Mat<3, 3> s;
  s(0, 1) = -v(2); s(0, 2) = v(1);
  s(1, 0) = v(2);  s(1, 2) = -v(0);
  s(2, 0) = -v(1); s(2, 1) = v(0);
  return s;
}

// Eigen Quaterniond(AngleAxisd(|r|, r.normalized())) (src/Utils.cpp:34-38)
inline Quaterniond angleAxisToQuaternion(double angle, const Vector3d & v)
{
  const double n2 = v.squaredNorm();
  const double n = n2 > 0.0 ? std::sqrt(n2) : 1.0;  // Eigen normalized(): v itself when |v| == 0
  const double s = std::sin(0.5 * angle);
  Quaterniond q;
  q.x = s * (v(0) / n); q.y = s * (v(1) / n); q.z = s * (v(2) / n);
  q.w = std::cos(0.5 * angle);
  return q;
}
inline Quaterniond rotationVectorToQuaternion(const Vector3d & r)
{
  return angleAxisToQuaternion(std::sqrt(r.squaredNorm()), r);
}

inline Quaterniond multiply(const Quaterniond & a, const Quaterniond & b)
{
  Quaterniond q;
  q.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  q.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  q.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  q.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return q;
}

// Utils::rotationMatrixToVector (src/Utils.cpp:22-26): Eigen AngleAxisd(R)
// = AngleAxisd(Quaterniond(R))
inline Vector3d rotationMatrixToVector(const Matrix3d & R)
{
  double q[4];  // x y z w
  double t = R.trace();
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R(2, 1) - R(1, 2)) * t;
    q[1] = (R(0, 2) - R(2, 0)) * t;
    q[2] = (R(1, 0) - R(0, 1)) * t;
  } else {
    int i = 0;
    if (R(1, 1) > R(0, 0)) {i = 1;}
    if (R(2, 2) > R(i, i)) {i = 2;}
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R(k, j) - R(j, k)) * t;
    q[j] = (R(j, i) + R(i, j)) * t;
    q[k] = (R(k, i) + R(i, k)) * t;
  }
  double n = std::sqrt((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]);
  if (n == 0.0) {return Vector3d();}
  const double angle = 2.0 * std::atan2(n, std::fabs(q[3]));
  if (q[3] < 0.0) {n = -n;}
  return Vector3d(angle * (q[0] / n), angle * (q[1] / n), angle * (q[2] / n));
}
}  // namespace Utils

class ErrorStateKF
{
public:
  using Mat18d = Mat<18, 18>;
  using Vec18d = Mat<18, 1>;
  using Mat12d = Mat<12, 12>;
  using Mat6d = Mat<6, 6>;

  // ErrorStateKF::ErrorStateKF (src/ErrorStateKF.cpp:8-60)
  explicit ErrorStateKF(const Config & config)
  : icp_(std::make_shared<ICP>(config))
  {
    constexpr double GRAVITY_MAGNITUDE = 9.81;
    const auto & imu = config.imu;
    State initState;
    initState.biasAccel = Vector3d(imu.bias_a[0], imu.bias_a[1], imu.bias_a[2]);
    initState.biasGyro = Vector3d(imu.bias_g[0], imu.bias_g[1], imu.bias_g[2]);
    initState.gravity = Vector3d(imu.gravity[0], imu.gravity[1], imu.gravity[2]);
    states_.push_back(initState);

    const double rootRate = std::sqrt(imu.update_rate);
    const double sigmaGyroNoise = imu.gyro_noise_density * rootRate * M_PI / 180.0;
    const double sigmaAccelWork = imu.accel_zero_g_offset * rootRate * 1e-3 * GRAVITY_MAGNITUDE;
    const double sigmaGyroWork = imu.gyro_zero_rate_offset * rootRate * M_PI / 180.0;
    for (int i = 0; i < 3; ++i) {
      const double sigmaAccelNoise = imu.accel_noise_density[i] * GRAVITY_MAGNITUDE * rootRate;
      Q_(i, i) = sigmaAccelNoise * sigmaAccelNoise;
      Q_(3 + i, 3 + i) = std::pow(sigmaGyroNoise, 2.0);
      Q_(6 + i, 6 + i) = std::pow(sigmaAccelWork, 2.0);
      Q_(9 + i, 9 + i) = std::pow(sigmaGyroWork, 2.0);
    }
    for (int i = 0; i < 3; ++i) {
      V_(i, i) = config.kalman_filter.translation_noise;
      V_(3 + i, 3 + i) = config.kalman_filter.rotation_noise;
    }
  }

  const std::deque<State> & getStates() const {return states_;}
  double getLastStateTime() const {return states_.back().timestamp;}
  void feedImu(ImuMeasurementPtr imu) {ImuMeasurements_.push_back(std::move(imu));}

  // ErrorStateKF::initialize (:62-74)
  void initialize(double lidarEndTime)
  {
    states_[0].timestamp = lidarEndTime;
    replayImu(lidarEndTime);
  }

  // ErrorStateKF::process (:76-113)
  void process(const ImuMeasurementPtr & imu)
  {
    const State & prevState = states_.back();
    const double dt = imu->timestamp - prevState.timestamp;
    if (dt < 0.0) {return;}
    State newState = prevState;
    newState.timestamp = imu->timestamp;
    const Matrix3d R = prevState.attitude.toRotationMatrix();
    Vector3d acceleration, angularVelocity;
    for (int i = 0; i < 3; ++i) {
      acceleration(i) = imu->acceleration(i) - prevState.biasAccel(i);
      angularVelocity(i) = imu->angularVelocity(i) - prevState.biasGyro(i);
    }
    const Quaterniond angleDiff = Utils::angleAxisToQuaternion(
      std::sqrt(angularVelocity.squaredNorm()) * dt, angularVelocity);

    const double dt2 = dt * dt;
    for (int i = 0; i < 3; ++i) {
      const double aWorld = ((R(i, 0) * acceleration(0) + R(i, 1) * acceleration(1)) +
        R(i, 2) * acceleration(2)) + prevState.gravity(i);
      newState.position(i) = prevState.position(i) + prevState.velocity(i) * dt + 0.5 * aWorld * dt2;
      newState.velocity(i) = prevState.velocity(i) + aWorld * dt;
    }
    newState.attitude = Utils::multiply(prevState.attitude, angleDiff);

    Mat18d F_x = Mat18d::Identity();
    const Mat<3, 3> Rm = toMat(R);
    const Mat<3, 3> RS = Rm * Utils::skewSymmetric(acceleration);
    Quaterniond conj = angleDiff;
    conj.x = -conj.x; conj.y = -conj.y; conj.z = -conj.z;
    const Mat<3, 3> Rc = toMat(conj.toRotationMatrix());
    for (int i = 0; i < 3; ++i) {
      F_x(i, 3 + i) = dt;
      F_x(3 + i, 15 + i) = dt;
      F_x(6 + i, 12 + i) = -dt;
      for (int j = 0; j < 3; ++j) {
        F_x(3 + i, 6 + j) = -RS(i, j) * dt;
        F_x(3 + i, 9 + j) = -Rm(i, j) * dt;
        F_x(6 + i, 6 + j) = Rc(i, j);
      }
    }
    // F P F^T as (F (F P)^T)^T: F_x is Identity plus a handful of 3x3 blocks and
    // operator* skips the zero entries of its LEFT operand, so both products cost
    // ~40 rows of 18 instead of 18^3 (same terms in the same order as the dense form)
    const Mat18d FP = F_x * toMat18(prevState.P);
    Mat18d P = (F_x * FP.transpose()).transpose();
    // + F_i Q_i F_i^T : the 12 noise terms enter error-state rows 3..14 (:45-46)
    for (int i = 0; i < 12; ++i) {
      for (int j = 0; j < 12; ++j) {
        double q = Q_(i, j);
        if (i < 6 && j < 6) {q *= dt2;} else if (i >= 6 && j >= 6) {q *= dt;}
        P(3 + i, 3 + j) += q;
      }
    }
    fromMat18(P, newState.P);
    states_.push_back(std::move(newState));
  }

  // ErrorStateKF::update (:115-162)
  Isometry3d update(const LidarMeasurement & lidar, const LocalMap & localMap)
  {
    const double lidarEndTime = lidar.endTime;
    while (!states_.empty() && states_.back().timestamp > lidarEndTime) {states_.pop_back();}

    State newState = states_.back();
    newState.timestamp = lidarEndTime;
    const State & prevState = states_.back();
    Isometry3d guess;
    guess.linear() = prevState.attitude.toRotationMatrix();
    guess.translation() = prevState.position;
    // icp_->align (:130) in two halves: the gain below does not depend on the observation, so
    // it is computed while the Gauss-Newton kernel runs (same operations, same results)
    icp_->alignBegin(*lidar.cloud, localMap, guess);

    // H (6x18): I at (0,0) and (3,6)  (:55-57)
    Mat<6, 18> H;
    for (int i = 0; i < 3; ++i) {
      H(i, i) = 1.0;
      H(3 + i, 6 + i) = 1.0;
    }
    const Mat18d P = toMat18(prevState.P);
    const Mat<18, 6> PHt = P * H.transpose();
    Mat6d S = H * PHt;
    for (int i = 0; i < 36; ++i) {S.a[i] += V_.a[i];}
    const Mat<18, 6> K = PHt * inverse6(S);
    Mat18d I_KH = Mat18d::Identity();
    const Mat18d KH = K * H;
    for (int i = 0; i < 18 * 18; ++i) {I_KH.a[i] -= KH.a[i];}
    Mat18d Pn = I_KH * P;  // :142 (the Joseph form is commented out in the reference)

    const Isometry3d observation = icp_->alignEnd();

    Mat<6, 1> residual;
    for (int i = 0; i < 3; ++i) {residual(i, 0) = observation.t(i) - guess.t(i);}
    Matrix3d Rrel;  // guess.linear().transpose() * observation.linear()
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {
        Rrel(i, j) = (guess.R(0, i) * observation.R(0, j) + guess.R(1, i) * observation.R(1, j)) +
          guess.R(2, i) * observation.R(2, j);
      }
    }
    const Vector3d rv = Utils::rotationMatrixToVector(Rrel);
    for (int i = 0; i < 3; ++i) {residual(3 + i, 0) = rv(i);}

    const Vec18d errorState = K * residual;
    injectError(newState, errorState);
    // reset (:174-180)
    Mat18d G = Mat18d::Identity();
    const Mat<3, 3> skew = Utils::skewSymmetric(
      Vector3d(errorState(6, 0), errorState(7, 0), errorState(8, 0)));
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {G(6 + i, 6 + j) = (i == j ? 1.0 : 0.0) - 0.5 * skew(i, j);}
    }
    // G P G^T as (G (G P)^T)^T: both products have the sparse G on the left (24 non-zeros; the matrix
    // product skips zero left entries), same terms in the same order as the dense form
    Pn = (G * (G * Pn).transpose()).transpose();
    fromMat18(Pn, newState.P);

    states_.push_back(newState);
    replayImu(lidarEndTime);  // :148-155

    Isometry3d transform;
    transform.linear() = newState.attitude.toRotationMatrix();
    transform.translation() = newState.position;
    return transform;
  }

  int lastIcpIterations() const {return icp_->lastIterations();}  // not in the reference

private:
  ErrorStateKF() = delete;

  // :164-172
  void injectError(State & state, const Vec18d & e) const
  {
    for (int i = 0; i < 3; ++i) {
      state.position(i) += e(i, 0);
      state.velocity(i) += e(3 + i, 0);
      state.biasAccel(i) += e(9 + i, 0);
      state.biasGyro(i) += e(12 + i, 0);
      state.gravity(i) += e(15 + i, 0);
    }
    state.attitude = Utils::multiply(
      state.attitude, Utils::rotationVectorToQuaternion(Vector3d(e(6, 0), e(7, 0), e(8, 0))));
  }

  // drop the IMU samples before lidarEndTime, re-propagate the rest (:66-73, :148-155)
  void replayImu(double lidarEndTime)
  {
    while (!ImuMeasurements_.empty() && ImuMeasurements_.front()->timestamp < lidarEndTime) {
      ImuMeasurements_.pop_front();
    }
    for (auto & imu : ImuMeasurements_) {process(imu);}
  }

  static Mat<3, 3> toMat(const Matrix3d & R)
  {
    Mat<3, 3> m;
    for (int i = 0; i < 9; ++i) {m.a[i] = R.m[i];}
    return m;
  }
  static Mat18d toMat18(const std::array<double, 324> & p)
  {
    Mat18d m;
    for (int i = 0; i < 324; ++i) {m.a[i] = p[i];}
    return m;
  }
  static void fromMat18(const Mat18d & m, std::array<double, 324> & p)
  {
    for (int i = 0; i < 324; ++i) {p[i] = m.a[i];}
  }

  // Eigen's fixed-size inverse() of a 6x6 is a partial-pivot LU solve
  static Mat6d inverse6(const Mat6d & A)
  {
    double w[6][12];
    for (int i = 0; i < 6; ++i) {
      for (int j = 0; j < 6; ++j) {
        w[i][j] = A(i, j);
        w[i][6 + j] = i == j ? 1.0 : 0.0;
      }
    }
    for (int c = 0; c < 6; ++c) {
      int piv = c;
      for (int r = c + 1; r < 6; ++r) {
        if (std::fabs(w[r][c]) > std::fabs(w[piv][c])) {piv = r;}
      }
      if (piv != c) {
        for (int j = 0; j < 12; ++j) {std::swap(w[c][j], w[piv][j]);}
      }
      const double d = w[c][c];
      for (int j = 0; j < 12; ++j) {w[c][j] /= d;}
      for (int r = 0; r < 6; ++r) {
        if (r == c || w[r][c] == 0.0) {continue;}
        const double f = w[r][c];
        for (int j = 0; j < 12; ++j) {w[r][j] -= f * w[c][j];}
      }
    }
    Mat6d out;
    for (int i = 0; i < 6; ++i) {
      for (int j = 0; j < 6; ++j) {out(i, j) = w[i][6 + j];}
    }
    return out;
  }

  std::shared_ptr<ICP> icp_;
  std::deque<State> states_;
  std::deque<ImuMeasurementPtr> ImuMeasurements_;
  Mat12d Q_;
  Mat6d V_;
};

}  // namespace ESKF_LIO

#endif  // ESKF_LIO_B200_ERROR_STATE_KALMAN_FILTER_HPP_
