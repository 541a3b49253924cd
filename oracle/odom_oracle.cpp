// odom_oracle.cpp — CPU ORACLE (test infrastructure; see eskf_oracle.h) for
// the callers either side of the hot path: the 18-state error-state Kalman
// filter (src/ErrorStateKF.cpp) and the per-frame call order of Odometry::run
// (src/Odometry.cpp:16-98), ROS-free.  PARITY UNPINNED by the reference (no
// tests exist); pinned by tests/test_eskf_oracle.py (NumPy restatement + KATs).
//
// Flat fp64 arrays, no Eigen.  Eigen internals restated from their published
// algorithms: AngleAxisd <-> Quaterniond <-> Matrix3d conversions, quaternion
// product, normalized(), fixed-size inverse() of a 6x6 (partial-pivot LU).
//
// Deviations forced by undefined behaviour in the reference:
//   - Q_ (ErrorStateKF.hpp:43) is never zero-initialised; only its four
//     diagonal 3x3 blocks are assigned (:37-40).  Off-diagonal blocks = 0 here.
//   - LocalMap::prevTransform_ is uninitialised on the first updateLocalMap
//     (Odometry.cpp:61 passes initialize=false): the first frame is inserted
//     with initialize=true.
//   - the eviction period is tested against omp_get_wtime() (LocalMap.cpp:60);
//     the clock here is the LiDAR end time (sim time), so runs are reproducible.
#include <omp.h>

#include <cmath>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <vector>

#include "eskf_oracle.h"

namespace {

constexpr int N = 18;

struct OState {  // ESKF_LIO::State, include/ESKF_LIO/Types.hpp:31-40
  double t = 0.0;
  double p[3] = {0, 0, 0}, v[3] = {0, 0, 0};
  double q[4] = {0, 0, 0, 1};  // x y z w
  double ba[3] = {0, 0, 0}, bg[3] = {0, 0, 0}, g[3] = {0, 0, 0};
  double P[N * N];
  OState() {
    std::memset(P, 0, sizeof P);
    for (int i = 0; i < N; ++i) P[i * N + i] = 1e-3;
  }
};

struct Imu {
  double t, w[3], a[3];
};

struct Scan {
  std::vector<double> xyz, time;
  double start, end;
};

inline double dot3(const double* a, const double* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

void quat_to_R(const double* q, double* R) { orc_quat_to_matrix(q, R); }

// Eigen Quaterniond(AngleAxisd(angle, axis)): w = cos(a/2), vec = sin(a/2) axis
void angle_axis_to_quat(double angle, const double* axis, double* q) {
  const double ha = 0.5 * angle, s = std::sin(ha);
  q[0] = s * axis[0];
  q[1] = s * axis[1];
  q[2] = s * axis[2];
  q[3] = std::cos(ha);
}

// Eigen normalized(): v / |v| when |v|^2 > 0, else v
void normalized(const double* v, double* out) {
  const double z = dot3(v, v);
  const double n = z > 0.0 ? std::sqrt(z) : 1.0;
  for (int i = 0; i < 3; ++i) out[i] = z > 0.0 ? v[i] / n : v[i];
}

// Eigen quaternion product a * b (no normalisation)
void quat_mul(const double* a, const double* b, double* o) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  o[3] = aw * bw - ax * bx - ay * by - az * bz;
  o[0] = aw * bx + ax * bw + ay * bz - az * by;
  o[1] = aw * by + ay * bw + az * bx - ax * bz;
  o[2] = aw * bz + az * bw + ax * by - ay * bx;
}

// Eigen Quaterniond(Matrix3d) (Shepperd's method as in Eigen/Geometry/Quaternion.h)
void R_to_quat(const double* R, double* q) {
  double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
}

// Utils::rotationMatrixToVector (src/Utils.cpp:22-26): AngleAxisd(R) goes
// through Quaterniond(R); angle = 2 atan2(|vec|, |w|), axis = vec / (+-|vec|)
void R_to_rotvec(const double* R, double* r) {
  double q[4];
  R_to_quat(R, q);
  double n = std::sqrt(dot3(q, q));
  if (n != 0.0) {
    const double angle = 2.0 * std::atan2(n, std::fabs(q[3]));
    if (q[3] < 0.0) n = -n;
    for (int i = 0; i < 3; ++i) r[i] = angle * (q[i] / n);
  } else {
    r[0] = r[1] = r[2] = 0.0;  // angle 0 about (1,0,0)
  }
}

void matmul(const double* A, const double* B, double* C, int n, int k, int m) {  // (n x k)(k x m)
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) {
      double s = 0.0;
      for (int l = 0; l < k; ++l) s += A[i * k + l] * B[l * m + j];
      C[i * m + j] = s;
    }
}

void matmul_bt(const double* A, const double* B, double* C, int n, int k, int m) {  // A (n x k) * B^T, B (m x k)
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) {
      double s = 0.0;
      for (int l = 0; l < k; ++l) s += A[i * k + l] * B[j * k + l];
      C[i * m + j] = s;
    }
}

// 6x6 inverse, partial-pivot Gauss-Jordan (Eigen: PartialPivLU for sizes > 4)
void inv6(const double* Ain, double* out) {
  double A[6][12];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      A[i][j] = Ain[6 * i + j];
      A[i][6 + j] = i == j ? 1.0 : 0.0;
    }
  for (int c = 0; c < 6; ++c) {
    int piv = c;
    for (int r = c + 1; r < 6; ++r)
      if (std::fabs(A[r][c]) > std::fabs(A[piv][c])) piv = r;
    if (piv != c)
      for (int j = 0; j < 12; ++j) std::swap(A[c][j], A[piv][j]);
    const double d = A[c][c];
    for (int j = 0; j < 12; ++j) A[c][j] /= d;
    for (int r = 0; r < 6; ++r) {
      if (r == c) continue;
      const double f = A[r][c];
      if (f == 0.0) continue;
      for (int j = 0; j < 12; ++j) A[r][j] -= f * A[c][j];
    }
  }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) out[6 * i + j] = A[i][6 + j];
}

}  // namespace

struct orc_odom {
  orc_odom_config cfg;
  std::deque<OState> states;   // ErrorStateKF::states_
  std::deque<Imu> imus;        // ErrorStateKF::ImuMeasurements_
  std::deque<Imu> imu_buffer;  // Odometry::imuBuffer_
  std::deque<std::unique_ptr<Scan>> cloud_buffer;  // Odometry::cloudBuffer_
  std::unique_ptr<Scan> lidar;  // Odometry::lidarMeas_
  double Q[12 * 12];
  double V[36];
  orc_map* map = nullptr;
  bool initialized = false;
  double T_il[16];
  double last_pose[16];
  int last_iterations = 0;
  uint64_t frames = 0, last_kept = 0, last_removed = 0;
  double stage_sum[3] = {0, 0, 0}, stage_max[3] = {0, 0, 0};
  int last_inserted = 0;
};

namespace {

// ErrorStateKF::process (src/ErrorStateKF.cpp:76-113)
void kf_process(orc_odom* o, const Imu& imu) {
  const OState& prev = o->states.back();
  const double dt = imu.t - prev.t;
  if (dt < 0.0) return;  // :80-82
  OState ns = prev;
  ns.t = imu.t;
  double R[9];
  quat_to_R(prev.q, R);
  double acc[3], w[3];
  for (int i = 0; i < 3; ++i) {
    acc[i] = imu.a[i] - prev.ba[i];
    w[i] = imu.w[i] - prev.bg[i];
  }
  double axis[3], dq[4];
  normalized(w, axis);
  angle_axis_to_quat(std::sqrt(dot3(w, w)) * dt, axis, dq);  // :88-90
  const double dt2 = dt * dt;
  double aw[3];
  for (int i = 0; i < 3; ++i) aw[i] = dot3(R + 3 * i, acc) + prev.g[i];  // R*a + g
  for (int i = 0; i < 3; ++i) {
    ns.p[i] = prev.p[i] + prev.v[i] * dt + 0.5 * aw[i] * dt2;  // :93-94
    ns.v[i] = prev.v[i] + aw[i] * dt;                          // :95
  }
  quat_mul(prev.q, dq, ns.q);  // :96

  // F_x (member, Identity + the blocks of :102-107)
  double F[N * N];
  std::memset(F, 0, sizeof F);
  for (int i = 0; i < N; ++i) F[i * N + i] = 1.0;
  const double S[9] = {0.0, -acc[2], acc[1], acc[2], 0.0, -acc[0], -acc[1], acc[0], 0.0};
  double RS[9];
  matmul(R, S, RS, 3, 3, 3);
  double dqc[4] = {-dq[0], -dq[1], -dq[2], dq[3]}, Rc[9];
  quat_to_R(dqc, Rc);
  for (int i = 0; i < 3; ++i) {
    F[i * N + 3 + i] = dt;        // (0,3)  I dt
    F[(3 + i) * N + 15 + i] = dt; // (3,15) I dt
    F[(6 + i) * N + 12 + i] = -dt; // (6,12) -I dt
    for (int j = 0; j < 3; ++j) {
      F[(3 + i) * N + 6 + j] = -RS[3 * i + j] * dt;  // (3,6)  -R [a]x dt
      F[(3 + i) * N + 9 + j] = -R[3 * i + j] * dt;   // (3,9)  -R dt
      F[(6 + i) * N + 6 + j] = Rc[3 * i + j];        // (6,6)  dq^* as a matrix
    }
  }
  // P = F P F^T + F_i Q_i F_i^T  (:109); F_i maps the 12 noise terms to rows 3..14
  double FP[N * N];
  matmul(F, prev.P, FP, N, N, N);
  matmul_bt(FP, F, ns.P, N, N, N);
  for (int i = 0; i < 12; ++i)
    for (int j = 0; j < 12; ++j) {
      double qv = o->Q[i * 12 + j];
      if (i < 6 && j < 6) qv *= dt2;       // :99
      else if (i >= 6 && j >= 6) qv *= dt; // :100
      ns.P[(3 + i) * N + 3 + j] += qv;
    }
  o->states.push_back(ns);
}

void replay(orc_odom* o, double lidar_end) {
  while (!o->imus.empty() && o->imus.front().t < lidar_end) o->imus.pop_front();
  // (copy: kf_process never touches o->imus)
  for (const Imu& m : o->imus) kf_process(o, m);
}

// ErrorStateKF::update (src/ErrorStateKF.cpp:115-162), first half: roll the
// states back to the scan end and form the ICP initial guess (:118-129)
void kf_rollback_guess(orc_odom* o, double lidar_end, double* guess) {
  while (!o->states.empty() && o->states.back().t > lidar_end) o->states.pop_back();  // :120-122
  const OState& prev = o->states.back();
  double Rg[9];
  quat_to_R(prev.q, Rg);
  std::memset(guess, 0, 16 * sizeof(double));
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) guess[4 * i + j] = Rg[3 * i + j];
    guess[4 * i + 3] = prev.p[i];
  }
  guess[15] = 1.0;
}

// second half (:132-161): Kalman update with the ICP pose `obs` as the
// observation, error injection, covariance reset, IMU replay
void kf_measurement_update(orc_odom* o, double lidar_end, const double* guess, const double* obs,
                           double* T_out) {
  OState ns = o->states.back();
  ns.t = lidar_end;
  const OState& prev = o->states.back();
  double res[6], Rg[9], Ro[9], Rrel[9];
  for (int i = 0; i < 3; ++i) {
    res[i] = obs[4 * i + 3] - guess[4 * i + 3];  // :133
    for (int j = 0; j < 3; ++j) {
      Ro[3 * i + j] = obs[4 * i + j];
      Rg[3 * i + j] = guess[4 * i + j];
    }
  }
  for (int i = 0; i < 3; ++i)  // guess.linear()^T * observation.linear()
    for (int j = 0; j < 3; ++j)
      Rrel[3 * i + j] = (Rg[i] * Ro[j] + Rg[3 + i] * Ro[3 + j]) + Rg[6 + i] * Ro[6 + j];
  R_to_rotvec(Rrel, res + 3);  // :134-135
  // K = P H^T (H P H^T + V)^-1 ; H picks error-state rows 0..2 and 6..8 (:55-57)
  const int hidx[6] = {0, 1, 2, 6, 7, 8};
  double PHt[N * 6], Sm[36], Si[36], K[N * 6];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < 6; ++j) PHt[i * 6 + j] = prev.P[i * N + hidx[j]];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) Sm[6 * i + j] = PHt[hidx[i] * 6 + j] + o->V[6 * i + j];
  inv6(Sm, Si);
  matmul(PHt, Si, K, N, 6, 6);
  double err[N];
  matmul(K, res, err, N, 6, 1);
  // P = (I - K H) P  (:142)
  double IKH[N * N];
  std::memset(IKH, 0, sizeof IKH);
  for (int i = 0; i < N; ++i) {
    IKH[i * N + i] = 1.0;
    for (int j = 0; j < 6; ++j) IKH[i * N + hidx[j]] -= K[i * 6 + j];
  }
  matmul(IKH, prev.P, ns.P, N, N, N);
  // injectError (:164-172)
  for (int i = 0; i < 3; ++i) {
    ns.p[i] += err[i];
    ns.v[i] += err[3 + i];
    ns.ba[i] += err[9 + i];
    ns.bg[i] += err[12 + i];
    ns.g[i] += err[15 + i];
  }
  double ax[3], dq[4], qn[4];
  normalized(err + 6, ax);
  angle_axis_to_quat(std::sqrt(dot3(err + 6, err + 6)), ax, dq);  // Utils.cpp:34-38
  quat_mul(ns.q, dq, qn);
  std::memcpy(ns.q, qn, sizeof qn);
  // reset (:174-180): G = I except G(6..8,6..8) = I - 0.5 [dtheta]x ; P = G P G^T
  double G[N * N];
  std::memset(G, 0, sizeof G);
  for (int i = 0; i < N; ++i) G[i * N + i] = 1.0;
  const double* e = err + 6;
  const double sk[9] = {0.0, -e[2], e[1], e[2], 0.0, -e[0], -e[1], e[0], 0.0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) G[(6 + i) * N + 6 + j] = (i == j ? 1.0 : 0.0) - 0.5 * sk[3 * i + j];
  double GP[N * N], Pn[N * N];
  matmul(G, ns.P, GP, N, N, N);
  matmul_bt(GP, G, Pn, N, N, N);
  std::memcpy(ns.P, Pn, sizeof Pn);
  o->states.push_back(ns);  // :146
  replay(o, lidar_end);     // :148-155
  double Rn[9];
  quat_to_R(ns.q, Rn);
  std::memset(T_out, 0, 16 * sizeof(double));
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T_out[4 * i + j] = Rn[3 * i + j];
    T_out[4 * i + 3] = ns.p[i];
  }
  T_out[15] = 1.0;
}

void kf_update(orc_odom* o, const double* xyz, const double* cov, size_t n, double lidar_end,
               double* T_out) {
  double guess[16], obs[16];
  kf_rollback_guess(o, lidar_end, guess);
  orc_icp_params prm = {o->cfg.max_iteration, o->cfg.neighbor_mode, o->cfg.icp_translation_sq_threshold,
                        o->cfg.icp_cosine_threshold};
  orc_align_info info;
  orc_align(o->map, xyz, cov, n, guess, &prm, obs, &info, nullptr, nullptr, nullptr, nullptr);  // :130
  o->last_iterations = info.iterations;
  kf_measurement_update(o, lidar_end, guess, obs, T_out);
}

// CloudPreprocessor::process over the whole state deque (src/Odometry.cpp:74)
long preprocess(orc_odom* o, Scan& s, bool with_states, std::vector<double>* xyz_out,
                std::vector<double>* cov_out) {
  const size_t n = s.time.size();
  std::vector<orc_state> st;
  if (with_states) {
    st.resize(o->states.size());
    for (size_t i = 0; i < st.size(); ++i) {
      st[i].timestamp = o->states[i].t;
      std::memcpy(st[i].position, o->states[i].p, sizeof st[i].position);
      std::memcpy(st[i].attitude_xyzw, o->states[i].q, sizeof st[i].attitude_xyzw);
    }
  }
  xyz_out->resize(3 * n);
  cov_out->resize(9 * n);
  std::vector<uint32_t> src(n);
  const long m = orc_preprocess(s.xyz.data(), s.time.data(), n, o->T_il, st.data(), st.size(),
                                o->cfg.preprocess_voxel_size, xyz_out->data(), cov_out->data(), src.data());
  if (m >= 0) {
    xyz_out->resize(3 * static_cast<size_t>(m));
    cov_out->resize(9 * static_cast<size_t>(m));
  }
  return m;
}

}  // namespace

extern "C" {

void orc_odom_default_config(orc_odom_config* c) {
  // config/hilti_config.yaml
  c->imu_update_rate = 400.0;
  const double ba[3] = {0.06080652138668933, 0.08353074835853214, 0.057072968234636895};
  const double bg[3] = {-0.0015351229643790084, -0.0013449146576507546, 0.00030127855524786183};
  const double g[3] = {0.01165152782783894, -0.008749296634685332, 9.804989173462031};
  const double and_[3] = {105.0, 105.0, 135.0};
  for (int i = 0; i < 3; ++i) {
    c->bias_a[i] = ba[i];
    c->bias_g[i] = bg[i];
    c->gravity[i] = g[i];
    c->accel_noise_density[i] = and_[i];
  }
  c->accel_zero_g_offset = 20.0;
  c->gyro_noise_density = 0.014;
  c->gyro_zero_rate_offset = 1.0;
  c->translation_noise = 1.0e-6;
  c->rotation_noise = 1.0e-6;
  const double q[4] = {0.7071068, -0.7071068, 0.0, 0.0};
  const double t[3] = {-0.001, -0.00855, 0.055};
  std::memcpy(c->lidar_quaternion_xyzw, q, sizeof q);
  std::memcpy(c->lidar_translation, t, sizeof t);
  c->map_voxel_size = 0.3;
  c->max_points_per_voxel = 1000;
  c->update_translation_sq_threshold = 1.0e-2;
  c->update_cosine_threshold = 0.985;
  c->remove_enabled = 1;
  c->remove_distance_threshold = 100.0;
  c->remove_period = 10.0;
  c->preprocess_voxel_size = 0.3;
  c->max_iteration = 100;
  c->neighbor_mode = 1;
  c->icp_translation_sq_threshold = 1.0e-6;
  c->icp_cosine_threshold = 0.9999;
}

// ErrorStateKF::ErrorStateKF (src/ErrorStateKF.cpp:8-60) + Odometry ctor (Odometry.hpp:22-34)
orc_odom* orc_odom_create(const orc_odom_config* cfg) {
  orc_odom* o = new orc_odom;
  o->cfg = *cfg;
  constexpr double kG = 9.81;  // GRAVITY_MAGNITUDE :11
  OState s0;
  for (int i = 0; i < 3; ++i) {
    s0.ba[i] = cfg->bias_a[i];
    s0.bg[i] = cfg->bias_g[i];
    s0.g[i] = cfg->gravity[i];
  }
  o->states.push_back(s0);
  const double sr = std::sqrt(cfg->imu_update_rate);
  std::memset(o->Q, 0, sizeof o->Q);
  for (int i = 0; i < 3; ++i) {
    const double sa = cfg->accel_noise_density[i] * kG * sr;                 // :29-32
    o->Q[i * 12 + i] = sa * sa;                                              // :38-39
    o->Q[(3 + i) * 12 + 3 + i] = std::pow(cfg->gyro_noise_density * sr * M_PI / 180.0, 2.0);        // :33,40
    o->Q[(6 + i) * 12 + 6 + i] = std::pow(cfg->accel_zero_g_offset * sr * 1e-3 * kG, 2.0);          // :34,41
    o->Q[(9 + i) * 12 + 9 + i] = std::pow(cfg->gyro_zero_rate_offset * sr * M_PI / 180.0, 2.0);     // :35,42
  }
  std::memset(o->V, 0, sizeof o->V);
  for (int i = 0; i < 3; ++i) {
    o->V[6 * i + i] = cfg->translation_noise;
    o->V[6 * (3 + i) + 3 + i] = cfg->rotation_noise;
  }
  double R[9];
  orc_quat_to_matrix(cfg->lidar_quaternion_xyzw, R);  // CloudPreprocessor.hpp:20-28
  std::memset(o->T_il, 0, sizeof o->T_il);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) o->T_il[4 * i + j] = R[3 * i + j];
    o->T_il[4 * i + 3] = cfg->lidar_translation[i];
  }
  o->T_il[15] = 1.0;
  o->map = orc_map_create(cfg->map_voxel_size, cfg->max_points_per_voxel);
  orc_map_set_update_params(o->map, cfg->update_translation_sq_threshold, cfg->update_cosine_threshold,
                            cfg->remove_enabled, cfg->remove_distance_threshold, cfg->remove_period);
  std::memset(o->last_pose, 0, sizeof o->last_pose);
  o->last_pose[0] = o->last_pose[5] = o->last_pose[10] = o->last_pose[15] = 1.0;
  return o;
}

void orc_odom_destroy(orc_odom* o) {
  if (!o) return;
  orc_map_destroy(o->map);
  delete o;
}

// ImuSubscriber callback -> imuBuffer (include/ESKF_LIO/Subscriber.hpp:38-52)
void orc_odom_feed_imu(orc_odom* o, double t, const double gyro[3], const double acc[3]) {
  Imu m;
  m.t = t;
  std::memcpy(m.w, gyro, sizeof m.w);
  std::memcpy(m.a, acc, sizeof m.a);
  o->imu_buffer.push_back(m);
}

// LidarSubscriber callback -> cloudBuffer (Subscriber.hpp:80-103)
void orc_odom_feed_lidar(orc_odom* o, const double* xyz, const double* point_time, size_t n,
                         double start_time, double end_time) {
  std::unique_ptr<Scan> s(new Scan);
  s->xyz.assign(xyz, xyz + 3 * n);
  s->time.assign(point_time, point_time + n);
  s->start = start_time;
  s->end = end_time;
  o->cloud_buffer.push_back(std::move(s));
}

// one trip of the while loop of Odometry::run (src/Odometry.cpp:16-98).
// Returns 1 when a LiDAR frame was consumed, 0 otherwise, -1 on a deskew error.
int orc_odom_spin_once(orc_odom* o) {
  // :23-41 drain the IMU queue
  while (!o->imu_buffer.empty()) {
    const Imu m = o->imu_buffer.front();
    o->imu_buffer.pop_front();
    if (o->initialized) kf_process(o, m);  // :30
    o->imus.push_back(m);                  // feedImu :31,36
  }
  if (!o->lidar && !o->cloud_buffer.empty()) {  // :44-49
    o->lidar = std::move(o->cloud_buffer.front());
    o->cloud_buffer.pop_front();
  }
  if (!o->lidar) return 0;
  const double lidar_end = o->lidar->end;
  std::vector<double> p, c;
  if (!o->initialized) {  // :55-63
    o->initialized = true;
    o->states[0].t = lidar_end;  // ErrorStateKF::initialize :62-74
    replay(o, lidar_end);
    std::unique_ptr<Scan> s = std::move(o->lidar);
    const long m = preprocess(o, *s, false, &p, &c);
    if (m < 0) return -1;
    double I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    o->last_inserted = orc_map_update(o->map, p.data(), c.data(), static_cast<size_t>(m), I16, 1,
                                      lidar_end, &o->last_removed);
    o->last_kept = static_cast<uint64_t>(m);
    return 1;
  }
  if (o->states.back().t < lidar_end) return 0;  // :65-69 wait for the next IMU sample
  const double t0 = omp_get_wtime();
  const long m = preprocess(o, *o->lidar, true, &p, &c);  // :74
  if (m < 0) return -1;
  const double t1 = omp_get_wtime();
  kf_update(o, p.data(), c.data(), static_cast<size_t>(m), lidar_end, o->last_pose);  // :79
  const double t2 = omp_get_wtime();
  o->lidar.reset();
  o->last_inserted = orc_map_update(o->map, p.data(), c.data(), static_cast<size_t>(m), o->last_pose, 0,
                                    lidar_end, &o->last_removed);  // :86
  const double t3 = omp_get_wtime();
  const double d[3] = {t1 - t0, t2 - t1, t3 - t2};
  for (int i = 0; i < 3; ++i) {
    o->stage_sum[i] += d[i];
    if (d[i] > o->stage_max[i]) o->stage_max[i] = d[i];
  }
  ++o->frames;
  o->last_kept = static_cast<uint64_t>(m);
  return 1;
}

void orc_odom_last_pose(const orc_odom* o, double T16[16]) { std::memcpy(T16, o->last_pose, 16 * sizeof(double)); }

void orc_odom_info(const orc_odom* o, orc_odom_info_t* out) {
  out->frames = o->frames;
  out->n_states = o->states.size();
  out->map_voxels = orc_map_size(o->map);
  out->last_kept = o->last_kept;
  out->last_removed = o->last_removed;
  out->last_iterations = o->last_iterations;
  out->last_inserted = o->last_inserted;
  for (int i = 0; i < 3; ++i) {
    out->stage_avg_ms[i] = o->frames ? 1e3 * o->stage_sum[i] / static_cast<double>(o->frames) : 0.0;
    out->stage_max_ms[i] = 1e3 * o->stage_max[i];
  }
}

// the newest state: t, p(3), v(3), q xyzw(4), ba(3), bg(3), g(3) = 20 doubles; P (nullable) 324
void orc_odom_last_state(const orc_odom* o, double out20[20], double* P324) {
  const OState& s = o->states.back();
  out20[0] = s.t;
  std::memcpy(out20 + 1, s.p, 24);
  std::memcpy(out20 + 4, s.v, 24);
  std::memcpy(out20 + 7, s.q, 32);
  std::memcpy(out20 + 11, s.ba, 24);
  std::memcpy(out20 + 14, s.bg, 24);
  std::memcpy(out20 + 17, s.g, 24);
  if (P324) std::memcpy(P324, s.P, sizeof s.P);
}

const orc_map* orc_odom_map(const orc_odom* o) { return o->map; }

// ---- filter pieces alone (KATs / cross-checks against the NumPy restatement)
void orc_kf_process(orc_odom* o, double t, const double gyro[3], const double acc[3]) {
  Imu m;
  m.t = t;
  std::memcpy(m.w, gyro, sizeof m.w);
  std::memcpy(m.a, acc, sizeof m.a);
  kf_process(o, m);
}

void orc_rotation_matrix_to_vector(const double R9[9], double r3[3]) { R_to_rotvec(R9, r3); }

// ErrorStateKF::update with the ICP result supplied by the caller
void orc_kf_update_with_observation(orc_odom* o, double lidar_end, const double obs16[16],
                                    double guess16_out[16], double T_out16[16]) {
  kf_rollback_guess(o, lidar_end, guess16_out);
  kf_measurement_update(o, lidar_end, guess16_out, obs16, T_out16);
}

}  // extern "C"
