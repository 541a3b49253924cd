// eskf_oracle.cpp — CPU ORACLE (test infrastructure; see eskf_oracle.h).
//
// Dependency-free fp64 restatement of the ESKF_LIO hot path.  PARITY UNPINNED
// by the reference (it has no tests); pinned by oracle/np_oracle.py, the KATs
// and tests/golden/.  Build: g++ -O3 -fopenmp -ffp-contract=off -shared -fPIC.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference).  Eigen / Open3D internals are restated from their
// published algorithms (both are un-vendored, unpinned dependencies).

#include "eskf_oracle.h"

#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <unordered_map>
#include <utility>
#include <vector>

namespace {

// ---------------------------------------------------------------- small math
struct M3 {
  double m[9];  // row-major
};

inline double dot3(double a0, double b0, double a1, double b1, double a2, double b2) {
  return (a0 * b0 + a1 * b1) + a2 * b2;
}

inline M3 mul33(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C.m[3 * i + j] = dot3(A.m[3 * i], B.m[j], A.m[3 * i + 1], B.m[3 + j], A.m[3 * i + 2], B.m[6 + j]);
  return C;
}

inline M3 transpose33(const M3& A) {
  M3 T;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T.m[3 * i + j] = A.m[3 * j + i];
  return T;
}

inline void mulvec3(const double* R, const double* v, double* out) {
  double x = v[0], y = v[1], z = v[2];
  out[0] = dot3(R[0], x, R[1], y, R[2], z);
  out[1] = dot3(R[3], x, R[4], y, R[5], z);
  out[2] = dot3(R[6], x, R[7], y, R[8], z);
}

struct Iso {
  double R[9];
  double t[3];
};

inline Iso iso_from16(const double* T) {
  Iso r;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) r.R[3 * i + j] = T[4 * i + j];
    r.t[i] = T[4 * i + 3];
  }
  return r;
}

inline void iso_to16(const Iso& a, double* T) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T[4 * i + j] = a.R[3 * i + j];
    T[4 * i + 3] = a.t[i];
  }
  T[12] = 0.0;
  T[13] = 0.0;
  T[14] = 0.0;
  T[15] = 1.0;
}

// Eigen Isometry3d * Isometry3d: linear = La*Lb ; translation = La*tb + ta
inline Iso iso_mul(const Iso& a, const Iso& b) {
  Iso r;
  M3 A, B;
  std::memcpy(A.m, a.R, sizeof A.m);
  std::memcpy(B.m, b.R, sizeof B.m);
  M3 C = mul33(A, B);
  std::memcpy(r.R, C.m, sizeof C.m);
  double v[3];
  mulvec3(a.R, b.t, v);
  for (int i = 0; i < 3; ++i) r.t[i] = v[i] + a.t[i];
  return r;
}

// Eigen Isometry3d::inverse(): linear = R^T ; translation = -(R^T t)
inline Iso iso_inv(const Iso& a) {
  Iso r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.R[3 * i + j] = a.R[3 * j + i];
  double v[3];
  mulvec3(r.R, a.t, v);
  for (int i = 0; i < 3; ++i) r.t[i] = -v[i];
  return r;
}

// Isometry3d * Vector3d  (Utils.cpp:13-20 transformPoints): (R p) + t
inline void iso_apply(const Iso& a, double* p) {
  double v[3];
  mulvec3(a.R, p, v);
  p[0] = v[0] + a.t[0];
  p[1] = v[1] + a.t[1];
  p[2] = v[2] + a.t[2];
}

// Eigen Matrix3d::inverse(): cofactors scaled by 1/det.
inline M3 inv33(const M3& A) {
  const double* a = A.m;
  double c00 = a[4] * a[8] - a[5] * a[7];
  double c01 = a[5] * a[6] - a[3] * a[8];
  double c02 = a[3] * a[7] - a[4] * a[6];
  double det = dot3(a[0], c00, a[1], c01, a[2], c02);
  double id = 1.0 / det;
  M3 R;
  R.m[0] = c00 * id;
  R.m[3] = c01 * id;
  R.m[6] = c02 * id;
  R.m[1] = (a[2] * a[7] - a[1] * a[8]) * id;
  R.m[4] = (a[0] * a[8] - a[2] * a[6]) * id;
  R.m[7] = (a[1] * a[6] - a[0] * a[7]) * id;
  R.m[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  R.m[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  R.m[8] = (a[0] * a[4] - a[1] * a[3]) * id;
  return R;
}

// Eigen AngleAxisd(angle, axis).toRotationMatrix()
inline void angle_axis_to_matrix(double angle, const double* ax, double* R) {
  double s = std::sin(angle), c = std::cos(angle);
  double sx = s * ax[0], sy = s * ax[1], sz = s * ax[2];
  double cx = (1.0 - c) * ax[0], cy = (1.0 - c) * ax[1], cz = (1.0 - c) * ax[2];
  double tmp;
  tmp = cx * ax[1];
  R[1] = tmp - sz;
  R[3] = tmp + sz;
  tmp = cx * ax[2];
  R[2] = tmp + sy;
  R[6] = tmp - sy;
  tmp = cy * ax[2];
  R[5] = tmp - sx;
  R[7] = tmp + sx;
  R[0] = cx * ax[0] + c;
  R[4] = cy * ax[1] + c;
  R[8] = cz * ax[2] + c;
}

inline double norm3(const double* v) { return std::sqrt(dot3(v[0], v[0], v[1], v[1], v[2], v[2])); }

// Eigen normalized(): v / |v| if |v|^2 > 0 else v
inline void normalized3(const double* v, double* out) {
  double z = dot3(v[0], v[0], v[1], v[1], v[2], v[2]);
  if (z > 0.0) {
    double n = std::sqrt(z);
    out[0] = v[0] / n;
    out[1] = v[1] / n;
    out[2] = v[2] / n;
  } else {
    out[0] = v[0];
    out[1] = v[1];
    out[2] = v[2];
  }
}

// ------------------------------------------------------------------ voxels
constexpr int64_t kBias = 1 << 20;

inline uint64_t pack_key(int32_t x, int32_t y, int32_t z) {
  return (static_cast<uint64_t>(x + kBias) << 42) | (static_cast<uint64_t>(y + kBias) << 21) |
         static_cast<uint64_t>(z + kBias);
}

inline void unpack_key(uint64_t k, int32_t* out) {
  out[0] = static_cast<int32_t>((k >> 42) & 0x1FFFFF) - static_cast<int32_t>(kBias);
  out[1] = static_cast<int32_t>((k >> 21) & 0x1FFFFF) - static_cast<int32_t>(kBias);
  out[2] = static_cast<int32_t>(k & 0x1FFFFF) - static_cast<int32_t>(kBias);
}

// (point / voxelSize).array().floor().cast<int>()
inline void voxel_index(const double* p, double v, int32_t* out) {
  out[0] = static_cast<int32_t>(std::floor(p[0] / v));
  out[1] = static_cast<int32_t>(std::floor(p[1] / v));
  out[2] = static_cast<int32_t>(std::floor(p[2] / v));
}

struct KeyHash {
  size_t operator()(uint64_t k) const {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return static_cast<size_t>(k);
  }
};

// LocalMap::Voxel (LocalMap.hpp:63-89) minus the raw point list (GUI/export only)
struct Voxel {
  uint64_t n;
  double mean[3];
  double cov[9];
};

}  // namespace

struct orc_map {
  double voxel_size;
  uint64_t cap;
  double upd_trans_sq = 1e-2;
  double upd_cos = 0.985;
  int remove_enabled = 0;
  double distance_thr = 100.0;
  double remove_period = 10.0;
  double current_remove_time = std::numeric_limits<double>::lowest();  // LocalMap.hpp:40
  Iso prev;  // prevTransform_ (uninitialised in the reference; identity here)
  std::unordered_map<uint64_t, Voxel, KeyHash> grid;
};

namespace {

// Voxel::Voxel / Voxel::addPoint  (LocalMap.hpp:72-87)
inline void voxel_add(orc_map* m, uint64_t key, const double* p, const double* C) {
  auto it = m->grid.find(key);
  if (it == m->grid.end()) {
    Voxel v;
    v.n = 1;
    std::memcpy(v.mean, p, sizeof v.mean);
    std::memcpy(v.cov, C, sizeof v.cov);
    m->grid.emplace(key, v);
    return;
  }
  Voxel& v = it->second;
  if (v.n < m->cap) {
    double n = static_cast<double>(v.n);
    double n1 = static_cast<double>(v.n + 1);
    for (int i = 0; i < 3; ++i) v.mean[i] = (n * v.mean[i] + p[i]) / n1;
    for (int i = 0; i < 9; ++i) v.cov[i] = (n * v.cov[i] + C[i]) / n1;
    ++v.n;
  }
}

// ICP::computeJTJAndJTr, dense exactly as written (Registration.cpp:83-102)
inline void jtj_jtr(const double* p, const double* mu, const double* C, double* H, double* b) {
  double J[18];  // 3x6 row-major: [I | -skew(p)]
  std::memset(J, 0, sizeof J);
  J[0] = 1.0;
  J[7] = 1.0;
  J[14] = 1.0;
  // -skew(p) = [0 pz -py; -pz 0 px; py -px 0]
  J[4] = p[2];
  J[5] = -p[1];
  J[9] = -p[2];
  J[11] = p[0];
  J[15] = p[1];
  J[16] = -p[0];
  M3 Cm;
  std::memcpy(Cm.m, C, sizeof Cm.m);
  M3 W = inv33(Cm);
  double JT[18];  // 6x3 = J^T * W
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 3; ++j)
      JT[3 * i + j] = dot3(J[i], W.m[j], J[6 + i], W.m[3 + j], J[12 + i], W.m[6 + j]);
  double r[3] = {p[0] - mu[0], p[1] - mu[1], p[2] - mu[2]};
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < 6; ++j)
      H[6 * i + j] = dot3(JT[3 * i], J[j], JT[3 * i + 1], J[6 + j], JT[3 * i + 2], J[12 + j]);
    b[i] = dot3(JT[3 * i], r[0], JT[3 * i + 1], r[1], JT[3 * i + 2], r[2]);
  }
}

// Eigen LDLT<Matrix6d>: diagonal pivoting, D pseudo-inverse with tolerance
// numeric_limits<double>::min() in solve().
void ldlt_solve6(const double* Hin, const double* bin, double* x) {
  constexpr int N = 6;
  double A[N][N];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) A[i][j] = Hin[N * i + j];
  int perm[N];
  for (int i = 0; i < N; ++i) perm[i] = i;
  double L[N][N];
  double D[N];
  for (int i = 0; i < N; ++i) {
    D[i] = 0.0;
    for (int j = 0; j < N; ++j) L[i][j] = (i == j) ? 1.0 : 0.0;
  }
  for (int k = 0; k < N; ++k) {
    int piv = k;
    double best = std::fabs(A[k][k]);
    for (int i = k + 1; i < N; ++i)
      if (std::fabs(A[i][i]) > best) {
        best = std::fabs(A[i][i]);
        piv = i;
      }
    if (piv != k) {
      for (int j = 0; j < N; ++j) std::swap(A[k][j], A[piv][j]);
      for (int i = 0; i < N; ++i) std::swap(A[i][k], A[i][piv]);
      for (int j = 0; j < k; ++j) std::swap(L[k][j], L[piv][j]);
      std::swap(perm[k], perm[piv]);
    }
    double d = A[k][k];
    D[k] = d;
    if (!(std::fabs(d) > 0.0)) {
      if (k == 0) break;  // whole diagonal is zero
      continue;
    }
    for (int i = k + 1; i < N; ++i) L[i][k] = A[i][k] / d;
    for (int i = k + 1; i < N; ++i)
      for (int j = k + 1; j < N; ++j) A[i][j] -= L[i][k] * d * L[j][k];
  }
  double y[N];
  for (int i = 0; i < N; ++i) y[i] = bin[perm[i]];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < i; ++j) y[i] -= L[i][j] * y[j];
  const double tol = std::numeric_limits<double>::min();
  for (int i = 0; i < N; ++i) y[i] = (std::fabs(D[i]) > tol) ? y[i] / D[i] : 0.0;
  for (int i = N - 1; i >= 0; --i)
    for (int j = i + 1; j < N; ++j) y[i] -= L[j][i] * y[j];
  for (int i = 0; i < N; ++i) x[perm[i]] = y[i];
}

void transform_cloud(double* xyz, double* cov, size_t n, const double* T) {
  Iso a = iso_from16(T);
  M3 R;
  std::memcpy(R.m, a.R, sizeof R.m);
  M3 Rt = transpose33(R);
  for (size_t i = 0; i < n; ++i) {
    double* p = xyz + 3 * i;
    double x = p[0], y = p[1], z = p[2];
    p[0] = dot3(a.R[0], x, a.R[1], y, a.R[2], z) + a.t[0];
    p[1] = dot3(a.R[3], x, a.R[4], y, a.R[5], z) + a.t[1];
    p[2] = dot3(a.R[6], x, a.R[7], y, a.R[8], z) + a.t[2];
    if (cov) {
      M3 C;
      std::memcpy(C.m, cov + 9 * i, sizeof C.m);
      M3 out = mul33(mul33(R, C), Rt);
      std::memcpy(cov + 9 * i, out.m, sizeof out.m);
    }
  }
}

const int kOff7[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};

size_t linearize(const orc_map* map, const double* xyz, const double* cov, size_t n, int mode,
                 double* H, double* b, uint8_t* hit) {
  const int nn = (mode == 7) ? 7 : 1;
  const int nt = omp_get_max_threads();
  std::vector<double> Hp(static_cast<size_t>(nt) * 36, 0.0), bp(static_cast<size_t>(nt) * 6, 0.0);
  std::vector<size_t> cp(nt, 0);
#pragma omp parallel num_threads(nt)
  {
    const int tid = omp_get_thread_num();
    double* Ht = Hp.data() + 36 * tid;
    double* bt = bp.data() + 6 * tid;
    size_t cnt = 0;
#pragma omp for schedule(static)
    for (long long i = 0; i < static_cast<long long>(n); ++i) {
      int32_t k[3];
      voxel_index(xyz + 3 * i, map->voxel_size, k);
      for (int o = 0; o < nn; ++o) {
        auto f = map->grid.find(pack_key(k[0] + kOff7[o][0], k[1] + kOff7[o][1], k[2] + kOff7[o][2]));
        bool found = f != map->grid.end();
        if (hit) hit[nn * i + o] = found ? 1 : 0;
        if (!found) continue;
        double C[9];
        for (int j = 0; j < 9; ++j) C[j] = cov[9 * i + j] + f->second.cov[j];
        double Hi[36], bi[6];
        jtj_jtr(xyz + 3 * i, f->second.mean, C, Hi, bi);
        for (int j = 0; j < 36; ++j) Ht[j] += Hi[j];
        for (int j = 0; j < 6; ++j) bt[j] += bi[j];
        ++cnt;
      }
    }
    cp[tid] = cnt;
  }
  // the reference merges thread-private sums under `omp critical` in arrival
  // order (Registration.cpp:71-75); merged in thread order here => deterministic
  std::memset(H, 0, 36 * sizeof(double));
  std::memset(b, 0, 6 * sizeof(double));
  size_t total = 0;
  for (int t = 0; t < nt; ++t) {
    for (int j = 0; j < 36; ++j) H[j] += Hp[36 * t + j];
    for (int j = 0; j < 6; ++j) b[j] += bp[6 * t + j];
    total += cp[t];
  }
  return total;
}

void se3_to_SE3(const double* se3, double* T) {
  double J[9];
  orc_compute_J(se3 + 3, J);
  Iso a;
  mulvec3(J, se3, a.t);
  orc_rotvec_to_matrix(se3 + 3, a.R);
  iso_to16(a, T);
}

// ---------------------------------------------------------------- KD-tree
struct KdNode {
  int32_t left, right;  // children, or -1
  int32_t begin, end;   // leaf range in idx
  int32_t dim;
  double split;
};

struct KdTree {
  const double* pts;
  std::vector<int32_t> idx;
  std::vector<KdNode> nodes;
  static constexpr int kLeaf = 15;  // nanoflann leaf_max_size used by Open3D

  int32_t build(int32_t b, int32_t e) {
    KdNode nd;
    nd.left = nd.right = -1;
    nd.begin = b;
    nd.end = e;
    nd.dim = 0;
    nd.split = 0.0;
    int32_t id = static_cast<int32_t>(nodes.size());
    nodes.push_back(nd);
    if (e - b <= kLeaf) return id;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int32_t i = b; i < e; ++i)
      for (int d = 0; d < 3; ++d) {
        double v = pts[3 * idx[i] + d];
        lo[d] = std::min(lo[d], v);
        hi[d] = std::max(hi[d], v);
      }
    int dim = 0;
    for (int d = 1; d < 3; ++d)
      if (hi[d] - lo[d] > hi[dim] - lo[dim]) dim = d;
    int32_t mid = b + (e - b) / 2;
    std::nth_element(idx.begin() + b, idx.begin() + mid, idx.begin() + e,
                     [&](int32_t a, int32_t c) { return pts[3 * a + dim] < pts[3 * c + dim]; });
    double split = pts[3 * idx[mid] + dim];
    int32_t l = build(b, mid);
    int32_t r = build(mid, e);
    nodes[id].left = l;
    nodes[id].right = r;
    nodes[id].dim = dim;
    nodes[id].split = split;
    return id;
  }

  void init(const double* p, size_t n) {
    pts = p;
    idx.resize(n);
    for (size_t i = 0; i < n; ++i) idx[i] = static_cast<int32_t>(i);
    nodes.clear();
    nodes.reserve(2 * n / kLeaf + 16);
    if (n) build(0, static_cast<int32_t>(n));
  }

  // bounded max-heap of (d2, idx)
  void search(int32_t node, const double* q, int k, std::vector<std::pair<double, int32_t>>& heap) const {
    const KdNode& nd = nodes[node];
    if (nd.left < 0) {
      for (int32_t i = nd.begin; i < nd.end; ++i) {
        const double* p = pts + 3 * idx[i];
        double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
        double d2 = (dx * dx + dy * dy) + dz * dz;
        if (static_cast<int>(heap.size()) < k) {
          heap.emplace_back(d2, idx[i]);
          std::push_heap(heap.begin(), heap.end());
        } else if (std::make_pair(d2, idx[i]) < heap.front()) {
          std::pop_heap(heap.begin(), heap.end());
          heap.back() = std::make_pair(d2, idx[i]);
          std::push_heap(heap.begin(), heap.end());
        }
      }
      return;
    }
    double diff = q[nd.dim] - nd.split;
    int32_t first = diff < 0 ? nd.left : nd.right;
    int32_t second = diff < 0 ? nd.right : nd.left;
    search(first, q, k, heap);
    if (static_cast<int>(heap.size()) < k || diff * diff <= heap.front().first) search(second, q, k, heap);
  }

  void knn(const double* q, int k, int32_t* out_idx, double* out_d2) const {
    std::vector<std::pair<double, int32_t>> heap;
    heap.reserve(k + 1);
    if (!nodes.empty()) search(0, q, k, heap);
    std::sort(heap.begin(), heap.end());
    for (int j = 0; j < k; ++j) {
      if (j < static_cast<int>(heap.size())) {
        out_idx[j] = heap[j].second;
        if (out_d2) out_d2[j] = heap[j].first;
      } else {
        out_idx[j] = -1;
        if (out_d2) out_d2[j] = -1.0;
      }
    }
  }
};

// symmetric 3x3 Jacobi eigen-decomposition; eigenvalues descending, V columns
void eig_sym3(const double* C, double* w, double* V) {
  double A[3][3] = {{C[0], 0.5 * (C[1] + C[3]), 0.5 * (C[2] + C[6])},
                    {0.5 * (C[1] + C[3]), C[4], 0.5 * (C[5] + C[7])},
                    {0.5 * (C[2] + C[6]), 0.5 * (C[5] + C[7]), C[8]}};
  double U[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-34 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double ukp = U[k][p], ukq = U[k][q];
          U[k][p] = c * ukp - s * ukq;
          U[k][q] = s * ukp + c * ukq;
        }
      }
  }
  int order[3] = {0, 1, 2};
  std::sort(order, order + 3, [&](int a, int b) { return A[a][a] > A[b][b]; });
  for (int j = 0; j < 3; ++j) {
    w[j] = A[order[j]][order[j]];
    for (int i = 0; i < 3; ++i) V[3 * i + j] = U[i][order[j]];
  }
}

}  // namespace

// =========================================================================
extern "C" {

int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }

// Utils.cpp:5-11
void orc_skew(const double v[3], double o[9]) {
  o[0] = 0.0;
  o[1] = -v[2];
  o[2] = v[1];
  o[3] = v[2];
  o[4] = 0.0;
  o[5] = -v[0];
  o[6] = -v[1];
  o[7] = v[0];
  o[8] = 0.0;
}

// Utils.cpp:40-54  SO(3) left Jacobian
void orc_compute_J(const double r[3], double J[9]) {
  double angle = norm3(r);
  double a[3];
  normalized3(r, a);
  if (angle < 1e-6) {
    for (int i = 0; i < 9; ++i) J[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  double f1 = std::sin(angle) / angle;
  double f2 = (1.0 - std::cos(angle)) / angle;
  double S[9];
  orc_skew(a, S);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double id = (i == j) ? 1.0 : 0.0;
      J[3 * i + j] = (f1 * id + (1.0 - f1) * a[i] * a[j]) + f2 * S[3 * i + j];
    }
}

// Utils.cpp:28-32
void orc_rotvec_to_matrix(const double r[3], double R[9]) {
  double a[3];
  normalized3(r, a);
  angle_axis_to_matrix(norm3(r), a, R);
}

// Utils.cpp:56-63
void orc_se3_to_SE3(const double se3[6], double T16[16]) { se3_to_SE3(se3, T16); }

// Eigen Quaterniond::toRotationMatrix()
void orc_quat_to_matrix(const double q[4], double R[9]) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  double twx = tx * w, twy = ty * w, twz = tz * w;
  double txx = tx * x, txy = ty * x, txz = tz * x;
  double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0 - (tyy + tzz);
  R[1] = txy - twz;
  R[2] = txz + twy;
  R[3] = txy + twz;
  R[4] = 1.0 - (txx + tzz);
  R[5] = tyz - twx;
  R[6] = txz - twy;
  R[7] = tyz + twx;
  R[8] = 1.0 - (txx + tyy);
}

// Utils.cpp:65-75 (Eigen Quaterniond::slerp restated)
void orc_interpolate_SE3(const orc_state* s1, const orc_state* s2, double t, double T16[16]) {
  double factor = (t - s1->timestamp) / (s2->timestamp - s1->timestamp + 1e-6);
  const double* a = s1->attitude_xyzw;
  const double* b = s2->attitude_xyzw;
  double d = (a[0] * b[0] + a[1] * b[1]) + (a[2] * b[2] + a[3] * b[3]);
  double absD = std::fabs(d);
  double s0, sc1;
  if (absD >= 1.0 - std::numeric_limits<double>::epsilon()) {
    s0 = 1.0 - factor;
    sc1 = factor;
  } else {
    double theta = std::acos(absD);
    double sinTheta = std::sin(theta);
    s0 = std::sin((1.0 - factor) * theta) / sinTheta;
    sc1 = std::sin(factor * theta) / sinTheta;
  }
  if (d < 0.0) sc1 = -sc1;
  double q[4];
  for (int i = 0; i < 4; ++i) q[i] = s0 * a[i] + sc1 * b[i];
  Iso r;
  orc_quat_to_matrix(q, r.R);
  for (int i = 0; i < 3; ++i) r.t[i] = s1->position[i] + factor * (s2->position[i] - s1->position[i]);
  iso_to16(r, T16);
}

void orc_transform_cloud(double* xyz, double* cov, size_t n, const double T16[16]) {
  transform_cloud(xyz, cov, n, T16);
}

void orc_voxel_index(const double* xyz, size_t n, double voxel_size, int32_t* out) {
  for (size_t i = 0; i < n; ++i) voxel_index(xyz + 3 * i, voxel_size, out + 3 * i);
}

void orc_jtj_jtr(const double p[3], const double mu[3], const double C9[9], double H36[36],
                 double b6[6]) {
  jtj_jtr(p, mu, C9, H36, b6);
}

void orc_ldlt_solve6(const double H36[36], const double b6[6], double x6[6]) {
  ldlt_solve6(H36, b6, x6);
}

// Registration.cpp:37-50
int orc_convergence_check(const double T16[16], double trans_sq_thr, double cos_thr) {
  double cosine = 0.5 * (((T16[0] + T16[5]) + T16[10]) - 1.0);
  if (cosine < cos_thr) return 0;
  double tsq = dot3(T16[3], T16[3], T16[7], T16[7], T16[11], T16[11]);
  if (tsq > trans_sq_thr) return 0;
  return 1;
}

size_t orc_linearize(const orc_map* map, const double* xyz, const double* cov, size_t n,
                     int neighbor_mode, double H36[36], double b6[6], uint8_t* hit) {
  return linearize(map, xyz, cov, n, neighbor_mode, H36, b6, hit);
}

// Registration.cpp:7-35
int orc_align(const orc_map* map, const double* xyz, const double* cov, size_t n,
              const double guess16[16], const orc_icp_params* prm, double T_out16[16],
              orc_align_info* info, double* trace_H, double* trace_b, uint64_t* trace_ncorr,
              double* trace_step) {
  std::vector<double> pts(xyz, xyz + 3 * n), covs(cov, cov + 9 * n);  // :11 deep copy
  Iso total = iso_from16(guess16);
  transform_cloud(pts.data(), covs.data(), n, guess16);  // :13
  int iters = 0, converged = 0;
  for (int i = 0; i < prm->max_iteration; ++i) {
    double H[36], b[6], nb[6], se3[6], step16[16];
    size_t nc = linearize(map, pts.data(), covs.data(), n, prm->neighbor_mode, H, b, nullptr);
    for (int j = 0; j < 6; ++j) nb[j] = -b[j];
    ldlt_solve6(H, nb, se3);  // :78
    se3_to_SE3(se3, step16);  // :79
    if (trace_H) std::memcpy(trace_H + 36 * i, H, sizeof H);
    if (trace_b) std::memcpy(trace_b + 6 * i, b, sizeof b);
    if (trace_ncorr) trace_ncorr[i] = nc;
    if (trace_step) std::memcpy(trace_step + 16 * i, step16, sizeof step16);
    total = iso_mul(iso_from16(step16), total);  // :20
    ++iters;
    if (orc_convergence_check(step16, prm->translation_sq_threshold, prm->cosine_threshold)) {
      converged = 1;  // :22-25
      break;
    }
    transform_cloud(pts.data(), covs.data(), n, step16);  // :27
  }
  iso_to16(total, T_out16);
  if (info) {
    info->iterations = iters;
    info->converged = converged;
  }
  return 0;
}

orc_map* orc_map_create(double voxel_size, uint64_t cap) {
  orc_map* m = new orc_map();
  m->voxel_size = voxel_size;
  m->cap = cap;
  for (int i = 0; i < 9; ++i) m->prev.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  m->prev.t[0] = m->prev.t[1] = m->prev.t[2] = 0.0;
  return m;
}

void orc_map_destroy(orc_map* m) { delete m; }

void orc_map_set_update_params(orc_map* m, double trans_sq_thr, double cos_thr, int remove_enabled,
                               double distance_thr, double remove_period) {
  m->upd_trans_sq = trans_sq_thr;
  m->upd_cos = cos_thr;
  m->remove_enabled = remove_enabled;
  m->distance_thr = distance_thr;
  m->remove_period = remove_period;
}

// LocalMap.cpp:132-147
int orc_needs_map_update(const double prev16[16], const double cur16[16], double trans_sq_thr,
                         double cos_thr) {
  Iso moved = iso_mul(iso_inv(iso_from16(prev16)), iso_from16(cur16));
  double cosine = 0.5 * (((moved.R[0] + moved.R[4]) + moved.R[8]) - 1.0);
  if (cosine < cos_thr) return 1;
  double tsq = dot3(moved.t[0], moved.t[0], moved.t[1], moved.t[1], moved.t[2], moved.t[2]);
  if (tsq > trans_sq_thr) return 1;
  return 0;
}

// LocalMap.cpp:47-58
void orc_map_insert(orc_map* m, const double* xyz, const double* cov, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    int32_t k[3];
    voxel_index(xyz + 3 * i, m->voxel_size, k);
    voxel_add(m, pack_key(k[0], k[1], k[2]), xyz + 3 * i, cov + 9 * i);
  }
}

// LocalMap.cpp:62-69 + needsPointRemoval :149-154
uint64_t orc_map_evict(orc_map* m, const double pos[3], double distance_thr) {
  uint64_t removed = 0;
  for (auto it = m->grid.begin(); it != m->grid.end();) {
    int32_t k[3];
    unpack_key(it->first, k);
    double c[3];
    for (int d = 0; d < 3; ++d) c[d] = (static_cast<double>(k[d]) + 0.5) * m->voxel_size - pos[d];
    double dist = std::sqrt(dot3(c[0], c[0], c[1], c[1], c[2], c[2]));
    if (dist > distance_thr) {
      it = m->grid.erase(it);
      ++removed;
    } else {
      ++it;
    }
  }
  return removed;
}

// LocalMap.cpp:10-76
int orc_map_update(orc_map* m, double* xyz, double* cov, size_t n, const double T16[16],
                   int initialize, double now, uint64_t* removed) {
  if (removed) *removed = 0;
  transform_cloud(xyz, cov, n, T16);  // :15
  Iso cur = iso_from16(T16);
  double prev16[16];
  iso_to16(m->prev, prev16);
  if (!initialize && !orc_needs_map_update(prev16, T16, m->upd_trans_sq, m->upd_cos)) {
    m->prev = cur;  // :40
    return 0;
  }
  orc_map_insert(m, xyz, cov, n);
  if (m->remove_enabled && now - m->current_remove_time > m->remove_period) {  // :60
    uint64_t r = orc_map_evict(m, cur.t, m->distance_thr);
    if (removed) *removed = r;
    m->current_remove_time = now;
  }
  m->prev = cur;  // :74
  return 1;
}

uint64_t orc_map_size(const orc_map* m) { return m->grid.size(); }

void orc_map_export(const orc_map* m, int32_t* keys, uint64_t* count, double* mean, double* cov) {
  std::vector<uint64_t> ks;
  ks.reserve(m->grid.size());
  for (const auto& kv : m->grid) ks.push_back(kv.first);
  std::sort(ks.begin(), ks.end());
  for (size_t i = 0; i < ks.size(); ++i) {
    const Voxel& v = m->grid.find(ks[i])->second;
    unpack_key(ks[i], keys + 3 * i);
    count[i] = v.n;
    std::memcpy(mean + 3 * i, v.mean, sizeof v.mean);
    std::memcpy(cov + 9 * i, v.cov, sizeof v.cov);
  }
}

void orc_map_query(const orc_map* m, const double* xyz, size_t n, int32_t* keys, uint8_t* hit,
                   uint64_t* count, double* mean, double* cov) {
  for (size_t i = 0; i < n; ++i) {
    int32_t k[3];
    voxel_index(xyz + 3 * i, m->voxel_size, k);
    if (keys) std::memcpy(keys + 3 * i, k, sizeof k);
    auto f = m->grid.find(pack_key(k[0], k[1], k[2]));
    bool found = f != m->grid.end();
    if (hit) hit[i] = found;
    if (count) count[i] = found ? f->second.n : 0;
    if (mean)
      for (int j = 0; j < 3; ++j) mean[3 * i + j] = found ? f->second.mean[j] : 0.0;
    if (cov)
      for (int j = 0; j < 9; ++j) cov[9 * i + j] = found ? f->second.cov[j] : 0.0;
  }
}

void orc_knn(const double* xyz, size_t n, const double* queries, size_t nq, int k, int32_t* out_idx,
             double* out_d2) {
  KdTree tree;
  tree.init(xyz, n);
#pragma omp parallel for schedule(dynamic, 64)
  for (long long i = 0; i < static_cast<long long>(nq); ++i)
    tree.knn(queries + 3 * i, k, out_idx + static_cast<size_t>(k) * i,
             out_d2 ? out_d2 + static_cast<size_t>(k) * i : nullptr);
}

void orc_knn_bruteforce(const double* xyz, size_t n, const double* queries, size_t nq, int k,
                        int32_t* out_idx, double* out_d2) {
#pragma omp parallel for schedule(dynamic, 16)
  for (long long qi = 0; qi < static_cast<long long>(nq); ++qi) {
    const double* q = queries + 3 * qi;
    std::vector<std::pair<double, int32_t>> d(n);
    for (size_t i = 0; i < n; ++i) {
      double dx = xyz[3 * i] - q[0], dy = xyz[3 * i + 1] - q[1], dz = xyz[3 * i + 2] - q[2];
      d[i] = std::make_pair((dx * dx + dy * dy) + dz * dz, static_cast<int32_t>(i));
    }
    size_t kk = std::min(static_cast<size_t>(k), n);
    std::partial_sort(d.begin(), d.begin() + kk, d.end());
    for (int j = 0; j < k; ++j) {
      bool ok = static_cast<size_t>(j) < kk;
      out_idx[static_cast<size_t>(k) * qi + j] = ok ? d[j].second : -1;
      if (out_d2) out_d2[static_cast<size_t>(k) * qi + j] = ok ? d[j].first : -1.0;
    }
  }
}

// Open3D utility::ComputeCovariance: 9 raw cumulants, population covariance,
// Identity for an empty index list.
void orc_cov_from_indices(const double* xyz, const int32_t* idx, int k, double C[9]) {
  double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int cnt = 0;
  for (int j = 0; j < k; ++j) {
    if (idx[j] < 0) continue;
    const double* p = xyz + 3 * static_cast<size_t>(idx[j]);
    c[0] += p[0];
    c[1] += p[1];
    c[2] += p[2];
    c[3] += p[0] * p[0];
    c[4] += p[0] * p[1];
    c[5] += p[0] * p[2];
    c[6] += p[1] * p[1];
    c[7] += p[1] * p[2];
    c[8] += p[2] * p[2];
    ++cnt;
  }
  if (cnt == 0) {
    for (int i = 0; i < 9; ++i) C[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  for (int i = 0; i < 9; ++i) c[i] /= static_cast<double>(cnt);
  C[0] = c[3] - c[0] * c[0];
  C[4] = c[6] - c[1] * c[1];
  C[8] = c[8] - c[2] * c[2];
  C[1] = C[3] = c[4] - c[0] * c[1];
  C[2] = C[6] = c[5] - c[0] * c[2];
  C[5] = C[7] = c[7] - c[1] * c[2];
}

// CloudPreprocessor.cpp:120-123: U * diag(1,1,1e-2) * V^T from JacobiSVD of a
// symmetric PSD matrix.  For such a matrix with sigma_3 > 0, U == V == the
// eigenvectors (descending), so the product equals U F U^T; restated with a
// symmetric Jacobi eigen-solver (np_oracle.py cross-checks against a true SVD).
void orc_regularize_cov(const double C9[9], double out[9]) {
  double w[3], V[9];
  eig_sym3(C9, w, V);
  const double f[3] = {1.0, 1.0, 1e-2};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      out[3 * i + j] = dot3(V[3 * i] * f[0], V[3 * j], V[3 * i + 1] * f[1], V[3 * j + 1],
                            V[3 * i + 2] * f[2], V[3 * j + 2]);
}

// CloudPreprocessor.cpp:76-127
size_t orc_downsample_cov(const double* xyz, size_t n, double voxel_size, double* out_xyz,
                          double* out_cov, uint32_t* out_src) {
  KdTree tree;
  tree.init(xyz, n);  // :79-80 KD-tree over ALL raw points
  std::unordered_map<uint64_t, uint32_t, KeyHash> first;
  first.reserve(n);
  for (size_t i = 0; i < n; ++i) {  // :87-92 first point per voxel wins
    int32_t k[3];
    voxel_index(xyz + 3 * i, voxel_size, k);
    first.emplace(pack_key(k[0], k[1], k[2]), static_cast<uint32_t>(i));
  }
  std::vector<std::pair<uint64_t, uint32_t>> kept(first.begin(), first.end());
  // the reference's output order is the unordered_map's iteration order
  // (:97-99, unspecified); fixed here to ascending source index
  std::sort(kept.begin(), kept.end(),
            [](const std::pair<uint64_t, uint32_t>& a, const std::pair<uint64_t, uint32_t>& b) {
              return a.second < b.second;
            });
  const long long m = static_cast<long long>(kept.size());
  constexpr int K = 30;  // KDTreeSearchParamKNN() default
#pragma omp parallel for schedule(dynamic, 64)
  for (long long i = 0; i < m; ++i) {  // :103-124
    uint32_t src = kept[i].second;
    const double* p = xyz + 3 * static_cast<size_t>(src);
    out_xyz[3 * i] = p[0];
    out_xyz[3 * i + 1] = p[1];
    out_xyz[3 * i + 2] = p[2];
    if (out_src) out_src[i] = src;
    int32_t nn[K];
    tree.knn(p, K, nn, nullptr);
    int found = 0;
    for (int j = 0; j < K; ++j) found += nn[j] >= 0;
    double C[9];
    if (found >= 3) {
      orc_cov_from_indices(xyz, nn, K, C);
    } else {
      for (int j = 0; j < 9; ++j) C[j] = (j % 4 == 0) ? 1.0 : 0.0;
    }
    orc_regularize_cov(C, out_cov + 9 * i);
  }
  return static_cast<size_t>(m);
}

// CloudPreprocessor.cpp:25-74
int orc_deskew(double* xyz, const double* point_time, size_t n, const orc_state* states,
               size_t n_states) {
  if (n == 0 || n_states == 0) return 0;
  const double end_time = point_time[n - 1];  // :33
  long before = static_cast<long>(n_states) - 1;  // :35-42 reverse scan
  while (before >= 0 && states[before].timestamp > end_time) --before;
  if (before < 0) return -1;
  // :44 `stateBeforeLidarEnd - 1` on a reverse iterator = the NEXT state in
  // time; undefined in the reference when none exists -> clamp to `before`.
  long after = std::min(before + 1, static_cast<long>(n_states) - 1);
  double Tend[16];
  orc_interpolate_SE3(&states[before], &states[after], end_time, Tend);  // :46-48
  Iso end_inv = iso_inv(iso_from16(Tend));                                // :49
  size_t start = 0, end = 0;
  for (long s = 0; s <= after; ++s) {  // :51 cbegin .. stateAfterLidarEnd.base()
    start = end;
    size_t i = start;
    while (i < n) {  // :54-61 (runs off the end => `end` not advanced: the
      if (point_time[i] < states[s].timestamp) {  // last segment stays untouched)
        ++i;
      } else {
        end = i;
        break;
      }
    }
    if (start == end) continue;  // :63-65
    Iso st;
    orc_quat_to_matrix(states[s].attitude_xyzw, st.R);
    std::memcpy(st.t, states[s].position, sizeof st.t);
    Iso tf = iso_mul(end_inv, st);  // :70
    for (size_t j = start; j < end; ++j) iso_apply(tf, xyz + 3 * j);  // :72
  }
  return 0;
}

// CloudPreprocessor.cpp:10-23
long orc_preprocess(double* xyz, const double* point_time, size_t n, const double T_il16[16],
                    const orc_state* states, size_t n_states, double voxel_size, double* out_xyz,
                    double* out_cov, uint32_t* out_src) {
  transform_cloud(xyz, nullptr, n, T_il16);  // :16
  if (n_states > 0) {
    if (orc_deskew(xyz, point_time, n, states, n_states) != 0) return -1;  // :17-19
  }
  return static_cast<long>(orc_downsample_cov(xyz, n, voxel_size, out_xyz, out_cov, out_src));
}

}  // extern "C"
