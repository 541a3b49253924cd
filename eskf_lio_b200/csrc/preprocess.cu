// preprocess.cu — CloudPreprocessor (src/CloudPreprocessor.cpp) on the device.
//
//   process (:10-23)      T_il transform + deskew + downsample/covariances
//   deskew  (:25-74)      the per-IMU-interval rigid transforms are tiny host
//                         work (<= ~45 poses per sweep); the 64k-point apply
//                         is fused into the voxelize kernel's first phase
//   voxelDownsampleAndEstimateCovariances (:76-127)
//                         first-point-per-voxel == run heads of the stable
//                         radix sort; 30-NN covariance = K4 below
//
// K4a knn_search_kernel: one warp per kept point.  The points are sorted by the
// Morton code of their voxel, so an aligned 2^L-voxel block is ONE contiguous
// range of the sorted array.  The warp looks up the 3x3x3 blocks around the
// query at level L (27 lanes, block-range tables built by the voxelize kernel;
// binary search only above kKnnHashLevels), selects the exact 30 nearest of
// their points and stops when the 30th distance is inside the searched
// neighbourhood; otherwise L += 1.  That is an exact k-NN (same set as the
// reference's KD-tree, Open3D KDTreeFlann with KDTreeSearchParamKNN() => k =
// 30) with bounded work in sparse regions.
// Selection without a serial insertion per candidate: pass 1 streams the
// candidates once keeping the two smallest distances of every lane; the 32nd
// smallest of those 64 values bounds the 30th nearest distance from above;
// pass 2 compacts the (few) candidates under the bound into shared memory and
// a rank-count sort on (distance, index) orders them.  The serial shuffle
// insertion is kept as the overflow path.
// K4b knn_finish_kernel: one THREAD per kept point: cumulants in distance
// order (bit-exact with the oracle), Jacobi eigen-decomposition,
// U diag(1,1,0.01) U^T, outputs.
#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <limits>
#include <type_traits>

#include "internal.h"

namespace eskf {

namespace {

constexpr int kKnn = 30;  // open3d::geometry::KDTreeSearchParamKNN() default

struct KnnParams {
  const uint64_t* key[2];
  const VoxelHeader* hdr;
  const double* sx;  // positions in sorted order
  const double* sy;
  const double* sz;
  const uint32_t* sidx[2];  // sorted position -> source index
  const uint32_t* kept_src;
  const uint32_t* kept_pos;
  const double* px;  // positions in source order
  const double* py;
  const double* pz;
  unsigned n;
  double voxel;
  double* ox;
  double* oy;
  double* oz;
  double* ocov;
  size_t opitch;
  uint32_t* osrc;
  const uint32_t* orig;   // nullable: source index -> index in the uncropped sweep (range crop)
  float4* oc4;
  float2* oc2;
  uint4* levels;          // block-range tables (common.cuh); read by the search, emptied again by the finish kernel
  unsigned level_stride;
  uint32_t* nbr;          // [kKnn][nbr_pitch] neighbour source indices, ascending distance
  uint32_t* nbr_cnt;      // [n_kept]
  size_t nbr_pitch;
  unsigned* knn_next;     // work counter of the search kernel (VoxelHeader::knn_next)
  unsigned buf_limit;     // <= kBuf: survivors beyond this take the serial-insertion path
  unsigned long long* stats;  // ESKF_TRACE: [0] queries [1] level attempts [2] pass-1 candidates
                              // [3] pass-2 candidates [4] survivors [5] overflows [6] levels skipped
};

__device__ __forceinline__ unsigned lower_bound_u64(const uint64_t* a, unsigned lo, unsigned hi,
                                                    uint64_t v) {
  while (lo < hi) {
    const unsigned mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// symmetric 3x3 Jacobi eigen-decomposition (cyclic sweeps), fully unrolled so
// everything stays in registers.  w descending, V columns = eigenvectors.
__device__ __forceinline__ void eig_sym3(const double* C, double* w, double* V) {
  double a00 = C[0], a01 = 0.5 * (C[1] + C[3]), a02 = 0.5 * (C[2] + C[6]);
  double a11 = C[4], a12 = 0.5 * (C[5] + C[7]), a22 = C[8];
  double u[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int sweep = 0; sweep < 60; ++sweep) {
    const double off = a01 * a01 + a02 * a02 + a12 * a12;
    const double diag = a00 * a00 + a11 * a11 + a22 * a22;
    if (off <= 1e-34 * diag || off == 0.0) break;
    // rotation in plane (p, q): A <- J^T A J
#define ESKF_JACOBI(app, aqq, apq, akp, akq, up0, uq0, up1, uq1, up2, uq2)          \
    if (apq != 0.0) {                                                                \
      const double theta = (aqq - app) / (2.0 * apq);                                \
      const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0)); \
      const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;                        \
      app = app - tt * apq;                                                          \
      aqq = aqq + tt * apq;                                                          \
      apq = 0.0;                                                                     \
      const double kp = akp, kq = akq;                                               \
      akp = c * kp - s * kq;                                                         \
      akq = s * kp + c * kq;                                                         \
      double x0 = up0, y0 = uq0; up0 = c * x0 - s * y0; uq0 = s * x0 + c * y0;       \
      double x1 = up1, y1 = uq1; up1 = c * x1 - s * y1; uq1 = s * x1 + c * y1;       \
      double x2 = up2, y2 = uq2; up2 = c * x2 - s * y2; uq2 = s * x2 + c * y2;       \
    }
    ESKF_JACOBI(a00, a11, a01, a02, a12, u[0], u[1], u[3], u[4], u[6], u[7])  // (0,1), k = 2
    ESKF_JACOBI(a00, a22, a02, a01, a12, u[0], u[2], u[3], u[5], u[6], u[8])  // (0,2), k = 1
    ESKF_JACOBI(a11, a22, a12, a01, a02, u[1], u[2], u[4], u[5], u[7], u[8])  // (1,2), k = 0
#undef ESKF_JACOBI
  }
  // sort descending (3 elements)
  double ev[3] = {a00, a11, a22};
  int o0 = 0, o1 = 1, o2 = 2;
  if (ev[o0] < ev[o1]) { int t = o0; o0 = o1; o1 = t; }
  if (ev[o1] < ev[o2]) { int t = o1; o1 = o2; o2 = t; }
  if (ev[o0] < ev[o1]) { int t = o0; o0 = o1; o1 = t; }
  const int ord[3] = {o0, o1, o2};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    w[j] = ord[j] == 0 ? a00 : (ord[j] == 1 ? a11 : a22);
#pragma unroll
    for (int i = 0; i < 3; ++i)
      V[3 * i + j] = ord[j] == 0 ? u[3 * i] : (ord[j] == 1 ? u[3 * i + 1] : u[3 * i + 2]);
  }
}

__global__ void __launch_bounds__(128) knn_cov_kernel(KnnParams P) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  const unsigned n_kept = P.hdr->n_out;
  const unsigned sel = P.hdr->sel;
  const uint64_t* keys = P.key[sel];
  const uint32_t* sidx = P.sidx[sel];
  const int m0 = P.hdr->mn[0], m1 = P.hdr->mn[1], m2 = P.hdr->mn[2];
  const int M0 = -P.hdr->nmx[0] - m0, M1 = -P.hdr->nmx[1] - m1, M2 = -P.hdr->nmx[2] - m2;
  const unsigned n = P.n;
  const double kInf = __longlong_as_double(0x7ff0000000000000LL);
  const int K = kKnn < static_cast<int>(n) ? kKnn : static_cast<int>(n);

  for (unsigned r = warp; r < n_kept; r += nwarps) {
    const unsigned j0 = P.kept_pos[r];
    const double qx = P.sx[j0], qy = P.sy[j0], qz = P.sz[j0];
    const uint64_t mk = __ldg(keys + j0);
    const int cx = static_cast<int>(compact3(mk >> 2));
    const int cy = static_cast<int>(compact3(mk >> 1));
    const int cz = static_cast<int>(compact3(mk));
    double ld2 = kInf;   // lane l: l-th nearest so far
    int lid = -1;
    int cnt = 0;
    double kth = kInf;
    // scan the points of the per-lane ranges [start, start + len) as ONE
    // flattened candidate list (32 candidates per step, next batch prefetched)
    auto scan_ranges = [&](unsigned start, unsigned len) {
      unsigned off = len;  // inclusive scan of the lengths
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, off, o);
        if (lane >= static_cast<unsigned>(o)) off += u;
      }
      const unsigned total = __shfl_sync(0xffffffffu, off, 31);
      const unsigned excl = off - len;
      auto fetch = [&](unsigned c, double& d, int& id) {
        unsigned pos = 0;  // first range whose inclusive offset exceeds c
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
          const unsigned v = __shfl_sync(0xffffffffu, off, pos + step - 1);
          if (v <= c) pos += step;
        }
        const unsigned s_t = __shfl_sync(0xffffffffu, start, pos);
        const unsigned e_t = __shfl_sync(0xffffffffu, excl, pos);
        d = kInf;
        id = -1;
        if (c < total) {
          const unsigned j = s_t + (c - e_t);
          const double dx = P.sx[j] - qx, dy = P.sy[j] - qy, dz = P.sz[j] - qz;
          d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
          id = static_cast<int>(__ldg(sidx + j));
        }
      };
      double d0, d1;
      int id0, id1;
      fetch(lane, d0, id0);
      for (unsigned c = 0; c < total; c += 32) {
        d1 = kInf;
        id1 = -1;
        if (c + 32 < total) fetch(c + 32 + lane, d1, id1);
        unsigned bal = __ballot_sync(0xffffffffu, id0 >= 0 && d0 <= kth);
        while (bal) {
          const int src = __ffs(bal) - 1;
          bal &= bal - 1;
          const double cd = __shfl_sync(0xffffffffu, d0, src);
          const int cid = __shfl_sync(0xffffffffu, id0, src);
          // (d2, index) lexicographic order, like the oracle's heap
          const int pos = __popc(__ballot_sync(0xffffffffu, ld2 < cd || (ld2 == cd && lid < cid)));
          if (pos < K) {
            const double up_d = __shfl_up_sync(0xffffffffu, ld2, 1);
            const int up_i = __shfl_up_sync(0xffffffffu, lid, 1);
            if (static_cast<int>(lane) > pos) {
              ld2 = up_d;
              lid = up_i;
            } else if (static_cast<int>(lane) == pos) {
              ld2 = cd;
              lid = cid;
            }
            if (static_cast<int>(lane) >= K) {
              ld2 = kInf;
              lid = -1;
            }
            if (cnt < K) ++cnt;
            kth = __shfl_sync(0xffffffffu, ld2, K - 1);
          }
        }
        d0 = d1;
        id0 = id1;
      }
    };
    for (int L = 0; L <= kKeyBits; ++L) {
      const int bx = cx >> L, by = cy >> L, bz = cz >> L;
      const double span = static_cast<double>(1 << L);
      unsigned start = 0, end = 0;
      double box2 = kInf;  // squared distance from q to this lane's block (AABB)
      if (lane < 27) {
        const int nx = bx + static_cast<int>(lane % 3) - 1;
        const int ny = by + static_cast<int>((lane / 3) % 3) - 1;
        const int nz = bz + static_cast<int>(lane / 9) - 1;
        if (nx >= 0 && ny >= 0 && nz >= 0 && nx <= (M0 >> L) && ny <= (M1 >> L) && nz <= (M2 >> L)) {
          const uint64_t lo = morton3(nx, ny, nz) << (3 * L);
          const uint64_t hi = lo + (1ull << (3 * L));
          start = lower_bound_u64(keys, 0, n, lo);
          end = lower_bound_u64(keys, start, n, hi);
          const double x0 = (static_cast<double>(m0) + static_cast<double>(nx) * span) * P.voxel;
          const double y0 = (static_cast<double>(m1) + static_cast<double>(ny) * span) * P.voxel;
          const double z0 = (static_cast<double>(m2) + static_cast<double>(nz) * span) * P.voxel;
          const double w = span * P.voxel;
          const double ex = fmax(fmax(x0 - qx, qx - (x0 + w)) - 1e-9, 0.0);
          const double ey = fmax(fmax(y0 - qy, qy - (y0 + w)) - 1e-9, 0.0);
          const double ez = fmax(fmax(z0 - qz, qz - (z0 + w)) - 1e-9, 0.0);
          box2 = ex * ex + ey * ey + ez * ez;
        }
      }
      const unsigned len = end - start;
      const unsigned total = warp_reduce_add(len);
      const bool all = (M0 >> L) == 0 && (M1 >> L) == 0 && (M2 >> L) == 0;
      // fewer than K candidates cannot finish the search at this level
      if (total < static_cast<unsigned>(K) && !all) continue;
      ld2 = kInf;
      lid = -1;
      cnt = 0;
      kth = kInf;
      // the query's own block first (lane 13), then only the blocks that can
      // still hold a point closer than the current K-th distance
      scan_ranges(start, lane == 13 ? len : 0u);
      scan_ranges(start, (lane != 13 && box2 <= kth) ? len : 0u);
      // the 3x3x3 neighbourhood already holds every point?
      if (all) break;
      if (cnt == K) {
        const double ax = (static_cast<double>(m0) + static_cast<double>(bx) * span);
        const double ay = (static_cast<double>(m1) + static_cast<double>(by) * span);
        const double az = (static_cast<double>(m2) + static_cast<double>(bz) * span);
        double bound = fmin(qx - (ax - span) * P.voxel, (ax + 2.0 * span) * P.voxel - qx);
        bound = fmin(bound, fmin(qy - (ay - span) * P.voxel, (ay + 2.0 * span) * P.voxel - qy));
        bound = fmin(bound, fmin(qz - (az - span) * P.voxel, (az + 2.0 * span) * P.voxel - qz));
        bound -= 1e-9;
        if (bound > 0.0 && kth <= bound * bound) break;
      }
    }
    // Open3D utility::ComputeCovariance over the neighbours in ascending
    // distance order, exact ops (src/CloudPreprocessor.cpp:111-118)
    double nxp = 0.0, nyp = 0.0, nzp = 0.0;
    if (lid >= 0) {
      nxp = P.px[lid];
      nyp = P.py[lid];
      nzp = P.pz[lid];
    }
    double C[9];
    if (cnt >= 3) {
      double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int l = 0; l < cnt; ++l) {
        const double x = __shfl_sync(0xffffffffu, nxp, l);
        const double y = __shfl_sync(0xffffffffu, nyp, l);
        const double z = __shfl_sync(0xffffffffu, nzp, l);
        c[0] = __dadd_rn(c[0], x);
        c[1] = __dadd_rn(c[1], y);
        c[2] = __dadd_rn(c[2], z);
        c[3] = __dadd_rn(c[3], __dmul_rn(x, x));
        c[4] = __dadd_rn(c[4], __dmul_rn(x, y));
        c[5] = __dadd_rn(c[5], __dmul_rn(x, z));
        c[6] = __dadd_rn(c[6], __dmul_rn(y, y));
        c[7] = __dadd_rn(c[7], __dmul_rn(y, z));
        c[8] = __dadd_rn(c[8], __dmul_rn(z, z));
      }
      const double nn = static_cast<double>(cnt);
#pragma unroll
      for (int k = 0; k < 9; ++k) c[k] = __ddiv_rn(c[k], nn);
      C[0] = __dadd_rn(c[3], -__dmul_rn(c[0], c[0]));
      C[4] = __dadd_rn(c[6], -__dmul_rn(c[1], c[1]));
      C[8] = __dadd_rn(c[8], -__dmul_rn(c[2], c[2]));
      C[1] = C[3] = __dadd_rn(c[4], -__dmul_rn(c[0], c[1]));
      C[2] = C[6] = __dadd_rn(c[5], -__dmul_rn(c[0], c[2]));
      C[5] = C[7] = __dadd_rn(c[7], -__dmul_rn(c[1], c[2]));
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) C[k] = (k % 4 == 0) ? 1.0 : 0.0;
    }
    // regularise: U diag(1,1,1e-2) V^T (src/CloudPreprocessor.cpp:120-123);
    // U == V for a symmetric PSD matrix
    double wv[3], V[9], R[9];
    eig_sym3(C, wv, V);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        R[3 * i + j] = (V[3 * i] * V[3 * j] + V[3 * i + 1] * V[3 * j + 1]) + 1e-2 * (V[3 * i + 2] * V[3 * j + 2]);
    if (lane == 0) {
      P.ox[r] = qx;
      P.oy[r] = qy;
      P.oz[r] = qz;
#pragma unroll
      for (int k = 0; k < 9; ++k) P.ocov[k * P.opitch + r] = R[k];
      P.osrc[r] = P.orig != nullptr ? P.orig[P.kept_src[r]] : P.kept_src[r];
      P.oc4[r] = make_float4(static_cast<float>(R[0]), static_cast<float>(R[1]),
                             static_cast<float>(R[2]), static_cast<float>(R[4]));
      P.oc2[r] = make_float2(static_cast<float>(R[5]), static_cast<float>(R[8]));
    }
  }
}

// ===================================================================== K4a
constexpr int kSearchThreads = 128;
constexpr int kSearchWarps = kSearchThreads / 32;
constexpr int kBuf = 128;  // survivors of pass 2 per warp (typically 35..70)

// the 27 neighbour ranges of a query, one per lane (lanes 27..31 hold len 0)
struct Ranges {
  unsigned start;
  unsigned len;
  double box2;  // squared distance from the query to the block's AABB
};

// Enumerates the points of the lanes' ranges [start, start + len) as ONE
// flattened candidate list, 32 candidates a step: candidate c belongs to the
// first range whose inclusive length prefix exceeds c.  f(valid, d2, id, j).
template <typename F>
__device__ __forceinline__ void for_each_candidate(const KnnParams& P, const uint32_t* sidx,
                                                   unsigned start, unsigned len, double qx, double qy,
                                                   double qz, unsigned lane, F&& f) {
  unsigned off = len;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned u = __shfl_up_sync(0xffffffffu, off, o);
    if (lane >= static_cast<unsigned>(o)) off += u;
  }
  const unsigned total = __shfl_sync(0xffffffffu, off, 31);
  const unsigned excl = off - len;
  const double kInf = __longlong_as_double(0x7ff0000000000000LL);
  for (unsigned c0 = 0; c0 < total; c0 += 32) {
    const unsigned c = c0 + lane;
    unsigned pos = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
      const unsigned v = __shfl_sync(0xffffffffu, off, pos + step - 1);
      if (v <= c) pos += step;
    }
    const unsigned s_t = __shfl_sync(0xffffffffu, start, pos & 31u);
    const unsigned e_t = __shfl_sync(0xffffffffu, excl, pos & 31u);
    const bool valid = c < total;
    double d = kInf;
    int id = -1;
    if (valid) {
      const unsigned j = s_t + (c - e_t);
      const double dx = P.sx[j] - qx, dy = P.sy[j] - qy, dz = P.sz[j] - qz;
      d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      id = static_cast<int>(__ldg(sidx + j));
    }
    f(valid, d, id);
  }
}

// ascending bitonic sort of one float per lane
__device__ __forceinline__ float warp_sort_f32(float v, unsigned lane) {
#pragma unroll
  for (unsigned k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (unsigned j = k >> 1; j > 0; j >>= 1) {
      const float o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool up = (lane & k) == 0u || k == 32u;
      const bool lower = (lane & j) == 0u;
      v = (lower == up) ? fminf(v, o) : fmaxf(v, o);
    }
  }
  return v;
}

// Upper bound of the K-th (K <= 32) smallest candidate distance from the two smallest
// distances every lane has seen: the K-th smallest of the 64 values.  The minima are
// tracked as floats rounded UP from the fp64 distances (an upper bound is all that is
// needed; the selection itself stays exact), which makes the two sorts a third as long.
__device__ __forceinline__ float bound_from_minima(float m1, float m2, unsigned lane, int K) {
  const float a = warp_sort_f32(m1, lane);
  const float b = warp_sort_f32(m2, lane);
  float c = fminf(a, __shfl_sync(0xffffffffu, b, 31u - lane));  // the 32 smallest of the union, bitonic
#pragma unroll
  for (unsigned j = 16; j > 0; j >>= 1) {  // bitonic merge -> ascending
    const float o = __shfl_xor_sync(0xffffffffu, c, j);
    c = (lane & j) == 0u ? fminf(c, o) : fmaxf(c, o);
  }
  return __shfl_sync(0xffffffffu, c, K - 1);
}

// Overflow path (and the reference behaviour the selection above must equal):
// exact top-K of the given ranges by shuffle insertion, one entry per lane,
// (d2, index) lexicographic order like the oracle's heap.
struct SerialTop {
  double ld2;
  int lid;
  int cnt;
  double kth;
};

__device__ __forceinline__ void serial_scan(const KnnParams& P, const uint32_t* sidx, unsigned start,
                                            unsigned len, double qx, double qy, double qz,
                                            unsigned lane, int K, SerialTop& T) {
  const double kInf = __longlong_as_double(0x7ff0000000000000LL);
  for_each_candidate(P, sidx, start, len, qx, qy, qz, lane, [&](bool valid, double d0, int id0) {
    unsigned bal = __ballot_sync(0xffffffffu, valid && d0 <= T.kth);
    while (bal) {
      const int src = __ffs(bal) - 1;
      bal &= bal - 1;
      const double cd = __shfl_sync(0xffffffffu, d0, src);
      const int cid = __shfl_sync(0xffffffffu, id0, src);
      const int pos = __popc(__ballot_sync(0xffffffffu, T.ld2 < cd || (T.ld2 == cd && T.lid < cid)));
      if (pos < K) {
        const double up_d = __shfl_up_sync(0xffffffffu, T.ld2, 1);
        const int up_i = __shfl_up_sync(0xffffffffu, T.lid, 1);
        if (static_cast<int>(lane) > pos) {
          T.ld2 = up_d;
          T.lid = up_i;
        } else if (static_cast<int>(lane) == pos) {
          T.ld2 = cd;
          T.lid = cid;
        }
        if (static_cast<int>(lane) >= K) {
          T.ld2 = kInf;
          T.lid = -1;
        }
        if (T.cnt < K) ++T.cnt;
        T.kth = __shfl_sync(0xffffffffu, T.ld2, K - 1);
      }
    }
  });
}

__global__ void __launch_bounds__(kSearchThreads) knn_search_kernel(KnnParams P) {
  __shared__ double s_d[kSearchWarps][kBuf];
  __shared__ int s_i[kSearchWarps][kBuf];
  __shared__ unsigned s_mask[kKnnHashLevels];
  const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (threadIdx.x < kKnnHashLevels)
    s_mask[threadIdx.x] = knn_level_slots(P.hdr->n_out, P.hdr->bits, static_cast<int>(threadIdx.x)) - 1u;
  __syncthreads();
  const unsigned n_kept = P.hdr->n_out;
  const unsigned sel = P.hdr->sel;
  const uint64_t* keys = P.key[sel];
  const uint32_t* sidx = P.sidx[sel];
  const int m0 = P.hdr->mn[0], m1 = P.hdr->mn[1], m2 = P.hdr->mn[2];
  const int M0 = -P.hdr->nmx[0] - m0, M1 = -P.hdr->nmx[1] - m1, M2 = -P.hdr->nmx[2] - m2;
  const unsigned n = P.n;
  const double kInf = __longlong_as_double(0x7ff0000000000000LL);
  const int K = kKnn < static_cast<int>(n) ? kKnn : static_cast<int>(n);
  double* bd = s_d[wib];
  int* bi = s_i[wib];

  // queries differ a lot in cost (dense near field vs sparse far field): the
  // resident warps pull them from a counter instead of striding
  for (;;) {
    unsigned r = 0;
    if (lane == 0) r = atomicAdd(P.knn_next, 1u);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= n_kept) break;
    const unsigned j0 = P.kept_pos[r];
    const double qx = P.sx[j0], qy = P.sy[j0], qz = P.sz[j0];
    const uint64_t mk = __ldg(keys + j0);
    const int cx = static_cast<int>(compact3(mk >> 2));
    const int cy = static_cast<int>(compact3(mk >> 1));
    const int cz = static_cast<int>(compact3(mk));
    int my_id = -1;  // lane l: source index of the l-th nearest neighbour
    int cnt = 0;
    for (int L = 0; L <= kKeyBits; ++L) {
      const int bx = cx >> L, by = cy >> L, bz = cz >> L;
      const double span = static_cast<double>(1 << L);
      unsigned start = 0, end = 0;
      double box2 = kInf;
      // Morton spreading once per neighbour coordinate (lanes 0..8: x-1 x x+1 y-1 ... z+1)
      unsigned long long spr = 0;
      {
        const int base = lane < 3 ? bx : (lane < 6 ? by : bz);
        const int cv = base + static_cast<int>(lane % 3) - 1;
        if (lane < 9 && cv >= 0) spr = spread3(static_cast<uint32_t>(cv));
      }
      const unsigned dxl = lane % 3, dyl = (lane / 3) % 3, dzl = (lane / 9) % 3;
      const unsigned long long sprx = __shfl_sync(0xffffffffu, spr, dxl);
      const unsigned long long spry = __shfl_sync(0xffffffffu, spr, 3u + dyl);
      const unsigned long long sprz = __shfl_sync(0xffffffffu, spr, 6u + dzl);
      if (lane < 27) {
        const int nx = bx + static_cast<int>(dxl) - 1;
        const int ny = by + static_cast<int>(dyl) - 1;
        const int nz = bz + static_cast<int>(dzl) - 1;
        if (nx >= 0 && ny >= 0 && nz >= 0 && nx <= (M0 >> L) && ny <= (M1 >> L) && nz <= (M2 >> L)) {
          const uint64_t bk = (sprx << 2) | (spry << 1) | sprz;  // morton3(nx, ny, nz)
          if (L < kKnnHashLevels) {
            if (!knn_level_find(P.levels + static_cast<size_t>(L) * P.level_stride, s_mask[L], bk, start, end))
              start = end = 0;
          } else {
            const uint64_t lo = bk << (3 * L);
            start = lower_bound_u64(keys, 0, n, lo);
            end = lower_bound_u64(keys, start, n, lo + (1ull << (3 * L)));
          }
        }
      }
      const unsigned len = end - start;
      const unsigned total = warp_reduce_add(len);
      const bool all = (M0 >> L) == 0 && (M1 >> L) == 0 && (M2 >> L) == 0;
      // Too few candidates to finish here.  (K is the hard limit; below ~2K
      // the K-th neighbour almost never lies inside the inscribed ball, and
      // starting one level up is cheaper than a failed attempt.)
      if (total < static_cast<unsigned>(2 * K) && !all) {
        if (P.stats && lane == 0) atomicAdd(P.stats + 6, 1ull);
        continue;
      }
      if (len != 0u) {  // squared distance from the query to this lane's block (only for levels that are searched)
        const int nx = bx + static_cast<int>(dxl) - 1;
        const int ny = by + static_cast<int>(dyl) - 1;
        const int nz = bz + static_cast<int>(dzl) - 1;
        const double x0 = (static_cast<double>(m0) + static_cast<double>(nx) * span) * P.voxel;
        const double y0 = (static_cast<double>(m1) + static_cast<double>(ny) * span) * P.voxel;
        const double z0 = (static_cast<double>(m2) + static_cast<double>(nz) * span) * P.voxel;
        const double w = span * P.voxel;
        const double ex = fmax(fmax(x0 - qx, qx - (x0 + w)) - 1e-9, 0.0);
        const double ey = fmax(fmax(y0 - qy, qy - (y0 + w)) - 1e-9, 0.0);
        const double ez = fmax(fmax(z0 - qz, qz - (z0 + w)) - 1e-9, 0.0);
        box2 = ex * ex + ey * ey + ez * ez;
      }

      // ---- pass 1: two smallest distances per lane, own block first
      const float kInfF = __int_as_float(0x7f800000);
      float mn1 = kInfF, mn2 = kInfF;  // rounded up: (double)mn >= the exact distance
      auto track = [&](bool valid, double d, int) {
        if (valid) {
          const float df = __double2float_ru(d);
          if (df < mn1) {
            mn2 = mn1;
            mn1 = df;
          } else if (df < mn2) {
            mn2 = df;
          }
        }
      };
      const unsigned own_len = __shfl_sync(0xffffffffu, len, 13);
      for_each_candidate(P, sidx, start, lane == 13 ? len : 0u, qx, qy, qz, lane, track);
      double b0 = kInf;  // 32 distinct candidates are <= max(mn1) once every lane has one
      if (own_len >= 32u) {
        float m = mn1;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        b0 = static_cast<double>(m);
      }
      const unsigned len1b = (lane != 13 && box2 <= b0) ? len : 0u;
      for_each_candidate(P, sidx, start, len1b, qx, qy, qz, lane, track);
      // with no more than 2 x 32 candidates everything fits the rank sort below: no bound needed
      const double b1 = own_len + warp_reduce_add(len1b) > 64u
                            ? static_cast<double>(bound_from_minima(mn1, mn2, lane, K)) : kInf;

      // ---- pass 2: compact the candidates under the bound
      unsigned S = 0;
      for_each_candidate(P, sidx, start, (lane == 13 || box2 <= b1) ? len : 0u, qx, qy, qz, lane,
                         [&](bool valid, double d, int id) {
                           const bool keep = valid && d <= b1;
                           const unsigned bal = __ballot_sync(0xffffffffu, keep);
                           const unsigned at = S + __popc(bal & ((1u << lane) - 1u));
                           if (keep && at < static_cast<unsigned>(kBuf)) {
                             bd[at] = d;
                             bi[at] = id;
                           }
                           S += __popc(bal);
                         });
      __syncwarp();
      if (P.stats) {
        const unsigned n2 = warp_reduce_add((lane == 13 || box2 <= b1) ? len : 0u);
        const unsigned n1 = own_len + warp_reduce_add(len1b);
        if (lane == 0) {
          atomicAdd(P.stats + 1, 1ull);
          atomicAdd(P.stats + 2, static_cast<unsigned long long>(n1));
          atomicAdd(P.stats + 3, static_cast<unsigned long long>(n2));
          atomicAdd(P.stats + 4, static_cast<unsigned long long>(S));
          if (S > P.buf_limit) atomicAdd(P.stats + 5, 1ull);
        }
      }
      double kth = kInf;
      if (S <= P.buf_limit) {
        // rank-count sort: entry e goes to position #{o : (d_o, id_o) < (d_e, id_e)}
        double ed[kBuf / 32];
        int ei[kBuf / 32], rk[kBuf / 32];
#pragma unroll
        for (int t = 0; t < kBuf / 32; ++t) {
          const unsigned e = lane + 32u * t;
          ed[t] = e < S ? bd[e] : kInf;
          ei[t] = e < S ? bi[e] : 0x7fffffff;
          rk[t] = 0;
        }
        auto rank_pass = [&](auto nt_tag) {
          constexpr int NT = decltype(nt_tag)::value;
#pragma unroll 4
          for (unsigned o = 0; o < S; ++o) {
            const double od = bd[o];
            const int oi = bi[o];
#pragma unroll
            for (int t = 0; t < NT; ++t) rk[t] += (od < ed[t] || (od == ed[t] && oi < ei[t])) ? 1 : 0;
          }
        };
        if (S <= 32u) rank_pass(std::integral_constant<int, 1>());
        else if (S <= 64u) rank_pass(std::integral_constant<int, 2>());
        else rank_pass(std::integral_constant<int, kBuf / 32>());
        __syncwarp();
        cnt = static_cast<int>(S) < K ? static_cast<int>(S) : K;
        // hand entry of rank l to lane l through the (now free) buffer
#pragma unroll
        for (int t = 0; t < kBuf / 32; ++t) {
          const unsigned e = lane + 32u * t;
          if (e < S && rk[t] < K) {
            bd[rk[t]] = ed[t];
            bi[rk[t]] = ei[t];
          }
        }
        __syncwarp();
        my_id = static_cast<int>(lane) < cnt ? bi[lane] : -1;
        if (cnt == K) kth = bd[K - 1];
        __syncwarp();
      } else {
        SerialTop T;
        T.ld2 = kInf;
        T.lid = -1;
        T.cnt = 0;
        T.kth = kInf;
        serial_scan(P, sidx, start, lane == 13 ? len : 0u, qx, qy, qz, lane, K, T);
        serial_scan(P, sidx, start, (lane != 13 && box2 <= T.kth) ? len : 0u, qx, qy, qz, lane, K, T);
        my_id = T.lid;
        cnt = T.cnt;
        kth = T.cnt == K ? T.kth : kInf;
      }
      // the 3x3x3 neighbourhood already holds every point?
      if (all) break;
      if (cnt == K) {
        const double ax = (static_cast<double>(m0) + static_cast<double>(bx) * span);
        const double ay = (static_cast<double>(m1) + static_cast<double>(by) * span);
        const double az = (static_cast<double>(m2) + static_cast<double>(bz) * span);
        double bound = fmin(qx - (ax - span) * P.voxel, (ax + 2.0 * span) * P.voxel - qx);
        bound = fmin(bound, fmin(qy - (ay - span) * P.voxel, (ay + 2.0 * span) * P.voxel - qy));
        bound = fmin(bound, fmin(qz - (az - span) * P.voxel, (az + 2.0 * span) * P.voxel - qz));
        bound -= 1e-9;
        if (bound > 0.0 && kth <= bound * bound) break;
      }
    }
    if (static_cast<int>(lane) < cnt) P.nbr[static_cast<size_t>(lane) * P.nbr_pitch + r] = static_cast<uint32_t>(my_id);
    if (lane == 0) {
      P.nbr_cnt[r] = static_cast<uint32_t>(cnt);
      if (P.stats) atomicAdd(P.stats, 1ull);
    }
  }
}

// ===================================================================== K4b
__global__ void __launch_bounds__(64) knn_finish_kernel(KnnParams P) {
  const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
  {
    // the search is done with the block-range tables: empty the entries this sweep could have used, so
    // that the next sweep's voxelize kernel need not (11 MB of stores for a 64k-point sweep, spread
    // here over n threads of which two thirds have nothing else to do)
    const unsigned bits = P.hdr->bits, nthreads = gridDim.x * blockDim.x;
    const uint4 empty = make_uint4(~0u, ~0u, ~0u, ~0u);
    for (int L = 0; L < kKnnHashLevels; ++L) {
      const unsigned slots = knn_level_slots(P.hdr->n_out, bits, L);
      uint4* tab = P.levels + static_cast<size_t>(L) * P.level_stride;
      for (unsigned k = r; k < slots; k += nthreads) tab[k] = empty;
    }
  }
  if (r >= P.hdr->n_out) return;
  const unsigned j0 = P.kept_pos[r];
  const int cnt = static_cast<int>(P.nbr_cnt[r]);
  // Open3D utility::ComputeCovariance over the neighbours in ascending
  // distance order, exact ops (src/CloudPreprocessor.cpp:111-118)
  double C[9];
  if (cnt >= 3) {
    double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int l = 0; l < cnt; ++l) {
      const uint32_t id = P.nbr[static_cast<size_t>(l) * P.nbr_pitch + r];
      const double x = P.px[id], y = P.py[id], z = P.pz[id];
      c[0] = __dadd_rn(c[0], x);
      c[1] = __dadd_rn(c[1], y);
      c[2] = __dadd_rn(c[2], z);
      c[3] = __dadd_rn(c[3], __dmul_rn(x, x));
      c[4] = __dadd_rn(c[4], __dmul_rn(x, y));
      c[5] = __dadd_rn(c[5], __dmul_rn(x, z));
      c[6] = __dadd_rn(c[6], __dmul_rn(y, y));
      c[7] = __dadd_rn(c[7], __dmul_rn(y, z));
      c[8] = __dadd_rn(c[8], __dmul_rn(z, z));
    }
    const double nn = static_cast<double>(cnt);
#pragma unroll
    for (int k = 0; k < 9; ++k) c[k] = __ddiv_rn(c[k], nn);
    C[0] = __dadd_rn(c[3], -__dmul_rn(c[0], c[0]));
    C[4] = __dadd_rn(c[6], -__dmul_rn(c[1], c[1]));
    C[8] = __dadd_rn(c[8], -__dmul_rn(c[2], c[2]));
    C[1] = C[3] = __dadd_rn(c[4], -__dmul_rn(c[0], c[1]));
    C[2] = C[6] = __dadd_rn(c[5], -__dmul_rn(c[0], c[2]));
    C[5] = C[7] = __dadd_rn(c[7], -__dmul_rn(c[1], c[2]));
  } else {
#pragma unroll
    for (int k = 0; k < 9; ++k) C[k] = (k % 4 == 0) ? 1.0 : 0.0;
  }
  // regularise: U diag(1,1,1e-2) V^T (src/CloudPreprocessor.cpp:120-123);
  // U == V for a symmetric PSD matrix
  double wv[3], V[9], R[9];
  eig_sym3(C, wv, V);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      R[3 * i + j] = (V[3 * i] * V[3 * j] + V[3 * i + 1] * V[3 * j + 1]) + 1e-2 * (V[3 * i + 2] * V[3 * j + 2]);
  P.ox[r] = P.sx[j0];
  P.oy[r] = P.sy[j0];
  P.oz[r] = P.sz[j0];
#pragma unroll
  for (int k = 0; k < 9; ++k) P.ocov[k * P.opitch + r] = R[k];
  P.osrc[r] = P.orig != nullptr ? P.orig[P.kept_src[r]] : P.kept_src[r];
  P.oc4[r] = make_float4(static_cast<float>(R[0]), static_cast<float>(R[1]), static_cast<float>(R[2]),
                         static_cast<float>(R[4]));
  P.oc2[r] = make_float2(static_cast<float>(R[5]), static_cast<float>(R[8]));
}

// ------------------------------------------------------------ host helpers
struct HIso {
  double R[9];
  double t[3];
};

inline double hdot3(double a0, double b0, double a1, double b1, double a2, double b2) {
  return (a0 * b0 + a1 * b1) + a2 * b2;
}

void quat_to_matrix(const double* q, double* R) {  // Eigen Quaterniond::toRotationMatrix
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}

HIso iso_mul(const HIso& a, const HIso& b) {
  HIso r;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      r.R[3 * i + j] = hdot3(a.R[3 * i], b.R[j], a.R[3 * i + 1], b.R[3 + j], a.R[3 * i + 2], b.R[6 + j]);
    r.t[i] = hdot3(a.R[3 * i], b.t[0], a.R[3 * i + 1], b.t[1], a.R[3 * i + 2], b.t[2]) + a.t[i];
  }
  return r;
}

HIso iso_inv(const HIso& a) {
  HIso r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.R[3 * i + j] = a.R[3 * j + i];
  for (int i = 0; i < 3; ++i)
    r.t[i] = -hdot3(r.R[3 * i], a.t[0], r.R[3 * i + 1], a.t[1], r.R[3 * i + 2], a.t[2]);
  return r;
}

// Utils::interpolateSE3 (src/Utils.cpp:65-75) with Eigen's slerp
HIso interpolate(const eskf_state& s1, const eskf_state& s2, double t) {
  const double f = (t - s1.timestamp) / (s2.timestamp - s1.timestamp + 1e-6);
  const double* a = s1.attitude_xyzw;
  const double* b = s2.attitude_xyzw;
  const double d = (a[0] * b[0] + a[1] * b[1]) + (a[2] * b[2] + a[3] * b[3]);
  const double ad = std::fabs(d);
  double s0, s1c;
  if (ad >= 1.0 - std::numeric_limits<double>::epsilon()) {
    s0 = 1.0 - f;
    s1c = f;
  } else {
    const double th = std::acos(ad), sn = std::sin(th);
    s0 = std::sin((1.0 - f) * th) / sn;
    s1c = std::sin(f * th) / sn;
  }
  if (d < 0.0) s1c = -s1c;
  double q[4];
  for (int i = 0; i < 4; ++i) q[i] = s0 * a[i] + s1c * b[i];
  HIso r;
  quat_to_matrix(q, r.R);
  for (int i = 0; i < 3; ++i) r.t[i] = s1.position[i] + f * (s2.position[i] - s1.position[i]);
  return r;
}

}  // namespace

// CloudPreprocessor::deskew (src/CloudPreprocessor.cpp:25-74) reduced to its
// segment table: [begin, end) point ranges and the rigid transform of each.
// Points outside every segment stay untouched — including the reference's
// last segment, whose inner scan runs off the end (:54-65).
int compute_deskew_segments(const double* point_time, size_t n, const eskf_state* states,
                            size_t n_states, std::vector<DeskewSeg>* out, int sorted_hint) {
  out->clear();
  if (n == 0 || n_states == 0) return ESKF_OK;
  const double end_time = point_time[n - 1];
  long before = static_cast<long>(n_states) - 1;
  while (before >= 0 && states[before].timestamp > end_time) --before;
  if (before < 0) {
    set_error("deskew: no state at or before the scan end time");
    return ESKF_ERR_INVALID;
  }
  const long after = before + 1 < static_cast<long>(n_states) ? before + 1 : before;
  const HIso end_inv = iso_inv(interpolate(states[before], states[after], end_time));
  // The reference finds a segment's end with a forward linear scan: the FIRST point at or after
  // `start` whose stamp is not below the state's (:54-61).  With non-decreasing stamps (the
  // reference's own precondition, :33) that is a binary search; ring-major or merged sweeps whose
  // stamps are not sorted take the reference's scan itself, so the segments stay identical.
  // (the check is one pass over the stamps, ~20 us for a 64k-point sweep: a caller that knows, e.g.
  // from eskf_stamps_sorted() run when the sweep arrived, says so with the option "stamps_sorted")
  const bool sorted = sorted_hint < 0 ? eskf_stamps_sorted(point_time, n) != 0 : sorted_hint != 0;
  size_t start = 0, end = 0;
  for (long s = 0; s <= after; ++s) {
    start = end;
    // first point at or after `start` whose time is not < the state's stamp
    const double ts = states[s].timestamp;
    size_t lo = start, hi = n;
    if (start < n && !(point_time[start] < ts)) {
      hi = start;  // common case for old states: nothing to consume
    } else if (!sorted) {
      hi = start;
      while (hi < n && point_time[hi] < ts) ++hi;
      if (hi >= n) continue;  // ran off the end: `end` is not advanced, the next state rescans (:54-65)
    } else {
      while (lo < hi) {
        const size_t mid = (lo + hi) / 2;
        if (point_time[mid] < ts) lo = mid + 1; else hi = mid;
      }
      hi = lo;
    }
    if (hi >= n) continue;  // ran off the end: `end` is not advanced (with time-ordered states no later one finds a point either)
    end = hi;
    if (start == end) continue;
    HIso st;
    quat_to_matrix(states[s].attitude_xyzw, st.R);
    for (int i = 0; i < 3; ++i) st.t[i] = states[s].position[i];
    const HIso tf = iso_mul(end_inv, st);
    DeskewSeg seg;
    seg.begin = static_cast<uint32_t>(start);
    seg.end = static_cast<uint32_t>(end);
    for (int i = 0; i < 9; ++i) seg.T[i] = tf.R[i];
    for (int i = 0; i < 3; ++i) seg.T[9 + i] = tf.t[i];
    out->push_back(seg);
  }
  return ESKF_OK;
}

namespace {

int check_header(eskf_ctx* ctx, unsigned* n_out) {
  VoxelHeader* h = nullptr;
  ESKF_TRY(ctx_pinned(ctx, sizeof(VoxelHeader), reinterpret_cast<void**>(&h)));
  ESKF_CUDA(cudaMemcpyAsync(h, ctx->hdr.p, sizeof(VoxelHeader), cudaMemcpyDeviceToHost, ctx->stream));
  ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h->gb.error) {
    set_error("grid barrier timeout in voxelize kernel");
    return ESKF_ERR_INTERNAL;
  }
  if (h->error & 1u) {
    set_error("voxel coordinate outside the 21-bit key range");
    return ESKF_ERR_RANGE;
  }
  *n_out = h->n_out;
  return ESKF_OK;
}

// raw (device, xyz only) -> out (device, xyz + cov + src index).  Launch only; preprocess_finish waits
// for the kept-point count.  keep_input: the transformed positions go to a scratch buffer and raw stays
// as delivered (so that the call can be repeated with other deskew segments).
int preprocess_launch(eskf_ctx* ctx, eskf_cloud* raw, const double* T_il,
                      const std::vector<DeskewSeg>& segs, double voxel, eskf_cloud* out,
                      const uint32_t* orig, bool keep_input, bool* mapped_out) {
  const unsigned n = static_cast<unsigned>(raw->n);
  ESKF_TRY(cloud_reserve(out, raw->n, true));
  double* tx = raw->x();
  double* ty = raw->y();
  double* tz = raw->z();
  if (keep_input) {
    const size_t pitch = (static_cast<size_t>(n) + 63) / 64 * 64;
    ESKF_TRY(ctx->xform.ensure(3 * pitch * sizeof(double)));
    tx = ctx->xform.as<double>();
    ty = tx + pitch;
    tz = ty + pitch;
  }
  VoxelizeArgs a;
  std::memset(&a, 0, sizeof a);
  a.in_x = raw->x();
  a.in_y = raw->y();
  a.in_z = raw->z();
  a.in_stride = 1;
  a.out_x = tx;
  a.out_y = ty;
  a.out_z = tz;
  a.cov = nullptr;
  a.n = n;
  a.voxel = voxel;
  a.has_T1 = T_il != nullptr;
  if (T_il)
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) a.T1[3 * i + j] = T_il[4 * i + j];
      a.T1[9 + i] = T_il[4 * i + 3];
    }
  if (!segs.empty()) {
    const size_t bytes = segs.size() * sizeof(DeskewSeg);
    ESKF_TRY(ctx->segs.ensure(bytes));
    void* hp = nullptr;
    ESKF_TRY(ctx_pinned(ctx, bytes, &hp));
    std::memcpy(hp, segs.data(), bytes);
    ESKF_CUDA(cudaMemcpyAsync(ctx->segs.p, hp, bytes, cudaMemcpyHostToDevice, ctx->stream));
    a.segs = ctx->segs.as<DeskewSeg>();
    a.n_segs = static_cast<int>(segs.size());
  }
  a.orig = orig;
  a.mode = 1;
  const bool mapped = ctx->opt_mapped_results && ctx->mail_h != nullptr && !ctx->opt_trace;
  if (mapped) {
    a.mail = ctx->mail_d;
    a.mail_seq = ++ctx->vox_seq;
  }
  ESKF_TRY(voxelize(ctx, a));
  SortView v = sort_view(ctx, n);
  KnnParams P;
  P.key[0] = v.key[0];
  P.key[1] = v.key[1];
  P.hdr = v.hdr;
  P.sx = v.sx;
  P.sy = v.sy;
  P.sz = v.sz;
  P.sidx[0] = v.idx[0];
  P.sidx[1] = v.idx[1];
  P.kept_src = v.kept_src;
  P.kept_pos = v.kept_pos;
  P.px = tx;
  P.py = ty;
  P.pz = tz;
  P.n = n;
  P.voxel = voxel;
  P.ox = out->x();
  P.oy = out->y();
  P.oz = out->z();
  P.ocov = out->cov;
  P.opitch = out->cap;
  P.osrc = out->src;
  P.orig = orig;
  P.oc4 = out->c4;
  P.oc2 = out->c2;
  P.levels = v.levels;
  P.level_stride = v.level_stride;
  const size_t pitch = (static_cast<size_t>(n) + 63) / 64 * 64;
  ESKF_TRY(ctx->knn_nbr.ensure((static_cast<size_t>(kKnn) + 1) * pitch * sizeof(uint32_t)));
  P.nbr = ctx->knn_nbr.as<uint32_t>();
  P.nbr_cnt = P.nbr + static_cast<size_t>(kKnn) * pitch;
  P.nbr_pitch = pitch;
  P.knn_next = &v.hdr->knn_next;
  P.buf_limit = static_cast<unsigned>(ctx->opt_knn_buffer);
  P.stats = nullptr;
  if (ctx->opt_trace) {
    ESKF_TRY(ctx->misc.ensure(256));
    P.stats = ctx->misc.as<unsigned long long>();
    ESKF_CUDA(cudaMemsetAsync(P.stats, 0, 8 * sizeof(unsigned long long), ctx->stream));
  }
  static const int legacy = [] {
    const char* e = getenv("ESKF_KNN_LEGACY");  // A/B knob: the one-kernel shuffle-insertion search
    return e ? atoi(e) : 0;
  }();
  if (legacy) {
    const unsigned blocks = std::min<unsigned>((n + 3) / 4, static_cast<unsigned>(ctx->sm_count) * 16u);
    knn_cov_kernel<<<blocks, 128, 0, ctx->stream>>>(P);
    ESKF_CUDA(cudaGetLastError());
    count_launch(ctx);
  } else {
    // the number of kept points is only known on the device: size the grids
    // for the worst case (every point kept); surplus warps / threads exit at once
    static int per_sm = 0;
    if (per_sm == 0) {
      ESKF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, knn_search_kernel, kSearchThreads, 0));
      if (per_sm < 1) per_sm = 1;
    }
    const unsigned sblocks = std::min<unsigned>((n + kSearchWarps - 1) / kSearchWarps,
                                                static_cast<unsigned>(ctx->sm_count * per_sm));
    knn_search_kernel<<<sblocks, kSearchThreads, 0, ctx->stream>>>(P);
    ESKF_CUDA(cudaGetLastError());
    trace_mark(ctx, "knn_search");
    knn_finish_kernel<<<(n + 63) / 64, 64, 0, ctx->stream>>>(P);
    ESKF_CUDA(cudaGetLastError());
    ctx->knn_levels_clean = true;
    trace_mark(ctx, "knn_finish");
    count_launch(ctx, 2);
  }
  *mapped_out = mapped;
  return ESKF_OK;
}

int preprocess_finish(eskf_ctx* ctx, eskf_cloud* out, bool mapped) {
  unsigned n_out = 0;
  int waited = 1;
  if (mapped) {
    // returns as soon as the voxelize kernel has published the count: the k-NN kernels keep
    // running behind (everything that follows is ordered on the same stream)
    waited = wait_mail(ctx, &ctx->mail_h->vox_seq, ctx->vox_seq);
    if (waited == ESKF_ERR_CUDA) return waited;
    if (waited == ESKF_OK) {
      if (ctx->mail_h->vox_gb_error) {
        set_error("grid barrier timeout in voxelize kernel");
        return ESKF_ERR_INTERNAL;
      }
      if (ctx->mail_h->vox_error & 1u) {
        set_error("voxel coordinate outside the 21-bit key range");
        return ESKF_ERR_RANGE;
      }
      n_out = ctx->mail_h->vox_n_out;
    }
  }
  if (waited != ESKF_OK) ESKF_TRY(check_header(ctx, &n_out));
  trace_flush(ctx, "preprocess");
  if (ctx->opt_trace) {
    unsigned long long h[8];
    if (cudaMemcpy(h, ctx->misc.p, sizeof h, cudaMemcpyDeviceToHost) == cudaSuccess && h[0] > 0)
      std::fprintf(stderr, "[eskf trace] knn: %llu queries, %.2f attempts/query (+%.2f levels skipped), per attempt: "
                   "%.0f pass-1 + %.0f pass-2 candidates, %.1f survivors, %llu overflows\n", h[0],
                   static_cast<double>(h[1]) / h[0], static_cast<double>(h[6]) / h[0],
                   static_cast<double>(h[2]) / (h[1] ? h[1] : 1), static_cast<double>(h[3]) / (h[1] ? h[1] : 1),
                   static_cast<double>(h[4]) / (h[1] ? h[1] : 1), h[5]);
  }
  out->n = n_out;
  out->has_cov = true;
  out->has_c32 = true;
  out->has_src = true;
  return ESKF_OK;
}

// ---- range crop ------------------------------------------------------------
// north_star names a range crop in the preprocessor; the reference has none
// (src/CloudPreprocessor.cpp:10-23), so the oracle defines it (oracle.preprocess):
// a point of the incoming sweep survives iff min^2 <= x^2 + y^2 + z^2 <= max^2 in
// the LiDAR frame (fp64, ((x*x + y*y) + z*z), no FMA); T_il and the deskew act on
// every point exactly as before (segments and the sweep's end time come from the
// full stamp array), and the cropped points are erased ahead of
// voxelDownsampleAndEstimateCovariances: they neither claim a voxel nor count as
// neighbours.  Here: a stable compaction of the raw sweep (two small kernels)
// into a scratch cloud + the surviving points' original indices, which the
// voxelize kernel uses to find a point's deskew segment and the k-NN kernels to
// report source indices.
constexpr int kCropT = 256;

__device__ __forceinline__ bool crop_keeps(double x, double y, double z, double min2, double max2) {
  const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
  return r2 >= min2 && r2 <= max2;
}

__global__ void __launch_bounds__(kCropT) crop_count_kernel(const double* x, const double* y, const double* z,
                                                           unsigned n, double min2, double max2, unsigned* counts) {
  const unsigned i = blockIdx.x * kCropT + threadIdx.x;
  const bool keep = i < n && crop_keeps(x[i], y[i], z[i], min2, max2);
  const unsigned c = __syncthreads_count(keep);
  if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(kCropT) crop_scatter_kernel(const double* x, const double* y, const double* z,
                                                             unsigned n, double min2, double max2,
                                                             const unsigned* counts, unsigned n_blocks, double* ox,
                                                             double* oy, double* oz, uint32_t* orig, unsigned* total) {
  __shared__ unsigned s_warp[kCropT / 32];
  __shared__ unsigned s_base;
  const unsigned t = threadIdx.x, lane = t & 31, w = t >> 5;
  // survivors in the blocks before this one
  unsigned part = 0;
  for (unsigned b = t; b < blockIdx.x; b += kCropT) part += counts[b];
  part = warp_reduce_add(part);
  if (lane == 0) s_warp[w] = part;
  __syncthreads();
  if (t == 0) {
    unsigned base = 0;
    for (int k = 0; k < kCropT / 32; ++k) base += s_warp[k];
    s_base = base;
  }
  __syncthreads();
  const unsigned i = blockIdx.x * kCropT + t;
  double px = 0.0, py = 0.0, pz = 0.0;
  bool keep = false;
  if (i < n) {
    px = x[i]; py = y[i]; pz = z[i];
    keep = crop_keeps(px, py, pz, min2, max2);
  }
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  __syncthreads();
  if (lane == 0) s_warp[w] = __popc(m);
  __syncthreads();
  unsigned off = s_base;
  for (unsigned k = 0; k < w; ++k) off += s_warp[k];
  if (keep) {
    const unsigned o = off + __popc(m & ((1u << lane) - 1u));
    ox[o] = px; oy[o] = py; oz[o] = pz;
    orig[o] = i;
  }
  if (blockIdx.x == n_blocks - 1 && t == 0) {
    unsigned tot = s_base;
    for (int k = 0; k < kCropT / 32; ++k) tot += s_warp[k];
    *total = tot;
  }
}

// raw -> ctx->crop_cloud (+ ctx->crop_orig); the survivor count comes back with one small copy
int crop_device(eskf_ctx* ctx, const eskf_cloud* raw) {
  const unsigned n = static_cast<unsigned>(raw->n);
  const unsigned blocks = (n + kCropT - 1) / kCropT;
  if (!ctx->crop_cloud) ESKF_TRY(eskf_cloud_create(ctx, raw->n, &ctx->crop_cloud));
  ESKF_TRY(cloud_reserve(ctx->crop_cloud, raw->n, false));
  ESKF_TRY(ctx->crop_orig.ensure(static_cast<size_t>(n) * sizeof(uint32_t)));
  ESKF_TRY(ctx->crop_cnt.ensure((static_cast<size_t>(blocks) + 1) * sizeof(unsigned)));
  unsigned* counts = ctx->crop_cnt.as<unsigned>();
  crop_count_kernel<<<blocks, kCropT, 0, ctx->stream>>>(raw->x(), raw->y(), raw->z(), n, ctx->crop_min2,
                                                        ctx->crop_max2, counts);
  ESKF_CUDA(cudaGetLastError());
  eskf_cloud* c = ctx->crop_cloud;
  crop_scatter_kernel<<<blocks, kCropT, 0, ctx->stream>>>(raw->x(), raw->y(), raw->z(), n, ctx->crop_min2,
                                                          ctx->crop_max2, counts, blocks, c->x(), c->y(), c->z(),
                                                          ctx->crop_orig.as<uint32_t>(), counts + blocks);
  ESKF_CUDA(cudaGetLastError());
  count_launch(ctx, 2);
  unsigned* h = nullptr;
  ESKF_TRY(ctx_pinned(ctx, sizeof(unsigned), reinterpret_cast<void**>(&h)));
  ESKF_CUDA(cudaMemcpyAsync(h, counts + blocks, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
  ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
  c->n = *h;
  c->has_cov = c->has_c32 = c->has_src = false;
  return ESKF_OK;
}

}  // namespace
}  // namespace eskf

using namespace eskf;

extern "C" {

namespace {
int preprocess_launch_any(eskf_ctx* ctx, eskf_cloud* raw, const double* T_il, const std::vector<DeskewSeg>& segs,
                          double voxel_size, eskf_cloud* out, bool keep_input, bool* mapped, bool* empty) {
  *empty = false;
  if (ctx->crop) {
    ESKF_TRY(crop_device(ctx, raw));  // (reads raw, writes the scratch cloud: raw stays as delivered)
    if (ctx->crop_cloud->n == 0) {
      *empty = true;
      return ESKF_OK;
    }
    return eskf::preprocess_launch(ctx, ctx->crop_cloud, T_il, segs, voxel_size, out, ctx->crop_orig.as<uint32_t>(),
                                   false, mapped);
  }
  return eskf::preprocess_launch(ctx, raw, T_il, segs, voxel_size, out, nullptr, keep_input, mapped);
}
}  // namespace

int eskf_preprocess_cloud(eskf_ctx* ctx, eskf_cloud* raw, const double* point_time,
                          const double T_il[16], const eskf_state* states, size_t n_states,
                          double voxel_size, eskf_cloud* out) {
  ESKF_REQUIRE(ctx && raw && out, "null argument");
  ESKF_REQUIRE(raw->ctx == ctx && out->ctx == ctx, "clouds belong to another context");
  ESKF_REQUIRE(raw != out, "raw and out must be different clouds");
  ESKF_REQUIRE(voxel_size > 0.0, "voxel_size must be positive");
  ESKF_REQUIRE(raw->n < (1ull << 31), "cloud too large");
  ESKF_CUDA(cudaSetDevice(ctx->device));
  if (raw->n == 0) {
    out->n = 0;
    return ESKF_OK;
  }
  // Deskew segments need to know whether the stamps are non-decreasing (compute_deskew_segments).
  // Unless the caller says (option "stamps_sorted"), the answer is taken for granted, the kernels are
  // launched, and the one pass over the stamps that settles it runs on the host WHILE the GPU works
  // (the host would only be waiting for the kept-point count); if the stamps turn out unsorted the call is
  // repeated with the segments of the reference's forward scan -- the first attempt left raw untouched.
  std::vector<DeskewSeg> segs;
  const bool optimistic = n_states > 0 && ctx->opt_stamps_sorted < 0;
  if (n_states > 0) {
    ESKF_REQUIRE(states && point_time, "deskew needs states and point_time");
    ESKF_TRY(compute_deskew_segments(point_time, raw->n, states, n_states, &segs, optimistic ? 1 : ctx->opt_stamps_sorted));
  }
  raw->has_cov = false;
  raw->has_c32 = false;
  bool mapped = false, empty = false;
  ESKF_TRY(preprocess_launch_any(ctx, raw, T_il, segs, voxel_size, out, optimistic, &mapped, &empty));
  if (optimistic && !eskf_stamps_sorted(point_time, raw->n)) {
    ESKF_TRY(compute_deskew_segments(point_time, raw->n, states, n_states, &segs, 0));
    ESKF_TRY(preprocess_launch_any(ctx, raw, T_il, segs, voxel_size, out, false, &mapped, &empty));
  }
  if (empty) {
    out->n = 0;
    return ESKF_OK;
  }
  return eskf::preprocess_finish(ctx, out, mapped);
}

int eskf_stamps_sorted(const double* point_time, size_t n) {
  if (!point_time) return 1;
  // branch-free blocks (the compiler vectorises the inner loop), early exit between blocks
  size_t i = 1;
  while (i < n) {
    const size_t e = std::min(n, i + 4096);
    int bad = 0;
    for (; i < e; ++i) bad |= point_time[i] < point_time[i - 1];
    if (bad) return 0;
  }
  return 1;
}

int eskf_ctx_set_range_crop(eskf_ctx* ctx, double min_range, double max_range) {
  ESKF_REQUIRE(ctx, "null ctx");
  ESKF_REQUIRE(min_range >= 0.0 && (max_range == 0.0 || max_range >= min_range), "range crop needs 0 <= min <= max (max 0: unbounded)");
  ctx->crop = min_range > 0.0 || max_range > 0.0;
  ctx->crop_min2 = min_range * min_range;
  ctx->crop_max2 = max_range > 0.0 ? max_range * max_range : 1.7976931348623157e308;
  return ESKF_OK;
}

int eskf_preprocess(eskf_ctx* ctx, const double* xyz, const double* point_time, size_t n,
                    const double T_il[16], const eskf_state* states, size_t n_states,
                    double voxel_size, size_t* n_out, double* xyz_out, double* cov_out,
                    uint32_t* src_index_out) {
  ESKF_REQUIRE(ctx && n_out, "null argument");
  *n_out = 0;
  if (n == 0) return ESKF_OK;
  ESKF_REQUIRE(xyz, "null xyz");
  if (!ctx->tmp_cloud[1]) ESKF_TRY(eskf_cloud_create(ctx, n, &ctx->tmp_cloud[1]));
  if (!ctx->tmp_cloud[2]) ESKF_TRY(eskf_cloud_create(ctx, n, &ctx->tmp_cloud[2]));
  ESKF_TRY(eskf_cloud_upload(ctx->tmp_cloud[1], xyz, nullptr, n));
  ESKF_TRY(eskf_preprocess_cloud(ctx, ctx->tmp_cloud[1], point_time, T_il, states, n_states,
                                 voxel_size, ctx->tmp_cloud[2]));
  return eskf_cloud_download(ctx->tmp_cloud[2], xyz_out, cov_out, src_index_out, n, n_out);
}

int eskf_downsample_cov(eskf_ctx* ctx, const double* xyz, size_t n, double voxel_size,
                        size_t* n_out, double* xyz_out, double* cov_out, uint32_t* src_index_out) {
  return eskf_preprocess(ctx, xyz, nullptr, n, nullptr, nullptr, 0, voxel_size, n_out, xyz_out,
                         cov_out, src_index_out);
}

}  // extern "C"
