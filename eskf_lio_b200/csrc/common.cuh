// common.cuh — device helpers shared by the hot-path kernels (sm_100a).
//
// Exact-arithmetic contract: everything that decides a voxel key or feeds the
// fp64 master statistics of the map uses round-to-nearest intrinsics
// (__dmul_rn/__dadd_rn/__ddiv_rn), which nvcc never contracts into FMAs, in
// the same evaluation order as the reference's Eigen/Open3D expressions built
// without FMA (CMakeLists.txt:6-9):  ((a0*b0 + a1*b1) + a2*b2) [+ t].
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace eskf {

constexpr int kKeyBits = 21;
constexpr int64_t kKeyBias = 1 << 20;
constexpr uint64_t kEmptyKey = ~0ull;

// ------------------------------------------------------------- exact fp64
__device__ __forceinline__ double dot3_rn(double a0, double b0, double a1, double b1, double a2,
                                          double b2) {
  return __dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2));
}

// Open3D PointCloud::Transform / Eigen Isometry3d * Vector3d:  (R p) + t.
// T points at 12 doubles: R row-major (9) then t (3).
__device__ __forceinline__ void transform_point_rn(const double* __restrict__ T, double& x,
                                                   double& y, double& z) {
  const double px = x, py = y, pz = z;
  x = __dadd_rn(dot3_rn(T[0], px, T[1], py, T[2], pz), T[9]);
  y = __dadd_rn(dot3_rn(T[3], px, T[4], py, T[5], pz), T[10]);
  z = __dadd_rn(dot3_rn(T[6], px, T[7], py, T[8], pz), T[11]);
}

// C <- (R C) R^T, row-major 3x3, same association as Eigen's R * C * R^T.
__device__ __forceinline__ void rotate_cov_rn(const double* __restrict__ R, double* C) {
  double A[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      A[3 * i + j] = dot3_rn(R[3 * i], C[j], R[3 * i + 1], C[3 + j], R[3 * i + 2], C[6 + j]);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = dot3_rn(A[3 * i], R[3 * j], A[3 * i + 1], R[3 * j + 1], A[3 * i + 2], R[3 * j + 2]);
}

// getVoxelIndex (src/LocalMap.cpp:114-118, src/CloudPreprocessor.cpp:129-133):
// floor(p / voxel) with a true IEEE division, then a truncating cast.
__device__ __forceinline__ int voxel_coord(double p, double voxel) {
  return static_cast<int>(floor(__ddiv_rn(p, voxel)));
}

// Same result as voxel_coord(p, voxel) without the fp64 division in the common
// case: q = p * (1/voxel) is within 3*2^-53 |q| of the correctly rounded
// quotient, so floor(q) can only differ when q sits that close to an integer;
// then (about once per 1e9 coordinates) the exact division decides.
__device__ __forceinline__ int voxel_coord(double p, double voxel, double inv_voxel) {
  const double q = p * inv_voxel;
  const double fl = floor(q);
  const double fr = q - fl;
  const double tol = fabs(q) * 4.5e-16 + 1e-300;
  if (fr < tol || 1.0 - fr < tol) return static_cast<int>(floor(__ddiv_rn(p, voxel)));
  return static_cast<int>(fl);
}

__device__ __forceinline__ bool coord_in_range(int k) { return k > -kKeyBias && k < kKeyBias; }

// ------------------------------------------------------------------- keys
// table key: three biased 21-bit fields, (kx, ky, kz) lexicographic
__host__ __device__ __forceinline__ uint64_t pack_key(int x, int y, int z) {
  return (static_cast<uint64_t>(x + kKeyBias) << 42) | (static_cast<uint64_t>(y + kKeyBias) << 21) |
         static_cast<uint64_t>(z + kKeyBias);
}

__host__ __device__ __forceinline__ void unpack_key(uint64_t k, int& x, int& y, int& z) {
  x = static_cast<int>((k >> 42) & 0x1FFFFF) - static_cast<int>(kKeyBias);
  y = static_cast<int>((k >> 21) & 0x1FFFFF) - static_cast<int>(kKeyBias);
  z = static_cast<int>(k & 0x1FFFFF) - static_cast<int>(kKeyBias);
}

// murmur3 fmix64
__host__ __device__ __forceinline__ uint64_t hash_key(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}

// Table addressing.  The table is probed through a compact array of 16-bit
// TAGS (0 = empty slot); the 64 B voxel record at the same index is touched
// only when the tag matches.  The tag array of a multi-million-voxel map
// (2 B/slot) fits the 126 MB L2, so a lookup that misses never goes to HBM and
// a hit costs exactly one 64 B record.  home = high-multiply of the low hash
// word into [0, n_slots) (n_slots need not be a power of two).
// (A 4x4x4 "brick" hash that keeps neighbouring voxels in adjacent slots was
// tried and measured SLOWER on B200 — 1.69 vs 1.20 ms per 10-iteration dense
// launch — linear probing through interleaved bricks lengthens miss chains.)
using tag_t = uint16_t;  // 2 B/slot: the tag array of 16M slots is 32 MB
// homes aligned to buckets of kBucket slots (a power of two <= 64; 1 = plain linear probing).  16 makes a
// probe window one aligned 16 B load, but measured a net loss on B200: the register-pipelined kernel got 5 %
// slower on the grown 0.1 m table and 9 % on the compact one (longer runs inside a bucket), the
// parked-candidate kernel was unchanged (profiles/r2_align_experiments.md)
#ifndef ESKF_BUCKET
#define ESKF_BUCKET 1
#endif
constexpr uint32_t kBucket = ESKF_BUCKET;
struct SlotAddr {
  uint32_t home;
  tag_t tag;
};
__host__ __device__ __forceinline__ SlotAddr slot_addr(uint64_t key, uint32_t n_slots) {
  const uint64_t h = hash_key(key);
  SlotAddr a;
  // homes are aligned to buckets of kBucket slots (n_slots is a multiple of 64): the keys of a bucket
  // fill it front to back, so a lookup's first probe WINDOW — 16 filter bytes = one aligned 16 B load,
  // or 8 tags = one aligned 16 B load — starts at the home and needs no offset arithmetic; at the load
  // factors the table runs at (<= 1/2) a bucket overflows into the next one for < 1 % of the keys
  a.home = static_cast<uint32_t>(((h & 0xffffffffull) * static_cast<uint64_t>(n_slots)) >> 32) & ~(kBucket - 1u);
  a.tag = static_cast<tag_t>((h >> 48) | 1u);
  return a;
}
__host__ __device__ __forceinline__ uint32_t next_slot(uint32_t h, uint32_t n_slots) {
  return h + 1u == n_slots ? 0u : h + 1u;
}

// 21-bit -> 63-bit Morton spreading (x bit i -> bit 3i)
__host__ __device__ __forceinline__ uint64_t spread3(uint32_t v) {
  uint64_t x = v & 0x1FFFFF;
  x = (x | (x << 32)) & 0x1F00000000FFFFull;
  x = (x | (x << 16)) & 0x1F0000FF0000FFull;
  x = (x | (x << 8)) & 0x100F00F00F00F00Full;
  x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

__host__ __device__ __forceinline__ uint32_t compact3(uint64_t x) {
  x &= 0x1249249249249249ull;
  x = (x | (x >> 2)) & 0x10C30C30C30C30C3ull;
  x = (x | (x >> 4)) & 0x100F00F00F00F00Full;
  x = (x | (x >> 8)) & 0x1F0000FF0000FFull;
  x = (x | (x >> 16)) & 0x1F00000000FFFFull;
  x = (x | (x >> 32)) & 0x1FFFFF;
  return static_cast<uint32_t>(x);
}

// Morton code of coordinates rebased to the batch minimum (all >= 0):
// x -> bits 3i+2, y -> 3i+1, z -> 3i
__host__ __device__ __forceinline__ uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) {
  return (spread3(x) << 2) | (spread3(y) << 1) | spread3(z);
}

// ----------------------------------------------- k-NN block-range tables
// After the Morton sort an aligned block of 2^L voxels per axis is ONE
// contiguous range of the sorted scan.  For the levels the 30-NN search visits
// most (L < kKnnHashLevels) the voxelize kernel records every occupied block's
// [start, end) in a small open-addressing table, so a neighbour-block lookup
// costs one or two 16 B loads instead of two 16-step binary searches.
// Entry (uint4): x,y = block key (Morton >> 3L) lo/hi, z = start, w = end;
// all-ones = empty.  Level L's table starts at L * level_stride entries and
// uses knn_level_slots() of them (a power of two >= 2 x the blocks it can hold; the kernels size it by
// the number of occupied voxels, the host reserves for one block per point).
constexpr int kKnnHashLevels = 6;

__host__ __device__ __forceinline__ unsigned knn_level_slots(unsigned n, unsigned bits, int L) {
  const unsigned b = bits > static_cast<unsigned>(L) ? bits - static_cast<unsigned>(L) : 0u;
  unsigned long long most = n;  // occupied blocks <= min(n, 8^b)
  if (b < 10u && (1ull << (3u * b)) < most) most = 1ull << (3u * b);
  const unsigned long long want = 2ull * most > 8ull ? 2ull * most : 8ull;  // next power of two >= want
#ifdef __CUDA_ARCH__
  return 1u << (64 - __clzll(static_cast<long long>(want - 1ull)));
#else
  return 1u << (64 - __builtin_clzll(want - 1ull));
#endif
}

__device__ __forceinline__ uint4* knn_level_claim_from(uint4* tab, unsigned mask, uint64_t bk, unsigned h) {
  for (;;) {
    unsigned long long* kp = reinterpret_cast<unsigned long long*>(tab + h);
    const unsigned long long old = atomicCAS(kp, ~0ull, static_cast<unsigned long long>(bk));
    if (old == ~0ull || old == bk) return tab + h;
    h = (h + 1u) & mask;
  }
}

__device__ __forceinline__ bool knn_level_find(const uint4* tab, unsigned mask, uint64_t bk,
                                               unsigned& start, unsigned& end) {
  unsigned h = static_cast<unsigned>(hash_key(bk)) & mask;
  for (unsigned probe = 0; probe <= mask; ++probe) {
    const uint4 e = __ldg(tab + h);
    const uint64_t k = (static_cast<uint64_t>(e.y) << 32) | e.x;
    if (k == bk) {
      start = e.z;
      end = e.w;
      return true;
    }
    if (k == ~0ull) return false;
    h = (h + 1u) & mask;
  }
  return false;
}

// --------------------------------------------------------- cache-global IO
template <typename T>
__device__ __forceinline__ T ld_cg(const T* p) {
  return __ldcg(p);
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------ grid barrier
// Software grid-wide barrier for persistent kernels launched cooperatively
// (all CTAs co-resident).  One acq_rel ticket atomic per CTA and one release
// store by the last arriver; no separate fences (bar.sync + the cumulative
// release of thread 0 publish the whole CTA's writes).  `count` only grows:
// barrier k owns tickets [k * nblocks, (k + 1) * nblocks).  Every spin is
// bounded: on timeout the error word is set and the caller must bail out, so
// a logic error can never hang the GPU.
struct GridBarrier {
  unsigned count;  // zeroed before the launch
  unsigned gen;    // number of completed barriers
  unsigned error;
  unsigned pad;
};

constexpr unsigned kSpinLimit = 1u << 24;  // x >= 32 ns sleeps: > 0.5 s

__device__ __forceinline__ unsigned atom_add_acq_rel_u32(unsigned* p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}

__device__ __forceinline__ bool grid_sync(GridBarrier* gb, unsigned nblocks) {
  __shared__ unsigned s_ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned ok = 1;
    if (nblocks > 1) {
      const unsigned ticket = atom_add_acq_rel_u32(&gb->count, 1u);
      const unsigned epoch = ticket / nblocks + 1u;
      if (ticket % nblocks == nblocks - 1u) {
        st_release_u32(&gb->gen, epoch);
      } else {
        unsigned spins = 0;
        while (static_cast<int>(ld_acquire_u32(&gb->gen) - epoch) < 0) {
          __nanosleep(32);
          if (++spins > kSpinLimit || ld_acquire_u32(&gb->error) != 0) {
            atomicExch(&gb->error, 1u);
            ok = 0;
            break;
          }
        }
      }
    }
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}

// ------------------------------------------------------------ block helpers
__device__ __forceinline__ int warp_reduce_min(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_reduce_max(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ unsigned warp_reduce_add(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_reduce_add(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace eskf
