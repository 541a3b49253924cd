// voxelize.cu — K1/K2/K3: transform + voxel key + radix sort by key + run heads,
// as ONE persistent cooperative kernel (software grid barriers between phases,
// so a 64k-point scan costs one launch instead of ~15).
//
// Replaces, on the device:
//   - the getVoxelIndex + unordered_map passes of
//     CloudPreprocessor::voxelDownsampleAndEstimateCovariances
//     (src/CloudPreprocessor.cpp:85-99,129-133): first point (lowest input
//     index) per voxel wins == head of each key run of a STABLE sort;
//   - the per-point find-or-emplace ordering of LocalMap::updateLocalMap
//     (src/LocalMap.cpp:47-58): points of one voxel are merged in input index
//     order == order inside a key run of a stable sort;
//   - PointCloud::Transform in front of both (src/CloudPreprocessor.cpp:16,
//     src/LocalMap.cpp:15) and the per-segment deskew transform
//     (src/CloudPreprocessor.cpp:67-72).
//
// Sort key: Morton code of the voxel coordinates rebased to the batch minimum
// (so only ceil(3*bits/8) 8-bit LSD passes run, and an aligned 2^L block of
// voxels is one contiguous key range — the k-NN search relies on that).
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

#include <cooperative_groups.h>

#include "internal.h"

namespace eskf {

namespace {

#ifndef ESKF_VOX_THREADS
#define ESKF_VOX_THREADS 256
#endif
constexpr int kT = ESKF_VOX_THREADS;  // threads per CTA (>= 256: thread t < 256 owns radix digit t)
constexpr int kW = kT / 32;
static_assert(kT >= 256 && kT % 32 == 0, "the digit-per-thread steps need at least 256 threads");

struct VoxParams {
  VoxelizeArgs a;
  uint64_t* key[2];
  uint32_t* idx[2];
  unsigned* hist;       // [2][G][256]
  unsigned* blk_count;  // [2][G]
  VoxelHeader* hdr;
  uint32_t* run_start;
  uint32_t* keep_flag;
  uint32_t* kept_src;
  uint32_t* kept_pos;
  double* sx;
  double* sy;
  double* sz;
  uint4* levels;          // mode 1: k-NN block-range tables (common.cuh)
  unsigned level_stride;
  unsigned chunk;  // elements per CTA, multiple of kT
  unsigned long long* stamps;  // ESKF_TRACE: %globaltimer at the phase boundaries of the one-cluster kernel (CTA 0)
};

// the same with the per-warp totals scanned by one warp (W <= 32): a third of the instructions for 32 warps
template <int W>
__device__ __forceinline__ unsigned block_exclusive_scan2(unsigned v, unsigned* s_tmp, unsigned* total = nullptr) {
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= static_cast<unsigned>(o)) inc += u;
  }
  __syncthreads();
  if (lane == 31) s_tmp[w] = inc;
  __syncthreads();
  if (w == 0) {
    const unsigned c = lane < static_cast<unsigned>(W) ? s_tmp[lane] : 0u;
    unsigned ci = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, ci, o);
      if (lane >= static_cast<unsigned>(o)) ci += u;
    }
    s_tmp[lane] = ci - c;             // exclusive prefix of the warps
    if (lane == 31) s_tmp[32] = ci;   // everything
  }
  __syncthreads();
  const unsigned woff = s_tmp[w];
  if (total) *total = s_tmp[32];
  __syncthreads();
  return woff + inc - v;
}

template <int W = kW>
__device__ __forceinline__ unsigned block_reduce_add(unsigned v, unsigned* s_tmp) {
  v = warp_reduce_add(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_tmp[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned r = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) r += s_tmp[i];
  return r;
}

// exclusive scan of one value per thread over the CTA (32 W threads)
template <int W = kW>
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* s_tmp, unsigned* total) {
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= static_cast<unsigned>(o)) inc += u;
  }
  __syncthreads();
  if (lane == 31) s_tmp[w] = inc;
  __syncthreads();
  unsigned woff = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    unsigned c = s_tmp[i];
    if (i < static_cast<int>(w)) woff += c;
    tot += c;
  }
  if (total) *total = tot;
  // all reads of s_tmp done before any warp runs ahead into code that may write shared memory
  // again (compute-sanitizer racecheck flagged a WAR against the next phase's stores)
  __syncthreads();
  return woff + inc - v;
}

__global__ void __launch_bounds__(kT) voxelize_kernel(VoxParams P) {
  const unsigned G = gridDim.x, b = blockIdx.x, t = threadIdx.x;
  const unsigned w = t >> 5, lane = t & 31;
  const unsigned n = P.a.n;
  GridBarrier* gb = &P.hdr->gb;

  const unsigned bbeg = min(n, b * P.chunk);
  const unsigned bend = min(n, bbeg + P.chunk);
  const unsigned wchunk = P.chunk / kW;
  const unsigned wbeg = min(bend, bbeg + w * wchunk);
  const unsigned wend = min(bend, wbeg + wchunk);

  __shared__ unsigned s_whist[kW][256];
  __shared__ unsigned s_off[kW][256];
  __shared__ unsigned s_tmp[kW];

  // ---------------------------------------------------------------- phase 0
  // transform (+deskew) + voxel coordinate + batch min/max
  {
    int mn0 = INT_MAX, mn1 = INT_MAX, mn2 = INT_MAX;
    int nm0 = INT_MAX, nm1 = INT_MAX, nm2 = INT_MAX;
    bool bad = false;
    const int stride = P.a.in_stride;
    for (unsigned i = bbeg + t; i < bend; i += kT) {
      double x = P.a.in_x[static_cast<size_t>(i) * stride];
      double y = P.a.in_y[static_cast<size_t>(i) * stride];
      double z = P.a.in_z[static_cast<size_t>(i) * stride];
      if (P.a.has_T1) transform_point_rn(P.a.T1, x, y, z);
      if (P.a.n_segs > 0) {
        const unsigned oi = P.a.orig != nullptr ? P.a.orig[i] : i;  // (segments index the uncropped sweep)
        int lo = 0, hi = P.a.n_segs;  // first segment with end > oi
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (P.a.segs[mid].end > oi) hi = mid; else lo = mid + 1;
        }
        if (lo < P.a.n_segs && oi >= P.a.segs[lo].begin) transform_point_rn(P.a.segs[lo].T, x, y, z);
      }
      P.a.out_x[i] = x;
      P.a.out_y[i] = y;
      P.a.out_z[i] = z;
      if (P.a.cov != nullptr && P.a.has_T1) {
        double C[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) C[k] = P.a.cov[k * P.a.cov_pitch + i];
        rotate_cov_rn(P.a.T1, C);
#pragma unroll
        for (int k = 0; k < 9; ++k) P.a.cov[k * P.a.cov_pitch + i] = C[k];
      }
      const int kx = voxel_coord(x, P.a.voxel);
      const int ky = voxel_coord(y, P.a.voxel);
      const int kz = voxel_coord(z, P.a.voxel);
      if (!(coord_in_range(kx) && coord_in_range(ky) && coord_in_range(kz))) bad = true;
      P.key[1][i] = pack_key(kx, ky, kz);  // absolute key parked in buffer B
      mn0 = min(mn0, kx); mn1 = min(mn1, ky); mn2 = min(mn2, kz);
      nm0 = min(nm0, -kx); nm1 = min(nm1, -ky); nm2 = min(nm2, -kz);
    }
    mn0 = warp_reduce_min(mn0); mn1 = warp_reduce_min(mn1); mn2 = warp_reduce_min(mn2);
    nm0 = warp_reduce_min(nm0); nm1 = warp_reduce_min(nm1); nm2 = warp_reduce_min(nm2);
    if (lane == 0 && mn0 != INT_MAX) {
      atomicMin(&P.hdr->mn[0], mn0); atomicMin(&P.hdr->mn[1], mn1); atomicMin(&P.hdr->mn[2], mn2);
      atomicMin(&P.hdr->nmx[0], nm0); atomicMin(&P.hdr->nmx[1], nm1); atomicMin(&P.hdr->nmx[2], nm2);
    }
    if (bad) atomicOr(&P.hdr->error, 1u);
  }
  if (!grid_sync(gb, G)) return;

  // ---------------------------------------------------------------- phase 1
  // rebased Morton key + identity payload (own chunk only: no grid barrier)
  const int m0 = ld_cg(&P.hdr->mn[0]), m1 = ld_cg(&P.hdr->mn[1]), m2 = ld_cg(&P.hdr->mn[2]);
  unsigned bits;
  {
    const int e0 = -ld_cg(&P.hdr->nmx[0]) - m0;
    const int e1 = -ld_cg(&P.hdr->nmx[1]) - m1;
    const int e2 = -ld_cg(&P.hdr->nmx[2]) - m2;
    const unsigned ex = static_cast<unsigned>(max(max(e0, e1), max(e2, 0)));
    bits = ex == 0 ? 0u : (32u - static_cast<unsigned>(__clz(ex)));
    if (bits > kKeyBits) bits = kKeyBits;
  }
  for (unsigned i = bbeg + t; i < bend; i += kT) {
    int kx, ky, kz;
    unpack_key(P.key[1][i], kx, ky, kz);
    P.key[0][i] = morton3(static_cast<uint32_t>(kx - m0), static_cast<uint32_t>(ky - m1),
                          static_cast<uint32_t>(kz - m2));
    P.idx[0][i] = i;
  }
  if (b == 0 && t == 0) P.hdr->bits = bits;
  // (mode 1: the k-NN block-range tables are all-empty here: knn_finish_kernel empties what its sweep
  // used, voxelize() the whole buffer after an allocation or a failed call)
  __syncthreads();

  // ------------------------------------------------------- LSD radix passes
  const unsigned npass = (3u * bits + 7u) / 8u;
  unsigned sel = 0;
  for (unsigned p = 0; p < npass; ++p) {
    const uint64_t* sk = P.key[sel];
    const uint32_t* si = P.idx[sel];
    const unsigned shift = 8u * p;
    // A: per-warp digit histograms of the warp's contiguous sub-chunk
    for (unsigned k = t; k < kW * 256; k += kT) (&s_whist[0][0])[k] = 0;
    __syncthreads();
    for (unsigned j0 = wbeg + lane; j0 < wend; j0 += 128) {
      uint64_t kk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) kk[u] = j0 + 32u * u < wend ? ld_cg(sk + j0 + 32u * u) : 0ull;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 + 32u * u < wend) atomicAdd(&s_whist[w][static_cast<unsigned>(kk[u] >> shift) & 255u], 1u);
    }
    __syncthreads();
    unsigned* histp = P.hist + static_cast<size_t>(p & 1u) * G * 256u;
    if (t < 256u) {
      unsigned s = 0;
#pragma unroll
      for (int i = 0; i < kW; ++i) s += s_whist[i][t];
      histp[b * 256u + t] = s;
    }
    if (!grid_sync(gb, G)) return;
    // B: global digit offsets for this CTA (every CTA redoes the tiny scan)
    // (16 independent loads in flight per thread: this scan is one L2 round
    // trip per batch, and it sits on the critical path of every pass)
    unsigned total = 0, pre = 0;
    for (unsigned b0 = 0; b0 < G && t < 256u; b0 += 16) {
      unsigned v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = b0 + k < G ? ld_cg(&histp[(b0 + k) * 256u + t]) : 0u;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (b0 + k < b) pre += v[k];
        total += v[k];
      }
    }
    // a digit shared by every key makes the pass the identity: skip it
    // (decision is identical in every CTA; hist is double-buffered so the
    // next pass may start writing without another barrier)
    if (__syncthreads_or(t < 256u && total == n)) continue;
    unsigned base = block_exclusive_scan(total, s_tmp, nullptr);
    if (t < 256u) {
      unsigned run = base + pre;
#pragma unroll
      for (int i = 0; i < kW; ++i) {
        s_off[i][t] = run;
        run += s_whist[i][t];
      }
    }
    __syncthreads();
    // C: stable scatter, one warp per contiguous sub-chunk, 32 keys a step
    uint64_t* dk = P.key[sel ^ 1u];
    uint32_t* di = P.idx[sel ^ 1u];
    uint64_t nkey = wbeg + lane < wend ? ld_cg(sk + wbeg + lane) : 0ull;
    uint32_t nid = wbeg + lane < wend ? ld_cg(si + wbeg + lane) : 0u;
    for (unsigned j0 = wbeg; j0 < wend; j0 += 32) {
      const unsigned j = j0 + lane;
      const bool valid = j < wend;
      const unsigned am = __ballot_sync(0xffffffffu, valid);
      const uint64_t key = nkey;
      const uint32_t id = nid;
      if (j + 32u < wend) {  // next step's loads overlap this step's ranking
        nkey = ld_cg(sk + j + 32u);
        nid = ld_cg(si + j + 32u);
      }
      if (valid) {
        const unsigned d = static_cast<unsigned>(key >> shift) & 255u;
        const unsigned mask = __match_any_sync(am, d);
        const unsigned leader = __ffs(mask) - 1;
        const unsigned off = s_off[w][d];
        __syncwarp(am);
        if (lane == leader) s_off[w][d] = off + __popc(mask);
        __syncwarp(am);
        const unsigned dst = off + __popc(mask & ((1u << lane) - 1u));
        dk[dst] = key;
        di[dst] = id;
      }
    }
    sel ^= 1u;
    if (!grid_sync(gb, G)) return;
  }
  const uint64_t* sk = P.key[sel];
  const uint32_t* si = P.idx[sel];
  if (b == 0 && t == 0) P.hdr->sel = sel;

  // ------------------------------------------------------------- run heads
  if (P.a.mode == 0) {
    // runs in sorted order: run_start[r] = first sorted position of run r
    unsigned cnt = 0;
    for (unsigned j = bbeg + t; j < bend; j += kT)
      cnt += (j == 0 || ld_cg(sk + j - 1) != ld_cg(sk + j)) ? 1u : 0u;
    cnt = block_reduce_add(cnt, s_tmp);
    if (t == 0) P.blk_count[b] = cnt;
    if (!grid_sync(gb, G)) return;
    unsigned base = 0;
    for (unsigned bb = t; bb < b; bb += kT) base += ld_cg(&P.blk_count[bb]);
    base = block_reduce_add(base, s_tmp);
    for (unsigned tile = bbeg; tile < bend; tile += kT) {
      const unsigned j = tile + t;
      const unsigned head = (j < bend && (j == 0 || ld_cg(sk + j - 1) != ld_cg(sk + j))) ? 1u : 0u;
      unsigned tot;
      const unsigned rank = block_exclusive_scan(head, s_tmp, &tot);
      if (head) P.run_start[base + rank] = j;
      base += tot;
    }
    if (b == G - 1 && t == 0) {
      P.hdr->n_out = base;
      P.run_start[base] = n;
    }
  } else {
    // kept points (run heads) listed in ascending SOURCE index, plus the
    // positions gathered into sorted order for the k-NN scan
    for (unsigned j = bbeg + t; j < bend; j += kT) {
      const bool head = (j == 0 || ld_cg(sk + j - 1) != ld_cg(sk + j));
      const uint32_t id = ld_cg(si + j);
      P.keep_flag[id] = head ? j + 1u : 0u;
      P.sx[j] = ld_cg(P.a.out_x + id);
      P.sy[j] = ld_cg(P.a.out_y + id);
      P.sz[j] = ld_cg(P.a.out_z + id);
    }
    if (!grid_sync(gb, G)) return;
    unsigned cnt = 0;
    for (unsigned i = bbeg + t; i < bend; i += kT) cnt += ld_cg(&P.keep_flag[i]) != 0u ? 1u : 0u;
    cnt = block_reduce_add(cnt, s_tmp);
    if (t == 0) P.blk_count[G + b] = cnt;
    if (!grid_sync(gb, G)) return;
    unsigned base = 0;
    for (unsigned bb = t; bb < b; bb += kT) base += ld_cg(&P.blk_count[G + bb]);
    base = block_reduce_add(base, s_tmp);
    for (unsigned tile = bbeg; tile < bend; tile += kT) {
      const unsigned i = tile + t;
      const unsigned f = i < bend ? ld_cg(&P.keep_flag[i]) : 0u;
      unsigned tot;
      const unsigned rank = block_exclusive_scan(f != 0u ? 1u : 0u, s_tmp, &tot);
      if (f != 0u) {
        P.kept_src[base + rank] = i;
        P.kept_pos[base + rank] = f - 1u;
      }
      base += tot;
    }
    if (b == G - 1 && t == 0) {
      P.hdr->n_out = base;
      if (P.a.mail != nullptr) {  // the host is polling for the kept-point count
        volatile HostMail* m = P.a.mail;
        m->vox_n_out = base;
        m->vox_error = ld_cg(&P.hdr->error);
        m->vox_gb_error = ld_cg(&P.hdr->gb.error);
        __threadfence_system();
        m->vox_seq = P.a.mail_seq;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// The same job for ONE sweep (mode 1, n <= 65536), as one thread-block cluster whose CTAs hold the
// (key, index) pairs in their shared memory for the whole sort: the scatter of a radix pass is a
// store into the owner CTA's shared memory (DSMEM), the digit histograms are read from the peers'
// shared memory, and the barriers between the phases are hardware cluster barriers.  The grid-wide
// kernel above pays ~3 us per software barrier (11 of them) plus an L2 round trip per dependent
// step; here nothing but the input, the outputs and the kept list touches L2.
//   element = (Morton key << 16) | source index   (one 8-byte word: one DSMEM store per element and pass)
//   wide keys (more than 16 bits per axis after rebasing): key and 16-bit index in separate arrays
// Global position j of the sorted order lives in CTA j / E at j % E, E = P.chunk (a power of two).
namespace cg = cooperative_groups;
constexpr int kTC = 1024;   // threads per CTA of the cluster
constexpr int kWC = kTC / 32;
constexpr int kHP = 257;    // pitch of the per-warp digit counts (conflict-free both by row and by column)
constexpr int kSegSmem = 128;  // deskew segment bounds staged in shared memory (more: searched in global memory)
constexpr unsigned kClusterMaxPoints = 65536;  // the source index travels in 16 bits
constexpr unsigned kClusterMaxChunk = 8192;    // elements per CTA (shared memory)

__host__ __device__ inline size_t cluster_smem_bytes(unsigned E) {
  return static_cast<size_t>(E) * (2 * 8 + 2 * 2) + (kWC * kHP + 3 * 256 + 2 * kSegSmem + 64) * sizeof(unsigned);
}

// Morton code of three coordinates below 2^10 with 32-bit operations (the 64-bit spread of common.cuh
// costs ~3x the instructions; same value)
__device__ __forceinline__ uint32_t spread3_10(uint32_t x) {
  x = (x | (x << 16)) & 0x030000FFu;
  x = (x | (x << 8)) & 0x0300F00Fu;
  x = (x | (x << 4)) & 0x030C30C3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}

__device__ __forceinline__ void vox_stamp(const VoxParams& P, int slot) {
  if (P.stamps != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.stamps[slot] = t;
  }
}

struct ClusterSmem {
  uint64_t* bufk;   // [2][E] elements (or keys, WIDE)
  uint16_t* bufi;   // [2][E] source indices (WIDE only)
  unsigned* whist;  // [32][kHP] per-warp digit counts -> exclusive prefix over the warps -> running scatter offsets
  unsigned* chist;  // [2][256] this CTA's digit counts (read by the peers)
  unsigned* dig;    // [256] global offset of this CTA's first element per digit
  unsigned* seg;    // [2][kSegSmem] deskew segment begin / end
  int* mm;          // [6] local min, [8..13] cluster min
  unsigned* cnt;    // kept points of this CTA (read by the peers)
  unsigned* tmp;    // [33]
};

template <bool WIDE>
__device__ __forceinline__ void cluster_sort_and_list(const VoxParams& P, cg::cluster_group& cl, const ClusterSmem& S,
                                                      unsigned bits, unsigned gbeg, unsigned cnt) {
  const unsigned CS = gridDim.x, rank = blockIdx.x, t = threadIdx.x, w = t >> 5, lane = t & 31;
  const unsigned n = P.a.n, E = P.chunk, logE = 31u - static_cast<unsigned>(__clz(E));
  const unsigned wch = E / kWC;
  const unsigned wbeg = min(cnt, w * wch), wend = min(cnt, wbeg + wch);
  const unsigned npass = (3u * bits + 7u) / 8u;
  unsigned* whist = S.whist;
  unsigned sel = 0;
  for (unsigned p = 0; p < npass; ++p) {
    const uint64_t* sk = S.bufk + sel * E;
    const uint16_t* si = S.bufi + sel * E;
    const unsigned shift = 8u * p + (WIDE ? 0u : 16u);
    // A: per-warp digit counts of the warp's contiguous sub-chunk ...
    for (unsigned k = t; k < kWC * kHP; k += kTC) whist[k] = 0;
    __syncthreads();
    for (unsigned j = wbeg + lane; j < wend; j += 32)
      atomicAdd(&whist[w * kHP + (static_cast<unsigned>(sk[j] >> shift) & 255u)], 1u);
    __syncthreads();
    // ... turned, per digit, into the exclusive prefix over the warps, and the CTA's count
    unsigned* mine = S.chist + (p & 1u) * 256u;
    if (t < 256u) {
      unsigned run = 0;
#pragma unroll
      for (int i = 0; i < kWC; ++i) {
        const unsigned c = whist[i * kHP + t];
        whist[i * kHP + t] = run;
        run += c;
      }
      mine[t] = run;
    }
    if (p == 0) vox_stamp(P, 4);
    cl.sync();
    if (p == 0) vox_stamp(P, 5);
    // B: this CTA's global offset per digit, from the peers' counts (one DSMEM round trip)
    unsigned total = 0, pre = 0;
    if (t < 256u) {
      unsigned v[16];
#pragma unroll
      for (unsigned r = 0; r < 16; ++r) v[r] = r < CS ? *cl.map_shared_rank(mine + t, r) : 0u;
#pragma unroll
      for (unsigned r = 0; r < 16; ++r) {
        if (r < rank) pre += v[r];
        total += v[r];
      }
    }
    // a digit shared by every key makes the pass the identity: skip it (same decision in every CTA;
    // the count buffers alternate, so the next pass may write its own while peers still read these)
    if (__syncthreads_or(t < 256u && total == n)) continue;
    const unsigned base = block_exclusive_scan2<kWC>(t < 256u ? total : 0u, S.tmp);
    if (t < 256u) S.dig[t] = base + pre;
    __syncthreads();
    if (p == 0) vox_stamp(P, 6);
    // C: stable scatter into the owners' shared memory
    uint64_t* dk = S.bufk + (sel ^ 1u) * E;
    uint16_t* di = S.bufi + (sel ^ 1u) * E;
    for (unsigned j0 = wbeg; j0 < wend; j0 += 32) {
      const unsigned j = j0 + lane;
      const bool valid = j < wend;
      const unsigned am = __ballot_sync(0xffffffffu, valid);
      if (valid) {
        const uint64_t key = sk[j];
        const unsigned d = static_cast<unsigned>(key >> shift) & 255u;
        const unsigned mask = __match_any_sync(am, d);
        const unsigned leader = __ffs(mask) - 1;
        const unsigned off = whist[w * kHP + d];
        const unsigned first = S.dig[d];
        __syncwarp(am);
        if (lane == leader) whist[w * kHP + d] = off + __popc(mask);
        __syncwarp(am);
        const unsigned dst = first + off + __popc(mask & ((1u << lane) - 1u));
        const unsigned owner = dst >> logE, loc = dst & (E - 1u);
        *cl.map_shared_rank(dk + loc, owner) = key;
        if (WIDE) *cl.map_shared_rank(di + loc, owner) = si[j];
      }
    }
    sel ^= 1u;
    if (p == 0) vox_stamp(P, 7);
    cl.sync();
    if (p == 0) vox_stamp(P, 8);
  }
  vox_stamp(P, 9);
  cl.sync();  // (also when no pass ran: the parked keys of every CTA have been consumed)

  // sorted (key, index) out; run heads flagged at their SOURCE position, in the owner's shared memory
  const uint64_t* sk = S.bufk + sel * E;
  const uint16_t* si = S.bufi + sel * E;
  unsigned* flag = reinterpret_cast<unsigned*>(S.bufk + (sel ^ 1u) * E);  // [E], the retired key buffer
  for (unsigned l = t; l < cnt; l += kTC) {
    const unsigned j = gbeg + l;
    const uint64_t k = sk[l];
    const uint64_t m = WIDE ? k : (k >> 16);
    const unsigned id = WIDE ? si[l] : (static_cast<unsigned>(k) & 0xffffu);
    bool head = j == 0;
    if (!head) {
      const uint64_t kp = l > 0 ? sk[l - 1] : *cl.map_shared_rank(sk + (E - 1u), rank - 1u);
      head = (WIDE ? kp : (kp >> 16)) != m;
    }
    P.key[0][j] = m;
    P.idx[0][j] = id;
    *cl.map_shared_rank(flag + (id & (E - 1u)), id >> logE) = head ? j + 1u : 0u;
  }
  vox_stamp(P, 10);
  cl.sync();
  vox_stamp(P, 11);
  // kept points (run heads) listed in ascending source index: thread t owns the `per` consecutive
  // source positions from t * per on
  const unsigned per = E / kTC;  // 1, 2, 4 or 8
  unsigned f[8];
  unsigned c = 0;
#pragma unroll
  for (unsigned k = 0; k < 8; ++k) {
    const unsigned l = t * per + k;
    f[k] = (k < per && l < cnt) ? flag[l] : 0u;
    c += f[k] != 0u ? 1u : 0u;
  }
  unsigned tot;
  unsigned at = block_exclusive_scan2<kWC>(c, S.tmp, &tot);
  if (t == 0) *S.cnt = tot;
  cl.sync();
  vox_stamp(P, 12);
  unsigned below = 0, all = 0;
  if (t < 32u) {
    const unsigned v = t < CS ? *cl.map_shared_rank(S.cnt, t) : 0u;
    below = warp_reduce_add(t < rank ? v : 0u);
    all = warp_reduce_add(v);
    if (t == 0) {
      S.tmp[0] = below;
      S.tmp[1] = all;
    }
  }
  __syncthreads();
  at += S.tmp[0];
  all = S.tmp[1];
#pragma unroll
  for (unsigned k = 0; k < 8; ++k)
    if (f[k] != 0u) {
      P.kept_src[at] = gbeg + t * per + k;
      P.kept_pos[at] = f[k] - 1u;
      ++at;
    }
  vox_stamp(P, 13);
  if (rank == 0 && t == 0) {
    P.hdr->sel = 0;
    P.hdr->n_out = all;
    if (P.a.mail != nullptr) {  // the host is polling for the kept-point count
      volatile HostMail* m = P.a.mail;
      m->vox_n_out = all;
      m->vox_error = ld_cg(&P.hdr->error);
      m->vox_gb_error = 0u;
      __threadfence_system();
      m->vox_seq = P.a.mail_seq;
    }
  }
}

__global__ void __launch_bounds__(kTC) voxelize_cluster_kernel(VoxParams P) {
  extern __shared__ __align__(16) unsigned char vox_smem[];
  cg::cluster_group cl = cg::this_cluster();
  const unsigned CS = gridDim.x, rank = blockIdx.x, t = threadIdx.x, lane = t & 31;
  const unsigned n = P.a.n, E = P.chunk;
  const unsigned gbeg = min(n, rank * E), cnt = min(n, gbeg + E) - gbeg;
  ClusterSmem S;
  S.bufk = reinterpret_cast<uint64_t*>(vox_smem);
  S.bufi = reinterpret_cast<uint16_t*>(S.bufk + 2 * E);
  S.whist = reinterpret_cast<unsigned*>(S.bufi + 2 * E);
  S.chist = S.whist + kWC * kHP;
  S.dig = S.chist + 512;
  S.seg = S.dig + 256;
  S.mm = reinterpret_cast<int*>(S.seg + 2 * kSegSmem);
  S.cnt = reinterpret_cast<unsigned*>(S.mm + 16);
  S.tmp = S.cnt + 8;
  uint64_t* park = S.bufk + E;  // absolute voxel keys until the batch minimum is known

  vox_stamp(P, 0);
  // transform (+deskew) + voxel coordinate + min/max
  const int n_segs = P.a.n_segs;
  const bool seg_smem = n_segs <= kSegSmem;
  if (t < 6u) S.mm[t] = INT_MAX;
  if (seg_smem)
    for (int k = static_cast<int>(t); k < n_segs; k += kTC) {
      S.seg[k] = P.a.segs[k].begin;
      S.seg[kSegSmem + k] = P.a.segs[k].end;
    }
  __syncthreads();
  {
    int mn0 = INT_MAX, mn1 = INT_MAX, mn2 = INT_MAX;
    int nm0 = INT_MAX, nm1 = INT_MAX, nm2 = INT_MAX;
    bool bad = false;
    const int stride = P.a.in_stride;
    const double inv_voxel = 1.0 / P.a.voxel;
    constexpr int U0 = 2;  // elements per thread and trip: their loads are issued together
    for (unsigned l0 = t; l0 < cnt; l0 += U0 * kTC) {
      double x[U0], y[U0], z[U0];
#pragma unroll
      for (int u = 0; u < U0; ++u) {
        const unsigned l = l0 + u * kTC;
        if (l < cnt) {
          const size_t i = static_cast<size_t>(gbeg + l) * stride;
          x[u] = P.a.in_x[i];
          y[u] = P.a.in_y[i];
          z[u] = P.a.in_z[i];
        }
      }
#pragma unroll
      for (int u = 0; u < U0; ++u) {
        const unsigned l = l0 + u * kTC;
        if (l >= cnt) continue;
        const unsigned i = gbeg + l;
        if (P.a.has_T1) transform_point_rn(P.a.T1, x[u], y[u], z[u]);
        if (n_segs > 0) {
          const unsigned oi = P.a.orig != nullptr ? P.a.orig[i] : i;  // (segments index the uncropped sweep)
          int lo = 0, hi = n_segs;  // first segment with end > oi
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const unsigned e = seg_smem ? S.seg[kSegSmem + mid] : P.a.segs[mid].end;
            if (e > oi) hi = mid; else lo = mid + 1;
          }
          if (lo < n_segs && oi >= (seg_smem ? S.seg[lo] : P.a.segs[lo].begin))
            transform_point_rn(P.a.segs[lo].T, x[u], y[u], z[u]);
        }
        P.a.out_x[i] = x[u];
        P.a.out_y[i] = y[u];
        P.a.out_z[i] = z[u];
        const int kx = voxel_coord(x[u], P.a.voxel, inv_voxel);  // (== the division, common.cuh)
        const int ky = voxel_coord(y[u], P.a.voxel, inv_voxel);
        const int kz = voxel_coord(z[u], P.a.voxel, inv_voxel);
        if (!(coord_in_range(kx) && coord_in_range(ky) && coord_in_range(kz))) bad = true;
        park[l] = pack_key(kx, ky, kz);
        mn0 = min(mn0, kx); mn1 = min(mn1, ky); mn2 = min(mn2, kz);
        nm0 = min(nm0, -kx); nm1 = min(nm1, -ky); nm2 = min(nm2, -kz);
      }
    }
    mn0 = warp_reduce_min(mn0); mn1 = warp_reduce_min(mn1); mn2 = warp_reduce_min(mn2);
    nm0 = warp_reduce_min(nm0); nm1 = warp_reduce_min(nm1); nm2 = warp_reduce_min(nm2);
    if (lane == 0 && mn0 != INT_MAX) {
      atomicMin(&S.mm[0], mn0); atomicMin(&S.mm[1], mn1); atomicMin(&S.mm[2], mn2);
      atomicMin(&S.mm[3], nm0); atomicMin(&S.mm[4], nm1); atomicMin(&S.mm[5], nm2);
    }
    if (bad) atomicOr(&P.hdr->error, 1u);
  }
  vox_stamp(P, 1);
  cl.sync();
  vox_stamp(P, 2);
  if (t < 6u) {
    int v[16];
#pragma unroll
    for (unsigned r = 0; r < 16; ++r) v[r] = r < CS ? *cl.map_shared_rank(S.mm + t, r) : INT_MAX;
    int m = INT_MAX;
#pragma unroll
    for (unsigned r = 0; r < 16; ++r) m = min(m, v[r]);
    S.mm[8 + t] = m;
    if (rank == 0) (t < 3u ? P.hdr->mn[t] : P.hdr->nmx[t - 3u]) = m;
  }
  __syncthreads();
  const int m0 = S.mm[8], m1 = S.mm[9], m2 = S.mm[10];
  unsigned bits;
  {
    const int e0 = -S.mm[11] - m0, e1 = -S.mm[12] - m1, e2 = -S.mm[13] - m2;
    const unsigned ex = static_cast<unsigned>(max(max(e0, e1), max(e2, 0)));
    bits = ex == 0 ? 0u : (32u - static_cast<unsigned>(__clz(ex)));
    if (bits > kKeyBits) bits = kKeyBits;
  }
  if (rank == 0 && t == 0) P.hdr->bits = bits;
  const bool wide = 3u * bits > 48u;
  for (unsigned l = t; l < cnt; l += kTC) {
    int kx, ky, kz;
    unpack_key(park[l], kx, ky, kz);
    const uint32_t ux = static_cast<uint32_t>(kx - m0), uy = static_cast<uint32_t>(ky - m1),
                   uz = static_cast<uint32_t>(kz - m2);
    const uint64_t m = bits <= 10u ? static_cast<uint64_t>((spread3_10(ux) << 2) | (spread3_10(uy) << 1) | spread3_10(uz))
                                   : morton3(ux, uy, uz);
    if (wide) {
      S.bufk[l] = m;
      S.bufi[l] = static_cast<uint16_t>(gbeg + l);
    } else {
      S.bufk[l] = (m << 16) | static_cast<uint64_t>(gbeg + l);
    }
  }
  __syncthreads();
  vox_stamp(P, 3);
  if (wide) cluster_sort_and_list<true>(P, cl, S, bits, gbeg, cnt);
  else cluster_sort_and_list<false>(P, cl, S, bits, gbeg, cnt);
  cl.sync();  // no CTA leaves while a peer may still read its shared memory
  vox_stamp(P, 15);
}

// k-NN block-range tables (common.cuh): one thread per sorted element; the
// first / last element of every occupied block of level L records the block's
// start / end.  A separate launch with n threads: inside the persistent kernel
// (4 elements per thread, up to 6 dependent L2 atomics each) it cost ~35 us.
struct GatherArgs {  // the one-cluster kernel leaves the positions in sorted order to this kernel's n threads
  const uint32_t* idx;  // nullptr: nothing to gather
  const double* px;
  const double* py;
  const double* pz;
  double* sx;
  double* sy;
  double* sz;
};
__global__ void __launch_bounds__(256) knn_levels_kernel(const uint64_t* key0, const uint64_t* key1,
                                                         const VoxelHeader* hdr, uint4* levels,
                                                         unsigned level_stride, unsigned n, GatherArgs ga) {
  const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  if (ga.idx != nullptr) {
    const uint32_t id = __ldg(ga.idx + j);
    ga.sx[j] = __ldg(ga.px + id);
    ga.sy[j] = __ldg(ga.py + id);
    ga.sz[j] = __ldg(ga.pz + id);
  }
  const uint64_t* sk = hdr->sel ? key1 : key0;
  const unsigned bits = hdr->bits;
  const uint64_t cur = __ldg(sk + j);
  const uint64_t dprev = j == 0 ? ~0ull : (__ldg(sk + j - 1) ^ cur);
  const uint64_t dnext = j + 1 == n ? ~0ull : (__ldg(sk + j + 1) ^ cur);
  // first claim attempt of every level this element heads / tails, all atomics in flight at once
  // (one L2 round trip instead of up to six dependent ones); collisions are resolved below
  bool head[kKnnHashLevels], tail[kKnnHashLevels];
  unsigned mask[kKnnHashLevels], slot[kKnnHashLevels];
  unsigned long long old[kKnnHashLevels];
#pragma unroll
  for (int L = 0; L < kKnnHashLevels; ++L) {
    head[L] = (dprev >> (3 * L)) != 0;
    tail[L] = (dnext >> (3 * L)) != 0;
    mask[L] = knn_level_slots(hdr->n_out, bits, L) - 1u;  // (occupied blocks of any level <= occupied voxels = kept points)
    const uint64_t bk = cur >> (3 * L);
    slot[L] = static_cast<unsigned>(hash_key(bk)) & mask[L];
    old[L] = 0ull;
    if (head[L] || tail[L]) {
      uint4* tab = levels + static_cast<size_t>(L) * level_stride;
      old[L] = atomicCAS(reinterpret_cast<unsigned long long*>(tab + slot[L]), ~0ull,
                         static_cast<unsigned long long>(bk));
    }
  }
#pragma unroll
  for (int L = 0; L < kKnnHashLevels; ++L) {
    if (!head[L] && !tail[L]) continue;
    uint4* tab = levels + static_cast<size_t>(L) * level_stride;
    const uint64_t bk = cur >> (3 * L);
    uint4* e = tab + slot[L];
    if (old[L] != ~0ull && old[L] != bk)  // somebody else's block sits there: keep probing
      e = knn_level_claim_from(tab, mask[L], bk, (slot[L] + 1u) & mask[L]);
    if (head[L]) e->z = j;
    if (tail[L]) e->w = j + 1u;
  }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

SortView sort_view(eskf_ctx* ctx, unsigned n) {
  SortView v;
  const size_t nn = align_up(static_cast<size_t>(n) + 1, 64);
  char* p = ctx->sortbuf.as<char>();
  v.key[0] = reinterpret_cast<uint64_t*>(p);
  v.key[1] = v.key[0] + nn;
  v.idx[0] = reinterpret_cast<uint32_t*>(v.key[1] + nn);
  v.idx[1] = v.idx[0] + nn;
  uint32_t* r = ctx->runs.as<uint32_t>();
  v.run_start = r;
  v.keep_flag = r + nn;
  v.kept_src = r + 2 * nn;
  v.kept_pos = r + 3 * nn;
  v.sx = ctx->sorted_xyz.as<double>();
  v.sy = v.sx + nn;
  v.sz = v.sx + 2 * nn;
  v.hdr = ctx->hdr.as<VoxelHeader>();
  v.levels = ctx->knn_levels.as<uint4>();
  v.level_stride = knn_level_slots(n, kKeyBits, 0);
  return v;
}

int voxelize(eskf_ctx* ctx, const VoxelizeArgs& a) {
  ESKF_REQUIRE(a.n > 0, "voxelize: empty input");
  ESKF_REQUIRE(a.voxel > 0.0, "voxel_size must be positive");
  const unsigned n = a.n;
  const size_t nn = align_up(static_cast<size_t>(n) + 1, 64);
  ESKF_TRY(ctx->sortbuf.ensure(nn * (8 + 8 + 4 + 4)));
  ESKF_TRY(ctx->runs.ensure(nn * 4 * 4));
  if (a.mode == 1) {
    ESKF_TRY(ctx->sorted_xyz.ensure(nn * 3 * 8));
    const void* before = ctx->knn_levels.p;
    ESKF_TRY(ctx->knn_levels.ensure(static_cast<size_t>(kKnnHashLevels) * knn_level_slots(n, kKeyBits, 0) *
                                    sizeof(uint4)));
    if (ctx->knn_levels.p != before) ctx->knn_levels_clean = false;
    if (!ctx->knn_levels_clean)  // all-ones = empty entries
      ESKF_CUDA(cudaMemsetAsync(ctx->knn_levels.p, 0xFF, ctx->knn_levels.bytes, ctx->stream));
    ctx->knn_levels_clean = false;  // (the caller says so again once the tables are emptied behind the search)
  }
  ESKF_TRY(ctx->hdr.ensure(sizeof(VoxelHeader)));

  // elements per CTA: small enough that a 64k-point scan spreads over ~64 SMs
  // (the phases are latency-bound), large enough to keep the barriers cheap
  static const unsigned epb = [] {
    const char* e = getenv("ESKF_VOX_EPB");  // tuning knob
    const int v = e ? atoi(e) : 0;
    return v >= 256 ? static_cast<unsigned>(v) : 4u * kT;
  }();
  int G = static_cast<int>((n + epb - 1) / epb);
  if (G > ctx->max_blocks_voxelize) G = ctx->max_blocks_voxelize;
  if (G < 1) G = 1;
  // a sweep goes through ONE cluster: power-of-two CTA count and elements per CTA
  unsigned cl_cap = static_cast<unsigned>(ctx->vox_cluster_max);
  if (ctx->opt_vox_cluster == 0) cl_cap = 0;
  else if (ctx->opt_vox_cluster > 1) cl_cap = std::min(cl_cap, static_cast<unsigned>(ctx->opt_vox_cluster));
  const bool cluster = a.mode == 1 && cl_cap > 0 && n <= kClusterMaxPoints && a.cov == nullptr &&
                       n <= static_cast<size_t>(cl_cap) * kClusterMaxChunk;
  unsigned E = 1024;
  if (cluster) {
    while (static_cast<size_t>(E) * cl_cap < n) E *= 2;
    G = 1;
    while (static_cast<size_t>(G) * E < n) G *= 2;
  }
  ESKF_TRY(ctx->hist.ensure(static_cast<size_t>(G) * (2 * 256 + 2) * sizeof(unsigned)));

  VoxParams P;
  P.a = a;
  SortView v = sort_view(ctx, n);
  P.key[0] = v.key[0];
  P.key[1] = v.key[1];
  P.idx[0] = v.idx[0];
  P.idx[1] = v.idx[1];
  P.hist = ctx->hist.as<unsigned>();
  P.blk_count = P.hist + static_cast<size_t>(G) * 2 * 256;
  P.hdr = v.hdr;
  P.run_start = v.run_start;
  P.keep_flag = v.keep_flag;
  P.kept_src = v.kept_src;
  P.kept_pos = v.kept_pos;
  P.sx = v.sx;
  P.sy = v.sy;
  P.sz = v.sz;
  P.levels = v.levels;
  P.level_stride = v.level_stride;
  P.chunk = cluster ? E : static_cast<unsigned>(align_up((n + G - 1) / G, kT));
  P.stamps = nullptr;
  if (cluster && ctx->opt_trace) {
    ESKF_TRY(ctx->vox_stamps.ensure(16 * sizeof(unsigned long long)));
    P.stamps = ctx->vox_stamps.as<unsigned long long>();
    ESKF_CUDA(cudaMemsetAsync(P.stamps, 0, 16 * sizeof(unsigned long long), ctx->stream));
  }

  // header: min/max words to 0x7F7F7F7F (+inf for in-range ints; the one-cluster kernel writes them itself), rest zero
  if (cluster) {
    ESKF_CUDA(cudaMemsetAsync(v.hdr, 0, sizeof(VoxelHeader), ctx->stream));
  } else {
    ESKF_CUDA(cudaMemsetAsync(v.hdr, 0x7F, offsetof(VoxelHeader, error), ctx->stream));
    ESKF_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(v.hdr) + offsetof(VoxelHeader, error), 0,
                              sizeof(VoxelHeader) - offsetof(VoxelHeader, error), ctx->stream));
  }
  void* args[] = {&P};
  trace_mark(ctx, "start");
  GatherArgs ga;
  std::memset(&ga, 0, sizeof ga);
  if (cluster) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(G);
    cfg.blockDim = dim3(kTC);
    cfg.dynamicSmemBytes = cluster_smem_bytes(E);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = static_cast<unsigned>(G);
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ESKF_CUDA(cudaLaunchKernelEx(&cfg, voxelize_cluster_kernel, P));
    if (P.stamps != nullptr) {
      unsigned long long h[16];
      if (cudaMemcpyAsync(h, P.stamps, sizeof h, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
          cudaStreamSynchronize(ctx->stream) == cudaSuccess) {
        std::fprintf(stderr, "[eskf trace] voxelize cluster (%d CTAs x %u), ns from the kernel's start:", G, E);
        for (int k = 1; k < 16; ++k)
          if (h[k] != 0) std::fprintf(stderr, " [%d] %llu", k, h[k] - h[0]);
        std::fprintf(stderr, "\n");
      }
    }
    ga.idx = v.idx[0];
    ga.px = a.out_x;
    ga.py = a.out_y;
    ga.pz = a.out_z;
    ga.sx = v.sx;
    ga.sy = v.sy;
    ga.sz = v.sz;
  } else {
    ESKF_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(voxelize_kernel), dim3(G), dim3(kT),
                                          args, 0, ctx->stream));
  }
  count_launch(ctx);
  trace_mark(ctx, "voxelize");
  if (a.mode == 1) {
    knn_levels_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(v.key[0], v.key[1], v.hdr, v.levels,
                                                                v.level_stride, n, ga);
    ESKF_CUDA(cudaGetLastError());
    count_launch(ctx);
    trace_mark(ctx, "knn_levels");
  }
  return ESKF_OK;
}

int voxelize_max_blocks(int sm_count, int* out) {
  int per_sm = 0;
  ESKF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, voxelize_kernel, kT, 0));
  if (per_sm > 4) per_sm = 4;
  *out = per_sm * sm_count;
  return ESKF_OK;
}

// largest cluster (CTAs of kTC threads with the shared memory of their share of 65536 points) the
// device can place for the one-cluster kernel: 16 (non-portable size, opt-in), else 8, else 0 = off
int voxelize_cluster_max(int* out) {
  *out = 0;
  if (cudaFuncSetAttribute(voxelize_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(cluster_smem_bytes(kClusterMaxChunk))) != cudaSuccess) {
    cudaGetLastError();
    return ESKF_OK;
  }
  const cudaError_t np = cudaFuncSetAttribute(voxelize_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (np != cudaSuccess) cudaGetLastError();
  for (unsigned cs = np == cudaSuccess ? 16u : 8u; cs >= 8u; cs /= 2u) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(cs);
    cfg.blockDim = dim3(kTC);
    cfg.dynamicSmemBytes = cluster_smem_bytes(kClusterMaxChunk);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, voxelize_cluster_kernel, &cfg) == cudaSuccess && nc >= 1) {
      *out = static_cast<int>(cs);
      return ESKF_OK;
    }
    cudaGetLastError();
  }
  return ESKF_OK;
}

}  // namespace eskf
