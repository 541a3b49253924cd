// local_map.cu — the voxel hash map of LocalMap (src/LocalMap.cpp,
// include/ESKF_LIO/LocalMap.hpp) as an open-addressing table in HBM.
//
// Layout: tags[n_slots] (2 B each, 0 = empty) is the array that is PROBED —
// small enough to live in the 126 MB L2; slots[n_slots] (64 B records: key,
// count, fp32 mean-relative-to-centre, fp32 covariance) is what the
// registration kernel gathers on a tag match; master[n_slots][12]
// keeps the fp64 running mean / covariance so that inserts reproduce
// Voxel::addPoint (LocalMap.hpp:79-87) bit for bit.  The raw per-voxel point
// list of the reference (LocalMap.hpp:67) is dropped: only save() and the GUI
// read it.
#include <algorithm>
#include <numeric>

#include "internal.h"

namespace eskf {

namespace {

__device__ __forceinline__ void write_slot_payload(VoxelSlot* s, uint32_t count, const double* mean,
                                                   const double* cov, int kx, int ky, int kz,
                                                   double voxel) {
  // centre exactly as needsPointRemoval forms it (src/LocalMap.cpp:151)
  const double cx = __dmul_rn(static_cast<double>(kx) + 0.5, voxel);
  const double cy = __dmul_rn(static_cast<double>(ky) + 0.5, voxel);
  const double cz = __dmul_rn(static_cast<double>(kz) + 0.5, voxel);
  s->count = count;
  s->pad0 = 0;
  s->mx = static_cast<float>(mean[0] - cx);
  s->my = static_cast<float>(mean[1] - cy);
  s->mz = static_cast<float>(mean[2] - cz);
  s->pad1 = 0.f;
  s->c00 = static_cast<float>(cov[0]);
  s->c01 = static_cast<float>(cov[1]);
  s->c02 = static_cast<float>(cov[2]);
  s->c11 = static_cast<float>(cov[4]);
  s->c12 = static_cast<float>(cov[5]);
  s->c22 = static_cast<float>(cov[8]);
  s->pad2 = 0.f;
  s->pad3 = 0.f;
}

constexpr uint32_t kNoSlot = 0xffffffffu;

// find the slot of `key`, claiming an empty one if absent.  Returns the slot
// index, or kNoSlot when the table is full.  *is_new tells which.
// Within one kernel every key is handled by exactly one thread (one thread per
// key run), so a tag match whose record key is not (yet) ours is another voxel.
__device__ __forceinline__ uint32_t find_or_claim(tag_t* tags, VoxelSlot* slots, uint32_t n_slots,
                                                  uint64_t key, bool* is_new) {
  const SlotAddr a = slot_addr(key, n_slots);
  uint32_t h = a.home;
  for (uint32_t probe = 0; probe < n_slots; ++probe) {
    tag_t t = *reinterpret_cast<volatile tag_t*>(tags + h);
    if (t == 0u) {
      t = atomicCAS(reinterpret_cast<unsigned short*>(tags + h), static_cast<unsigned short>(0), a.tag);
      if (t == 0u) {
        slots[h].key = key;
        *is_new = true;
        return h;
      }
    }
    if (t == a.tag && *reinterpret_cast<volatile uint64_t*>(&slots[h].key) == key) {
      *is_new = false;
      return h;
    }
    h = next_slot(h, n_slots);
  }
  return kNoSlot;
}

__device__ __forceinline__ uint32_t find_slot(const tag_t* tags, const VoxelSlot* slots,
                                              uint32_t n_slots, uint64_t key) {
  const SlotAddr a = slot_addr(key, n_slots);
  uint32_t h = a.home;
  for (uint32_t probe = 0; probe < n_slots; ++probe) {
    const tag_t t = tags[h];
    if (t == 0u) return kNoSlot;
    if (t == a.tag && slots[h].key == key) return h;
    h = next_slot(h, n_slots);
  }
  return kNoSlot;
}

__global__ void clear_slots_kernel(VoxelSlot* slots, tag_t* tags, uint64_t n) {
  const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  tags[i] = 0u;
  uint4* p = reinterpret_cast<uint4*>(slots + i);
  p[0] = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
  p[1] = make_uint4(0u, 0u, 0u, 0u);
  p[2] = make_uint4(0u, 0u, 0u, 0u);
  p[3] = make_uint4(0u, 0u, 0u, 0u);
}

// 16 slots per thread: 32 B of tags in, 16 B of filter out (+ the wrap-around pad)
__global__ void __launch_bounds__(256) build_filter_kernel(const tag_t* __restrict__ tags, uint64_t n_slots,
                                                           uint8_t* __restrict__ filt) {
  const uint64_t groups = n_slots / 16;  // n_slots is a multiple of 64
  for (uint64_t g = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; g < groups;
       g += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint4 a = __ldcs(reinterpret_cast<const uint4*>(tags) + 2 * g);
    const uint4 b = __ldcs(reinterpret_cast<const uint4*>(tags) + 2 * g + 1);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t v = 0;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const uint32_t t = (w[2 * k + (h >> 1)] >> (16 * (h & 1))) & 0xffffu;
        v |= (t != 0u ? filter_tag(t) : 0u) << (8 * h);
      }
      o[k] = v;
    }
    const uint4 out = make_uint4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<uint4*>(filt)[g] = out;
    if (g == 0) reinterpret_cast<uint4*>(filt + n_slots)[0] = out;  // kFilterPad = 16 entries
  }
}

// K5: one thread per key run.  The points of a run are folded in input-index
// order with the exact expression of Voxel::addPoint (LocalMap.hpp:79-87):
//   mean = (n * mean + p) / (n + 1) ; cov likewise ; hard cap on n.
struct InsertParams {
  tag_t* tags;
  VoxelSlot* slots;
  double* master;
  uint32_t n_slots;
  unsigned long long* d_count;
  const uint64_t* key[2];
  const uint32_t* idx[2];
  const uint32_t* run_start;
  const VoxelHeader* hdr;
  const double* x;
  const double* y;
  const double* z;
  const double* cov;
  size_t cov_pitch;
  double voxel;
  uint32_t cap_pts;
};

__global__ void __launch_bounds__(128) insert_runs_kernel(InsertParams P) {
  const unsigned n_runs = P.hdr->n_out;
  const unsigned sel = P.hdr->sel;
  const uint64_t* keys = P.key[sel];
  const uint32_t* idx = P.idx[sel];
  const int m0 = P.hdr->mn[0], m1 = P.hdr->mn[1], m2 = P.hdr->mn[2];
  for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += gridDim.x * blockDim.x) {
    const unsigned j0 = P.run_start[r], j1 = P.run_start[r + 1];
    const uint64_t mk = keys[j0];
    const int kx = static_cast<int>(compact3(mk >> 2)) + m0;
    const int ky = static_cast<int>(compact3(mk >> 1)) + m1;
    const int kz = static_cast<int>(compact3(mk)) + m2;
    if (!(coord_in_range(kx) && coord_in_range(ky) && coord_in_range(kz))) {
      // outside the 21-bit key range: pack_key would alias a real voxel.  Dropped and counted, like
      // the sort-free path does (eskf_map_size reports ESKF_ERR_RANGE)
      atomicAdd(P.d_count + 3, static_cast<unsigned long long>(j1 - j0));
      continue;
    }
    bool is_new;
    const uint32_t s = find_or_claim(P.tags, P.slots, P.n_slots, pack_key(kx, ky, kz), &is_new);
    if (s == kNoSlot) {
      atomicAdd(P.d_count + 1, 1ull);  // table full: reported by the host
      continue;
    }
    double* M = P.master + static_cast<size_t>(s) * kMasterStride;
    double mean[3], C[9];
    uint32_t cnt = 0;
    if (is_new) {
      atomicAdd(P.d_count, 1ull);
    } else {
      cnt = P.slots[s].count;
#pragma unroll
      for (int k = 0; k < 3; ++k) mean[k] = M[k];
#pragma unroll
      for (int k = 0; k < 9; ++k) C[k] = M[3 + k];
    }
    const uint32_t before = cnt;
    for (unsigned j = j0; j < j1 && cnt < P.cap_pts; ++j) {
      const uint32_t i = idx[j];
      const double px = P.x[i], py = P.y[i], pz = P.z[i];
      if (cnt == 0) {
        mean[0] = px; mean[1] = py; mean[2] = pz;
#pragma unroll
        for (int k = 0; k < 9; ++k) C[k] = P.cov[k * P.cov_pitch + i];
      } else {
        const double nn = static_cast<double>(cnt), n1 = static_cast<double>(cnt + 1);
        mean[0] = __ddiv_rn(__dadd_rn(__dmul_rn(nn, mean[0]), px), n1);
        mean[1] = __ddiv_rn(__dadd_rn(__dmul_rn(nn, mean[1]), py), n1);
        mean[2] = __ddiv_rn(__dadd_rn(__dmul_rn(nn, mean[2]), pz), n1);
#pragma unroll
        for (int k = 0; k < 9; ++k)
          C[k] = __ddiv_rn(__dadd_rn(__dmul_rn(nn, C[k]), P.cov[k * P.cov_pitch + i]), n1);
      }
      ++cnt;
    }
    if (cnt != before) {
#pragma unroll
      for (int k = 0; k < 3; ++k) M[k] = mean[k];
#pragma unroll
      for (int k = 0; k < 9; ++k) M[3 + k] = C[k];
      write_slot_payload(&P.slots[s], cnt, mean, C, kx, ky, kz, P.voxel);
    }
  }
}

// K5b: the per-frame insert (a downsampled sweep: tens of thousands of points,
// one to a few per map voxel) without the radix sort.  link: one thread per
// point transforms it (and its covariance) to the world frame, finds or claims
// its voxel's slot and pushes itself on the slot's pending list (atomicExch on
// head[slot]).  fold: the first arrival of every list (next == empty) collects
// the list, orders it by point index and folds the points exactly like K5 —
// same input-index order, same expression, so both paths give bit-identical
// maps.  Two small launches (~15 us) instead of the persistent sort kernel
// (~50 us for 20k points: 8+ grid barriers) + K5.
constexpr uint32_t kListEnd = 0xffffffffu;
constexpr int kListLocal = 32;
constexpr size_t kListInsertMax = 131072;  // larger batches amortise the sort and may hold long lists

struct ListInsertParams {
  tag_t* tags;
  VoxelSlot* slots;
  double* master;
  uint32_t n_slots;
  unsigned long long* d_count;  // [0] voxels, [1] table-full, [3] key-range errors
  uint32_t* head;               // [n_slots] pending-list heads, all kListEnd between batches
  uint32_t* next;               // [n]
  uint32_t* slot_of;            // [n]
  double* x;
  double* y;
  double* z;
  double* cov;
  size_t cov_pitch;
  unsigned n;
  double voxel;
  uint32_t cap_pts;
  double T[12];
};

// find-or-claim that tolerates several threads looking for the SAME key at once:
// a tag match whose record key is still empty belongs to a claimer that is about
// to publish it, so wait for the key instead of walking on.
__device__ __forceinline__ uint32_t find_or_claim_shared(tag_t* tags, VoxelSlot* slots, uint32_t n_slots,
                                                         uint64_t key) {
  const SlotAddr a = slot_addr(key, n_slots);
  uint32_t h = a.home;
  for (uint32_t probe = 0; probe < n_slots; ++probe) {
    tag_t t = *reinterpret_cast<volatile tag_t*>(tags + h);
    if (t == 0u) {
      t = atomicCAS(reinterpret_cast<unsigned short*>(tags + h), static_cast<unsigned short>(0), a.tag);
      if (t == 0u) {
        *reinterpret_cast<volatile uint64_t*>(&slots[h].key) = key;
        __threadfence();
        return h;
      }
    }
    if (t == a.tag) {
      uint64_t k = *reinterpret_cast<volatile uint64_t*>(&slots[h].key);
      for (unsigned spins = 0; k == kEmptyKey && spins < kSpinLimit; ++spins) {
        __nanosleep(20);
        k = *reinterpret_cast<volatile uint64_t*>(&slots[h].key);
      }
      if (k == key) return h;
    }
    h = next_slot(h, n_slots);
  }
  return kNoSlot;
}

__global__ void __launch_bounds__(128) link_points_kernel(ListInsertParams P) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  double x = P.x[i], y = P.y[i], z = P.z[i];
  transform_point_rn(P.T, x, y, z);  // LocalMap.cpp:15
  P.x[i] = x;
  P.y[i] = y;
  P.z[i] = z;
  double C[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) C[k] = P.cov[k * P.cov_pitch + i];
  rotate_cov_rn(P.T, C);
#pragma unroll
  for (int k = 0; k < 9; ++k) P.cov[k * P.cov_pitch + i] = C[k];
  const int kx = voxel_coord(x, P.voxel), ky = voxel_coord(y, P.voxel), kz = voxel_coord(z, P.voxel);
  P.slot_of[i] = kNoSlot;
  if (!(coord_in_range(kx) && coord_in_range(ky) && coord_in_range(kz))) {
    atomicAdd(P.d_count + 3, 1ull);
    return;
  }
  const uint32_t s = find_or_claim_shared(P.tags, P.slots, P.n_slots, pack_key(kx, ky, kz));
  if (s == kNoSlot) {
    atomicAdd(P.d_count + 1, 1ull);
    return;
  }
  P.next[i] = atomicExch(P.head + s, i);
  P.slot_of[i] = s;
}

__global__ void __launch_bounds__(128) fold_lists_kernel(ListInsertParams P) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  const uint32_t s = P.slot_of[i];
  if (s == kNoSlot || P.next[i] != kListEnd) return;  // only the first arrival of a list folds it
  // everything that only needs the slot index is requested at once (list head, record key +
  // count, fp64 master statistics): one HBM round trip instead of three dependent ones
  const uint32_t first = ld_cg(P.head + s);
  const uint4 rec0 = ld_cg(reinterpret_cast<const uint4*>(P.slots + s));  // key (8) count (4) pad (4)
  double* M = P.master + static_cast<size_t>(s) * kMasterStride;
  double mean[3], C[9];
#pragma unroll
  for (int k = 0; k < 3; ++k) mean[k] = ld_cg(M + k);
#pragma unroll
  for (int k = 0; k < 9; ++k) C[k] = ld_cg(M + 3 + k);
  P.head[s] = kListEnd;  // ready for the next batch
  uint32_t idx[kListLocal];
  unsigned L = 1;
  idx[0] = i;
  if (first != i) {  // more than this point in the voxel: collect the list
    L = 0;
    for (uint32_t h = first; h != kListEnd; h = P.next[h]) {
      if (L < kListLocal) idx[L] = h;
      ++L;
    }
    if (L <= kListLocal) {  // insertion sort, ascending point index (= the reference's input order)
      for (unsigned a = 1; a < L; ++a) {
        const uint32_t v = idx[a];
        unsigned b = a;
        while (b > 0 && idx[b - 1] > v) {
          idx[b] = idx[b - 1];
          --b;
        }
        idx[b] = v;
      }
    }
  }
  int kx, ky, kz;
  unpack_key((static_cast<uint64_t>(rec0.y) << 32) | rec0.x, kx, ky, kz);
  uint32_t cnt = rec0.z;
  if (cnt == 0) atomicAdd(P.d_count, 1ull);  // a voxel claimed by this batch (mean / C are overwritten below)
  const uint32_t before = cnt;
  uint32_t last = 0;
  for (unsigned j = 0; j < L && cnt < P.cap_pts; ++j) {
    uint32_t p;
    if (L <= kListLocal) {
      p = idx[j];
    } else {  // long list: next larger index by walking it again (rare: many points of ONE batch in one voxel)
      p = kListEnd;
      for (uint32_t h = first; h != kListEnd; h = P.next[h])
        if ((j == 0 || h > last) && h < p) p = h;
      last = p;
    }
    const double px = P.x[p], py = P.y[p], pz = P.z[p];
    if (cnt == 0) {
      mean[0] = px; mean[1] = py; mean[2] = pz;
#pragma unroll
      for (int k = 0; k < 9; ++k) C[k] = P.cov[k * P.cov_pitch + p];
    } else {
      const double nn = static_cast<double>(cnt), n1 = static_cast<double>(cnt + 1);
      mean[0] = __ddiv_rn(__dadd_rn(__dmul_rn(nn, mean[0]), px), n1);
      mean[1] = __ddiv_rn(__dadd_rn(__dmul_rn(nn, mean[1]), py), n1);
      mean[2] = __ddiv_rn(__dadd_rn(__dmul_rn(nn, mean[2]), pz), n1);
#pragma unroll
      for (int k = 0; k < 9; ++k)
        C[k] = __ddiv_rn(__dadd_rn(__dmul_rn(nn, C[k]), P.cov[k * P.cov_pitch + p]), n1);
    }
    ++cnt;
  }
  if (cnt != before) {
#pragma unroll
    for (int k = 0; k < 3; ++k) M[k] = mean[k];
#pragma unroll
    for (int k = 0; k < 9; ++k) M[3 + k] = C[k];
    write_slot_payload(&P.slots[s], cnt, mean, C, kx, ky, kz, P.voxel);
  }
}

// K6: rehash every surviving voxel of the old table into a fresh one.  With
// evict != 0 a voxel survives iff NOT needsPointRemoval (src/LocalMap.cpp:149-154):
//   |(k + 0.5) * voxel - pos| > dist_thresh  ->  erased.
struct RehashParams {
  const tag_t* old_tags;
  const VoxelSlot* old_slots;
  const double* old_master;
  uint64_t old_n;
  tag_t* tags;
  VoxelSlot* slots;
  double* master;
  uint32_t n_slots;
  unsigned long long* d_count;  // [0] survivors, [1] table-full errors, [2] removed
  int evict;
  double pos[3];
  double dist_thresh;
  double voxel;
};

__global__ void __launch_bounds__(256) rehash_kernel(RehashParams P) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < P.old_n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    if (P.old_tags[i] == 0u) continue;
    const VoxelSlot* o = P.old_slots + i;
    const uint64_t key = o->key;
    if (P.evict) {
      int kx, ky, kz;
      unpack_key(key, kx, ky, kz);
      const double dx = __dadd_rn(__dmul_rn(static_cast<double>(kx) + 0.5, P.voxel), -P.pos[0]);
      const double dy = __dadd_rn(__dmul_rn(static_cast<double>(ky) + 0.5, P.voxel), -P.pos[1]);
      const double dz = __dadd_rn(__dmul_rn(static_cast<double>(kz) + 0.5, P.voxel), -P.pos[2]);
      const double dist = sqrt(dot3_rn(dx, dx, dy, dy, dz, dz));
      if (dist > P.dist_thresh) {
        atomicAdd(P.d_count + 2, 1ull);
        continue;
      }
    }
    bool is_new;
    const uint32_t s = find_or_claim(P.tags, P.slots, P.n_slots, key, &is_new);
    if (s == kNoSlot) {
      atomicAdd(P.d_count + 1, 1ull);
      continue;
    }
    atomicAdd(P.d_count, 1ull);
    const uint4* src = reinterpret_cast<const uint4*>(o);
    uint4* dst = reinterpret_cast<uint4*>(P.slots + s);
    uint4 v0 = src[0];
    dst[1] = src[1];
    dst[2] = src[2];
    dst[3] = src[3];
    // key words already claimed; write count + pad
    reinterpret_cast<uint2*>(dst)[1] = make_uint2(v0.z, v0.w);
    const double* om = P.old_master + i * kMasterStride;
    double* nm = P.master + static_cast<size_t>(s) * kMasterStride;
#pragma unroll
    for (int k = 0; k < kMasterStride; ++k) nm[k] = om[k];
  }
}

__global__ void query_kernel(const tag_t* tags, const VoxelSlot* slots, const double* master,
                             uint32_t n_slots, double voxel, const double* xyz_aos, unsigned n, int32_t* key_xyz,
                             uint8_t* hit, uint32_t* count, double* mean, double* cov) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int kx = voxel_coord(xyz_aos[3 * i], voxel);
  const int ky = voxel_coord(xyz_aos[3 * i + 1], voxel);
  const int kz = voxel_coord(xyz_aos[3 * i + 2], voxel);
  if (key_xyz) {
    key_xyz[3 * i] = kx;
    key_xyz[3 * i + 1] = ky;
    key_xyz[3 * i + 2] = kz;
  }
  uint32_t s = kNoSlot;
  if (coord_in_range(kx) && coord_in_range(ky) && coord_in_range(kz))
    s = find_slot(tags, slots, n_slots, pack_key(kx, ky, kz));
  const bool found = s != kNoSlot;
  if (hit) hit[i] = found ? 1 : 0;
  if (count) count[i] = found ? slots[s].count : 0u;
  if (mean)
    for (int k = 0; k < 3; ++k)
      mean[3 * i + k] = found ? master[static_cast<size_t>(s) * kMasterStride + k] : 0.0;
  if (cov)
    for (int k = 0; k < 9; ++k)
      cov[9 * i + k] = found ? master[static_cast<size_t>(s) * kMasterStride + 3 + k] : 0.0;
}

__global__ void export_kernel(const tag_t* tags, const VoxelSlot* slots, const double* master,
                              uint64_t n_slots,
                              unsigned long long* cursor, uint64_t capacity, uint64_t* keys,
                              uint32_t* count, double* stats) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n_slots;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    if (tags[i] == 0u) continue;
    const uint64_t key = slots[i].key;
    const unsigned long long o = atomicAdd(cursor, 1ull);
    if (o >= capacity) continue;
    keys[o] = key;
    count[o] = slots[i].count;
    for (int k = 0; k < kMasterStride; ++k) stats[o * kMasterStride + k] = master[i * kMasterStride + k];
  }
}

int alloc_table(eskf_ctx* ctx, uint64_t n_slots, tag_t** tags, VoxelSlot** slots, double** master) {
  if (n_slots >= (1ull << 32) - 2) {
    set_error("voxel table of %llu slots exceeds the 32-bit slot index", static_cast<unsigned long long>(n_slots));
    return ESKF_ERR_CAPACITY;
  }
  *tags = nullptr;
  *slots = nullptr;
  *master = nullptr;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(tags), n_slots * sizeof(tag_t));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(slots), n_slots * sizeof(VoxelSlot));
  if (e == cudaSuccess)
    e = cudaMalloc(reinterpret_cast<void**>(master), n_slots * kMasterStride * sizeof(double));
  if (e != cudaSuccess) {
    if (*tags) cudaFree(*tags);
    if (*slots) cudaFree(*slots);
    *tags = nullptr;
    *slots = nullptr;
    set_error("cudaMalloc(voxel table, %llu slots) failed: %s", static_cast<unsigned long long>(n_slots),
              cudaGetErrorString(e));
    return ESKF_ERR_CUDA;
  }
  const unsigned blocks = static_cast<unsigned>((n_slots + 255) / 256);
  clear_slots_kernel<<<blocks, 256, 0, ctx->stream>>>(*slots, *tags, n_slots);
  ESKF_CUDA(cudaGetLastError());
  count_launch(ctx);
  return ESKF_OK;
}

// table size for `voxels` occupied voxels: load factor <= 1/2 keeps linear
// probing short; no power-of-two rounding so the tag array stays as small
// (= as L2-resident) as possible
uint64_t table_size_for(uint64_t voxels) {
  const uint64_t v = voxels < 1024 ? 1024 : voxels;
  return (2 * v + 63) / 64 * 64;
}

// move the map into a table of new_slots slots (optionally evicting)
int rebuild(eskf_map* m, uint64_t new_slots, int evict, const double* pos, double thresh,
            uint64_t* removed) {
  eskf_ctx* ctx = m->ctx;
  tag_t* nt = nullptr;
  VoxelSlot* ns = nullptr;
  double* nm = nullptr;
  if (m->spare_n == new_slots && m->spare_tags) {
    nt = m->spare_tags;
    ns = m->spare_slots;
    nm = m->spare_master;
    m->spare_tags = nullptr;
    m->spare_slots = nullptr;
    m->spare_master = nullptr;
    m->spare_n = 0;
    const unsigned cb = static_cast<unsigned>((new_slots + 255) / 256);
    clear_slots_kernel<<<cb, 256, 0, ctx->stream>>>(ns, nt, new_slots);
    ESKF_CUDA(cudaGetLastError());
    count_launch(ctx);
  } else {
    ESKF_TRY(alloc_table(ctx, new_slots, &nt, &ns, &nm));
  }
  // [0] survivors and [2] removed are recounted; [1] table-full and [3] key-range errors are sticky
  // until a caller has seen them (eskf_map_size): a rebuild must not swallow a pending overflow
  ESKF_CUDA(cudaMemsetAsync(m->d_count, 0, sizeof(unsigned long long), ctx->stream));
  ESKF_CUDA(cudaMemsetAsync(m->d_count + 2, 0, sizeof(unsigned long long), ctx->stream));
  RehashParams P;
  P.old_tags = m->tags;
  P.old_slots = m->slots;
  P.old_master = m->master;
  P.old_n = m->n_slots;
  P.tags = nt;
  P.slots = ns;
  P.master = nm;
  P.n_slots = static_cast<uint32_t>(new_slots);
  P.d_count = m->d_count;
  P.evict = evict;
  P.pos[0] = pos ? pos[0] : 0.0;
  P.pos[1] = pos ? pos[1] : 0.0;
  P.pos[2] = pos ? pos[2] : 0.0;
  P.dist_thresh = thresh;
  P.voxel = m->voxel;
  const unsigned blocks = static_cast<unsigned>(std::min<uint64_t>((m->n_slots + 255) / 256, 148ull * 16));
  rehash_kernel<<<blocks, 256, 0, ctx->stream>>>(P);
  ESKF_CUDA(cudaGetLastError());
  count_launch(ctx);
  unsigned long long* h = nullptr;
  ESKF_TRY(ctx_pinned(ctx, 4 * sizeof(unsigned long long), reinterpret_cast<void**>(&h)));
  ESKF_CUDA(cudaMemcpyAsync(h, m->d_count, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                            ctx->stream));
  ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (m->spare_tags) {  // a spare of another size is of no use any more
    cudaFree(m->spare_tags);
    cudaFree(m->spare_slots);
    cudaFree(m->spare_master);
    m->spare_tags = nullptr;
    m->spare_slots = nullptr;
    m->spare_master = nullptr;
    m->spare_n = 0;
  }
  if (m->n_slots == new_slots) {  // same-size sweep: keep the old table for the next one
    m->spare_tags = m->tags;
    m->spare_slots = m->slots;
    m->spare_master = m->master;
    m->spare_n = m->n_slots;
  } else {
    cudaFree(m->tags);
    cudaFree(m->slots);
    cudaFree(m->master);
  }
  m->tags = nt;
  m->slots = ns;
  m->master = nm;
  m->n_slots = new_slots;
  ++m->version;
  m->count_upper = h[0];
  if (removed) *removed = h[2];
  if (h[1] != 0) {  // (also an overflow of an earlier insert nobody has asked about: reported once)
    cudaMemsetAsync(m->d_count + 1, 0, sizeof(unsigned long long), ctx->stream);
    set_error("voxel table overflow (%llu runs dropped by an insert or by this rebuild)", h[1]);
    return ESKF_ERR_CAPACITY;
  }
  return ESKF_OK;
}

}  // namespace

int map_probe_filter(const eskf_map* m, const uint8_t** filt) {
  eskf_ctx* ctx = m->ctx;
  static_assert(kFilterPad == 16, "build_filter_kernel writes one 16-entry pad group");
  if (m->filt_slots != m->n_slots) {
    if (m->filt) {
      ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(m->filt);
    }
    m->filt = nullptr;
    m->filt_slots = 0;
    m->filt_version = 0;
    ESKF_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->filt), m->n_slots + kFilterPad));
    m->filt_slots = m->n_slots;
  }
  if (m->filt_version != m->version) {
    const uint64_t groups = m->n_slots / 16;
    const unsigned blocks = static_cast<unsigned>(std::min<uint64_t>((groups + 255) / 256, 148ull * 8));
    build_filter_kernel<<<blocks, 256, 0, ctx->stream>>>(m->tags, m->n_slots, m->filt);
    ESKF_CUDA(cudaGetLastError());
    count_launch(ctx);
    m->filt_version = m->version;
  }
  *filt = m->filt;
  return ESKF_OK;
}

// make room for up to `incoming` new voxels at load factor <= 1/2
int map_reserve(eskf_map* m, uint64_t incoming) {
  if ((m->count_upper + incoming) * 2 <= m->n_slots) return ESKF_OK;
  eskf_ctx* ctx = m->ctx;
  unsigned long long* h = nullptr;
  ESKF_TRY(ctx_pinned(ctx, sizeof(unsigned long long), reinterpret_cast<void**>(&h)));
  ESKF_CUDA(cudaMemcpyAsync(h, m->d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                            ctx->stream));
  ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
  m->count_upper = h[0];
  if ((m->count_upper + incoming) * 2 <= m->n_slots) return ESKF_OK;
  // grow geometrically: a rebuild allocates and frees device memory (milliseconds of host time),
  // so it must stay rare on a map that gains ~1k voxels per frame
  return rebuild(m, table_size_for(2 * (m->count_upper + incoming)), 0, nullptr, 0.0, nullptr);
}

}  // namespace eskf

using namespace eskf;

extern "C" {

int eskf_map_create(eskf_ctx* ctx, double voxel_size, uint32_t max_points_per_voxel,
                    uint64_t capacity_hint, eskf_map** out) {
  ESKF_REQUIRE(ctx && out, "null ctx/out");
  ESKF_REQUIRE(voxel_size > 0.0, "voxel_size must be positive");
  ESKF_REQUIRE(max_points_per_voxel > 0, "max_points_per_voxel must be positive");
  ESKF_CUDA(cudaSetDevice(ctx->device));
  eskf_map* m = new eskf_map();
  m->ctx = ctx;
  m->voxel = voxel_size;
  m->cap_pts = max_points_per_voxel;
  m->n_slots = table_size_for(capacity_hint);
  int st = alloc_table(ctx, m->n_slots, &m->tags, &m->slots, &m->master);
  if (st != ESKF_OK) {
    delete m;
    return st;
  }
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&m->d_count), 4 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemsetAsync(m->d_count, 0, 4 * sizeof(unsigned long long), ctx->stream);
  if (e != cudaSuccess) {
    set_error("map counters: %s", cudaGetErrorString(e));
    cudaFree(m->tags);
    cudaFree(m->slots);
    cudaFree(m->master);
    delete m;
    return ESKF_ERR_CUDA;
  }
  *out = m;
  return ESKF_OK;
}

int eskf_map_destroy(eskf_map* m) {
  if (!m) return ESKF_OK;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  cudaFree(m->tags);
  cudaFree(m->slots);
  cudaFree(m->master);
  if (m->spare_tags) {
    cudaFree(m->spare_tags);
    cudaFree(m->spare_slots);
    cudaFree(m->spare_master);
  }
  cudaFree(m->d_count);
  if (m->head) cudaFree(m->head);
  if (m->filt) cudaFree(m->filt);
  delete m;
  return ESKF_OK;
}

int eskf_map_insert_cloud(eskf_map* m, eskf_cloud* cloud, const double T[16]) {
  ESKF_REQUIRE(m && cloud && T, "null argument");
  ESKF_REQUIRE(cloud->ctx == m->ctx, "cloud and map belong to different contexts");
  ESKF_REQUIRE(cloud->has_cov, "map insert needs covariances");
  eskf_ctx* ctx = m->ctx;
  ESKF_CUDA(cudaSetDevice(ctx->device));
  if (cloud->n == 0) return ESKF_OK;
  ESKF_REQUIRE(cloud->n < (1ull << 31), "cloud too large");
  ESKF_TRY(map_reserve(m, cloud->n));
  ++m->version;  // (the probe filter of the dense registration is stale from here on)
  if (!ctx->opt_insert_sorted && cloud->n <= kListInsertMax) {
    // sort-free path: link every point to its voxel's pending list, fold the lists
    if (m->head_n != m->n_slots) {
      if (m->head) cudaFree(m->head);
      m->head = nullptr;
      m->head_n = 0;
      ESKF_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->head), m->n_slots * sizeof(uint32_t)));
      ESKF_CUDA(cudaMemsetAsync(m->head, 0xFF, m->n_slots * sizeof(uint32_t), ctx->stream));
      m->head_n = m->n_slots;
    }
    const unsigned n = static_cast<unsigned>(cloud->n);
    ESKF_TRY(ctx->link.ensure(static_cast<size_t>(n) * 2 * sizeof(uint32_t)));
    ListInsertParams L;
    L.tags = m->tags;
    L.slots = m->slots;
    L.master = m->master;
    L.n_slots = static_cast<uint32_t>(m->n_slots);
    L.d_count = m->d_count;
    L.head = m->head;
    L.next = ctx->link.as<uint32_t>();
    L.slot_of = L.next + n;
    L.x = cloud->x();
    L.y = cloud->y();
    L.z = cloud->z();
    L.cov = cloud->cov;
    L.cov_pitch = cloud->cap;
    L.n = n;
    L.voxel = m->voxel;
    L.cap_pts = m->cap_pts;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) L.T[3 * i + j] = T[4 * i + j];
      L.T[9 + i] = T[4 * i + 3];
    }
    trace_mark(ctx, "start");
    link_points_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(L);
    ESKF_CUDA(cudaGetLastError());
    trace_mark(ctx, "link");
    fold_lists_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(L);
    ESKF_CUDA(cudaGetLastError());
    trace_mark(ctx, "fold");
    trace_flush(ctx, "map_insert");
    count_launch(ctx, 2);
    cloud->has_c32 = false;  // fp32 mirror is stale after the in-place transform
    m->count_upper += cloud->n;
    return ESKF_OK;
  }
  VoxelizeArgs a;
  std::memset(&a, 0, sizeof a);
  a.in_x = cloud->x();
  a.in_y = cloud->y();
  a.in_z = cloud->z();
  a.in_stride = 1;
  a.out_x = cloud->x();
  a.out_y = cloud->y();
  a.out_z = cloud->z();
  a.cov = cloud->cov;
  a.cov_pitch = cloud->cap;
  a.n = static_cast<unsigned>(cloud->n);
  a.voxel = m->voxel;
  a.has_T1 = 1;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) a.T1[3 * i + j] = T[4 * i + j];
    a.T1[9 + i] = T[4 * i + 3];
  }
  a.mode = 0;
  ESKF_TRY(voxelize(ctx, a));
  cloud->has_c32 = false;  // fp32 mirror is stale after the in-place transform
  SortView v = sort_view(ctx, a.n);
  InsertParams P;
  P.tags = m->tags;
  P.slots = m->slots;
  P.master = m->master;
  P.n_slots = static_cast<uint32_t>(m->n_slots);
  P.d_count = m->d_count;
  P.key[0] = v.key[0];
  P.key[1] = v.key[1];
  P.idx[0] = v.idx[0];
  P.idx[1] = v.idx[1];
  P.run_start = v.run_start;
  P.hdr = v.hdr;
  P.x = cloud->x();
  P.y = cloud->y();
  P.z = cloud->z();
  P.cov = cloud->cov;
  P.cov_pitch = cloud->cap;
  P.voxel = m->voxel;
  P.cap_pts = m->cap_pts;
  const unsigned blocks = std::min<unsigned>((a.n + 127) / 128, 148u * 8);
  insert_runs_kernel<<<blocks, 128, 0, ctx->stream>>>(P);
  ESKF_CUDA(cudaGetLastError());
  count_launch(ctx);
  m->count_upper += cloud->n;
  return ESKF_OK;
}

int eskf_map_insert(eskf_map* m, const double* xyz, const double* cov, size_t n, const double T[16]) {
  ESKF_REQUIRE(m && T, "null argument");
  if (n == 0) return ESKF_OK;
  ESKF_REQUIRE(xyz && cov, "null xyz/cov");
  eskf_ctx* ctx = m->ctx;
  if (!ctx->tmp_cloud[0]) ESKF_TRY(eskf_cloud_create(ctx, n, &ctx->tmp_cloud[0]));
  ESKF_TRY(eskf_cloud_upload(ctx->tmp_cloud[0], xyz, cov, n));
  return eskf_map_insert_cloud(m, ctx->tmp_cloud[0], T);
}

int eskf_map_evict(eskf_map* m, const double pos[3], double dist_thresh, uint64_t* removed) {
  ESKF_REQUIRE(m && pos, "null argument");
  ESKF_CUDA(cudaSetDevice(m->ctx->device));
  return rebuild(m, m->n_slots, 1, pos, dist_thresh, removed);
}

int eskf_map_size(eskf_map* m, uint64_t* n_voxels) {
  ESKF_REQUIRE(m && n_voxels, "null argument");
  eskf_ctx* ctx = m->ctx;
  ESKF_CUDA(cudaSetDevice(ctx->device));
  unsigned long long* h = nullptr;
  ESKF_TRY(ctx_pinned(ctx, 4 * sizeof(unsigned long long), reinterpret_cast<void**>(&h)));
  ESKF_CUDA(cudaMemcpyAsync(h, m->d_count, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                            ctx->stream));
  ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
  *n_voxels = h[0];
  m->count_upper = h[0];
  if (h[1] != 0) {
    cudaMemsetAsync(m->d_count + 1, 0, sizeof(unsigned long long), ctx->stream);  // reported once
    set_error("voxel table overflowed (%llu runs dropped)", h[1]);
    return ESKF_ERR_CAPACITY;
  }
  if (h[3] != 0) {
    set_error("voxel coordinate outside the 21-bit key range (%llu points dropped)", h[3]);
    return ESKF_ERR_RANGE;
  }
  // also surface voxelize errors of the last insert
  if (ctx->hdr.p) {
    VoxelHeader* hh = nullptr;
    ESKF_TRY(ctx_pinned(ctx, sizeof(VoxelHeader), reinterpret_cast<void**>(&hh)));
    ESKF_CUDA(cudaMemcpyAsync(hh, ctx->hdr.p, sizeof(VoxelHeader), cudaMemcpyDeviceToHost, ctx->stream));
    ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hh->gb.error) {
      set_error("grid barrier timeout in voxelize kernel");
      return ESKF_ERR_INTERNAL;
    }
    if (hh->error & 1u) {
      set_error("voxel coordinate outside the 21-bit key range");
      return ESKF_ERR_RANGE;
    }
  }
  return ESKF_OK;
}

int eskf_map_capacity(eskf_map* m, uint64_t* n_slots) {
  ESKF_REQUIRE(m && n_slots, "null argument");
  *n_slots = m->n_slots;
  return ESKF_OK;
}

int eskf_map_compact(eskf_map* m) {
  ESKF_REQUIRE(m, "null map");
  uint64_t n = 0;
  ESKF_TRY(eskf_map_size(m, &n));
  const uint64_t want = table_size_for(n);
  if (want == m->n_slots) return ESKF_OK;
  return rebuild(m, want, 0, nullptr, 0.0, nullptr);
}

int eskf_map_query(eskf_map* m, const double* xyz, size_t n, int32_t* key_xyz, uint8_t* hit,
                   uint32_t* count, double* mean, double* cov) {
  ESKF_REQUIRE(m, "null map");
  if (n == 0) return ESKF_OK;
  ESKF_REQUIRE(xyz, "null xyz");
  eskf_ctx* ctx = m->ctx;
  ESKF_CUDA(cudaSetDevice(ctx->device));
  // staging: xyz | key | hit | count | mean | cov
  const size_t o_key = n * 24, o_hit = o_key + n * 12, o_cnt = (o_hit + n + 7) / 8 * 8;
  const size_t o_mean = o_cnt + n * 4 + (n % 2) * 4, o_cov = o_mean + n * 24, total = o_cov + n * 72;
  ESKF_TRY(ctx->stage.ensure(total));
  char* d = ctx->stage.as<char>();
  ESKF_CUDA(cudaMemcpyAsync(d, xyz, n * 24, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned blocks = static_cast<unsigned>((n + 127) / 128);
  query_kernel<<<blocks, 128, 0, ctx->stream>>>(
      m->tags, m->slots, m->master, static_cast<uint32_t>(m->n_slots), m->voxel,
      reinterpret_cast<const double*>(d),
      static_cast<unsigned>(n), reinterpret_cast<int32_t*>(d + o_key),
      reinterpret_cast<uint8_t*>(d + o_hit), reinterpret_cast<uint32_t*>(d + o_cnt),
      reinterpret_cast<double*>(d + o_mean), reinterpret_cast<double*>(d + o_cov));
  ESKF_CUDA(cudaGetLastError());
  count_launch(ctx);
  if (key_xyz) ESKF_CUDA(cudaMemcpyAsync(key_xyz, d + o_key, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
  if (hit) ESKF_CUDA(cudaMemcpyAsync(hit, d + o_hit, n, cudaMemcpyDeviceToHost, ctx->stream));
  if (count) ESKF_CUDA(cudaMemcpyAsync(count, d + o_cnt, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (mean) ESKF_CUDA(cudaMemcpyAsync(mean, d + o_mean, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
  if (cov) ESKF_CUDA(cudaMemcpyAsync(cov, d + o_cov, n * 72, cudaMemcpyDeviceToHost, ctx->stream));
  ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
  return ESKF_OK;
}

int eskf_map_export(eskf_map* m, size_t capacity, size_t* n, int32_t* key_xyz, uint32_t* count,
                    double* mean, double* cov) {
  ESKF_REQUIRE(m && n, "null argument");
  eskf_ctx* ctx = m->ctx;
  uint64_t nv = 0;
  ESKF_TRY(eskf_map_size(m, &nv));
  *n = nv;
  if (nv == 0) return ESKF_OK;
  if (capacity < nv) {
    set_error("export capacity %zu < %llu voxels", capacity, static_cast<unsigned long long>(nv));
    return ESKF_ERR_CAPACITY;
  }
  const size_t o_cnt = nv * 8, o_stats = (o_cnt + nv * 4 + 7) / 8 * 8, o_cur = o_stats + nv * 96;
  ESKF_TRY(ctx->stage.ensure(o_cur + 8));
  char* d = ctx->stage.as<char>();
  ESKF_CUDA(cudaMemsetAsync(d + o_cur, 0, 8, ctx->stream));
  const unsigned blocks = static_cast<unsigned>(std::min<uint64_t>((m->n_slots + 255) / 256, 148ull * 16));
  export_kernel<<<blocks, 256, 0, ctx->stream>>>(
      m->tags, m->slots, m->master, m->n_slots, reinterpret_cast<unsigned long long*>(d + o_cur), nv,
      reinterpret_cast<uint64_t*>(d), reinterpret_cast<uint32_t*>(d + o_cnt),
      reinterpret_cast<double*>(d + o_stats));
  ESKF_CUDA(cudaGetLastError());
  count_launch(ctx);
  std::vector<uint64_t> hk(nv);
  std::vector<uint32_t> hc(nv);
  std::vector<double> hs(nv * kMasterStride);
  ESKF_CUDA(cudaMemcpyAsync(hk.data(), d, nv * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ESKF_CUDA(cudaMemcpyAsync(hc.data(), d + o_cnt, nv * 4, cudaMemcpyDeviceToHost, ctx->stream));
  ESKF_CUDA(cudaMemcpyAsync(hs.data(), d + o_stats, nv * 96, cudaMemcpyDeviceToHost, ctx->stream));
  ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<size_t> order(nv);
  std::iota(order.begin(), order.end(), size_t{0});
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return hk[a] < hk[b]; });
  for (size_t o = 0; o < nv; ++o) {
    const size_t s = order[o];
    if (key_xyz) {
      int x, y, z;
      unpack_key(hk[s], x, y, z);
      key_xyz[3 * o] = x;
      key_xyz[3 * o + 1] = y;
      key_xyz[3 * o + 2] = z;
    }
    if (count) count[o] = hc[s];
    if (mean) std::memcpy(mean + 3 * o, hs.data() + s * kMasterStride, 3 * sizeof(double));
    if (cov) std::memcpy(cov + 9 * o, hs.data() + s * kMasterStride + 3, 9 * sizeof(double));
  }
  return ESKF_OK;
}

}  // extern "C"
