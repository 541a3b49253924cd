// registration.cu — K7/K8: ICP::align (src/Registration.cpp:7-35) as ONE
// persistent cooperative kernel: every Gauss-Newton iteration
//   transform (Registration.cpp:13,27)  ->  voxel lookup (LocalMap.cpp:93-100)
//   ->  per-point J^T W J / J^T W r (Registration.cpp:83-102)
//   ->  reduction (Registration.cpp:60-76)  ->  6x6 LDLT + se3ToSE3 (:78-79)
//   ->  compose + convergence test (:20-25)
// runs on the device with no host round trip; CTAs meet once per iteration
// through a "last CTA solves, everyone waits on an epoch word" hand-off.
//
// Numerics (BASELINE.json north_star): the position pipeline is fp64 with the
// reference's exact evaluation order (so voxel keys are bit-exact); the
// per-point 3x3 algebra is fp32 (fp64 on request); the 27 partial sums are
// accumulated in fp64 in a fixed order (deterministic for a given grid).
//
// HBM traffic per source point per iteration (1-neighbour): 24 B position
// read + 24 B position write + 24 B source covariance + 64 B voxel slot
// = 136 B.
#include <cfloat>
#include <cstdlib>
#include <memory>

#include "internal.h"

namespace eskf {

namespace {

constexpr int kT = 256;   // default CTA size (small clouds; the 7-neighbour / fp64 variants)
constexpr int kMaxT = 768;  // largest CTA size a variant may use (one CTA of 24 warps per SM)
constexpr int kAcc = 28;  // A(6) B(9) D(6) b(6) + correspondence count
constexpr int kMaxWorld = ESKF_MAX_WORLD;
constexpr int kMailStride = 64;  // 8-byte words per (ring slot, source rank): 28 sums + flag, or 56 flagged words
// measured on B200 (dense config, us per GN iteration at 0.1 m / 0.5 m voxels): FOLD 1 with
// 3 CTAs/SM 86.7 / 160.2; FOLD 4 or 8 need 2 CTAs/SM (the partial sums stay live across the
// load phase) 96 / 155; FOLD > 1 at 3 CTAs/SM spills: 192 / 255.
#ifndef ESKF_ALIGN_FOLD
#define ESKF_ALIGN_FOLD 1
#endif
constexpr int kFold = ESKF_ALIGN_FOLD;  // warp tiles (x 32 points) summed per lane in fp32 between reduce-scatters
#ifndef ESKF_TERMS_ASSIGN
#define ESKF_TERMS_ASSIGN 1
#endif
constexpr bool kAssign = kFold == 1 && ESKF_TERMS_ASSIGN != 0;

#ifndef ESKF_PIPELINED
#define ESKF_PIPELINED 1
#endif
#ifndef ESKF_COV_PREFETCH
#define ESKF_COV_PREFETCH 0
#endif
#ifndef ESKF_POS_PREFETCH
// 4-deep rotation: L2-prefetch the positions of the (statically dealt) tile this many trips ahead (0 = off).
// Round 2, 512-thread CTAs on the 8-bit filter: 0 -> 86.5, 4 -> 85.4 us per dense iteration.
#define ESKF_POS_PREFETCH 4
#endif
#ifndef ESKF_POS_AHEAD
#define ESKF_POS_AHEAD 0  // 4-deep rotation: 1 = keep two tiles of raw positions in flight instead of one
#endif
#ifndef ESKF_FAT_DEPTH
#define ESKF_FAT_DEPTH 3  // pipeline depth of the fat-CTA variants: 3 = issue and consume in the same trip; 4 = consume one trip later
#endif

struct AlignState {
  double T_total[12];  // R row-major (9) + t (3)
  double T_step[12];   // to be applied to the working cloud by the next iteration
  float Rf[12];        // fp32 copy of T_total's rotation
  int iter;
  int converged;
  int done;
  int pad0;
  unsigned long long n_corr;
  unsigned block_counter;
  unsigned epoch;
  unsigned error;
  unsigned tile_counter;  // dynamic tile scheduling (reset by the solver every iteration)
};

struct AlignParams {
  const tag_t* tags;  // probed (L2-resident); 0 = empty
  const VoxelSlot* slots;
  uint32_t n_slots;
  uint32_t pad_slots;
  double voxel;
  const double* x0;
  const double* y0;
  const double* z0;
  const float4* c4;
  const float2* c2;
  double* wx;
  double* wy;
  double* wz;
  unsigned n;
  double guess[12];
  int max_iteration;
  int neighbor_mode;
  double trans_sq_thr;
  double cos_thr;
  int fixed_iterations;
  int dynamic_tiles;  // 1: warps pull tiles from a counter (load-balanced, summation order varies)
  int ticket_chunk;   // tiles per ticket in that dynamic tail (1, 2 or 4)
  int dyn16;          // sixteenths of a pass dealt by tickets (eskf_ctx option align_dyn16, default 3)
  AlignState* st;
  double* partials;  // [G][kAcc]
  double* sums;      // [kAcc] (single_pass output / solve input)
  double* trace_H;
  double* trace_b;
  unsigned long long* trace_ncorr;
  double* trace_step;
  uint8_t* hit;
  // multi-GPU (one registration sharded by point range): the 28
  // sums of every rank meet in peer-mapped mailboxes over NVLink, see
  // exchange_sums().  world == 1: single GPU, no exchange.
  int world;
  int rank;
  unsigned seq;             // per-call sequence number (stale flags never match)
  double* peers[kMaxWorld]; // mailbox base of every rank (peers[rank] = the local one)
  HostMail* mail;           // nullable: host-mapped result words (internal.h)
  unsigned mail_seq;
  // depth-5 kernel (accumulate_points_resident)
  const uint8_t* filt;      // 8-bit probe filter of the table (internal.h, map_probe_filter)
  int k_res;                // warp tiles per warp whose working positions live in shared memory
  int ll;                   // 1: the next pose reaches the CTAs as flagged words (LL), no epoch round trip
  unsigned flags;           // depth 5 cache-policy experiments (eskf_ctx option "align_flags", see kFlag*)
  int n_cons;               // depth 6: consumer warps per CTA (0 = follow the hit rate)
  int xchg_ll;              // multi-GPU: 1 = flagged-word exchange of the sums, 0 = data + release flag
  unsigned long long* stamps;  // nullable: [max_it][8] globaltimer stamps of the iteration hand-off (ESKF_ALIGN_STAMPS=1)
  uint32_t* spill;          // depth 7: [G][10][spill_cap] words, hit-list entries beyond shared memory
  unsigned spill_cap;
  unsigned long long* llbox;  // [2][kLLWords] flagged words
};

// final pose + bookkeeping straight into host-mapped memory (the host polls align_seq)
__device__ __forceinline__ void publish_result(const AlignParams& P, const double* tot, int iterations,
                                               int converged, unsigned long long n_corr) {
  volatile HostMail* m = P.mail;
#pragma unroll
  for (int i = 0; i < 12; ++i) m->align_T[i] = tot[i];
  m->align_iter = iterations;
  m->align_converged = converged;
  m->align_ncorr = n_corr;
  m->align_error = 0u;
  __threadfence_system();
  m->align_seq = P.mail_seq;
}

// ------------------------------------------------------------ per point math
// adds the 27 unique terms of J^T W J / J^T W r (+ a correspondence count) of
// one correspondence into v[0..27]  (SET: stores them instead: v starts undefined)
template <typename F, bool SET = false>
__device__ __forceinline__ void point_terms(F px, F py, F pz, F rx, F ry, F rz, F m00, F m01,
                                            F m02, F m11, F m12, F m22, F* v) {
  // W = M^-1 by cofactors (Eigen Matrix3d::inverse(), Registration.cpp:95)
  const F c00 = m11 * m22 - m12 * m12;
  const F c01 = m02 * m12 - m01 * m22;
  const F c02 = m01 * m12 - m02 * m11;
  const F det = m00 * c00 + m01 * c01 + m02 * c02;
  const F inv = F(1) / det;
  const F w00 = c00 * inv, w01 = c01 * inv, w02 = c02 * inv;
  const F w11 = (m00 * m22 - m02 * m02) * inv;
  const F w12 = (m01 * m02 - m00 * m12) * inv;
  const F w22 = (m00 * m11 - m01 * m01) * inv;
  // J = [I | -skew(p)]  =>  H = [[W, B], [B^T, D]],  B_i = p x W_i,  D = skew(p) B
  const F b00 = py * w02 - pz * w01, b01 = pz * w00 - px * w02, b02 = px * w01 - py * w00;
  const F b10 = py * w12 - pz * w11, b11 = pz * w01 - px * w12, b12 = px * w11 - py * w01;
  const F b20 = py * w22 - pz * w12, b21 = pz * w02 - px * w22, b22 = px * w12 - py * w02;
  const F d00 = py * b20 - pz * b10, d01 = py * b21 - pz * b11, d02 = py * b22 - pz * b12;
  const F d11 = pz * b01 - px * b21, d12 = pz * b02 - px * b22;
  const F d22 = px * b12 - py * b02;
  // J^T W r = [W r ; p x (W r)]
  const F g0 = w00 * rx + w01 * ry + w02 * rz;
  const F g1 = w01 * rx + w11 * ry + w12 * rz;
  const F g2 = w02 * rx + w12 * ry + w22 * rz;
  const F g3 = py * g2 - pz * g1, g4 = pz * g0 - px * g2, g5 = px * g1 - py * g0;
  const F t[28] = {w00, w01, w02, w11, w12, w22, b00, b01, b02, b10, b11, b12, b20, b21,
                   b22, d00, d01, d02, d11, d12, d22, g0,  g1,  g2,  g3,  g4,  g5,  F(1)};
#pragma unroll
  for (int k = 0; k < 28; ++k) v[k] = SET ? t[k] : v[k] + t[k];
}

// C' = R S R^T for symmetric S, 6 unique outputs
template <typename F>
__device__ __forceinline__ void rotate_sym(const F* R, F s00, F s01, F s02, F s11, F s12, F s22,
                                           F* o) {
  F t[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    t[3 * i + 0] = R[3 * i] * s00 + R[3 * i + 1] * s01 + R[3 * i + 2] * s02;
    t[3 * i + 1] = R[3 * i] * s01 + R[3 * i + 1] * s11 + R[3 * i + 2] * s12;
    t[3 * i + 2] = R[3 * i] * s02 + R[3 * i + 1] * s12 + R[3 * i + 2] * s22;
  }
  o[0] = t[0] * R[0] + t[1] * R[1] + t[2] * R[2];
  o[1] = t[0] * R[3] + t[1] * R[4] + t[2] * R[5];
  o[2] = t[0] * R[6] + t[1] * R[7] + t[2] * R[8];
  o[3] = t[3] * R[3] + t[4] * R[4] + t[5] * R[5];
  o[4] = t[3] * R[6] + t[4] * R[7] + t[5] * R[8];
  o[5] = t[6] * R[6] + t[7] * R[7] + t[8] * R[8];
}

// Warp "reduce-scatter": every lane holds 32 partial terms; after 31 shuffles
// lane l holds the warp total of term l.  Fixed order => deterministic.
template <typename F>
__device__ __forceinline__ F warp_reduce_scatter32(F* v, unsigned lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < o; ++j) {
      const F a = v[j], b = v[j + o];
      const F send = upper ? a : b;
      const F keep = upper ? b : a;
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// ticket / epoch hand-off primitives: the CTA's partial sums are published by
// bar.sync + ONE acq_rel atomic of thread 0 (cumulative release), consumed by
// the last CTA through the same atomic (acquire) + bar.sync.
__device__ __forceinline__ unsigned atom_add_acq_rel(unsigned* p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}

__constant__ int c_off7[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0},
                                 {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};

__device__ __forceinline__ uint64_t load_key(const VoxelSlot* s) {
  const uint2 kw = __ldg(reinterpret_cast<const uint2*>(s));
  return (static_cast<uint64_t>(kw.y) << 32) | kw.x;
}

// finish a lookup whose first tag probe (slot h) returned `t`: walk the tag
// array (L2) and touch a 64 B record (HBM) only on a tag match
__device__ __forceinline__ const VoxelSlot* resolve_probe(const tag_t* tags,
                                                          const VoxelSlot* slots, uint32_t n_slots,
                                                          uint64_t key, SlotAddr a, tag_t t) {
  uint32_t h = a.home;
  for (uint32_t probe = 0; probe < n_slots; ++probe) {
    if (t == 0u) return nullptr;
    if (t == a.tag && load_key(slots + h) == key) return slots + h;
    h = next_slot(h, n_slots);
    t = __ldg(tags + h);
  }
  return nullptr;
}

// One pass over this CTA's points: transform, look up, accumulate.
// Work is dealt in warp tiles of 32*U consecutive points; the loads of the U
// points of a lane are issued back to back (position -> first probe -> voxel
// payload) so U independent dependent-load chains are in flight per thread.
// Per tile the 28 per-lane partial terms (fp32) are folded with a warp
// reduce-scatter and added to ONE fp64 accumulator per lane (lane l = term l).
template <typename F, int U, int NN, int NW>
__device__ __forceinline__ double accumulate_points(const AlignParams& P, const double* sT,
                                                    const F* sR, bool first, bool write_hit) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned tile_pts = 32u * U;
  const unsigned wglobal = blockIdx.x * NW + (threadIdx.x >> 5);
  const unsigned wstride = gridDim.x * NW;
  const double* sx = first ? P.x0 : P.wx;
  const double* sy = first ? P.y0 : P.wy;
  const double* sz = first ? P.z0 : P.wz;
  const unsigned n_tiles = (P.n + tile_pts - 1) / tile_pts;
  const double inv_voxel = 1.0 / P.voxel;
  double acc = 0.0;
  for (unsigned tile = wglobal; tile < n_tiles; tile += wstride) {
    const unsigned start = tile * tile_pts + lane;
    F v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = F(0);
    double x[U], y[U], z[U];
    bool valid[U];
    // stage 1: positions
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned i = start + 32u * u;
      valid[u] = i < P.n;
      x[u] = y[u] = z[u] = 0.0;
      if (valid[u]) {
        // streamed once per iteration: evict-first, so the position stream does
        // not push the map's tags / touched voxel records out of L2
        x[u] = __ldcs(sx + i);
        y[u] = __ldcs(sy + i);
        z[u] = __ldcs(sz + i);
      }
    }
    // stage 2: transform, write back, keys, first probes
    int kx[U], ky[U], kz[U];
    uint64_t key[U][NN];
    SlotAddr addr[U][NN];
    tag_t tag0[U][NN];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned i = start + 32u * u;
      if (valid[u]) {
        transform_point_rn(sT, x[u], y[u], z[u]);
        __stcs(P.wx + i, x[u]);
        __stcs(P.wy + i, y[u]);
        __stcs(P.wz + i, z[u]);
      }
      kx[u] = voxel_coord(x[u], P.voxel, inv_voxel);  // (== the division, common.cuh)
      ky[u] = voxel_coord(y[u], P.voxel, inv_voxel);
      kz[u] = voxel_coord(z[u], P.voxel, inv_voxel);
#pragma unroll
      for (int o = 0; o < NN; ++o) {
        const int vx = kx[u] + c_off7[o][0], vy = ky[u] + c_off7[o][1], vz = kz[u] + c_off7[o][2];
        const bool ok = valid[u] && coord_in_range(vx) && coord_in_range(vy) && coord_in_range(vz);
        key[u][o] = pack_key(vx, vy, vz);
        addr[u][o] = slot_addr(key[u][o], P.n_slots);
#if defined(ESKF_ABLATE) && ESKF_ABLATE == 1
        tag0[u][o] = static_cast<tag_t>(ok ? (addr[u][o].home & 1u) : 0u);
#else
        tag0[u][o] = ok ? __ldg(P.tags + addr[u][o].home) : static_cast<tag_t>(0);
#endif
      }
    }
    // stage 3: resolve the lookups
    const VoxelSlot* slot[U][NN];
#if !defined(ESKF_ABLATE)
    if (NN > 1 && (P.flags & 4096u) != 0u) {  // (align_flags bit 4096: the round-1 order, one lookup after the other: 5-10 % slower)
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int o = 0; o < NN; ++o)
          slot[u][o] = resolve_probe(P.tags, P.slots, P.n_slots, key[u][o], addr[u][o], tag0[u][o]);
    } else if (NN > 1) {
      // the neighbourhood's lookups together: the key words of every home slot whose tag matches are
      // requested back to back (one HBM round trip for up to NN records instead of NN dependent ones);
      // a lookup that has to walk on (tag of another key at home, or a 2^-16 tag collision) takes the
      // serial path from where it stands
      uint64_t hk[U][NN];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int o = 0; o < NN; ++o) {
          hk[u][o] = kEmptyKey;
          if (tag0[u][o] != 0u && tag0[u][o] == addr[u][o].tag) {
            hk[u][o] = load_key(P.slots + addr[u][o].home);
            // (the record's second 32 B sector, read once the key is confirmed)
            if (P.flags & 8192u)  // (bit 8192; measured: no gain, 30 % slower on a table spread over 640 MB)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(P.slots + addr[u][o].home) + 32));
          }
        }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int o = 0; o < NN; ++o) {
          if (tag0[u][o] == 0u) slot[u][o] = nullptr;
          else if (tag0[u][o] == addr[u][o].tag && hk[u][o] == key[u][o]) slot[u][o] = P.slots + addr[u][o].home;
          else slot[u][o] = resolve_probe(P.tags, P.slots, P.n_slots, key[u][o], addr[u][o], tag0[u][o]);
        }
    }
#endif
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int o = 0; o < NN; ++o) {
#if defined(ESKF_ABLATE) && ESKF_ABLATE == 1   // experiment: no table access at all
        slot[u][o] = nullptr;
        if (tag0[u][o] == 12345u) v[27] += F(1);
#elif defined(ESKF_ABLATE) && ESKF_ABLATE == 2  // experiment: tag probes only, no records
        {
          uint32_t h = addr[u][o].home;
          tag_t t = tag0[u][o];
          int found = 0;
          for (uint32_t probe = 0; probe < P.n_slots; ++probe) {
            if (t == 0u) break;
            if (t == addr[u][o].tag) { found = 1; break; }
            h = next_slot(h, P.n_slots);
            t = __ldg(P.tags + h);
          }
          v[27] += F(found);
          slot[u][o] = nullptr;
        }
#else
        if (NN == 1) slot[u][o] = resolve_probe(P.tags, P.slots, P.n_slots, key[u][o], addr[u][o], tag0[u][o]);
#endif
        if (write_hit && valid[u])
          P.hit[static_cast<size_t>(NN) * (start + 32u * u) + o] = slot[u][o] != nullptr ? 1 : 0;
      }
    // stage 4/5: payloads + per-point algebra
    if (NN == 1) {
      float4 pa[U], pc[U], pd[U], s4[U];
      float2 s2[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (slot[u][0] != nullptr) {
          pa[u] = __ldg(reinterpret_cast<const float4*>(slot[u][0]) + 1);  // mx my mz -
          pc[u] = __ldg(reinterpret_cast<const float4*>(slot[u][0]) + 2);  // c00 c01 c02 c11
          pd[u] = __ldg(reinterpret_cast<const float4*>(slot[u][0]) + 3);  // c12 c22 - -
          s4[u] = __ldcs(P.c4 + start + 32u * u);
          s2[u] = __ldcs(P.c2 + start + 32u * u);
        }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (slot[u][0] != nullptr) {
          F cr[6];
          rotate_sym<F>(sR, F(s4[u].x), F(s4[u].y), F(s4[u].z), F(s4[u].w), F(s2[u].x), F(s2[u].y), cr);
          // residual against the voxel mean, formed relative to the voxel centre
          const double cx = __dmul_rn(static_cast<double>(kx[u]) + 0.5, P.voxel);
          const double cy = __dmul_rn(static_cast<double>(ky[u]) + 0.5, P.voxel);
          const double cz = __dmul_rn(static_cast<double>(kz[u]) + 0.5, P.voxel);
          const F rx = F(x[u] - cx) - F(pa[u].x), ry = F(y[u] - cy) - F(pa[u].y),
                  rz = F(z[u] - cz) - F(pa[u].z);
          point_terms<F>(F(x[u]), F(y[u]), F(z[u]), rx, ry, rz, cr[0] + F(pc[u].x),
                         cr[1] + F(pc[u].y), cr[2] + F(pc[u].z), cr[3] + F(pc[u].w),
                         cr[4] + F(pd[u].x), cr[5] + F(pd[u].y), v);
        }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        bool any = false;
#pragma unroll
        for (int o = 0; o < NN; ++o) any = any || slot[u][o] != nullptr;
        if (!any) continue;
        const float4 s4 = __ldcs(P.c4 + start + 32u * u);
        const float2 s2 = __ldcs(P.c2 + start + 32u * u);
        F cr[6];
        rotate_sym<F>(sR, F(s4.x), F(s4.y), F(s4.z), F(s4.w), F(s2.x), F(s2.y), cr);
#pragma unroll
        for (int o = 0; o < NN; ++o) {
          if (slot[u][o] == nullptr) continue;
          const float4 a = __ldg(reinterpret_cast<const float4*>(slot[u][o]) + 1);
          const float4 c = __ldg(reinterpret_cast<const float4*>(slot[u][o]) + 2);
          const float4 d = __ldg(reinterpret_cast<const float4*>(slot[u][o]) + 3);
          const int vx = kx[u] + c_off7[o][0], vy = ky[u] + c_off7[o][1], vz = kz[u] + c_off7[o][2];
          const double cx = __dmul_rn(static_cast<double>(vx) + 0.5, P.voxel);
          const double cy = __dmul_rn(static_cast<double>(vy) + 0.5, P.voxel);
          const double cz = __dmul_rn(static_cast<double>(vz) + 0.5, P.voxel);
          const F rx = F(x[u] - cx) - F(a.x), ry = F(y[u] - cy) - F(a.y), rz = F(z[u] - cz) - F(a.z);
          point_terms<F>(F(x[u]), F(y[u]), F(z[u]), rx, ry, rz, cr[0] + F(c.x), cr[1] + F(c.y),
                         cr[2] + F(c.z), cr[3] + F(c.w), cr[4] + F(d.x), cr[5] + F(d.y), v);
        }
      }
    }
    acc += static_cast<double>(warp_reduce_scatter32<F>(v, lane));
  }
  return acc;
}

// ------------------------------------------------------------------------
// Hot path (1-neighbour): the same work as accumulate_points as a 3-stage
// software pipeline over warp tiles of 32 points.  Every loop trip issues, back
// to back and mutually independent,
//     the position loads of tile t+2          (HBM / L2 stream)
//     the 16 B tag-group load of tile t+1     (L2-resident tag array)
//     the 64 B voxel record + source covariance of tile t   (HBM / L2)
// and only then consumes them: scan the tags of t+1 -> candidate slot (the
// record is prefetched, evict_last, so the voxels a registration keeps touching
// stay in L2 across Gauss-Newton iterations), fp32 algebra of tile t, fp64
// transform + key + hash of tile t+2.  A trip therefore waits for ONE memory
// round trip instead of three dependent ones.
struct PtState {  // a transformed point waiting for its lookup
  double x, y, z;
  int kx, ky, kz;
  uint32_t home;
  uint32_t tag;   // 0 = no lookup (invalid lane / out of key range)
  uint32_t cand;  // candidate slot after the tag scan, kNoCand if none
};
constexpr uint32_t kNoCand = 0xffffffffu;

__device__ __forceinline__ void prefetch_record(const void* p) {
  asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p));
}

// A lookup first sees a WINDOW of 8 tags: slots [b, b+4) and the 4 slots that follow them in
// probe order, b = home & ~(kTagAlign - 1).  With kTagAlign = 4 (two 8 B loads) the window always
// holds >= 5 tags from `home` on and ~1e-3 of the lookups need a second, dependent L2 round trip;
// with one 16 B-aligned group (kTagAlign = 8) 3/8 of the lookups start with < 4 tags left, ~5 % of
// them run off the group, and a warp of 32 lookups paid the second round trip on most of its trips.
#ifndef ESKF_TAG_ALIGN
#define ESKF_TAG_ALIGN 4
#endif
constexpr uint32_t kTagAlign = ESKF_TAG_ALIGN;
static_assert(kTagAlign == 4 || kTagAlign == 8, "tag window alignment");

// scan the window's tags from position k0 on: position of the first tag equal to `tag`,
// kNoCand on an empty slot, kMore if the window is exhausted
constexpr uint32_t kMore = 0xfffffffeu;
__device__ __forceinline__ uint32_t scan_tag_window(uint4 w, uint32_t k0, uint32_t tag) {
  const uint64_t lo = (static_cast<uint64_t>(w.y) << 32) | w.x;
  const uint64_t hi = (static_cast<uint64_t>(w.w) << 32) | w.z;
  for (uint32_t k = k0; k < 8u; ++k) {
    const uint32_t t = static_cast<uint32_t>((k < 4u ? lo >> (16u * k) : hi >> (16u * (k - 4u))) & 0xffffu);
    if (t == 0u) return kNoCand;
    if (t == tag) return k;
  }
  return kMore;
}

// the window that starts at slot b (a multiple of 4; n_slots is a multiple of 64, so neither half
// straddles the end of the table); b2 = first slot of its second half
__device__ __forceinline__ uint4 load_tag_window(const tag_t* tags, uint32_t b, uint32_t b2) {
  if (kTagAlign == 8) return __ldg(reinterpret_cast<const uint4*>(tags + b));
  const uint2 lo = __ldg(reinterpret_cast<const uint2*>(tags + b));
  const uint2 hi = __ldg(reinterpret_cast<const uint2*>(tags + b2));
  return make_uint4(lo.x, lo.y, hi.x, hi.y);
}

// ---- async-proxy, cache-policy and probe-filter helpers (used by depths 4 .. 7)
constexpr int kRing = 3;
constexpr unsigned kTileBytes = 3u * 32u * sizeof(double);  // x[32] y[32] z[32]

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0u;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> shared bulk copy (async proxy; SASS: UBLKCP), completion counted on `bar`
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, unsigned bytes, uint64_t* bar,
                                          uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy(unsigned sel) {  // 0 normal, 1 evict_first, 2 evict_last
  return sel == 1u ? l2_policy_evict_first() : sel == 2u ? l2_policy_evict_last() : l2_policy_evict_normal();
}
// read-only loads / stores that carry an L2 eviction policy
__device__ __forceinline__ uint2 ldg_u2_hint(const void* p, uint64_t pol) {
  uint2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint4 ldg_u4_hint(const void* p, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
// the 16 filter bytes from slot (home & ~7) on (the filter is padded by 16 entries: no wrap): one aligned
// 16 B load when homes are bucket-aligned (common.cuh kBucket), else two 8 B loads
__device__ __forceinline__ uint4 ldg_u4_hint(const void* p, uint64_t pol);
__device__ __forceinline__ uint4 filter_window(const uint8_t* filt, uint32_t home, uint64_t pol) {
  if constexpr (kBucket >= 16u) {
    return ldg_u4_hint(filt + home, pol);
  } else {
    const uint8_t* w = filt + (home & ~7u);
    const uint2 lo = ldg_u2_hint(w, pol);
    const uint2 hi = ldg_u2_hint(w + 8, pol);
    return make_uint4(lo.x, lo.y, hi.x, hi.y);
  }
}
__device__ __forceinline__ float2 ldg_f2_hint(const void* p, uint64_t pol) {
  float2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float4 ldg_f4_hint(const void* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_f64_hint(double* p, double v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
// eskf_ctx option "align_flags" (experiments on what the L2 keeps between Gauss-Newton iterations)
constexpr unsigned kFlagRecPolicyMask = 3u;    // voxel records: 0 evict_normal, 1 evict_first, 2 evict_last
constexpr unsigned kFlagNoRecPrefetch = 4u;    // no prefetch.global.L2::evict_last of the record
constexpr unsigned kFlagFiltPolicyShift = 3u;  // bits 3-4: filter windows: 0 evict_normal, 1 evict_first, 2 evict_last
constexpr unsigned kFlagStoreFirst = 32u;      // position write-back marked evict_first
constexpr unsigned kFlagCovAllLanes = 64u;     // every lane loads its source covariance (a coalesced stream)
// depth 8 ablations (timing experiments only: results are wrong with any of them set)
constexpr unsigned kFlagNoProbe = 256u;        // no filter probes: every point misses
constexpr unsigned kFlagNoGather = 512u;       // probes, but candidates are dropped: no record / covariance gathers, no algebra
constexpr unsigned kFlagNoStore = 1024u;       // transformed positions are not written back
constexpr unsigned kFlagNoPrefetch = 2048u;    // no L2 prefetch of the positions

// generic-proxy writes (other SMs' position stores of the previous iteration, acquired through the
// iteration hand-off) -> this thread's async-proxy reads
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// scan the 16 filter bytes of a window from position k0 (< 8) on: position of the first byte equal to
// t8, kNoCand when an empty slot (0) comes first, kMore when the window is exhausted.  Byte-parallel:
// (v - 0x01..) & ~v & 0x80.. flags the zero bytes of v, exactly so at the lowest flagged position.
__device__ __forceinline__ uint32_t scan_filter_window(uint4 w, uint32_t k0, uint32_t t8) {
  const uint64_t ones = 0x0101010101010101ull, highs = 0x8080808080808080ull;
  const uint64_t pat = ones * t8;
  const uint64_t lo = (static_cast<uint64_t>(w.y) << 32) | w.x;
  const uint64_t hi = (static_cast<uint64_t>(w.w) << 32) | w.z;
  const uint64_t skip = (1ull << (8u * k0)) - 1ull;  // bytes before k0 never terminate the scan
  {
    const uint64_t e = lo | skip, m = (lo ^ pat) | skip;
    const uint64_t ze = (e - ones) & ~e & highs, zm = (m - ones) & ~m & highs;
    const uint64_t any = ze | zm;
    if (any != 0ull) {
      const int bit = __ffsll(static_cast<long long>(any)) - 1;
      return ((zm >> bit) & 1ull) ? static_cast<uint32_t>(bit >> 3) : kNoCand;
    }
  }
  {
    const uint64_t e = hi, m = hi ^ pat;
    const uint64_t ze = (e - ones) & ~e & highs, zm = (m - ones) & ~m & highs;
    const uint64_t any = ze | zm;
    if (any != 0ull) {
      const int bit = __ffsll(static_cast<long long>(any)) - 1;
      return ((zm >> bit) & 1ull) ? 8u + static_cast<uint32_t>(bit >> 3) : kNoCand;
    }
  }
  return kMore;
}

// slow path after a filter match whose record holds another key (1/255 per occupied probe)
__device__ __noinline__ const VoxelSlot* resolve_probe_filter(const uint8_t* filt, const VoxelSlot* slots,
                                                              uint32_t n_slots, uint64_t key, uint32_t h,
                                                              uint32_t t8) {
  for (uint32_t probe = 0; probe < n_slots; ++probe) {
    const uint32_t t = __ldg(filt + h);
    if (t == 0u) return nullptr;
    if (t == t8 && load_key(slots + h) == key) return slots + h;
    h = next_slot(h, n_slots);
  }
  return nullptr;
}

template <typename F, int NW, int DEPTH>
__device__ __forceinline__ double accumulate_points_pipelined(const AlignParams& P,
                                                              const double* sT, const F* sR,
                                                              bool first, bool write_hit) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned wglobal = blockIdx.x * NW + (threadIdx.x >> 5);
  const unsigned wstride = gridDim.x * NW;
  const double* sx = first ? P.x0 : P.wx;
  const double* sy = first ? P.y0 : P.wy;
  const double* sz = first ? P.z0 : P.wz;
  const unsigned n_tiles = (P.n + 31u) / 32u;
  const double inv_voxel = 1.0 / P.voxel;

  // fp64 transform, write-back, voxel key, table address of one raw point
  auto xform = [&](unsigned tile, double rx, double ry, double rz, PtState& s) {
    const unsigned i = tile * 32u + lane;
    s.x = rx; s.y = ry; s.z = rz;
    s.kx = s.ky = s.kz = 0;
    s.home = 0;
    s.tag = 0;
    s.cand = kNoCand;
    if (tile >= n_tiles || i >= P.n) return;
    transform_point_rn(sT, s.x, s.y, s.z);
    P.wx[i] = s.x;
    P.wy[i] = s.y;
    P.wz[i] = s.z;
    s.kx = voxel_coord(s.x, P.voxel, inv_voxel);
    s.ky = voxel_coord(s.y, P.voxel, inv_voxel);
    s.kz = voxel_coord(s.z, P.voxel, inv_voxel);
    if (coord_in_range(s.kx) && coord_in_range(s.ky) && coord_in_range(s.kz)) {
      const SlotAddr ad = slot_addr(pack_key(s.kx, s.ky, s.kz), P.n_slots);
      s.home = ad.home;
      s.tag = ad.tag;
    }
  };
  auto load_pos = [&](unsigned tile, double& rx, double& ry, double& rz) {
    const unsigned i = tile * 32u + lane;
    rx = ry = rz = 0.0;
    if (tile < n_tiles && i < P.n) {
      rx = first ? __ldcs(sx + i) : __ldcg(sx + i);
      ry = first ? __ldcs(sy + i) : __ldcg(sy + i);
      rz = first ? __ldcs(sz + i) : __ldcg(sz + i);
    }
  };
  auto wrap4 = [&](uint32_t b) -> uint32_t { return b + 4u == P.n_slots ? 0u : b + 4u; };
  // large clouds probe the map's 8-bit filter (16 slots per 16 B window, held in L2 by an eviction
  // hint) instead of the 16-bit tags: eskf_ctx option align_filter, fill_params
  // (DEPTH 4 decides at run time; the 3-stage loop exists with the filter compiled in, DEPTH 2, for
  // large clouds, and without it, DEPTH 3, which is what every frame runs: no extra registers there)
  const bool use_filter = DEPTH == 4 ? P.filt != nullptr : DEPTH == 2;
  const uint64_t pol_filt = l2_policy((P.flags >> kFlagFiltPolicyShift) & 3u);
  // first tag window of a transformed point (zeros = "empty" when it has no lookup)
  auto first_window = [&](const PtState& s) -> uint4 {
    if (s.tag == 0u) return make_uint4(0u, 0u, 0u, 0u);
    if (use_filter) return filter_window(P.filt, s.home, pol_filt);
    const uint32_t b = s.home & ~(kTagAlign - 1u);
    return load_tag_window(P.tags, b, wrap4(b));
  };
  // finish the tag scan of s given its first window w (rarely needs more windows)
  auto finish_scan = [&](PtState& s, uint4 w) {
    if (s.tag == 0u) return;
    if (use_filter) {
      uint32_t b0 = s.home & ~7u;
      uint32_t r = scan_filter_window(w, s.home & 7u, filter_tag(s.tag));
      uint32_t scanned = 16u - (s.home & 7u);
      while (r == kMore && scanned < P.n_slots) {
        b0 += 16u;
        if (b0 >= P.n_slots) b0 -= P.n_slots;
        const uint2 lo = __ldg(reinterpret_cast<const uint2*>(P.filt + b0));
        const uint2 hi = __ldg(reinterpret_cast<const uint2*>(P.filt + b0 + 8));
        r = scan_filter_window(make_uint4(lo.x, lo.y, hi.x, hi.y), 0u, filter_tag(s.tag));
        scanned += 16u;
      }
      if (r < 16u) {
        uint32_t c = b0 + r;
        if (c >= P.n_slots) c -= P.n_slots;
        s.cand = c;
        prefetch_record(P.slots + s.cand);
      }
      return;
    }
    uint32_t b = s.home & ~(kTagAlign - 1u), b2 = wrap4(b);
    uint32_t r = scan_tag_window(w, s.home & (kTagAlign - 1u), s.tag);
    uint32_t scanned = 8u - (s.home & (kTagAlign - 1u));
    while (r == kMore && scanned < P.n_slots) {
      b = wrap4(b2);
      b2 = wrap4(b);
      r = scan_tag_window(load_tag_window(P.tags, b, b2), 0u, s.tag);
      scanned += 8u;
    }
    if (r < 8u) {
      s.cand = r < 4u ? b + r : b2 + (r - 4u);
      prefetch_record(P.slots + s.cand);
    }
  };

  double acc = 0.0;
  // Tiles: a fixed stride per warp for the first 13/16 of every pass, the rest
  // pulled from a global counter.  The spread of per-warp progress (HBM channel
  // luck) otherwise leaves most warps waiting ~20 % of the pass at the
  // iteration barrier for the slowest one.  A ticket is requested at least one
  // trip before it is used (lane 0 keeps the raw value, the broadcast happens
  // at the use), so the atomic's latency stays off the pipeline.
  // Measured: 189 -> 160 us per iteration at 0.5 m voxels (every point hits).
  // Clouds with fewer than 8 tiles per warp, or dynamic_tiles == 0
  // (ESKF_ALIGN_DYNAMIC=0), run fully static: bit-reproducible summation order.
  const bool dynamic = P.dynamic_tiles != 0 && n_tiles >= 8u * wstride;  // small clouds: nothing to balance
  const unsigned k_static = dynamic ? (n_tiles - n_tiles * static_cast<unsigned>(P.dyn16) / 16u) / wstride : 0xffffffffu;
  const unsigned dyn_base = dynamic ? k_static * wstride : 0u;
  // A ticket is worth `chunk` consecutive tiles (fewer same-address atomics, and chunk trips of
  // lead time): lane 0 keeps the chunk in use and the one requested ahead.
  const unsigned chunk = static_cast<unsigned>(P.ticket_chunk);
  unsigned k_next = 0;
  unsigned tk_cur = 0, tk_next = 0;
  auto next_tile = [&]() -> unsigned {
    unsigned t;
    if (k_next < k_static) {
      t = wglobal + k_next * wstride;
    } else {
      const unsigned sub = (k_next - k_static) & (chunk - 1u);
      if (sub == 0u) {
        tk_cur = tk_next;  // (waits for the atomic issued >= one call ago)
        if (lane == 0) tk_next = atomicAdd(&P.st->tile_counter, chunk);
      }
      t = dyn_base + __shfl_sync(0xffffffffu, tk_cur, 0) + sub;
    }
    if (t > n_tiles) t = n_tiles;  // past the end (static mode, or the last tickets of a pass)
    ++k_next;
    if (k_next == k_static && lane == 0) tk_next = atomicAdd(&P.st->tile_counter, chunk);  // first chunk
    return t;
  };
  // the fp32 source covariances of a tile (512 + 256 B, consumed two trips after the tile id is
  // known) are pulled into L2 ahead of time, 64 B per lane of lanes 0..11
  auto prefetch_cov = [&](unsigned t) {
#if ESKF_COV_PREFETCH
    if (t < n_tiles && lane < 12u) {
      const char* a = lane < 8u ? reinterpret_cast<const char*>(P.c4 + static_cast<size_t>(t) * 32u) + 64u * lane
                                : reinterpret_cast<const char*>(P.c2 + static_cast<size_t>(t) * 32u) + 64u * (lane - 8u);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    }
#else
    (void)t;
#endif
  };
  F v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = F(0);
  int folded = 0;
  PtState cur, nxt;
  if constexpr (DEPTH == 4) {
  // ---- 4-deep rotation: every load is consumed ONE TRIP AFTER it was issued.  Across the loop's
  // back edge a warp has in flight: the 64 B record + source covariance of `cur` (tile t), the tag
  // window of `nxt` (tile t+1) and the raw positions of tile t+2; a trip consumes them in that
  // order and, right after each consumption, issues the same loads one tile further on into the
  // registers it just freed.  (The 3-stage loop below issues and consumes in the same trip: one
  // exposed memory round trip per trip, 50 % of the warp-stall samples on the dense config.)
  // A point waiting for its lookup keeps only what the algebra needs: the position and its
  // offset from the voxel centre already rounded to F, and the packed key.
  struct Slim {
    F px, py, pz;     // transformed position
    F dx, dy, dz;     // position - voxel centre (formed in fp64)
    uint32_t klo, khi;  // packed voxel key
    uint32_t home;
    uint32_t tag;     // 0 = no lookup (invalid lane / out of key range)
    uint32_t cand;    // candidate slot after the tag scan, kNoCand if none
  };
  struct RecRegs {
    uint2 key;
    float4 pa, pc;
    float2 pd;
    float4 s4;
    float2 s2;
  };
  auto xform4 = [&](unsigned tile, double x, double y, double z, Slim& q) {
    const unsigned i = tile * 32u + lane;
    q.px = q.py = q.pz = q.dx = q.dy = q.dz = F(0);
    q.klo = q.khi = 0u;
    q.home = 0u;
    q.tag = 0u;
    q.cand = kNoCand;
    if (tile >= n_tiles || i >= P.n) return;
    transform_point_rn(sT, x, y, z);
    P.wx[i] = x;
    P.wy[i] = y;
    P.wz[i] = z;
    const int kx = voxel_coord(x, P.voxel, inv_voxel);
    const int ky = voxel_coord(y, P.voxel, inv_voxel);
    const int kz = voxel_coord(z, P.voxel, inv_voxel);
    if (coord_in_range(kx) && coord_in_range(ky) && coord_in_range(kz)) {
      const uint64_t key = pack_key(kx, ky, kz);
      const SlotAddr ad = slot_addr(key, P.n_slots);
      q.klo = static_cast<uint32_t>(key);
      q.khi = static_cast<uint32_t>(key >> 32);
      q.home = ad.home;
      q.tag = ad.tag;
      // residual against the voxel mean is formed relative to the voxel centre
      q.px = F(x); q.py = F(y); q.pz = F(z);
      q.dx = F(x - __dmul_rn(static_cast<double>(kx) + 0.5, P.voxel));
      q.dy = F(y - __dmul_rn(static_cast<double>(ky) + 0.5, P.voxel));
      q.dz = F(z - __dmul_rn(static_cast<double>(kz) + 0.5, P.voxel));
    }
  };
  auto window4 = [&](const Slim& q) -> uint4 {
    if (q.tag == 0u) return make_uint4(0u, 0u, 0u, 0u);
    if (use_filter) {
      return filter_window(P.filt, q.home, pol_filt);
    }
    const uint32_t b0 = q.home & ~(kTagAlign - 1u);
    return load_tag_window(P.tags, b0, wrap4(b0));
  };
  auto scan4 = [&](Slim& q, uint4 w) {
    if (q.tag == 0u) return;
    if (use_filter) {
      uint32_t b0 = q.home & ~7u;
      uint32_t r = scan_filter_window(w, q.home & 7u, filter_tag(q.tag));
      uint32_t scanned = 16u - (q.home & 7u);
      while (r == kMore && scanned < P.n_slots) {
        b0 += 16u;
        if (b0 >= P.n_slots) b0 -= P.n_slots;
        const uint2 lo = __ldg(reinterpret_cast<const uint2*>(P.filt + b0));
        const uint2 hi = __ldg(reinterpret_cast<const uint2*>(P.filt + b0 + 8));
        r = scan_filter_window(make_uint4(lo.x, lo.y, hi.x, hi.y), 0u, filter_tag(q.tag));
        scanned += 16u;
      }
      if (r < 16u) {
        uint32_t c = b0 + r;
        if (c >= P.n_slots) c -= P.n_slots;
        q.cand = c;
      }
      return;
    }
    uint32_t b0 = q.home & ~(kTagAlign - 1u), b2 = wrap4(b0);
    uint32_t r = scan_tag_window(w, q.home & (kTagAlign - 1u), q.tag);
    uint32_t scanned = 8u - (q.home & (kTagAlign - 1u));
    while (r == kMore && scanned < P.n_slots) {
      b0 = wrap4(b2);
      b2 = wrap4(b0);
      r = scan_tag_window(load_tag_window(P.tags, b0, b2), 0u, q.tag);
      scanned += 8u;
    }
    if (r < 8u) q.cand = r < 4u ? b0 + r : b2 + (r - 4u);
  };
  auto issue_record = [&](const Slim& q, unsigned tile_of_q, RecRegs& r) {
    r.key = make_uint2(0u, 0u);
    if (q.cand != kNoCand) {
      const float4* rec = reinterpret_cast<const float4*>(P.slots + q.cand);
      const unsigned i = tile_of_q * 32u + lane;
      prefetch_record(rec);  // (evict_last: the voxels a registration keeps touching stay in L2)
      r.key = __ldg(reinterpret_cast<const uint2*>(rec));     // key
      r.pa = __ldg(rec + 1);                                   // mx my mz -
      r.pc = __ldg(rec + 2);                                   // c00 c01 c02 c11
      r.pd = __ldg(reinterpret_cast<const float2*>(rec + 3));  // c12 c22
      r.s4 = __ldcs(P.c4 + i);
      r.s2 = __ldcs(P.c2 + i);
    }
  };
  Slim c4s, n4s;
  unsigned tile = next_tile(), tile_n = next_tile(), tile_r = next_tile();
  RecRegs rec;
  uint4 tagw;
  double rx, ry, rz;
#if ESKF_POS_AHEAD
  // experiment (default off): a second tile of raw positions in flight (t+3), i.e. the HBM stream is
  // requested two trips before it is consumed; +6 registers
  unsigned tile_q = 0;
  double qx = 0.0, qy = 0.0, qz = 0.0;
#endif
  {
    double ax, ay, az, bx, by, bz;
    load_pos(tile, ax, ay, az);
    load_pos(tile_n, bx, by, bz);
    load_pos(tile_r, rx, ry, rz);
#if ESKF_POS_AHEAD
    tile_q = next_tile();
    load_pos(tile_q, qx, qy, qz);
#endif
    xform4(tile, ax, ay, az, c4s);
    scan4(c4s, window4(c4s));
    issue_record(c4s, tile, rec);
    xform4(tile_n, bx, by, bz, n4s);
    tagw = window4(n4s);
  }
  while (tile < n_tiles) {
    const unsigned i = tile * 32u + lane;
    // ---- 1. the record + covariance of the current tile
    float4 pa = rec.pa, pc = rec.pc;
    float2 pd = rec.pd;
    bool hit = false;
    if (c4s.cand != kNoCand) {
      hit = rec.key.x == c4s.klo && rec.key.y == c4s.khi;
      if (!hit) {
        // 16-bit tag collision (1/65536 per occupied probe): walk on, slowly
        const uint64_t key = (static_cast<uint64_t>(c4s.khi) << 32) | c4s.klo;
        SlotAddr rest;
        rest.home = next_slot(c4s.cand, P.n_slots);
        rest.tag = static_cast<tag_t>(c4s.tag);
        const VoxelSlot* far = resolve_probe(P.tags, P.slots, P.n_slots, key, rest, __ldg(P.tags + rest.home));
        if (far != nullptr) {
          hit = true;
          pa = __ldg(reinterpret_cast<const float4*>(far) + 1);
          pc = __ldg(reinterpret_cast<const float4*>(far) + 2);
          pd = __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float4*>(far) + 3));
        }
      }
    }
    if (write_hit && i < P.n) P.hit[i] = hit ? 1 : 0;
    if (hit) {
      F cr[6];
      rotate_sym<F>(sR, F(rec.s4.x), F(rec.s4.y), F(rec.s4.z), F(rec.s4.w), F(rec.s2.x), F(rec.s2.y), cr);
      point_terms<F, kAssign>(c4s.px, c4s.py, c4s.pz, c4s.dx - F(pa.x), c4s.dy - F(pa.y), c4s.dz - F(pa.z),
                              cr[0] + F(pc.x), cr[1] + F(pc.y), cr[2] + F(pc.z), cr[3] + F(pc.w),
                              cr[4] + F(pd.x), cr[5] + F(pd.y), v);
    } else if (kAssign) {
#pragma unroll
      for (int k = 0; k < 28; ++k) v[k] = F(0);
    }
    if (++folded == kFold) {
      acc += static_cast<double>(warp_reduce_scatter32<F>(v, lane));
      if (!kAssign) {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = F(0);
      }
      folded = 0;
    }
    // ---- 2. the tag window of the next tile; its record + covariance loads go out
    scan4(n4s, tagw);
    issue_record(n4s, tile_n, rec);
    c4s = n4s;
    // ---- 3. the positions of tile t+2; its tag window load goes out
    xform4(tile_r, rx, ry, rz, n4s);
    tagw = window4(n4s);
    tile = tile_n;
    tile_n = tile_r;
    // ---- 4. positions of the tile after that
#if ESKF_POS_AHEAD
    tile_r = tile_q;
    rx = qx; ry = qy; rz = qz;
    tile_q = next_tile();
    load_pos(tile_q, qx, qy, qz);
#else
    tile_r = next_tile();
    load_pos(tile_r, rx, ry, rz);
#endif
#if ESKF_POS_PREFETCH
    // statically dealt tiles are known ahead: pull the positions of the tile ESKF_POS_PREFETCH
    // trips further on into L2 (3 x 256 B, 64 B per lane of lanes 0..11)
    {
      const unsigned kp = k_next - 1u + ESKF_POS_PREFETCH;
      const unsigned tp = wglobal + kp * wstride;
      if (kp < k_static && tp < n_tiles && lane < 12u) {
        const double* base = lane < 4u ? sx : lane < 8u ? sy : sz;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + static_cast<size_t>(tp) * 32u + 8u * (lane & 3u)));
      }
    }
#endif
  }
  (void)cur;
  (void)nxt;
  (void)finish_scan;
  (void)first_window;
  (void)xform;
  (void)prefetch_cov;
  } else {
  // ---- prologue: cur = first tile (scanned), nxt = second tile (transformed)
  unsigned tile = next_tile(), tile_n = next_tile();
  prefetch_cov(tile_n);
  {
    double rx, ry, rz, qx, qy, qz;
    load_pos(tile, rx, ry, rz);
    load_pos(tile_n, qx, qy, qz);
    xform(tile, rx, ry, rz, cur);
    finish_scan(cur, first_window(cur));
    xform(tile_n, qx, qy, qz, nxt);
  }
  while (tile < n_tiles) {
    const unsigned i = tile * 32u + lane;
    const unsigned tile_r = next_tile();
    prefetch_cov(tile_r);
    // ---- issue: three independent groups of loads
    double rx, ry, rz;
    load_pos(tile_r, rx, ry, rz);
    const uint4 tagw = first_window(nxt);
    uint4 r0 = make_uint4(0u, 0u, 0u, 0u);
    float4 pa, pc, pd, s4;
    float2 s2;
    const bool has_cand = cur.cand != kNoCand;
    if (has_cand) {
      const VoxelSlot* rec = P.slots + cur.cand;
      r0 = __ldg(reinterpret_cast<const uint4*>(rec));       // key, count
      pa = __ldg(reinterpret_cast<const float4*>(rec) + 1);  // mx my mz -
      pc = __ldg(reinterpret_cast<const float4*>(rec) + 2);  // c00 c01 c02 c11
      pd = __ldg(reinterpret_cast<const float4*>(rec) + 3);  // c12 c22 - -
      s4 = __ldcs(P.c4 + i);
      s2 = __ldcs(P.c2 + i);
    }
    // ---- consume
    finish_scan(nxt, tagw);
    bool hit = false;
    if (has_cand) {
      const uint64_t key = pack_key(cur.kx, cur.ky, cur.kz);
      hit = ((static_cast<uint64_t>(r0.y) << 32) | r0.x) == key;
      if (!hit) {
        // 16-bit tag collision (1/65536 per occupied probe): walk on, slowly
        SlotAddr rest;
        rest.home = next_slot(cur.cand, P.n_slots);
        rest.tag = static_cast<tag_t>(cur.tag);
        const VoxelSlot* rec = resolve_probe(P.tags, P.slots, P.n_slots, key, rest, __ldg(P.tags + rest.home));
        if (rec != nullptr) {
          hit = true;
          pa = __ldg(reinterpret_cast<const float4*>(rec) + 1);
          pc = __ldg(reinterpret_cast<const float4*>(rec) + 2);
          pd = __ldg(reinterpret_cast<const float4*>(rec) + 3);
        }
      }
    }
    if (write_hit && i < P.n) P.hit[i] = hit ? 1 : 0;
    if (hit) {
      F cr[6];
      rotate_sym<F>(sR, F(s4.x), F(s4.y), F(s4.z), F(s4.w), F(s2.x), F(s2.y), cr);
      // residual against the voxel mean, formed relative to the voxel centre
      const double cx = __dmul_rn(static_cast<double>(cur.kx) + 0.5, P.voxel);
      const double cy = __dmul_rn(static_cast<double>(cur.ky) + 0.5, P.voxel);
      const double cz = __dmul_rn(static_cast<double>(cur.kz) + 0.5, P.voxel);
      const F ex = F(cur.x - cx) - F(pa.x), ey = F(cur.y - cy) - F(pa.y), ez = F(cur.z - cz) - F(pa.z);
      // kFold == 1: v is dead between trips, so a hit lane stores its terms and a miss lane
      // zeros instead of "zero all, then add" (28 FADDs and the zeroing of hit lanes less per trip)
      point_terms<F, kAssign>(F(cur.x), F(cur.y), F(cur.z), ex, ey, ez, cr[0] + F(pc.x), cr[1] + F(pc.y),
                              cr[2] + F(pc.z), cr[3] + F(pc.w), cr[4] + F(pd.x), cr[5] + F(pd.y), v);
    } else if (kAssign) {
#pragma unroll
      for (int k = 0; k < 28; ++k) v[k] = F(0);
    }
    // the 31-shuffle reduce-scatter runs once per kFold tiles (the shuffles
    // were ~17 % of the issue slots): lanes keep fp32 partial sums in between
    if (++folded == kFold) {
      acc += static_cast<double>(warp_reduce_scatter32<F>(v, lane));
      if (!kAssign) {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = F(0);
      }
      folded = 0;
    }
    cur = nxt;
    xform(tile_r, rx, ry, rz, nxt);
    tile = tile_n;
    tile_n = tile_r;
  }
  }
  if (folded != 0) acc += static_cast<double>(warp_reduce_scatter32<F>(v, lane));
  return acc;
}


// ------------------------------------------------------------------------
// Large clouds (depth 5): SM-resident working positions + 8-bit probe filter
// + a bulk-copy ring.
//
// The reference transforms its working copy of the cloud in place, once per
// Gauss-Newton iteration (Registration.cpp:13,27): 24 B read + 24 B written per
// point and iteration, the largest stream of the round-1 kernel.  Here the
// kernel is persistent over the whole loop, so a warp keeps the transformed
// positions of its first k_res (statically dealt) tiles in its own slice of
// shared memory from one iteration to the next — lane l only ever touches
// entry l of a tile, so no synchronisation at all — and only the tiles beyond
// that (and every tile of the first iteration, which reads the caller's cloud)
// stream through HBM.  Those arrive through a per-warp ring of kRing tiles
// filled by cp.async.bulk (3 x 256 B per tile, completion on an mbarrier,
// L2 evict-first), requested kRing trips before they are consumed and without
// occupying a register; their transformed positions go back with plain
// stores.  With 20 warps x (11 resident + 3 ring) tiles x 768 B = 210 KB of
// shared memory a 2 M-point cloud keeps 52 % of its positions on the SMs.
//
// Lookups probe the map's 8-bit FILTER (1 B per slot, L2-resident) instead
// of the 16-bit tag array: one 16-entry window = two 8 B loads from the same
// or adjacent sectors, scanned with byte-parallel arithmetic.
struct RingState {   // per warp, lives across the passes of one launch
  unsigned issued;   // bulk loads issued so far   (slot = n % kRing, parity = (n / kRing) & 1)
  unsigned consumed;
  unsigned phase;    // depth 9: bit s = the parity slot s's mbarrier completes next (flips per LOADED tile)
};

template <typename F, int NW>
__device__ __forceinline__ double accumulate_points_resident(const AlignParams& P, const double* sT,
                                                             const F* sR, bool first, bool write_hit,
                                                             double* wsm, uint64_t* wbar, RingState& ring) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned wglobal = blockIdx.x * NW + (threadIdx.x >> 5);
  const unsigned wstride = gridDim.x * NW;
  const double* sx = first ? P.x0 : P.wx;
  const double* sy = first ? P.y0 : P.wy;
  const double* sz = first ? P.z0 : P.wz;
  const unsigned n_tiles = (P.n + 31u) / 32u;
  const double inv_voxel = 1.0 / P.voxel;
  const unsigned K = static_cast<unsigned>(P.k_res);
  double* const ring_sm = wsm + static_cast<size_t>(K) * 96u;  // after the resident tiles
  const uint64_t policy = l2_policy_evict_first();
  const uint64_t pol_rec = l2_policy(P.flags & kFlagRecPolicyMask);
  const uint64_t pol_filt = l2_policy((P.flags >> kFlagFiltPolicyShift) & 3u);
  const uint64_t pol_store = l2_policy((P.flags & kFlagStoreFirst) ? 1u : 0u);
  const bool rec_prefetch = (P.flags & kFlagNoRecPrefetch) == 0u;
  const bool cov_all = (P.flags & kFlagCovAllLanes) != 0u;

  // tiles: a fixed stride per warp for the first 13/16 of every pass (at least the resident ones),
  // the rest pulled from a global counter (see accumulate_points_pipelined)
  const bool dynamic = P.dynamic_tiles != 0 && n_tiles >= 8u * wstride;
  unsigned k_static = 0xffffffffu;
  if (dynamic) {
    k_static = (n_tiles - n_tiles * static_cast<unsigned>(P.dyn16) / 16u) / wstride;
    if (k_static < K) k_static = K;
  }
  const unsigned dyn_base = dynamic ? k_static * wstride : 0u;
  const unsigned chunk = static_cast<unsigned>(P.ticket_chunk);
  unsigned k_next = first ? 0u : K;  // the generator only deals the tiles that come through the ring
  unsigned tk_cur = 0, tk_next = 0;
  if (dynamic && k_next == k_static && lane == 0) tk_next = atomicAdd(&P.st->tile_counter, chunk);
  auto gen = [&]() -> unsigned {
    unsigned t;
    if (k_next < k_static) {
      t = k_next < 0x7fffffffu / wstride ? wglobal + k_next * wstride : n_tiles;
    } else {
      const unsigned sub = (k_next - k_static) & (chunk - 1u);
      if (sub == 0u) {
        tk_cur = tk_next;  // (waits for the atomic issued >= one call ago)
        if (lane == 0) tk_next = atomicAdd(&P.st->tile_counter, chunk);
      }
      t = dyn_base + __shfl_sync(0xffffffffu, tk_cur, 0) + sub;
    }
    if (t > n_tiles) t = n_tiles;
    ++k_next;
    if (k_next == k_static && lane == 0) tk_next = atomicAdd(&P.st->tile_counter, chunk);  // first chunk
    return t;
  };
  // ring: request the positions of tile `t` (no-op past the end)
  auto issue = [&](unsigned t) {
    if (t >= n_tiles) return;
    if (lane == 0) {
      const unsigned slot = ring.issued % kRing;
      double* dst = ring_sm + slot * 96u;
      uint64_t* bar = wbar + slot;
      mbar_arrive_expect_tx(bar, kTileBytes);
      bulk_load(dst, sx + static_cast<size_t>(t) * 32u, 256u, bar, policy);
      bulk_load(dst + 32, sy + static_cast<size_t>(t) * 32u, 256u, bar, policy);
      bulk_load(dst + 64, sz + static_cast<size_t>(t) * 32u, 256u, bar, policy);
    }
    ++ring.issued;
  };
  // ring: wait for the oldest outstanding tile and read this lane's point
  auto consume = [&](double& x, double& y, double& z) {
    const unsigned slot = ring.consumed % kRing;
    const unsigned parity = (ring.consumed / kRing) & 1u;
    uint64_t* bar = wbar + slot;
    if (!mbar_try_wait(bar, parity)) {
      unsigned spins = 0;
      while (!mbar_try_wait(bar, parity)) {
        __nanosleep(20);
        if (++spins > kSpinLimit) {
          atomicExch(&P.st->error, 3u);
          break;
        }
      }
    }
    const double* src = ring_sm + slot * 96u;
    x = src[lane];
    y = src[32 + lane];
    z = src[64 + lane];
    ++ring.consumed;
  };

  struct Slim {
    F px, py, pz;       // transformed position
    F dx, dy, dz;       // position - voxel centre (formed in fp64)
    uint32_t klo, khi;  // packed voxel key
    uint32_t home;
    uint32_t tag;       // 8-bit filter tag; 0 = no lookup (invalid lane / out of key range)
    uint32_t cand;      // candidate slot after the filter scan, kNoCand if none
  };
  struct RecRegs {
    uint2 key;
    float4 pa, pc;
    float2 pd;
    float4 s4;
    float2 s2;
  };
  // fp64 transform, write-back (shared memory for a resident tile, HBM otherwise), key, table address
  auto xform = [&](unsigned tile, unsigned j, double x, double y, double z, Slim& q) {
    const unsigned i = tile * 32u + lane;
    q.px = q.py = q.pz = q.dx = q.dy = q.dz = F(0);
    q.klo = q.khi = 0u;
    q.home = 0u;
    q.tag = 0u;
    q.cand = kNoCand;
    if (tile >= n_tiles || i >= P.n) return;
    transform_point_rn(sT, x, y, z);
    if (j < K) {
      double* dst = wsm + j * 96u;
      dst[lane] = x;
      dst[32 + lane] = y;
      dst[64 + lane] = z;
    } else {
      st_f64_hint(P.wx + i, x, pol_store);
      st_f64_hint(P.wy + i, y, pol_store);
      st_f64_hint(P.wz + i, z, pol_store);
    }
    const int kx = voxel_coord(x, P.voxel, inv_voxel);
    const int ky = voxel_coord(y, P.voxel, inv_voxel);
    const int kz = voxel_coord(z, P.voxel, inv_voxel);
    if (coord_in_range(kx) && coord_in_range(ky) && coord_in_range(kz)) {
      const uint64_t key = pack_key(kx, ky, kz);
      const SlotAddr ad = slot_addr(key, P.n_slots);
      q.klo = static_cast<uint32_t>(key);
      q.khi = static_cast<uint32_t>(key >> 32);
      q.home = ad.home;
      q.tag = filter_tag(ad.tag);
      // residual against the voxel mean is formed relative to the voxel centre
      q.px = F(x); q.py = F(y); q.pz = F(z);
      q.dx = F(x - __dmul_rn(static_cast<double>(kx) + 0.5, P.voxel));
      q.dy = F(y - __dmul_rn(static_cast<double>(ky) + 0.5, P.voxel));
      q.dz = F(z - __dmul_rn(static_cast<double>(kz) + 0.5, P.voxel));
    }
  };
  auto window = [&](const Slim& q) -> uint4 {
    if (q.tag == 0u) return make_uint4(0u, 0u, 0u, 0u);
    return filter_window(P.filt, q.home, pol_filt);
  };
  auto scan = [&](Slim& q, uint4 w) {
    if (q.tag == 0u) return;
    uint32_t b0 = q.home & ~7u;
    uint32_t r = scan_filter_window(w, q.home & 7u, q.tag);
    uint32_t scanned = 16u - (q.home & 7u);
    while (r == kMore && scanned < P.n_slots) {  // (about 1e-6 of the lookups at load factor 1/4)
      b0 += 16u;
      if (b0 >= P.n_slots) b0 -= P.n_slots;
      const uint2 lo = __ldg(reinterpret_cast<const uint2*>(P.filt + b0));
      const uint2 hi = __ldg(reinterpret_cast<const uint2*>(P.filt + b0 + 8));
      r = scan_filter_window(make_uint4(lo.x, lo.y, hi.x, hi.y), 0u, q.tag);
      scanned += 16u;
    }
    if (r < 16u) {
      uint32_t c = b0 + r;
      if (c >= P.n_slots) c -= P.n_slots;
      q.cand = c;
    }
  };
  auto issue_record = [&](const Slim& q, unsigned tile_of_q, RecRegs& r) {
    r.key = make_uint2(0u, 0u);
    const unsigned i = tile_of_q * 32u + lane;
    if (q.cand != kNoCand) {
      const float4* rec = reinterpret_cast<const float4*>(P.slots + q.cand);
      if (rec_prefetch) prefetch_record(rec);  // (evict_last: the voxels a registration keeps touching stay in L2)
      r.key = ldg_u2_hint(rec, pol_rec);      // key
      r.pa = ldg_f4_hint(rec + 1, pol_rec);   // mx my mz -
      r.pc = ldg_f4_hint(rec + 2, pol_rec);   // c00 c01 c02 c11
      r.pd = ldg_f2_hint(rec + 3, pol_rec);   // c12 c22
    }
    if (q.cand != kNoCand || (cov_all && tile_of_q < n_tiles && i < P.n)) {
      r.s4 = __ldcs(P.c4 + i);
      r.s2 = __ldcs(P.c2 + i);
    }
  };
  // raw positions of the j-th tile of this warp's pass: the resident copy, or the head of the ring
  // (whose slot is then refilled with the tile kRing places further on)
  unsigned q0, q1, q2;  // tiles in the ring, oldest first
  static_assert(kRing == 3, "the ring's tile queue is three registers");
  auto stage_in = [&](unsigned j, unsigned& tile, Slim& q) {
    double x = 0.0, y = 0.0, z = 0.0;
    const bool resident = !first && j < K;
    if (resident) {
      tile = j < 0x7fffffffu / wstride ? wglobal + j * wstride : n_tiles;
      if (tile < n_tiles) {
        const double* src = wsm + j * 96u;
        x = src[lane];
        y = src[32 + lane];
        z = src[64 + lane];
      }
    } else {
      tile = q0;
      if (tile < n_tiles) consume(x, y, z);
    }
    xform(tile, j, x, y, z, q);
    if (!resident) {
      q0 = q1;
      q1 = q2;
      q2 = gen();
      __syncwarp();  // every lane has read (and used) its entry of the slot that is refilled now
      issue(q2);
    }
  };

  double acc = 0.0;
  F v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = F(0);
  // cross-proxy: the positions this pass bulk-loads were stored through the generic proxy
  // (by any SM) before the hand-off this thread has just acquired
  if (!first && lane == 0) fence_proxy_async();
  q0 = gen(); issue(q0);
  q1 = gen(); issue(q1);
  q2 = gen(); issue(q2);
  Slim cur, nxt;
  RecRegs rec;
  uint4 tagw;
  unsigned tile = 0, tile_n = 0, j = 0;
  stage_in(j++, tile, cur);
  scan(cur, window(cur));
  issue_record(cur, tile, rec);
  stage_in(j++, tile_n, nxt);
  tagw = window(nxt);
  while (tile < n_tiles) {
    const unsigned i = tile * 32u + lane;
    // ---- 1. the record + covariance of the current tile
    float4 pa = rec.pa, pc = rec.pc;
    float2 pd = rec.pd;
    bool hit = false;
    if (cur.cand != kNoCand) {
      hit = rec.key.x == cur.klo && rec.key.y == cur.khi;
      if (!hit) {
        const uint64_t key = (static_cast<uint64_t>(cur.khi) << 32) | cur.klo;
        const VoxelSlot* far = resolve_probe_filter(P.filt, P.slots, P.n_slots, key,
                                                    next_slot(cur.cand, P.n_slots), cur.tag);
        if (far != nullptr) {
          hit = true;
          pa = __ldg(reinterpret_cast<const float4*>(far) + 1);
          pc = __ldg(reinterpret_cast<const float4*>(far) + 2);
          pd = __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float4*>(far) + 3));
        }
      }
    }
    if (write_hit && i < P.n) P.hit[i] = hit ? 1 : 0;
    if (hit) {
      F cr[6];
      rotate_sym<F>(sR, F(rec.s4.x), F(rec.s4.y), F(rec.s4.z), F(rec.s4.w), F(rec.s2.x), F(rec.s2.y), cr);
      point_terms<F, true>(cur.px, cur.py, cur.pz, cur.dx - F(pa.x), cur.dy - F(pa.y), cur.dz - F(pa.z),
                           cr[0] + F(pc.x), cr[1] + F(pc.y), cr[2] + F(pc.z), cr[3] + F(pc.w),
                           cr[4] + F(pd.x), cr[5] + F(pd.y), v);
    } else {
#pragma unroll
      for (int k = 0; k < 28; ++k) v[k] = F(0);
    }
    acc += static_cast<double>(warp_reduce_scatter32<F>(v, lane));
    // ---- 2. the filter window of the next tile; its record + covariance loads go out
    scan(nxt, tagw);
    issue_record(nxt, tile_n, rec);
    cur = nxt;
    tile = tile_n;
    // ---- 3. the tile after that: positions (shared memory), transform, key; its window load goes out
    stage_in(j++, tile_n, nxt);
    tagw = window(nxt);
  }
  return acc;
}


// ------------------------------------------------------------------------
// Large clouds (depth 6): role-specialised warps around a shared-memory hit
// queue.
//
// With one thread per point and 27 % of the points finding a voxel (the dense
// 0.1 m config), a warp that does "look up, then linearise" runs the ~450
// instructions of the per-point algebra + reduce-scatter for 32 lanes of which
// 9 have a correspondence; profiles/r2_prof_align_d5.md: 820 warp instructions
// per 32-point tile, issue slots 41 % busy, DRAM 20 % of peak — the kernel is
// bound by instruction issue and dependent-load latency, not by HBM.
// Here the warps of a CTA split into
//   producers  transform a tile (fp64, exact order), form keys, probe the 8-bit
//              filter, and push every candidate correspondence — position,
//              offset from the voxel centre, key, slot, point index: 40 B —
//              into a ring in shared memory (one warp-aggregated reservation);
//   consumers  pop DENSE batches of 32 candidates, gather the 64 B records and
//              the source covariances (next batch's loads in flight while the
//              current one is linearised), verify keys, accumulate J^T W J / J^T W r.
// The ring is 32 batch slots x 32 entries (SoA, conflict free); a slot carries a
// `ready` count (entries written) and a generation (times consumed), so
// producers and consumers only ever spin on shared memory, with CTA-scope
// fences.  Producers that run out of tiles turn into consumers; the share of
// consumer warps follows the hit rate the CTA saw in the previous iteration.
// Working positions of the CTA's first k_res tiles stay in shared memory across
// iterations (as in depth 5, but owned by the CTA: any producer warp takes the
// next tile from a shared counter).
constexpr unsigned kQ = 1024;   // ring entries
constexpr unsigned kNB = kQ / 32;  // batch slots

struct HitQueue {
  float px[kQ], py[kQ], pz[kQ], dx[kQ], dy[kQ], dz[kQ];
  uint32_t klo[kQ], khi[kQ], cand[kQ], idx[kQ];
  unsigned ready[kNB];  // entries written into the batch that occupies the slot
  unsigned gen[kNB];    // times the slot has been consumed in this pass
  unsigned tail;        // entries reserved
  unsigned head;        // batches claimed
  unsigned prod_done;   // producer warps that have finished the pass
  unsigned next_m;      // next of the CTA's tiles to hand out
  unsigned n_cons;      // consumer warps of the current pass
  unsigned pad[3];
};
static_assert(sizeof(HitQueue) % 16 == 0, "HitQueue is followed by 16 B aligned data");

__device__ __forceinline__ unsigned ld_volatile_shared(const unsigned* p) {
  return *reinterpret_cast<const volatile unsigned*>(p);
}

template <typename F, int NW>
__device__ __forceinline__ double accumulate_points_roles(const AlignParams& P, const double* sT, const F* sR,
                                                          bool first, bool write_hit, double* res_sm,
                                                          HitQueue* Q) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned G = gridDim.x, b = blockIdx.x;
  const unsigned n_tiles = (P.n + 31u) / 32u;
  const unsigned M = b < n_tiles ? (n_tiles - b + G - 1u) / G : 0u;  // this CTA's tiles: b, b + G, ...
  const unsigned R = static_cast<unsigned>(P.k_res);                 // of which the first R are resident
  const double inv_voxel = 1.0 / P.voxel;
  const unsigned n_cons = ld_volatile_shared(&Q->n_cons);
  const unsigned n_prod = NW - n_cons;
  const double* sx = first ? P.x0 : P.wx;
  const double* sy = first ? P.y0 : P.wy;
  const double* sz = first ? P.z0 : P.wz;
  const uint64_t pol_rec = l2_policy(P.flags & kFlagRecPolicyMask);
  const uint64_t pol_filt = l2_policy((P.flags >> kFlagFiltPolicyShift) & 3u);
  const bool rec_prefetch = (P.flags & kFlagNoRecPrefetch) == 0u;
  double acc = 0.0;

  if (warp >= n_cons) {
    // ======================================================== producer
    struct Slim {
      F px, py, pz, dx, dy, dz;
      uint32_t klo, khi, home, tag;
      unsigned i;  // point index
    };
    auto claim = [&]() -> unsigned {  // index m of the next tile of this CTA (>= M: none left)
      unsigned m = 0;
      if (lane == 0) m = atomicAdd(&Q->next_m, 1u);
      return __shfl_sync(0xffffffffu, m, 0);
    };
    auto load_pos = [&](unsigned m, double& x, double& y, double& z) {
      x = y = z = 0.0;
      if (m >= M) return;
      const unsigned i = (b + G * m) * 32u + lane;
      if (!first && m < R) {
        const double* src = res_sm + static_cast<size_t>(m) * 96u;
        x = src[lane];
        y = src[32 + lane];
        z = src[64 + lane];
      } else if (i < P.n) {
        x = first ? __ldcs(sx + i) : __ldcg(sx + i);
        y = first ? __ldcs(sy + i) : __ldcg(sy + i);
        z = first ? __ldcs(sz + i) : __ldcg(sz + i);
      }
    };
    auto xform = [&](unsigned m, double x, double y, double z, Slim& q) {
      q.tag = 0u;
      q.home = 0u;
      if (m >= M) return;
      const unsigned i = (b + G * m) * 32u + lane;
      q.i = i;
      if (i >= P.n) return;
      transform_point_rn(sT, x, y, z);
      if (m < R) {
        double* dst = res_sm + static_cast<size_t>(m) * 96u;
        dst[lane] = x;
        dst[32 + lane] = y;
        dst[64 + lane] = z;
      } else {
        P.wx[i] = x;
        P.wy[i] = y;
        P.wz[i] = z;
      }
      const int kx = voxel_coord(x, P.voxel, inv_voxel);
      const int ky = voxel_coord(y, P.voxel, inv_voxel);
      const int kz = voxel_coord(z, P.voxel, inv_voxel);
      if (coord_in_range(kx) && coord_in_range(ky) && coord_in_range(kz)) {
        const uint64_t key = pack_key(kx, ky, kz);
        const SlotAddr ad = slot_addr(key, P.n_slots);
        q.klo = static_cast<uint32_t>(key);
        q.khi = static_cast<uint32_t>(key >> 32);
        q.home = ad.home;
        q.tag = filter_tag(ad.tag);
        q.px = F(x); q.py = F(y); q.pz = F(z);
        // the residual against the voxel mean is formed relative to the voxel centre
        q.dx = F(x - __dmul_rn(static_cast<double>(kx) + 0.5, P.voxel));
        q.dy = F(y - __dmul_rn(static_cast<double>(ky) + 0.5, P.voxel));
        q.dz = F(z - __dmul_rn(static_cast<double>(kz) + 0.5, P.voxel));
      }
    };
    auto window = [&](const Slim& q) -> uint4 {
      if (q.tag == 0u) return make_uint4(0u, 0u, 0u, 0u);
      return filter_window(P.filt, q.home, pol_filt);
    };
    // candidate slot of q given its first window (kNoCand: the voxel is not in the map)
    auto scan = [&](const Slim& q, uint4 w) -> uint32_t {
      if (q.tag == 0u) return kNoCand;
      uint32_t b0 = q.home & ~7u;
      uint32_t r = scan_filter_window(w, q.home & 7u, q.tag);
      uint32_t scanned = 16u - (q.home & 7u);
      while (r == kMore && scanned < P.n_slots) {
        b0 += 16u;
        if (b0 >= P.n_slots) b0 -= P.n_slots;
        const uint2 lo = __ldg(reinterpret_cast<const uint2*>(P.filt + b0));
        const uint2 hi = __ldg(reinterpret_cast<const uint2*>(P.filt + b0 + 8));
        r = scan_filter_window(make_uint4(lo.x, lo.y, hi.x, hi.y), 0u, q.tag);
        scanned += 16u;
      }
      if (r >= 16u) return kNoCand;
      uint32_t c = b0 + r;
      if (c >= P.n_slots) c -= P.n_slots;
      return c;
    };
    // push the candidates of a tile: one reservation per warp, entries of a batch become visible
    // to the consumers through the batch slot's `ready` count
    auto push = [&](const Slim& q, uint32_t cand) {
      const bool has = cand != kNoCand;
      const unsigned mask = __ballot_sync(0xffffffffu, has);
      if (write_hit && !has && q.i < P.n) P.hit[q.i] = 0;
      if (mask == 0u) return;
      const unsigned count = __popc(mask);
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(&Q->tail, count);
      base = __shfl_sync(0xffffffffu, base, 0);
      const unsigned e = base + __popc(mask & ((1u << lane) - 1u));  // this lane's entry
      const unsigned batch = e >> 5, slot = batch & (kNB - 1u), g = batch / kNB;
      if (has) {
        unsigned spins = 0;
        while (ld_volatile_shared(&Q->gen[slot]) != g) {  // the slot still holds an unconsumed older batch
          __nanosleep(20);
          if (++spins > kSpinLimit) {
            atomicExch(&P.st->error, 4u);
            break;
          }
        }
        __threadfence_block();
        const unsigned pos = e & (kQ - 1u);
        Q->px[pos] = static_cast<float>(q.px);
        Q->py[pos] = static_cast<float>(q.py);
        Q->pz[pos] = static_cast<float>(q.pz);
        Q->dx[pos] = static_cast<float>(q.dx);
        Q->dy[pos] = static_cast<float>(q.dy);
        Q->dz[pos] = static_cast<float>(q.dz);
        Q->klo[pos] = q.klo;
        Q->khi[pos] = q.khi;
        Q->cand[pos] = cand;
        Q->idx[pos] = q.i;
        __threadfence_block();  // entries before the count below
      }
      __syncwarp();
      // the reservation touches at most two batches
      const unsigned b_first = base >> 5, b_last = (base + count - 1u) >> 5;
      if (lane == 0) {
        const unsigned n_first = b_first == b_last ? count : 32u - (base & 31u);
        atomicAdd(&Q->ready[b_first & (kNB - 1u)], n_first);
        if (b_last != b_first) atomicAdd(&Q->ready[b_last & (kNB - 1u)], count - n_first);
      }
    };

    // three tiles in flight per warp: positions requested (c), filter window requested (w), scanned now
    unsigned m_c = claim();
    double rx, ry, rz;
    load_pos(m_c, rx, ry, rz);
    Slim qw;
    qw.tag = 0u; qw.home = 0u; qw.i = 0xffffffffu;
    qw.px = qw.py = qw.pz = qw.dx = qw.dy = qw.dz = F(0);
    qw.klo = qw.khi = 0u;
    uint4 tagw = make_uint4(0u, 0u, 0u, 0u);
    bool have_w = false;
    for (;;) {
      const unsigned m_x = m_c;  // tile whose positions have arrived
      double x = rx, y = ry, z = rz;
      const bool have_x = m_x < M;
      if (have_x) {
        m_c = claim();
        load_pos(m_c, rx, ry, rz);
      }
      // finish the tile whose window was requested a trip ago
      if (have_w) push(qw, scan(qw, tagw));
      if (!have_x) break;
      xform(m_x, x, y, z, qw);
      tagw = window(qw);
      have_w = true;
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicAdd(&Q->prod_done, 1u);
  }

  // ========================================================== consumer
  // (every warp ends up here: the consumer warps from the start, the producers once the CTA's
  // tiles are handed out)
  {
    struct Entry {
      F px, py, pz, dx, dy, dz;
      uint32_t klo, khi, cand, idx;
      bool valid;
    };
    struct RecRegs {
      uint2 key;
      float4 pa, pc;
      float2 pd;
      float4 s4;
      float2 s2;
    };
    auto claim_batch = [&]() -> unsigned {
      unsigned c = 0;
      if (lane == 0) c = atomicAdd(&Q->head, 1u);
      return __shfl_sync(0xffffffffu, c, 0);
    };
    // entries of batch c: > 0 when they can be read, 0 when the pass has no such batch, -1 when
    // not yet (only returned if !block)
    auto poll = [&](unsigned c, bool block) -> int {
      const unsigned slot = c & (kNB - 1u), g = c / kNB;
      unsigned spins = 0;
      for (;;) {
        int res = -1;
        if (lane == 0) {
          if (ld_volatile_shared(&Q->gen[slot]) == g) {
            const unsigned r = ld_volatile_shared(&Q->ready[slot]);
            if (r == 32u) {
              res = 32;
            } else if (ld_volatile_shared(&Q->prod_done) == n_prod) {
              __threadfence_block();
              const unsigned t = ld_volatile_shared(&Q->tail);
              if (c * 32u >= t) res = 0;
              else {
                const unsigned need = t - c * 32u < 32u ? t - c * 32u : 32u;
                if (ld_volatile_shared(&Q->ready[slot]) == need) res = static_cast<int>(need);
              }
            }
          } else if (ld_volatile_shared(&Q->prod_done) == n_prod) {
            __threadfence_block();
            if (c * 32u >= ld_volatile_shared(&Q->tail)) res = 0;
          }
        }
        res = __shfl_sync(0xffffffffu, res, 0);
        if (res >= 0 || !block) {
          if (res > 0) __threadfence_block();
          return res;
        }
        __nanosleep(40);
        if (++spins > kSpinLimit) {
          atomicExch(&P.st->error, 5u);
          return 0;
        }
      }
    };
    auto read_entries = [&](unsigned c, int need, Entry& e) {
      const unsigned pos = (c * 32u + lane) & (kQ - 1u);
      e.valid = static_cast<int>(lane) < need;
      e.px = F(Q->px[pos]); e.py = F(Q->py[pos]); e.pz = F(Q->pz[pos]);
      e.dx = F(Q->dx[pos]); e.dy = F(Q->dy[pos]); e.dz = F(Q->dz[pos]);
      e.klo = Q->klo[pos]; e.khi = Q->khi[pos];
      e.cand = e.valid ? Q->cand[pos] : kNoCand;
      e.idx = Q->idx[pos];
      __syncwarp();
      if (lane == 0) {  // hand the slot back to the producers
        const unsigned slot = c & (kNB - 1u);
        Q->ready[slot] = 0u;
        __threadfence_block();
        *reinterpret_cast<volatile unsigned*>(&Q->gen[slot]) = c / kNB + 1u;
      }
    };
    auto issue = [&](const Entry& e, RecRegs& r) {
      r.key = make_uint2(0u, 0u);
      if (e.cand != kNoCand) {
        const float4* rec = reinterpret_cast<const float4*>(P.slots + e.cand);
        if (rec_prefetch) prefetch_record(rec);
        r.key = ldg_u2_hint(rec, pol_rec);
        r.pa = ldg_f4_hint(rec + 1, pol_rec);
        r.pc = ldg_f4_hint(rec + 2, pol_rec);
        r.pd = ldg_f2_hint(rec + 3, pol_rec);
        r.s4 = __ldcs(P.c4 + e.idx);
        r.s2 = __ldcs(P.c2 + e.idx);
      }
    };
    F v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = F(0);
    auto compute = [&](const Entry& e, const RecRegs& r) {
      float4 pa = r.pa, pc = r.pc;
      float2 pd = r.pd;
      bool hit = false;
      if (e.cand != kNoCand) {
        hit = r.key.x == e.klo && r.key.y == e.khi;
        if (!hit) {  // 8-bit filter collision: walk on, slowly
          const uint64_t key = (static_cast<uint64_t>(e.khi) << 32) | e.klo;
          const SlotAddr ad = slot_addr(key, P.n_slots);
          const VoxelSlot* far = resolve_probe_filter(P.filt, P.slots, P.n_slots, key,
                                                      next_slot(e.cand, P.n_slots), filter_tag(ad.tag));
          if (far != nullptr) {
            hit = true;
            pa = __ldg(reinterpret_cast<const float4*>(far) + 1);
            pc = __ldg(reinterpret_cast<const float4*>(far) + 2);
            pd = __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float4*>(far) + 3));
          }
        }
        if (write_hit) P.hit[e.idx] = hit ? 1 : 0;
      }
      if (hit) {
        F cr[6];
        rotate_sym<F>(sR, F(r.s4.x), F(r.s4.y), F(r.s4.z), F(r.s4.w), F(r.s2.x), F(r.s2.y), cr);
        point_terms<F, true>(e.px, e.py, e.pz, e.dx - F(pa.x), e.dy - F(pa.y), e.dz - F(pa.z),
                             cr[0] + F(pc.x), cr[1] + F(pc.y), cr[2] + F(pc.z), cr[3] + F(pc.w),
                             cr[4] + F(pd.x), cr[5] + F(pd.y), v);
      } else {
#pragma unroll
        for (int k = 0; k < 28; ++k) v[k] = F(0);
      }
      acc += static_cast<double>(warp_reduce_scatter32<F>(v, lane));
    };

    Entry cur, nxt;
    RecRegs rec, nrec;
    bool have_cur = false;
    unsigned c = claim_batch();
    for (;;) {
      const int need = poll(c, !have_cur);
      if (need > 0) {
        read_entries(c, need, nxt);
        issue(nxt, nrec);
      }
      if (have_cur) compute(cur, rec);
      if (need > 0) {
        cur = nxt;
        rec = nrec;
        have_cur = true;
        c = claim_batch();
      } else if (need == 0) {
        break;
      } else {
        have_cur = false;
      }
    }
  }
  return acc;
}


// ------------------------------------------------------------------------
// Large clouds (depth 7): phase-split passes around a per-CTA hit list.
//
// Depth 6 showed that the producer / consumer split is right (dense algebra,
// landing registers only for points that found a voxel) and that its cost was
// the synchronisation: shared-memory atomics, fences and polling in every trip.
// Here a pass has two phases separated by ONE __syncthreads:
//   produce   every warp walks its share of the CTA's tiles, U tiles per trip
//             (the loads of a trip are issued one trip ahead: positions, then
//             filter windows), and appends every candidate correspondence —
//             40 B — to the CTA's hit list with one shared-memory atomic per trip;
//   consume   every warp takes dense batches of 32 entries (static deal), the
//             next batch's record + covariance loads in flight while the current
//             one is linearised.
// The list lives in shared memory (5.6 k entries: a hit rate of up to ~40 % of a
// 2 M-point cloud's 13.5 k points per CTA); what does not fit spills to a
// CTA-private region of HBM (the all-hit regime), read back through L2.
struct HitList {  // SoA words: field k of entry e at base[k * pitch + e]; fields: px py pz dx dy dz klo khi cand idx
  uint32_t* sm;      // shared memory, `cap` entries per field
  uint32_t* gm;      // the CTA's spill region, `gcap` entries per field
  unsigned cap, gcap;
};

template <typename F, int NW, int U>
__device__ __forceinline__ double accumulate_points_split(const AlignParams& P, const double* sT, const F* sR,
                                                          bool first, bool write_hit, const HitList& L,
                                                          unsigned* s_tail) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned G = gridDim.x, b = blockIdx.x;
  const unsigned n_tiles = (P.n + 31u) / 32u;
  const unsigned M = b < n_tiles ? (n_tiles - b + G - 1u) / G : 0u;  // this CTA's tiles: b, b + G, ...
  const double inv_voxel = 1.0 / P.voxel;
  const double* sx = first ? P.x0 : P.wx;
  const double* sy = first ? P.y0 : P.wy;
  const double* sz = first ? P.z0 : P.wz;
  const uint64_t pol_rec = l2_policy(P.flags & kFlagRecPolicyMask);
  const uint64_t pol_filt = l2_policy((P.flags >> kFlagFiltPolicyShift) & 3u);
  const bool rec_prefetch = (P.flags & kFlagNoRecPrefetch) == 0u;

  // =========================================================== produce
  {
    struct Slim {
      F px, py, pz, dx, dy, dz;
      uint32_t klo, khi, home, tag;
      unsigned i;  // point index (0xffffffff: no such point)
    };
    // trip k of this warp covers the CTA's tiles m = (k * NW + warp) * U + u
    auto tile_of = [&](unsigned k, int u) -> unsigned {
      const unsigned m = (k * NW + warp) * U + static_cast<unsigned>(u);
      return m < M ? b + G * m : 0xffffffffu;
    };
    auto load_pos = [&](unsigned tile, double& x, double& y, double& z) {
      x = y = z = 0.0;
      const unsigned i = tile * 32u + lane;
      if (tile != 0xffffffffu && i < P.n) {
        x = first ? __ldcs(sx + i) : __ldcg(sx + i);
        y = first ? __ldcs(sy + i) : __ldcg(sy + i);
        z = first ? __ldcs(sz + i) : __ldcg(sz + i);
      }
    };
    auto xform = [&](unsigned tile, double x, double y, double z, Slim& q) {
      q.tag = 0u;
      q.home = 0u;
      q.i = 0xffffffffu;
      const unsigned i = tile * 32u + lane;
      if (tile == 0xffffffffu || i >= P.n) return;
      q.i = i;
      transform_point_rn(sT, x, y, z);
      P.wx[i] = x;
      P.wy[i] = y;
      P.wz[i] = z;
      const int kx = voxel_coord(x, P.voxel, inv_voxel);
      const int ky = voxel_coord(y, P.voxel, inv_voxel);
      const int kz = voxel_coord(z, P.voxel, inv_voxel);
      if (coord_in_range(kx) && coord_in_range(ky) && coord_in_range(kz)) {
        const uint64_t key = pack_key(kx, ky, kz);
        const SlotAddr ad = slot_addr(key, P.n_slots);
        q.klo = static_cast<uint32_t>(key);
        q.khi = static_cast<uint32_t>(key >> 32);
        q.home = ad.home;
        q.tag = filter_tag(ad.tag);
        q.px = F(x); q.py = F(y); q.pz = F(z);
        // the residual against the voxel mean is formed relative to the voxel centre
        q.dx = F(x - __dmul_rn(static_cast<double>(kx) + 0.5, P.voxel));
        q.dy = F(y - __dmul_rn(static_cast<double>(ky) + 0.5, P.voxel));
        q.dz = F(z - __dmul_rn(static_cast<double>(kz) + 0.5, P.voxel));
      }
    };
    auto window = [&](const Slim& q) -> uint4 {
      if (q.tag == 0u) return make_uint4(0u, 0u, 0u, 0u);
      return filter_window(P.filt, q.home, pol_filt);
    };
    auto scan = [&](const Slim& q, uint4 w) -> uint32_t {
      if (q.tag == 0u) return kNoCand;
      uint32_t b0 = q.home & ~7u;
      uint32_t r = scan_filter_window(w, q.home & 7u, q.tag);
      uint32_t scanned = 16u - (q.home & 7u);
      while (r == kMore && scanned < P.n_slots) {
        b0 += 16u;
        if (b0 >= P.n_slots) b0 -= P.n_slots;
        const uint2 lo = __ldg(reinterpret_cast<const uint2*>(P.filt + b0));
        const uint2 hi = __ldg(reinterpret_cast<const uint2*>(P.filt + b0 + 8));
        r = scan_filter_window(make_uint4(lo.x, lo.y, hi.x, hi.y), 0u, q.tag);
        scanned += 16u;
      }
      if (r >= 16u) return kNoCand;
      uint32_t c = b0 + r;
      if (c >= P.n_slots) c -= P.n_slots;
      return c;
    };
    auto store_entry = [&](unsigned e, const Slim& q, uint32_t cand) {
      uint32_t* w = L.sm + e;
      unsigned pitch = L.cap;
      if (e >= L.cap) {
        w = L.gm + (e - L.cap);
        pitch = L.gcap;
      }
      w[0] = __float_as_uint(static_cast<float>(q.px));
      w[pitch] = __float_as_uint(static_cast<float>(q.py));
      w[2 * pitch] = __float_as_uint(static_cast<float>(q.pz));
      w[3 * pitch] = __float_as_uint(static_cast<float>(q.dx));
      w[4 * pitch] = __float_as_uint(static_cast<float>(q.dy));
      w[5 * pitch] = __float_as_uint(static_cast<float>(q.dz));
      w[6 * pitch] = q.klo;
      w[7 * pitch] = q.khi;
      w[8 * pitch] = cand;
      w[9 * pitch] = q.i;
    };

    const unsigned trips = (M + NW * U - 1u) / (NW * U);
    double rx[U], ry[U], rz[U];
    Slim q[U];
    uint4 tw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      load_pos(tile_of(0, u), rx[u], ry[u], rz[u]);
      q[u].tag = 0u; q[u].home = 0u; q[u].i = 0xffffffffu;
      q[u].px = q[u].py = q[u].pz = q[u].dx = q[u].dy = q[u].dz = F(0);
      q[u].klo = q[u].khi = 0u;
      tw[u] = make_uint4(0u, 0u, 0u, 0u);
    }
    // trip k: scan + push the tiles of trip k - 1 (windows requested a trip ago), transform the tiles
    // of trip k (positions requested a trip ago) and request their windows, request trip k + 1's positions
    for (unsigned k = 0; k <= trips; ++k) {
      uint32_t cand[U];
      unsigned mask[U], count = 0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        cand[u] = scan(q[u], tw[u]);
        mask[u] = __ballot_sync(0xffffffffu, cand[u] != kNoCand);
        count += __popc(mask[u]);
        if (write_hit && cand[u] == kNoCand && q[u].i < P.n) P.hit[q[u].i] = 0;
      }
      if (count != 0u) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(s_tail, count);
        base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (cand[u] != kNoCand) store_entry(base + __popc(mask[u] & ((1u << lane) - 1u)), q[u], cand[u]);
          base += __popc(mask[u]);
        }
      }
      if (k == trips) break;
      double x[U], y[U], z[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        x[u] = rx[u]; y[u] = ry[u]; z[u] = rz[u];
      }
#pragma unroll
      for (int u = 0; u < U; ++u) load_pos(k + 1 < trips ? tile_of(k + 1, u) : 0xffffffffu, rx[u], ry[u], rz[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        xform(tile_of(k, u), x[u], y[u], z[u], q[u]);
        tw[u] = window(q[u]);
      }
    }
  }
  __syncthreads();

  // =========================================================== consume
  double acc = 0.0;
  {
    struct Entry {
      F px, py, pz, dx, dy, dz;
      uint32_t klo, khi, cand, idx;
    };
    struct RecRegs {
      uint2 key;
      float4 pa, pc;
      float2 pd;
      float4 s4;
      float2 s2;
    };
    const unsigned total = *reinterpret_cast<volatile unsigned*>(s_tail);
    const unsigned n_batches = (total + 31u) / 32u;
    auto fetch = [&](unsigned j, Entry& e, RecRegs& r) {  // entries of batch j + their record / covariance loads
      unsigned idx = j * 32u + lane;
      e.cand = kNoCand;
      r.key = make_uint2(0u, 0u);
      if (j >= n_batches || idx >= total) return;
      const uint32_t* w = L.sm + idx;
      unsigned pitch = L.cap;
      if (idx >= L.cap) {
        w = L.gm + (idx - L.cap);
        pitch = L.gcap;
      }
      e.px = F(__uint_as_float(w[0])); e.py = F(__uint_as_float(w[pitch])); e.pz = F(__uint_as_float(w[2 * pitch]));
      e.dx = F(__uint_as_float(w[3 * pitch])); e.dy = F(__uint_as_float(w[4 * pitch]));
      e.dz = F(__uint_as_float(w[5 * pitch]));
      e.klo = w[6 * pitch]; e.khi = w[7 * pitch]; e.cand = w[8 * pitch]; e.idx = w[9 * pitch];
      const float4* rec = reinterpret_cast<const float4*>(P.slots + e.cand);
      if (rec_prefetch) prefetch_record(rec);
      r.key = ldg_u2_hint(rec, pol_rec);
      r.pa = ldg_f4_hint(rec + 1, pol_rec);
      r.pc = ldg_f4_hint(rec + 2, pol_rec);
      r.pd = ldg_f2_hint(rec + 3, pol_rec);
      r.s4 = __ldcs(P.c4 + e.idx);
      r.s2 = __ldcs(P.c2 + e.idx);
    };
    F v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = F(0);
    auto compute = [&](const Entry& e, const RecRegs& r) {
      float4 pa = r.pa, pc = r.pc;
      float2 pd = r.pd;
      bool hit = false;
      if (e.cand != kNoCand) {
        hit = r.key.x == e.klo && r.key.y == e.khi;
        if (!hit) {  // 8-bit filter collision: walk on, slowly
          const uint64_t key = (static_cast<uint64_t>(e.khi) << 32) | e.klo;
          const SlotAddr ad = slot_addr(key, P.n_slots);
          const VoxelSlot* far = resolve_probe_filter(P.filt, P.slots, P.n_slots, key,
                                                      next_slot(e.cand, P.n_slots), filter_tag(ad.tag));
          if (far != nullptr) {
            hit = true;
            pa = __ldg(reinterpret_cast<const float4*>(far) + 1);
            pc = __ldg(reinterpret_cast<const float4*>(far) + 2);
            pd = __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float4*>(far) + 3));
          }
        }
        if (write_hit) P.hit[e.idx] = hit ? 1 : 0;
      }
      if (hit) {
        F cr[6];
        rotate_sym<F>(sR, F(r.s4.x), F(r.s4.y), F(r.s4.z), F(r.s4.w), F(r.s2.x), F(r.s2.y), cr);
        point_terms<F, true>(e.px, e.py, e.pz, e.dx - F(pa.x), e.dy - F(pa.y), e.dz - F(pa.z),
                             cr[0] + F(pc.x), cr[1] + F(pc.y), cr[2] + F(pc.z), cr[3] + F(pc.w),
                             cr[4] + F(pd.x), cr[5] + F(pd.y), v);
      } else {
#pragma unroll
        for (int k = 0; k < 28; ++k) v[k] = F(0);
      }
      acc += static_cast<double>(warp_reduce_scatter32<F>(v, lane));
    };
    Entry cur, nxt;
    RecRegs rec, nrec;
    unsigned j = warp;
    fetch(j, cur, rec);
    while (j < n_batches) {
      fetch(j + NW, nxt, nrec);
      compute(cur, rec);
      cur = nxt;
      rec = nrec;
      j += NW;
    }
  }
  return acc;
}


// ------------------------------------------------------------------------
// Large clouds (depth 8): the depth-4 register pipeline with the candidate
// correspondences PARKED per warp until there are 32 of them.
//
// What depths 5-7 taught (profiles/r2_align_experiments.md): the kernel is bound
// by instruction issue and latency, not by HBM; the per-point algebra + the
// reduce-scatter (~430 of the ~700 warp instructions of a depth-4 trip) run for
// 32 lanes of which 27 % have a correspondence on the dense config; and every
// form of cross-warp hand-over (queues, fences, barriers) costs more than the
// dense algebra saves.  So every warp stays on its own, as in depth 4, but a
// trip only LOOKS UP its tile: transform, key, 8-bit filter window (requested a
// trip ahead), scan.  Lanes with a candidate park 10 words (position, offset
// from the voxel centre, key, slot, point index) in the warp's own ring in
// shared memory — no atomics, no fences: only this warp touches it — and once 32
// are parked the warp pops them as ONE dense batch: record + source covariance
// loads go out, and the batch is linearised when the next one is ready to go (or
// at the end of the pass), i.e. ~4 trips later on the dense config.  The landing
// registers of the gathers are therefore always full, the algebra and the
// 31-shuffle reduce-scatter run once per 32 correspondences instead of once per
// 32 points.
constexpr unsigned kPark = 128;  // entries of a warp's ring (a trip parks at most 32, a pop takes 32)
constexpr int kRing9 = 6;        // depth 9: tiles of positions in flight per warp
constexpr unsigned kWarpBytes9 = 10u * kPark * 4u + kRing9 * kTileBytes + 32u;  // parked words | ring | tile ids

// RING > 0: the raw positions arrive through a per-warp ring of RING tiles filled by cp.async.bulk
// (3 x 256 B per tile, completion on an mbarrier, L2 evict-first) and requested RING trips before they
// are used: with lookup-only trips (~0.5 us) one tile per warp in flight (12 KB per SM) held the
// position stream to ~1.2 TB/s (ablation, profiles/r2_align_experiments.md).  RING == 0: register loads
// one trip ahead + an L2 prefetch.
template <typename F, int NW, int RING, int U = 1>
__device__ __forceinline__ double accumulate_points_parked(const AlignParams& P, const double* sT, const F* sR,
                                                           bool first, bool write_hit, uint32_t* park /* [10][kPark] */,
                                                           double* ring_sm /* [RING][96] */, unsigned* tq /* [RING] */,
                                                           uint64_t* wbar /* [RING] */, RingState& ring) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned wglobal = blockIdx.x * NW + (threadIdx.x >> 5);
  const unsigned wstride = gridDim.x * NW;
  const double* sx = first ? P.x0 : P.wx;
  const double* sy = first ? P.y0 : P.wy;
  const double* sz = first ? P.z0 : P.wz;
  const unsigned n_tiles = (P.n + 31u) / 32u;
  const double inv_voxel = 1.0 / P.voxel;
  const uint64_t pol_rec = l2_policy(P.flags & kFlagRecPolicyMask);
  const uint64_t pol_filt = l2_policy((P.flags >> kFlagFiltPolicyShift) & 3u);
  const bool rec_prefetch = (P.flags & kFlagNoRecPrefetch) == 0u;

  // tiles: fixed stride for the first 13/16 of a pass, then tickets (see accumulate_points_pipelined)
  const bool dynamic = P.dynamic_tiles != 0 && n_tiles >= 8u * wstride;
  const unsigned k_static = dynamic ? (n_tiles - n_tiles * static_cast<unsigned>(P.dyn16) / 16u) / wstride : 0xffffffffu;
  const unsigned dyn_base = dynamic ? k_static * wstride : 0u;
  const unsigned chunk = static_cast<unsigned>(P.ticket_chunk);
  unsigned k_next = 0;
  unsigned tk_cur = 0, tk_next = 0;
  auto next_tile = [&]() -> unsigned {
    unsigned t;
    if (k_next < k_static) {
      t = wglobal + k_next * wstride;
    } else {
      const unsigned sub = (k_next - k_static) & (chunk - 1u);
      if (sub == 0u) {
        tk_cur = tk_next;  // (waits for the atomic issued >= one call ago)
        if (lane == 0) tk_next = atomicAdd(&P.st->tile_counter, chunk);
      }
      t = dyn_base + __shfl_sync(0xffffffffu, tk_cur, 0) + sub;
    }
    if (t > n_tiles) t = n_tiles;
    ++k_next;
    if (k_next == k_static && lane == 0) tk_next = atomicAdd(&P.st->tile_counter, chunk);  // first chunk
    return t;
  };
  // the positions of a statically dealt tile are pulled into L2 a few trips ahead (3 x 256 B, 64 B per
  // lane of lanes 0..11): a lookup trip is shorter than an HBM round trip
  const bool abl_noprobe = (P.flags & kFlagNoProbe) != 0u, abl_nogather = (P.flags & kFlagNoGather) != 0u;
  const bool abl_nostore = (P.flags & kFlagNoStore) != 0u, abl_noprefetch = (P.flags & kFlagNoPrefetch) != 0u;
  auto prefetch_pos = [&](unsigned k) {
    const unsigned tp = wglobal + k * wstride;
    if (!abl_noprefetch && k < k_static && tp < n_tiles && lane < 24u) {  // one 32 B sector per lane: 3 x 256 B
      const double* base = lane < 8u ? sx : lane < 16u ? sy : sz;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + static_cast<size_t>(tp) * 32u + 4u * (lane & 7u)));
    }
  };

  struct Slim {
    F px, py, pz, dx, dy, dz;
    uint32_t klo, khi, home, tag;  // tag: 8-bit filter tag, 0 = no lookup
    unsigned i;                    // point index (>= P.n: none)
  };
  struct Entry {
    F px, py, pz, dx, dy, dz;
    uint32_t klo, khi, cand, idx;
  };
  struct RecRegs {
    uint2 key;
    float4 pa, pc;
    float2 pd;
    float4 s4;
    float2 s2;
  };
  auto load_pos = [&](unsigned tile, double& x, double& y, double& z) {
    const unsigned i = tile * 32u + lane;
    x = y = z = 0.0;
    if (tile < n_tiles && i < P.n) {
      x = first ? __ldcs(sx + i) : __ldcg(sx + i);
      y = first ? __ldcs(sy + i) : __ldcg(sy + i);
      z = first ? __ldcs(sz + i) : __ldcg(sz + i);
    }
  };
  auto xform = [&](unsigned tile, double x, double y, double z, Slim& q) {
    const unsigned i = tile * 32u + lane;
    q.tag = 0u;
    q.home = 0u;
    q.i = 0xffffffffu;
    if (tile >= n_tiles || i >= P.n) return;
    q.i = i;
    transform_point_rn(sT, x, y, z);
    if (!abl_nostore) {
      P.wx[i] = x;
      P.wy[i] = y;
      P.wz[i] = z;
    }
    const int kx = voxel_coord(x, P.voxel, inv_voxel);
    const int ky = voxel_coord(y, P.voxel, inv_voxel);
    const int kz = voxel_coord(z, P.voxel, inv_voxel);
    if (coord_in_range(kx) && coord_in_range(ky) && coord_in_range(kz)) {
      const uint64_t key = pack_key(kx, ky, kz);
      const SlotAddr ad = slot_addr(key, P.n_slots);
      q.klo = static_cast<uint32_t>(key);
      q.khi = static_cast<uint32_t>(key >> 32);
      q.home = ad.home;
      q.tag = filter_tag(ad.tag);
      q.px = F(x); q.py = F(y); q.pz = F(z);
      // the residual against the voxel mean is formed relative to the voxel centre
      q.dx = F(x - __dmul_rn(static_cast<double>(kx) + 0.5, P.voxel));
      q.dy = F(y - __dmul_rn(static_cast<double>(ky) + 0.5, P.voxel));
      q.dz = F(z - __dmul_rn(static_cast<double>(kz) + 0.5, P.voxel));
    }
  };
  auto window = [&](const Slim& q) -> uint4 {
    if (q.tag == 0u) return make_uint4(0u, 0u, 0u, 0u);
    if (abl_noprobe) return make_uint4(q.klo & 0u, 0u, 0u, 0u);
    return filter_window(P.filt, q.home, pol_filt);
  };
  auto scan = [&](const Slim& q, uint4 w) -> uint32_t {
    if (q.tag == 0u) return kNoCand;
    uint32_t b0 = q.home & ~7u;
    uint32_t r = scan_filter_window(w, q.home & 7u, q.tag);
    uint32_t scanned = 16u - (q.home & 7u);
    while (r == kMore && scanned < P.n_slots) {
      b0 += 16u;
      if (b0 >= P.n_slots) b0 -= P.n_slots;
      const uint2 lo = __ldg(reinterpret_cast<const uint2*>(P.filt + b0));
      const uint2 hi = __ldg(reinterpret_cast<const uint2*>(P.filt + b0 + 8));
      r = scan_filter_window(make_uint4(lo.x, lo.y, hi.x, hi.y), 0u, q.tag);
      scanned += 16u;
    }
    if (r >= 16u) return kNoCand;
    uint32_t c = b0 + r;
    if (c >= P.n_slots) c -= P.n_slots;
    return c;
  };

  // ---- the warp's ring of parked candidates (SoA: field f of entry e at park[f * kPark + e % kPark])
  unsigned head = 0, tail = 0;  // entries [head, tail) are parked (warp-uniform)
  auto park_hits = [&](const Slim& q, uint32_t cand) {
    const bool has = cand != kNoCand;
    const unsigned mask = __ballot_sync(0xffffffffu, has);
    if (write_hit && !has && q.i < P.n) P.hit[q.i] = 0;
    if (has) {
      uint32_t* w = park + ((tail + __popc(mask & ((1u << lane) - 1u))) & (kPark - 1u));
      w[0] = __float_as_uint(static_cast<float>(q.px));
      w[kPark] = __float_as_uint(static_cast<float>(q.py));
      w[2 * kPark] = __float_as_uint(static_cast<float>(q.pz));
      w[3 * kPark] = __float_as_uint(static_cast<float>(q.dx));
      w[4 * kPark] = __float_as_uint(static_cast<float>(q.dy));
      w[5 * kPark] = __float_as_uint(static_cast<float>(q.dz));
      w[6 * kPark] = q.klo;
      w[7 * kPark] = q.khi;
      w[8 * kPark] = cand;
      w[9 * kPark] = q.i;
    }
    tail += __popc(mask);
    __syncwarp();  // parked entries are read by other lanes of this warp
  };
  // pop up to 32 parked entries as a dense batch and send its gathers off
  auto pop_batch = [&](Entry& e, RecRegs& r) {
    const unsigned avail = tail - head;
    const unsigned take = avail < 32u ? avail : 32u;
    e.cand = kNoCand;
    r.key = make_uint2(0u, 0u);
    if (lane < take) {
      const uint32_t* w = park + ((head + lane) & (kPark - 1u));
      e.px = F(__uint_as_float(w[0])); e.py = F(__uint_as_float(w[kPark])); e.pz = F(__uint_as_float(w[2 * kPark]));
      e.dx = F(__uint_as_float(w[3 * kPark])); e.dy = F(__uint_as_float(w[4 * kPark]));
      e.dz = F(__uint_as_float(w[5 * kPark]));
      e.klo = w[6 * kPark]; e.khi = w[7 * kPark]; e.cand = w[8 * kPark]; e.idx = w[9 * kPark];
      const float4* rec = reinterpret_cast<const float4*>(P.slots + e.cand);
      if (rec_prefetch) prefetch_record(rec);  // (evict_last: the voxels a registration keeps touching stay in L2)
      r.key = ldg_u2_hint(rec, pol_rec);
      r.pa = ldg_f4_hint(rec + 1, pol_rec);
      r.pc = ldg_f4_hint(rec + 2, pol_rec);
      r.pd = ldg_f2_hint(rec + 3, pol_rec);
      r.s4 = __ldcs(P.c4 + e.idx);
      r.s2 = __ldcs(P.c2 + e.idx);
    }
    head += take;
    __syncwarp();  // the slots may be parked over from here on
  };
  double acc = 0.0;
  F v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = F(0);
  auto compute = [&](const Entry& e, const RecRegs& r) {
    float4 pa = r.pa, pc = r.pc;
    float2 pd = r.pd;
    bool hit = false;
    if (e.cand != kNoCand) {
      hit = r.key.x == e.klo && r.key.y == e.khi;
      if (!hit) {  // 8-bit filter collision (1/255 per occupied probe): walk on, slowly
        const uint64_t key = (static_cast<uint64_t>(e.khi) << 32) | e.klo;
        const SlotAddr ad = slot_addr(key, P.n_slots);
        const VoxelSlot* far = resolve_probe_filter(P.filt, P.slots, P.n_slots, key,
                                                    next_slot(e.cand, P.n_slots), filter_tag(ad.tag));
        if (far != nullptr) {
          hit = true;
          pa = __ldg(reinterpret_cast<const float4*>(far) + 1);
          pc = __ldg(reinterpret_cast<const float4*>(far) + 2);
          pd = __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float4*>(far) + 3));
        }
      }
      if (write_hit) P.hit[e.idx] = hit ? 1 : 0;
    }
    if (hit) {
      F cr[6];
      rotate_sym<F>(sR, F(r.s4.x), F(r.s4.y), F(r.s4.z), F(r.s4.w), F(r.s2.x), F(r.s2.y), cr);
      point_terms<F, true>(e.px, e.py, e.pz, e.dx - F(pa.x), e.dy - F(pa.y), e.dz - F(pa.z),
                           cr[0] + F(pc.x), cr[1] + F(pc.y), cr[2] + F(pc.z), cr[3] + F(pc.w),
                           cr[4] + F(pd.x), cr[5] + F(pd.y), v);
    } else {
#pragma unroll
      for (int k = 0; k < 28; ++k) v[k] = F(0);
    }
    acc += static_cast<double>(warp_reduce_scatter32<F>(v, lane));
  };

  if constexpr (U > 1) {
    // ---- U tiles per trip (depth 11): the U filter windows and the U position sets a trip requests are
    // mutually independent, so a trip — now U x as long — covers a loaded memory round trip with
    // instruction-level parallelism inside the warp instead of with more warps (of which the register
    // file allows 16)
    static_assert(RING == 0, "U > 1 uses register loads");
    unsigned tile_a[U], tile_b[U];
    Slim qa[U], qb[U];
    uint4 ta[U], tb[U];
    double rx[U], ry[U], rz[U];
#pragma unroll
    for (int u = 0; u < U; ++u) tile_a[u] = next_tile();
#pragma unroll
    for (int u = 0; u < U; ++u) tile_b[u] = next_tile();
    {
      double ax[U], ay[U], az[U];
#pragma unroll
      for (int u = 0; u < U; ++u) load_pos(tile_a[u], ax[u], ay[u], az[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) load_pos(tile_b[u], rx[u], ry[u], rz[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        xform(tile_a[u], ax[u], ay[u], az[u], qa[u]);
        ta[u] = window(qa[u]);
      }
    }
    Entry be;
    RecRegs br;
    bool in_flight = false;
    while (tile_a[0] < n_tiles) {
      // 1. tiles B: positions have arrived; transform, key, request their windows
#pragma unroll
      for (int u = 0; u < U; ++u) {
        xform(tile_b[u], rx[u], ry[u], rz[u], qb[u]);
        tb[u] = window(qb[u]);
      }
      // 2. positions of the tiles after those
      unsigned tile_c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        tile_c[u] = next_tile();
        load_pos(tile_c[u], rx[u], ry[u], rz[u]);
      }
      // 3. tiles A: windows requested a trip ago; scan, park; a dense batch whenever 32 are parked
#pragma unroll
      for (int u = 0; u < U; ++u) {
        uint32_t cand = scan(qa[u], ta[u]);
        if (abl_nogather) {
          if (cand != kNoCand) acc += 1.0;
          cand = kNoCand;
        }
        park_hits(qa[u], cand);
        if (tail - head >= 32u) {
          if (in_flight) compute(be, br);
          pop_batch(be, br);
          in_flight = true;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        qa[u] = qb[u];
        ta[u] = tb[u];
        tile_a[u] = tile_b[u];
        tile_b[u] = tile_c[u];
      }
    }
    if (in_flight) compute(be, br);
    while (tail != head) {
      pop_batch(be, br);
      compute(be, br);
    }
    return acc;
  }
  // ---- the lookup pipeline.  Across the loop's back edge a warp has in flight: the filter window of
  // tile A (requested a full trip ago, scanned this trip), the raw positions of tile B (register loads
  // issued a trip ago, or the head of the bulk-copy ring) and possibly a batch of gathers.
  const uint64_t pol_stream = l2_policy_evict_first();
  auto ring_issue = [&](unsigned t) {  // request the positions of tile t into the ring's next slot
    const unsigned slot = ring.issued % (RING > 0 ? RING : 1);
    if (lane == 0) tq[slot] = t;
    if (t < n_tiles) {
      if (lane == 0) {
        double* dst = ring_sm + slot * 96u;
        uint64_t* bar = wbar + slot;
        mbar_arrive_expect_tx(bar, kTileBytes);
        bulk_load(dst, sx + static_cast<size_t>(t) * 32u, 256u, bar, pol_stream);
        bulk_load(dst + 32, sy + static_cast<size_t>(t) * 32u, 256u, bar, pol_stream);
        bulk_load(dst + 64, sz + static_cast<size_t>(t) * 32u, 256u, bar, pol_stream);
      }
    }
    ++ring.issued;
  };
  auto ring_consume = [&](unsigned& t, double& x, double& y, double& z) {  // oldest outstanding tile
    const unsigned slot = ring.consumed % (RING > 0 ? RING : 1);
    const unsigned parity = (ring.phase >> slot) & 1u;
    __syncwarp();
    t = tq[slot];
    x = y = z = 0.0;
    if (t < n_tiles) {
      ring.phase ^= 1u << slot;
      uint64_t* bar = wbar + slot;
      if (!mbar_try_wait(bar, parity)) {
        unsigned spins = 0;
        while (!mbar_try_wait(bar, parity)) {
          __nanosleep(20);
          if (++spins > kSpinLimit) {
            atomicExch(&P.st->error, 3u);
            break;
          }
        }
      }
      const double* src = ring_sm + slot * 96u;
      x = src[lane];
      y = src[32 + lane];
      z = src[64 + lane];
    }
    ++ring.consumed;
  };
  constexpr unsigned kAhead = 4;  // RING == 0: trips the L2 prefetch of the positions runs ahead
  unsigned tile_a, tile_b = 0;
  Slim qa, qb;
  uint4 ta, tb;
  double rx = 0.0, ry = 0.0, rz = 0.0;
  unsigned trips = 2;
  if constexpr (RING > 0) {
    // (a slot's mbarrier completes a phase only when a load was issued for it: ring.phase tracks the
    // parity per slot; slots that carried a tile id past the end change nothing)
    if (!first && lane == 0) fence_proxy_async();  // other SMs' generic-proxy position stores -> these bulk reads
    for (int d = 0; d < RING; ++d) ring_issue(next_tile());
    double ax, ay, az;
    ring_consume(tile_a, ax, ay, az);
    xform(tile_a, ax, ay, az, qa);
    ta = window(qa);
    __syncwarp();
    ring_issue(next_tile());
  } else {
#pragma unroll
    for (unsigned k = 0; k < kAhead; ++k) prefetch_pos(k);
    tile_a = next_tile();
    tile_b = next_tile();
    double ax, ay, az;
    load_pos(tile_a, ax, ay, az);
    load_pos(tile_b, rx, ry, rz);
    xform(tile_a, ax, ay, az, qa);
    ta = window(qa);
  }
  Entry be;
  RecRegs br;
  bool in_flight = false;
  while (tile_a < n_tiles) {
    // ---- 1. tile B: positions have arrived; transform, key, request its window; then the positions of
    // a tile further on
    if constexpr (RING > 0) {
      ring_consume(tile_b, rx, ry, rz);
      xform(tile_b, rx, ry, rz, qb);
      tb = window(qb);
      __syncwarp();  // every lane has read (and used) its entry of the slot that is refilled now
      ring_issue(next_tile());
    } else {
      xform(tile_b, rx, ry, rz, qb);
      tb = window(qb);
    }
    unsigned tile_c = 0;
    if constexpr (RING == 0) {
      tile_c = next_tile();
      load_pos(tile_c, rx, ry, rz);
      prefetch_pos(trips + kAhead - 1u);
      ++trips;
    }
    // ---- 3. tile A: its window was requested a trip ago; scan, park the candidates
    {
      uint32_t cand = scan(qa, ta);
      if (abl_nogather) {  // (ablation: the probe's result must stay live)
        if (cand != kNoCand) acc += 1.0;
        cand = kNoCand;
      }
      park_hits(qa, cand);
    }
    // ---- 4. 32 candidates parked: linearise the batch in flight, send the next one off
    if (tail - head >= 32u) {
      if (in_flight) compute(be, br);
      pop_batch(be, br);
      in_flight = true;
    }
    qa = qb;
    ta = tb;
    tile_a = tile_b;
    if constexpr (RING == 0) tile_b = tile_c;
  }
  if constexpr (RING > 0) {
    // drain: the ring still holds RING slots whose tile ids are past the end (nothing was loaded)
    for (int d = 0; d < RING; ++d) {
      unsigned t_;
      double x_, y_, z_;
      ring_consume(t_, x_, y_, z_);
    }
  }
  // ---- end of the pass: what is in flight, then what is still parked
  if (in_flight) compute(be, br);
  while (tail != head) {
    pop_batch(be, br);
    compute(be, br);
  }
  return acc;
}

// deterministic CTA reduction: lane l of every warp holds term l
template <int NW>
__device__ __forceinline__ void block_reduce_store(double acc, double (*s_part)[32], double* out) {
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  s_part[w][lane] = acc;
  __syncthreads();
  if (threadIdx.x < kAcc) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i) s += s_part[i][threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

// sum partials[0..G) in a fixed order -> s_sum[kAcc]  (whole CTA cooperates)
template <int NW>
__device__ __forceinline__ void final_reduce(const double* partials, unsigned G,
                                             double (*s_part)[32], double* s_sum) {
  const unsigned term = threadIdx.x & 31, grp = threadIdx.x >> 5;
  double s = 0.0;
  if (term < kAcc) {
    // 16 independent loads in flight per thread (a dependent chain of ~G/8 L2
    // round trips used to cost ~20 us per iteration): one round trip up to 128 CTAs
    for (unsigned bb = grp; bb < G; bb += NW * 16) {
      double v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const unsigned b2 = bb + NW * k;
        v[k] = b2 < G ? ld_cg(partials + static_cast<size_t>(b2) * kAcc + term) : 0.0;
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) s += v[k];
    }
  }
  s_part[grp][term] = s;
  __syncthreads();
  if (threadIdx.x < kAcc) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i) t += s_part[i][threadIdx.x];
    s_sum[threadIdx.x] = t;
  }
  __syncthreads();
}

// fast path of the 6x6 solve: LDL^T without pivoting, fully unrolled so the
// matrix lives in registers.  Returns false (caller falls back to the pivoted
// restatement of Eigen's LDLT below) unless every pivot is safely positive.
__device__ __forceinline__ bool ldlt_solve6_nopivot(const double* H, const double* b, double* x) {
  double A[6][6], D[6], y[6];
  double maxd = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) A[i][j] = H[6 * i + j];
    maxd = fmax(maxd, fabs(H[7 * i]));
  }
  const double tol = maxd * 1e-12;
  bool ok = maxd > 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double d = A[k][k];
#pragma unroll
    for (int j = 0; j < k; ++j) d -= A[k][j] * A[k][j] * D[j];
    ok = ok && (d > tol);
    D[k] = d;
    const double inv = 1.0 / d;
#pragma unroll
    for (int i = k + 1; i < 6; ++i) {
      double s = A[i][k];
#pragma unroll
      for (int j = 0; j < k; ++j) s -= A[i][j] * A[k][j] * D[j];
      A[i][k] = s * inv;
    }
  }
  if (!ok) return false;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
#pragma unroll
    for (int j = 0; j < i; ++j) s -= A[i][j] * y[j];
    y[i] = s;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) y[i] /= D[i];
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
#pragma unroll
    for (int j = i + 1; j < 6; ++j) s -= A[j][i] * x[j];
    x[i] = s;
  }
  return true;
}

// Eigen LDLT<Matrix6d>::solve restated: diagonal pivoting, zero pivots -> 0
__device__ __noinline__ void ldlt_solve6(const double* Hin, const double* bin, double* x) {
  double A[6][6], L[6][6], D[6], y[6];
  int perm[6];
  for (int i = 0; i < 6; ++i) {
    perm[i] = i;
    D[i] = 0.0;
    for (int j = 0; j < 6; ++j) {
      A[i][j] = Hin[6 * i + j];
      L[i][j] = (i == j) ? 1.0 : 0.0;
    }
  }
  for (int k = 0; k < 6; ++k) {
    int piv = k;
    double best = fabs(A[k][k]);
    for (int i = k + 1; i < 6; ++i)
      if (fabs(A[i][i]) > best) {
        best = fabs(A[i][i]);
        piv = i;
      }
    if (piv != k) {
      for (int j = 0; j < 6; ++j) { double t = A[k][j]; A[k][j] = A[piv][j]; A[piv][j] = t; }
      for (int i = 0; i < 6; ++i) { double t = A[i][k]; A[i][k] = A[i][piv]; A[i][piv] = t; }
      for (int j = 0; j < k; ++j) { double t = L[k][j]; L[k][j] = L[piv][j]; L[piv][j] = t; }
      int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
    }
    const double d = A[k][k];
    D[k] = d;
    if (!(fabs(d) > 0.0)) {
      if (k == 0) break;
      continue;
    }
    for (int i = k + 1; i < 6; ++i) L[i][k] = A[i][k] / d;
    for (int i = k + 1; i < 6; ++i)
      for (int j = k + 1; j < 6; ++j) A[i][j] -= L[i][k] * d * L[j][k];
  }
  for (int i = 0; i < 6; ++i) y[i] = bin[perm[i]];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < i; ++j) y[i] -= L[i][j] * y[j];
  for (int i = 0; i < 6; ++i) y[i] = (fabs(D[i]) > DBL_MIN) ? y[i] / D[i] : 0.0;
  for (int i = 5; i >= 0; --i)
    for (int j = i + 1; j < 6; ++j) y[i] -= L[j][i] * y[j];
  for (int i = 0; i < 6; ++i) x[perm[i]] = y[i];
}

// Utils::se3ToSE3 (src/Utils.cpp:56-63) with computeJ (:40-54) and
// rotationVectorToMatrix (:28-32, Eigen AngleAxisd::toRotationMatrix)
__device__ void se3_to_SE3(const double* se3, double* T /* R(9) t(3) */) {
  const double rx = se3[3], ry = se3[4], rz = se3[5];
  const double n2 = rx * rx + ry * ry + rz * rz;
  const double angle = sqrt(n2);
  double ax = rx, ay = ry, az = rz;
  if (n2 > 0.0) {
    ax = rx / angle;
    ay = ry / angle;
    az = rz / angle;
  }
  double s, c;
  sincos(angle, &s, &c);
  double J[9];
  if (angle < 1e-6) {
    J[0] = 1; J[1] = 0; J[2] = 0; J[3] = 0; J[4] = 1; J[5] = 0; J[6] = 0; J[7] = 0; J[8] = 1;
  } else {
    const double f1 = s / angle, f2 = (1.0 - c) / angle, g = 1.0 - f1;
    J[0] = f1 + g * ax * ax;      J[1] = g * ax * ay - f2 * az; J[2] = g * ax * az + f2 * ay;
    J[3] = g * ay * ax + f2 * az; J[4] = f1 + g * ay * ay;      J[5] = g * ay * az - f2 * ax;
    J[6] = g * az * ax - f2 * ay; J[7] = g * az * ay + f2 * ax; J[8] = f1 + g * az * az;
  }
  T[9] = J[0] * se3[0] + J[1] * se3[1] + J[2] * se3[2];
  T[10] = J[3] * se3[0] + J[4] * se3[1] + J[5] * se3[2];
  T[11] = J[6] * se3[0] + J[7] * se3[1] + J[8] * se3[2];
  const double cx = (1.0 - c) * ax, cy = (1.0 - c) * ay, cz = (1.0 - c) * az;
  const double sxv = s * ax, syv = s * ay, szv = s * az;
  double tmp = cx * ay;
  T[1] = tmp - szv; T[3] = tmp + szv;
  tmp = cx * az;
  T[2] = tmp + syv; T[6] = tmp - syv;
  tmp = cy * az;
  T[5] = tmp - sxv; T[7] = tmp + sxv;
  T[0] = cx * ax + c; T[4] = cy * ay + c; T[8] = cz * az + c;
}

// H(6x6, row-major) entry -> index of its unique term in the 27 sums
__constant__ unsigned char c_hmap[36] = {0, 1, 2,  6,  7,  8,  1, 3,  4,  9,  10, 11,
                                         2, 4, 5,  12, 13, 14, 6, 9,  12, 15, 16, 17,
                                         7, 10, 13, 16, 18, 19, 8, 11, 14, 17, 19, 20};

// element e (0..8 rotation, 9..11 translation) of step * old: explicit round-to-nearest operations, so
// that the solver and every CTA's local copy of the total pose agree bit for bit
__device__ __forceinline__ double compose_elem(const double* step, const double* old, int e) {
  if (e < 9) {
    const int i = e / 3, j = e % 3;
    return dot3_rn(step[3 * i], old[j], step[3 * i + 1], old[3 + j], step[3 * i + 2], old[6 + j]);
  }
  const int i = e - 9;
  return __dadd_rn(dot3_rn(step[3 * i], old[9], step[3 * i + 1], old[10], step[3 * i + 2], old[11]), step[9 + i]);
}

// ---- the per-iteration solve, fast path -----------------------------------
// Round 1 ran a warp-cooperative LDL^T over shared memory (a __syncwarp and a
// shared-memory round trip per elimination step, an IEEE division per lane and
// step): ESKF_ALIGN_STAMPS measured 12-13 us per Gauss-Newton iteration between
// "sums reduced" and "pose published" — most of the per-iteration hand-off, and
// more than an 8-way shard of the dense config computes in.  The system is 6x6:
// one thread factorises it in registers (fully unrolled, one reciprocal per
// pivot; ~1.5 k dependent-issue cycles), builds the step and the new total pose
// and leaves them in shared memory; the lanes of its warp then write the state
// and trace words in parallel.  `Told` is the total pose before this step
// (every CTA tracks it in shared memory, see align_kernel: no global load).
// Falls back to the pivoted restatement of Eigen's LDLT when a pivot is not
// safely positive.
// Utils::se3ToSE3 (src/Utils.cpp:56-63) for a Gauss-Newton step: |phi| < 0.5 rad takes sin / cos from their
// series (|error| < 1e-22, no library call on the solver's critical path), larger angles the library path
__device__ __forceinline__ void se3_to_SE3_small(const double* se3, double* T) {
  const double rx = se3[3], ry = se3[4], rz = se3[5];
  const double n2 = rx * rx + ry * ry + rz * rz;
  if (n2 >= 0.25) {
    se3_to_SE3(se3, T);
    return;
  }
  const double angle = sqrt(n2);
  double ax = rx, ay = ry, az = rz;
  if (n2 > 0.0) {
    const double ia = 1.0 / angle;
    ax *= ia; ay *= ia; az *= ia;
  }
  // s = x (1 - n2/6 (1 - n2/20 (1 - ...))),  c = 1 - n2/2 (1 - n2/12 (1 - ...))
  double ps = 1.0, pc = 1.0;
#pragma unroll
  for (int k = 9; k >= 1; --k) {
    ps = 1.0 - n2 * (1.0 / static_cast<double>((2 * k) * (2 * k + 1))) * ps;
    pc = 1.0 - n2 * (1.0 / static_cast<double>((2 * k - 1) * (2 * k))) * pc;
  }
  const double s = angle * ps, c = pc;
  double J[9];
  if (angle < 1e-6) {
    J[0] = 1; J[1] = 0; J[2] = 0; J[3] = 0; J[4] = 1; J[5] = 0; J[6] = 0; J[7] = 0; J[8] = 1;
  } else {
    const double f1 = ps, f2 = (1.0 - c) / angle, g = 1.0 - f1;
    J[0] = f1 + g * ax * ax;      J[1] = g * ax * ay - f2 * az; J[2] = g * ax * az + f2 * ay;
    J[3] = g * ay * ax + f2 * az; J[4] = f1 + g * ay * ay;      J[5] = g * ay * az - f2 * ax;
    J[6] = g * az * ax - f2 * ay; J[7] = g * az * ay + f2 * ax; J[8] = f1 + g * az * az;
  }
  T[9] = J[0] * se3[0] + J[1] * se3[1] + J[2] * se3[2];
  T[10] = J[3] * se3[0] + J[4] * se3[1] + J[5] * se3[2];
  T[11] = J[6] * se3[0] + J[7] * se3[1] + J[8] * se3[2];
  const double cx = (1.0 - c) * ax, cy = (1.0 - c) * ay, cz = (1.0 - c) * az;
  const double sxv = s * ax, syv = s * ay, szv = s * az;
  double tmp = cx * ay;
  T[1] = tmp - szv; T[3] = tmp + szv;
  tmp = cx * az;
  T[2] = tmp + syv; T[6] = tmp - syv;
  tmp = cy * az;
  T[5] = tmp - sxv; T[7] = tmp + sxv;
  T[0] = cx * ax + c; T[4] = cy * ay + c; T[8] = cz * az + c;
}

constexpr int kStampSlots = 12;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define ESKF_SUBSTAMP(slot) \
  do { if (P.stamps != nullptr) P.stamps[kStampSlots * it + (slot)] = globaltimer_ns(); } while (0)

// the math: lane 0 of the calling warp; results in sm: [36..47] step, [60..71] new total, [84] done, [85] converged
__device__ __noinline__ void solve_compute(const AlignParams& P, const double* S, const double* Told, int it,
                                           double* sm) {
  double* stp = sm + 36;
  double* tot = sm + 60;
  if ((threadIdx.x & 31) == 0) {
    // (index map known at compile time: shared-memory loads with immediate offsets)
    constexpr unsigned char kHmap[36] = {0, 1, 2,  6,  7,  8,  1, 3,  4,  9,  10, 11, 2, 4, 5,  12, 13, 14,
                                         6, 9, 12, 15, 16, 17, 7, 10, 13, 16, 18, 19, 8, 11, 14, 17, 19, 20};
    double H[36], nb[6], se3[6], step[12];
#pragma unroll
    for (int e = 0; e < 36; ++e) H[e] = S[kHmap[e]];
#pragma unroll
    for (int i = 0; i < 6; ++i) nb[i] = -S[21 + i];
    ESKF_SUBSTAMP(8);
    // JTJ.ldlt().solve(-JTr), Registration.cpp:78
    if (!ldlt_solve6_nopivot(H, nb, se3)) ldlt_solve6(H, nb, se3);
    ESKF_SUBSTAMP(9);
    se3_to_SE3_small(se3, step);
    ESKF_SUBSTAMP(10);
    // totalTransform = transformIter * totalTransform (Registration.cpp:20)
#pragma unroll
    for (int e = 0; e < 12; ++e) tot[e] = compose_elem(step, Told, e);
#pragma unroll
    for (int i = 0; i < 12; ++i) stp[i] = step[i];
    // convergenceCheck (Registration.cpp:37-50)
    const double cosine = 0.5 * (((step[0] + step[4]) + step[8]) - 1.0);
    const double tsq = step[9] * step[9] + step[10] * step[10] + step[11] * step[11];
    const int conv = (cosine >= P.cos_thr && tsq <= P.trans_sq_thr) ? 1 : 0;
    const int done = P.fixed_iterations > 0 ? (it + 1 >= P.fixed_iterations) : (conv || it + 1 >= P.max_iteration);
    sm[84] = static_cast<double>(done);
    sm[85] = static_cast<double>(conv);
    ESKF_SUBSTAMP(11);
  }
  __syncwarp();
}

// the bookkeeping: state + trace words, one per lane.  Nothing here is needed by the other CTAs to start
// their next pass (the dynamic-tile counter is reset ahead of the pose broadcast), so in the persistent
// kernels it runs AFTER the broadcast.
__device__ __forceinline__ void solve_store(const AlignParams& P, const double* S, int it, const double* sm) {
  const unsigned lane = threadIdx.x & 31;
  const double* stp = sm + 36;
  const double* tot = sm + 60;
  AlignState* st = P.st;
  for (int e = lane; e < 36; e += 32)
    if (P.trace_H) P.trace_H[36 * it + e] = S[c_hmap[e]];
  if (lane < 6 && P.trace_b) P.trace_b[6 * it + lane] = S[21 + lane];
  if (lane < 12) {
    st->T_total[lane] = tot[lane];
    st->T_step[lane] = stp[lane];
    if (P.trace_step) P.trace_step[12 * it + lane] = stp[lane];
  }
  if (lane < 9) st->Rf[lane] = static_cast<float>(tot[lane]);
  if (lane == 12) {
    const unsigned long long nc = static_cast<unsigned long long>(S[27]);
    if (P.trace_ncorr) P.trace_ncorr[it] = nc;
    st->n_corr = nc;
    st->iter = it + 1;
    st->converged = static_cast<int>(sm[85]);
    st->done = static_cast<int>(sm[84]);
  }
  if (lane == 0 && sm[84] != 0.0 && P.mail != nullptr)
    publish_result(P, tot, it + 1, static_cast<int>(sm[85]), static_cast<unsigned long long>(S[27]));
  __syncwarp();
}

__device__ __forceinline__ void solve_and_update_fast(const AlignParams& P, const double* S, const double* Told,
                                                      int it, double* sm) {
  solve_compute(P, S, Told, it, sm);
  if ((threadIdx.x & 31) == 12) P.st->tile_counter = 0u;
  solve_store(P, S, it, sm);
}

// ---- flagged-word ("LL") broadcast of the next pose ------------------------
// The solver warp publishes T_step (12 fp64 = 24 words), the fp32 rotation of
// T_total (9 words) and the done flag as 34 eight-byte words {payload, it + 1};
// a 64-bit store is single-copy atomic, so a poller that sees the flag has the
// payload: one L2 round trip per CTA instead of "poll the epoch word, then
// fetch the pose".  The box is zeroed with the rest of the state before every
// launch; the solver of iteration it + 1 cannot run before every CTA has read
// iteration it's words (it needs their tickets), so one buffer suffices.
constexpr int kLLWords = 25;
__device__ __forceinline__ void ll_publish(const AlignParams& P, const double* sm, int it) {
  const unsigned lane = threadIdx.x & 31;
  const double* stp = sm + 36;
  // release: the tile-counter reset and, cumulatively, every CTA's position stores acquired with the
  // tickets, before the words below (fence.acq_rel: __threadfence() is the costlier fence.sc)
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  volatile unsigned long long* box = P.llbox;
  const unsigned long long flag = static_cast<unsigned long long>(it + 1) << 32;
  if (lane < kLLWords) {
    unsigned payload;
    if (lane < 24) {
      const double d = stp[lane >> 1];
      payload = (lane & 1) ? static_cast<unsigned>(__double2hiint(d)) : static_cast<unsigned>(__double2loint(d));
    } else {
      payload = static_cast<unsigned>(sm[84] != 0.0);
    }
    box[lane] = flag | payload;
  }
}

// warp 0 of every CTA: wait for iteration `it`'s words; the step lands in s_T, the done flag in *s_done
template <typename F>
__device__ __forceinline__ int ll_receive(const AlignParams& P, int it, double* s_T, int* s_done) {
  const unsigned lane = threadIdx.x & 31;
  const volatile unsigned long long* box = P.llbox;
  const unsigned want = static_cast<unsigned>(it + 1);
  unsigned long long a = lane < kLLWords ? box[lane] : (static_cast<unsigned long long>(want) << 32);
  unsigned spins = 0;
  int ok = 1;
  while (static_cast<unsigned>(a >> 32) != want) {
    __nanosleep(20);
    a = box[lane];
    if (++spins > kSpinLimit || ((spins & 63u) == 0u && ld_acquire_u32(&P.st->error) != 0)) {
      atomicCAS(&P.st->error, 0u, 1u);
      ok = 0;
      break;
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  asm volatile("fence.acq_rel.gpu;" ::: "memory");  // acquire side: the other CTAs' position stores before this CTA's next pass
  const unsigned pa = static_cast<unsigned>(a);
  const unsigned src = (2u * lane) & 31u;
  const unsigned lo = __shfl_sync(0xffffffffu, pa, src), hi = __shfl_sync(0xffffffffu, pa, src + 1u);
  const unsigned dn = __shfl_sync(0xffffffffu, pa, 24);
  if (lane < 12) s_T[lane] = __hiloint2double(static_cast<int>(hi), static_cast<int>(lo));
  if (lane == 0) *s_done = (ok && dn == 0u) ? 0 : 1;
  return ok;
}

// ---- fused H/b exchange over NVLink / NVSwitch peer memory ---------------
// Called by the whole last CTA of a rank once its local sums sit in s_sum.
// Every rank stores its 28 sums into slot [parity][rank] of EVERY rank's
// mailbox (plain stores to peer-mapped memory), fences at system scope and
// then raises the slot's flag; it then waits for the `world` flags of its own
// mailbox and adds the contributions in rank order, so every rank forms the
// bit-identical total (and therefore the identical step) without a collective
// call or a host round trip: 216 B per peer per iteration is pure latency.
// Two parities: a rank can run at most one iteration ahead of a peer (it needs
// the peer's next contribution to go further).  Spins are bounded.
__device__ __forceinline__ double ld_acquire_sys_f64(const double* p) {
  double v;
  asm volatile("ld.acquire.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_f64(double* p, double v) {
  asm volatile("st.release.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

__device__ __forceinline__ bool exchange_sums(const AlignParams& P, int it, double* s_sum, int* s_flag) {
  const unsigned t = threadIdx.x;
  const int W = P.world;
  // 4-slot ring: (call parity, iteration parity).  With the iteration parity alone, a rank that has
  // started the NEXT call could overwrite slot 0 of a peer that was descheduled between seeing the flags
  // of this call's last (even) iteration and reading the data words.
  const size_t ring = ((static_cast<size_t>(P.seq) & 1u) << 1) | static_cast<size_t>(it & 1);
  const size_t slot = (ring * W + P.rank) * kMailStride;
  const double flag = static_cast<double>(P.seq) * 65536.0 + static_cast<double>(it + 1);
  for (unsigned k = t; k < static_cast<unsigned>(kAcc * W); k += blockDim.x)
    P.peers[k / kAcc][slot + k % kAcc] = s_sum[k % kAcc];
  __syncthreads();
  // (st.release.sys orders the CTA's data stores, made visible to this thread by
  // the bar.sync above, before the flag: no separate system fence)
  if (t < static_cast<unsigned>(W)) st_release_sys_f64(P.peers[t] + slot + kAcc, flag);
  if (t == 0) *s_flag = 1;
  __syncthreads();
  if (t < static_cast<unsigned>(W)) {
    const double* f = P.peers[P.rank] + (ring * W + t) * kMailStride + kAcc;
    unsigned spins = 0;
    while (ld_acquire_sys_f64(f) != flag) {
      __nanosleep(20);
      if (++spins > 32u * kSpinLimit || ld_acquire_u32(&P.st->error) != 0) {  // several seconds
        atomicExch(&P.st->error, 2u);
        *s_flag = 0;
        break;
      }
    }
  }
  __syncthreads();
  if (t < kAcc) {
    // the flags were acquired at system scope above (+ bar.sync): the data can be
    // read with plain L1-bypassing loads, all of them in flight at once
    const volatile double* box = P.peers[P.rank] + ring * W * kMailStride + t;
    double v[kMaxWorld];
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r) v[r] = r < W ? box[static_cast<size_t>(r) * kMailStride] : 0.0;
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r) s += v[r];  // rank order; absent ranks add +0.0
    s_sum[t] = s;
  }
  __syncthreads();
  return *s_flag != 0;
}

// The same exchange with flagged words (the default, eskf_ctx option align_xchg_ll): every sum travels
// as two 8-byte words {half of the fp64, tag of this call and iteration}; an 8-byte store is single-copy
// atomic, so a word whose tag matches carries its payload and nothing needs a release fence (a
// st.release.sys waits for the acknowledgement of every earlier peer write: a full NVLink round trip per
// iteration) nor a second, dependent read of the data after the flag.  Thread k of the last CTA sends
// word k % 56 to peer k / 56, then polls word k % 56 of source rank k / 56 in its own mailbox.
constexpr int kXWords = 2 * kAcc;  // 56
static_assert(kXWords <= kMailStride, "flagged words of one contribution fit its mailbox slot");
__device__ __forceinline__ bool exchange_sums_ll(const AlignParams& P, int it, double* s_sum, int* s_flag,
                                                 unsigned* s_words /* [kMaxWorld][kXWords] */) {
  const unsigned t = threadIdx.x;
  const int W = P.world;
  const size_t ring = ((static_cast<size_t>(P.seq) & 1u) << 1) | static_cast<size_t>(it & 1);
  const unsigned tag = ((P.seq & 0xffffu) << 16) | (static_cast<unsigned>(it + 1) & 0xffffu);
  if (t == 0) *s_flag = 1;
  for (unsigned k = t; k < static_cast<unsigned>(kXWords * W); k += blockDim.x) {
    const unsigned w = k % kXWords, peer = k / kXWords;
    const double d = s_sum[w >> 1];
    const unsigned payload = (w & 1u) ? static_cast<unsigned>(__double2hiint(d)) : static_cast<unsigned>(__double2loint(d));
    volatile unsigned long long* dst = reinterpret_cast<volatile unsigned long long*>(
        P.peers[peer] + (ring * W + P.rank) * kMailStride + w);
    *dst = (static_cast<unsigned long long>(tag) << 32) | payload;
  }
  __syncthreads();  // (s_flag; every thread's sends are issued)
  for (unsigned k = t; k < static_cast<unsigned>(kXWords * W); k += blockDim.x) {
    const unsigned w = k % kXWords, src = k / kXWords;
    const volatile unsigned long long* box = reinterpret_cast<const volatile unsigned long long*>(
        P.peers[P.rank] + (ring * W + src) * kMailStride + w);
    unsigned long long v = *box;
    unsigned spins = 0;
    while (static_cast<unsigned>(v >> 32) != tag) {
      __nanosleep(20);
      v = *box;
      if (++spins > 32u * kSpinLimit || ((spins & 255u) == 0u && ld_acquire_u32(&P.st->error) != 0)) {  // several seconds
        atomicExch(&P.st->error, 2u);
        *s_flag = 0;
        break;
      }
    }
    s_words[src * kXWords + w] = static_cast<unsigned>(v);
  }
  __syncthreads();
  if (t < kAcc) {
    double s = 0.0;
    for (int r = 0; r < W; ++r)  // rank order: every rank forms the bit-identical total
      s += __hiloint2double(static_cast<int>(s_words[r * kXWords + 2 * t + 1]), static_cast<int>(s_words[r * kXWords + 2 * t]));
    s_sum[t] = s;
  }
  __syncthreads();
  return *s_flag != 0;
}

#define ESKF_STAMP(cond, slot) \
  do { if (P.stamps != nullptr && (cond)) P.stamps[kStampSlots * it + (slot)] = globaltimer_ns(); } while (0)

template <typename F, int U, int NN, int MINB, int T = kT, int DEPTH = 3>
__global__ void __launch_bounds__(T, MINB) align_kernel(AlignParams P) {
  constexpr int NW = T / 32;
  __shared__ double s_T[12];
  __shared__ double s_Ttot[12];  // the total pose so far (every CTA composes the steps itself)
  __shared__ F s_R[9];
  __shared__ double s_part[NW][32];
  __shared__ double s_sum[kAcc];
  __shared__ double s_solve[96];
  __shared__ int s_last, s_done, s_xchg;
  __shared__ unsigned s_xw[kMaxWorld * 2 * kAcc];  // flagged words received from every rank
  extern __shared__ __align__(128) unsigned char s_dyn[];  // depth 5: [NW][k_res + kRing] tiles, then the ring's mbarriers
  const unsigned G = gridDim.x, t = threadIdx.x;
  AlignState* st = P.st;
  const int max_it = P.fixed_iterations > 0 ? P.fixed_iterations : P.max_iteration;

  RingState ring = {0u, 0u, 0u};
  double* wsm = nullptr;
  uint64_t* wbar = nullptr;
  HitQueue* hq = nullptr;
  HitList hl;
  __shared__ unsigned s_tail;
  if constexpr (DEPTH == 7) {
    // [10 fields][cap] words of shared memory + the CTA's spill region of P.spill_cap entries per field
    hl.sm = reinterpret_cast<uint32_t*>(s_dyn);
    hl.gm = P.spill + static_cast<size_t>(blockIdx.x) * 10u * P.spill_cap;
    hl.cap = static_cast<unsigned>(P.k_res);
    hl.gcap = P.spill_cap;
    if (t == 0) s_tail = 0u;
  }
  if constexpr (DEPTH == 6) {
    wsm = reinterpret_cast<double*>(s_dyn);  // the CTA's resident tiles
    hq = reinterpret_cast<HitQueue*>(s_dyn + static_cast<size_t>(P.k_res) * kTileBytes);
    if (t == 0) hq->tail = 0u;
  }
  if constexpr (DEPTH == 9) {
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_dyn + static_cast<size_t>(NW) * kWarpBytes9);
    if (t < NW * kRing9) mbar_init(bars + t, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if constexpr (DEPTH == 5) {
    const unsigned per_warp = static_cast<unsigned>(P.k_res + kRing) * 96u;  // doubles
    double* base = reinterpret_cast<double*>(s_dyn);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + static_cast<size_t>(NW) * per_warp);
    wsm = base + static_cast<size_t>(t >> 5) * per_warp;
    wbar = bars + (t >> 5) * kRing;
    if (t < NW * kRing) mbar_init(bars + t, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  // pose of the first iteration: the guess (later ones arrive with the hand-off below)
  if (t < 12) {
    s_T[t] = P.guess[t];
    s_Ttot[t] = P.guess[t];
  }
  if (t < 9) s_R[t] = F(P.guess[t]);
  __syncthreads();
  for (int it = 0; it < max_it; ++it) {
    ESKF_STAMP(blockIdx.x == 0 && t == 0, 0);
    double acc;
    if constexpr (DEPTH == 9) {
      // per warp: [10][kPark] parked words | [kRing9][96] doubles | [kRing9] tile ids (padded to 16 B); the
      // rings' mbarriers after the last warp's area
      unsigned char* wbase = s_dyn + static_cast<size_t>(t >> 5) * kWarpBytes9;
      acc = accumulate_points_parked<F, NW, kRing9>(
          P, s_T, s_R, it == 0, P.hit != nullptr && it == 0, reinterpret_cast<uint32_t*>(wbase),
          reinterpret_cast<double*>(wbase + 10u * kPark * 4u),
          reinterpret_cast<unsigned*>(wbase + 10u * kPark * 4u + kRing9 * kTileBytes),
          reinterpret_cast<uint64_t*>(s_dyn + static_cast<size_t>(NW) * kWarpBytes9) + (t >> 5) * kRing9, ring);
    } else if constexpr (DEPTH == 11) {
      acc = accumulate_points_parked<F, NW, 0, 2>(P, s_T, s_R, it == 0, P.hit != nullptr && it == 0,
                                                          reinterpret_cast<uint32_t*>(s_dyn) + static_cast<size_t>(t >> 5) * 10u * kPark,
                                                          nullptr, nullptr, nullptr, ring);
    } else if constexpr (DEPTH == 8) {
      acc = accumulate_points_parked<F, NW, 0>(P, s_T, s_R, it == 0, P.hit != nullptr && it == 0,
                                               reinterpret_cast<uint32_t*>(s_dyn) + static_cast<size_t>(t >> 5) * 10u * kPark,
                                               nullptr, nullptr, nullptr, ring);
    } else if constexpr (DEPTH == 7) {
      acc = accumulate_points_split<F, NW, 2>(P, s_T, s_R, it == 0, P.hit != nullptr && it == 0, hl, &s_tail);
    } else if constexpr (DEPTH == 6) {
      // pass set-up: empty queue; the share of consumer warps follows the hit rate of the last pass
      if (t < kNB) {
        hq->ready[t] = 0u;
        hq->gen[t] = 0u;
      }
      if (t == 0) {
        const unsigned n_tiles = (P.n + 31u) / 32u;
        const unsigned M = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + G - 1u) / G : 0u;
        unsigned nc = static_cast<unsigned>(P.n_cons);
        if (nc == 0u) {
          // producer trip ~ 1, consumer batch ~ 1.6 (instruction counts): consumers get h * 1.6 / (1 + h * 1.6)
          const float h = (it == 0 || M == 0u) ? 0.33f : static_cast<float>(hq->tail) / static_cast<float>(M * 32u);
          nc = static_cast<unsigned>(static_cast<float>(NW) * (1.6f * h) / (1.0f + 1.6f * h) + 0.5f);
        }
        if (nc < 2u) nc = 2u;
        if (nc > static_cast<unsigned>(NW) - 4u) nc = static_cast<unsigned>(NW) - 4u;
        hq->n_cons = nc;
        hq->tail = 0u;
        hq->head = 0u;
        hq->prod_done = 0u;
        hq->next_m = 0u;
      }
      __syncthreads();
      acc = accumulate_points_roles<F, NW>(P, s_T, s_R, it == 0, P.hit != nullptr && it == 0, wsm, hq);
    } else if constexpr (DEPTH == 5) {
      acc = accumulate_points_resident<F, NW>(P, s_T, s_R, it == 0, P.hit != nullptr && it == 0, wsm, wbar, ring);
    } else {
      acc = (NN == 1 && U == 1 && ESKF_PIPELINED)
                ? accumulate_points_pipelined<F, NW, DEPTH>(P, s_T, s_R, it == 0, P.hit != nullptr && it == 0)
                : accumulate_points<F, U, NN, NW>(P, s_T, s_R, it == 0, P.hit != nullptr && it == 0);
    }
    ESKF_STAMP(blockIdx.x == 0 && t == 0, 1);
    block_reduce_store<NW>(acc, s_part, P.partials + static_cast<size_t>(blockIdx.x) * kAcc);

    // last CTA to arrive reduces the partials and solves
    if (t == 0) {
      if constexpr (DEPTH == 7) s_tail = 0u;  // (every warp has read the pass's entry count before the block reduce)
      const unsigned ticket = atom_add_acq_rel(&st->block_counter, 1u);
      s_last = (ticket == G * static_cast<unsigned>(it + 1) - 1u) ? 1 : 0;
    }
    __syncthreads();
    const bool ll = P.ll != 0;
    if (s_last) {
      ESKF_STAMP(t == 0, 2);
      final_reduce<NW>(P.partials, G, s_part, s_sum);
      ESKF_STAMP(t == 0, 3);
      bool ok = true;
      if (P.world > 1) ok = P.xchg_ll ? exchange_sums_ll(P, it, s_sum, &s_xchg, s_xw) : exchange_sums(P, it, s_sum, &s_xchg);
      ESKF_STAMP(t == 0, 4);
      if (t < 32) {
        // (on an exchange timeout st->error is set: the waiters below bail out)
        if (ok) {
          solve_compute(P, s_sum, s_Ttot, it, s_solve);
          if (t == 12) st->tile_counter = 0u;  // (the one store the next pass needs, ahead of the broadcast)
        }
        ESKF_STAMP(t == 0, 5);
        if (ll) {
          if (ok) {
            ll_publish(P, s_solve, it);       // the other CTAs go on from here ...
            ESKF_STAMP(t == 0, 6);
            solve_store(P, s_sum, it, s_solve);  // ... while the state and trace words are written
          }
        } else {
          if (ok) solve_store(P, s_sum, it, s_solve);
          if (t == 0) st_release_u32(&st->epoch, static_cast<unsigned>(it + 1));
          ESKF_STAMP(t == 0, 6);
        }
      }
    }
    // everyone waits for the solver (bounded spin).  LL: the lanes of warp 0 poll the flagged words
    // that carry the next pose.  Otherwise warp 0 polls the epoch word, then its lanes fetch the next
    // pose and the done flag in one more round trip.
    if (t < 32) {
      if (ll) {
        ll_receive<F>(P, it, s_T, &s_done);
      } else {
        unsigned spins = 0;
        int ok = 1;
        while (ld_acquire_u32(&st->epoch) < static_cast<unsigned>(it + 1)) {
          __nanosleep(20);
          if (++spins > kSpinLimit || ld_acquire_u32(&st->error) != 0) {
            atomicCAS(&st->error, 0u, 1u);
            ok = 0;
            break;
          }
        }
        ok = __all_sync(0xffffffffu, ok);
        if (t < 12) s_T[t] = ld_cg(&st->T_step[t]);
        else if (t == 21) s_done = (ok && ld_acquire_u32(&st->error) == 0) ? ld_cg(&st->done) : 1;
      }
      // total pose <- step * total pose (the same expression as the solver's), and its rotation for the
      // per-point algebra
      __syncwarp();
      double nt = 0.0;
      if (t < 12) nt = compose_elem(s_T, s_Ttot, static_cast<int>(t));
      __syncwarp();
      if (t < 12) s_Ttot[t] = nt;
      if (t < 9) s_R[t] = F(nt);
    }
    __syncthreads();
    ESKF_STAMP(blockIdx.x == 0 && t == 0, 7);
    if (s_done) return;
  }
}

// sharded mode, after the caller's all-reduce of `sums`: one thread solves
__global__ void solve_kernel(AlignParams P, int it) {
  __shared__ double s_solve[96];
  __shared__ double s_in[kAcc];
  if (threadIdx.x < kAcc) s_in[threadIdx.x] = P.sums[threadIdx.x];
  __syncthreads();
  __shared__ double s_old[12];
  if (threadIdx.x < 12) s_old[threadIdx.x] = (it == 0) ? P.guess[threadIdx.x] : P.st->T_total[threadIdx.x];
  __syncthreads();
  if (threadIdx.x < 32) solve_and_update_fast(P, s_in, s_old, it, s_solve);
}

// sharded mode: ONE linearisation of this rank's point range; the 28 sums go
// to P.sums for the caller's all-reduce, solve_kernel follows
template <typename F, int U, int NN, int MINB, int T = kT, int DEPTH = 3>
__global__ void __launch_bounds__(T, MINB) linearize_pass_kernel(AlignParams P, int it) {
  constexpr int NW = T / 32;
  __shared__ double s_T[12];
  __shared__ F s_R[9];
  __shared__ double s_part[NW][32];
  __shared__ double s_sum[kAcc];
  __shared__ int s_flag;
  const unsigned G = gridDim.x, t = threadIdx.x;
  AlignState* st = P.st;
  if (t < 12) s_T[t] = (it == 0) ? P.guess[t] : ld_cg(&st->T_step[t]);
  if (t < 9) s_R[t] = (it == 0) ? F(P.guess[t]) : F(ld_cg(&st->Rf[t]));
  __syncthreads();
  const double acc = (NN == 1 && U == 1 && ESKF_PIPELINED)
                         ? accumulate_points_pipelined<F, NW, DEPTH>(P, s_T, s_R, it == 0, false)
                         : accumulate_points<F, U, NN, NW>(P, s_T, s_R, it == 0, false);
  block_reduce_store<NW>(acc, s_part, P.partials + static_cast<size_t>(blockIdx.x) * kAcc);
  if (t == 0) {
    const unsigned ticket = atom_add_acq_rel(&st->block_counter, 1u);
    s_flag = (ticket == G * static_cast<unsigned>(it + 1) - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (s_flag) {
    final_reduce<NW>(P.partials, G, s_part, s_sum);
    if (t < kAcc) P.sums[t] = s_sum[t];
  }
}


// ------------------------------------------------------------------------
// Large clouds (depth 10): one Gauss-Newton iteration = three launches.
//
// Every persistent variant above ends up with 16-20 warps per SM (96-128
// registers for the gathers' landing zone and the algebra) and ~13 stall cycles
// per issued instruction per warp: 26-41 % of the issue slots
// (profiles/r2_align_experiments.md).  The lookup half of the work — transform,
// key, hash, one 16 B filter probe — needs ~40 registers.  So the iteration is cut
// where the register needs differ:
//   mk_lookup_kernel   1536 threads per SM.  A CTA round = 16 tiles (one per warp);
//                      candidates are staged in shared memory, then appended to a
//                      hit list in HBM with ONE global atomic per round and
//                      coalesced 4-byte-per-lane stores (SoA, 40 B per candidate);
//   mk_gather_kernel   dense batches of 32 candidates from the list: records +
//                      source covariances (next batch in flight while the current
//                      one is linearised), block reduce, last CTA sums the partials;
//   solve_kernel       [NVLink exchange of the sums] 6x6 solve, pose, convergence.
// No cooperative launch, no in-kernel hand-off; the launches of an iteration are
// queued back to back (the host looks at the convergence word every few iterations
// only: kernels of iterations past convergence return at once).
constexpr int kMkLookupT = 512;
constexpr int kMkGatherT = 640;
constexpr unsigned kMkStage = kMkLookupT;  // candidates a CTA round can produce

struct MkList {  // the hit list: field f of entry e at words[f * cap + e]
  uint32_t* words;
  unsigned* count;  // entries appended in this iteration (reset by the solve kernel)
  unsigned cap;
};

__global__ void __launch_bounds__(kMkLookupT, 3) mk_lookup_kernel(AlignParams P, MkList L, int it) {
  __shared__ double s_T[12];
  __shared__ uint32_t s_stage[10][kMkStage];
  __shared__ unsigned s_n, s_base;
  const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;
  constexpr unsigned NW = kMkLookupT / 32;
  const AlignState* st = P.st;
  if (it > 0 && ld_cg(&st->done) != 0) return;
  const bool first = it == 0;
  if (t < 12) s_T[t] = first ? P.guess[t] : ld_cg(&st->T_step[t]);
  if (t == 0) s_n = 0u;
  __syncthreads();
  const double* sx = first ? P.x0 : P.wx;
  const double* sy = first ? P.y0 : P.wy;
  const double* sz = first ? P.z0 : P.wz;
  const unsigned n_tiles = (P.n + 31u) / 32u;
  const double inv_voxel = 1.0 / P.voxel;
  const uint64_t pol_filt = l2_policy((P.flags >> kFlagFiltPolicyShift) & 3u);
  const bool write_hit = P.hit != nullptr && first;
  for (unsigned base_tile = blockIdx.x * NW; base_tile < n_tiles; base_tile += gridDim.x * NW) {
    const unsigned tile = base_tile + warp;
    const unsigned i = tile * 32u + lane;
    uint32_t cand = kNoCand, klo = 0u, khi = 0u;
    float px = 0.f, py = 0.f, pz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
    if (tile < n_tiles && i < P.n) {
      double x = first ? __ldcs(sx + i) : __ldcg(sx + i);
      double y = first ? __ldcs(sy + i) : __ldcg(sy + i);
      double z = first ? __ldcs(sz + i) : __ldcg(sz + i);
      transform_point_rn(s_T, x, y, z);
      P.wx[i] = x;
      P.wy[i] = y;
      P.wz[i] = z;
      const int kx = voxel_coord(x, P.voxel, inv_voxel);
      const int ky = voxel_coord(y, P.voxel, inv_voxel);
      const int kz = voxel_coord(z, P.voxel, inv_voxel);
      if (coord_in_range(kx) && coord_in_range(ky) && coord_in_range(kz)) {
        const uint64_t key = pack_key(kx, ky, kz);
        const SlotAddr ad = slot_addr(key, P.n_slots);
        const uint32_t t8 = filter_tag(ad.tag);
        uint32_t b0 = ad.home & ~7u;
        uint32_t r = scan_filter_window(filter_window(P.filt, ad.home, pol_filt), ad.home & 7u, t8);
        uint32_t scanned = 16u - (ad.home & 7u);
        while (r == kMore && scanned < P.n_slots) {
          b0 += 16u;
          if (b0 >= P.n_slots) b0 -= P.n_slots;
          r = scan_filter_window(filter_window(P.filt, b0, pol_filt), 0u, t8);
          scanned += 16u;
        }
        if (r < 16u) {
          cand = b0 + r;
          if (cand >= P.n_slots) cand -= P.n_slots;
          klo = static_cast<uint32_t>(key);
          khi = static_cast<uint32_t>(key >> 32);
          px = static_cast<float>(x); py = static_cast<float>(y); pz = static_cast<float>(z);
          // the residual against the voxel mean is formed relative to the voxel centre
          dx = static_cast<float>(x - __dmul_rn(static_cast<double>(kx) + 0.5, P.voxel));
          dy = static_cast<float>(y - __dmul_rn(static_cast<double>(ky) + 0.5, P.voxel));
          dz = static_cast<float>(z - __dmul_rn(static_cast<double>(kz) + 0.5, P.voxel));
        }
      }
      if (write_hit && cand == kNoCand) P.hit[i] = 0;
    }
    // stage the round's candidates in shared memory (one shared-memory atomic per warp)
    const unsigned mask = __ballot_sync(0xffffffffu, cand != kNoCand);
    unsigned wbase = 0;
    if (lane == 0 && mask != 0u) wbase = atomicAdd(&s_n, static_cast<unsigned>(__popc(mask)));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (cand != kNoCand) {
      const unsigned e = wbase + __popc(mask & ((1u << lane) - 1u));
      s_stage[0][e] = __float_as_uint(px); s_stage[1][e] = __float_as_uint(py); s_stage[2][e] = __float_as_uint(pz);
      s_stage[3][e] = __float_as_uint(dx); s_stage[4][e] = __float_as_uint(dy); s_stage[5][e] = __float_as_uint(dz);
      s_stage[6][e] = klo; s_stage[7][e] = khi; s_stage[8][e] = cand; s_stage[9][e] = i;
    }
    __syncthreads();
    // ... and append them to the list: one global atomic per round, coalesced stores
    const unsigned cnt = s_n;
    if (t == 0) s_base = cnt != 0u ? atomicAdd(L.count, cnt) : 0u;
    __syncthreads();
    const unsigned gbase = s_base;
    if (t < cnt && gbase + t < L.cap) {
#pragma unroll
      for (int f = 0; f < 10; ++f) L.words[static_cast<size_t>(f) * L.cap + gbase + t] = s_stage[f][t];
    }
    if (t == 0) s_n = 0u;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kMkGatherT, 1) mk_gather_kernel(AlignParams P, MkList L, int it) {
  constexpr int NW = kMkGatherT / 32;
  __shared__ float s_R[9];
  __shared__ double s_part[NW][32];
  __shared__ double s_sum[kAcc];
  __shared__ int s_flag;
  const unsigned t = threadIdx.x, lane = t & 31, G = gridDim.x;
  AlignState* st = P.st;
  if (it > 0 && ld_cg(&st->done) != 0) return;
  if (t < 9) s_R[t] = it == 0 ? static_cast<float>(P.guess[t]) : ld_cg(&st->Rf[t]);
  __syncthreads();
  unsigned total = ld_cg(L.count);
  if (total > L.cap) total = L.cap;
  const unsigned n_batches = (total + 31u) / 32u;
  const uint64_t pol_rec = l2_policy(P.flags & kFlagRecPolicyMask);
  const bool rec_prefetch = (P.flags & kFlagNoRecPrefetch) == 0u;
  const bool write_hit = P.hit != nullptr && it == 0;
  struct Entry {
    float px, py, pz, dx, dy, dz;
    uint32_t klo, khi, cand, idx;
  };
  struct RecRegs {
    uint2 key;
    float4 pa, pc;
    float2 pd;
    float4 s4;
    float2 s2;
  };
  auto fetch = [&](unsigned j, Entry& e, RecRegs& r) {
    const unsigned idx = j * 32u + lane;
    e.cand = kNoCand;
    r.key = make_uint2(0u, 0u);
    if (j >= n_batches || idx >= total) return;
    const uint32_t* w = L.words + idx;
    const size_t pitch = L.cap;
    e.px = __uint_as_float(ld_cg(w)); e.py = __uint_as_float(ld_cg(w + pitch)); e.pz = __uint_as_float(ld_cg(w + 2 * pitch));
    e.dx = __uint_as_float(ld_cg(w + 3 * pitch)); e.dy = __uint_as_float(ld_cg(w + 4 * pitch));
    e.dz = __uint_as_float(ld_cg(w + 5 * pitch));
    e.klo = ld_cg(w + 6 * pitch); e.khi = ld_cg(w + 7 * pitch); e.cand = ld_cg(w + 8 * pitch); e.idx = ld_cg(w + 9 * pitch);
    const float4* rec = reinterpret_cast<const float4*>(P.slots + e.cand);
    if (rec_prefetch) prefetch_record(rec);
    r.key = ldg_u2_hint(rec, pol_rec);
    r.pa = ldg_f4_hint(rec + 1, pol_rec);
    r.pc = ldg_f4_hint(rec + 2, pol_rec);
    r.pd = ldg_f2_hint(rec + 3, pol_rec);
    r.s4 = __ldcs(P.c4 + e.idx);
    r.s2 = __ldcs(P.c2 + e.idx);
  };
  double acc = 0.0;
  float v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = 0.f;
  auto compute = [&](const Entry& e, const RecRegs& r) {
    float4 pa = r.pa, pc = r.pc;
    float2 pd = r.pd;
    bool hit = false;
    if (e.cand != kNoCand) {
      hit = r.key.x == e.klo && r.key.y == e.khi;
      if (!hit) {  // 8-bit filter collision: walk on, slowly
        const uint64_t key = (static_cast<uint64_t>(e.khi) << 32) | e.klo;
        const SlotAddr ad = slot_addr(key, P.n_slots);
        const VoxelSlot* far = resolve_probe_filter(P.filt, P.slots, P.n_slots, key, next_slot(e.cand, P.n_slots),
                                                    filter_tag(ad.tag));
        if (far != nullptr) {
          hit = true;
          pa = __ldg(reinterpret_cast<const float4*>(far) + 1);
          pc = __ldg(reinterpret_cast<const float4*>(far) + 2);
          pd = __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float4*>(far) + 3));
        }
      }
      if (write_hit) P.hit[e.idx] = hit ? 1 : 0;
    }
    if (hit) {
      float cr[6];
      rotate_sym<float>(s_R, r.s4.x, r.s4.y, r.s4.z, r.s4.w, r.s2.x, r.s2.y, cr);
      point_terms<float, true>(e.px, e.py, e.pz, e.dx - pa.x, e.dy - pa.y, e.dz - pa.z, cr[0] + pc.x, cr[1] + pc.y,
                               cr[2] + pc.z, cr[3] + pc.w, cr[4] + pd.x, cr[5] + pd.y, v);
    } else {
#pragma unroll
      for (int k = 0; k < 28; ++k) v[k] = 0.f;
    }
    acc += static_cast<double>(warp_reduce_scatter32<float>(v, lane));
  };
  const unsigned wglobal = blockIdx.x * NW + (t >> 5), wstride = G * NW;
  Entry cur, nxt;
  RecRegs rec, nrec;
  unsigned j = wglobal;
  fetch(j, cur, rec);
  while (j < n_batches) {
    fetch(j + wstride, nxt, nrec);
    compute(cur, rec);
    cur = nxt;
    rec = nrec;
    j += wstride;
  }
  block_reduce_store<NW>(acc, s_part, P.partials + static_cast<size_t>(blockIdx.x) * kAcc);
  if (t == 0) {
    const unsigned ticket = atom_add_acq_rel(&st->block_counter, 1u);
    s_flag = (ticket == G * static_cast<unsigned>(it + 1) - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (s_flag) {
    final_reduce<NW>(P.partials, G, s_part, s_sum);
    if (t < kAcc) P.sums[t] = s_sum[t];
  }
}

// [NVLink exchange of the sums,] solve, pose, convergence; resets the hit list for the next iteration
__global__ void __launch_bounds__(kMkGatherT) mk_solve_kernel(AlignParams P, MkList L, int it) {
  __shared__ double s_solve[96];
  __shared__ double s_in[kAcc];
  __shared__ double s_old[12];
  __shared__ unsigned s_xw[kMaxWorld * 2 * kAcc];
  __shared__ int s_xchg;
  const unsigned t = threadIdx.x;
  if (it > 0 && ld_cg(&P.st->done) != 0) return;
  if (t < kAcc) s_in[t] = ld_cg(P.sums + t);
  if (t < 12) s_old[t] = (it == 0) ? P.guess[t] : ld_cg(&P.st->T_total[t]);
  if (t == 0) *L.count = 0u;
  __syncthreads();
  bool ok = true;
  if (P.world > 1) ok = P.xchg_ll ? exchange_sums_ll(P, it, s_in, &s_xchg, s_xw) : exchange_sums(P, it, s_in, &s_xchg);
  if (t < 32 && ok) solve_and_update_fast(P, s_in, s_old, it, s_solve);
}

// kernel variants: (math type, points per lane U, neighbourhood, min CTAs/SM)
struct Variant {
  void* align;
  void* pass;
  int threads;  // CTA size
  int pts_per_block;
  int per_sm;  // filled by align_max_blocks()
  int depth5;  // 1: accumulate_points_resident, 2: accumulate_points_roles (dynamic shared memory, probe filter)
  size_t max_dyn_smem;  // depth 5: opt-in dynamic shared memory available to the kernel (align_max_blocks)
};

// The 1-neighbour fp32 kernel exists at three CTA shapes with the same 24 warps per SM:
// 3 x 256, 2 x 384 and 1 x 768 threads.  Fewer, fatter CTAs shorten the per-iteration
// hand-off of a large cloud (148 instead of 444 partial sums to reduce, tickets to take and
// epoch pollers); small clouds keep 256-thread CTAs so that they spread over more SMs.
// The fat shapes also exist with the 4-deep load rotation (accumulate_points_pipelined): at 768
// threads it has to live in 80 registers; 640 threads (20 warps) get 96 and 512 (16 warps) 128
// (ptxas budgets registers for the CTA size rounded up to a multiple of 128 threads).
enum { V_F32_N1 = 0, V_F32_N7, V_F64_N1, V_F64_N7, V_F32_N1_T384, V_F32_N1_T768, V_F32_N1_T768D4,
       V_F32_N1_T640D4, V_F32_N1_T512D4, V_F32_N1_T640R, V_F32_N1_T512R, V_F32_N1_T640Q, V_F32_N1_T512Q,
       V_F32_N1_T640S, V_F32_N1_T512S, V_F32_N1_T640P, V_F32_N1_T512P, V_F32_N1_T768P, V_F32_N1_T640B, V_F32_N1_T512B, V_F32_N1_T448D4, V_F32_N1_T384D4, V_F32_N1_T512U2, V_F32_N1_T384U2, V_F32_N1_F, V_COUNT };

// measured on B200, dense config (2M pts, 10 iterations per launch):
//   U=1/3 CTAs 1.267 ms, U=1/4 CTAs 1.246 ms, U=2/3 CTAs 1.303 ms,
//   U=2/2 CTAs 1.396 ms, U=4/2 CTAs 1.417 ms  
#ifndef ESKF_ALIGN_U
#define ESKF_ALIGN_U 1
#endif
#ifndef ESKF_ALIGN_MINB
#define ESKF_ALIGN_MINB 3
#endif
#ifndef ESKF_ALIGN_FAT_T
// measured on B200, dense config, us per GN iteration at 0.1 m / 0.5 m voxels (tables as grown by the
// bulk inserts), 2-tile tickets; first box: 3 x 256 threads 86.9 / 163.9, 2 x 384 85.6 / 161.7,
// 1 x 768 85.1 / 158.2; a slower box: 3 x 256 105.5 / 162.8, 1 x 768 101.3 / 158.8, 1 x 640 with the
// 4-deep rotation 93.6 / 148.7, 1 x 512 4-deep 102.3 / 147.8, 1 x 768 4-deep (spills at 80 registers)
// 177.7 / 241.1  (profiles/r1_align_ab.md)
#define ESKF_ALIGN_FAT_T 512  // CTA size of the large-cloud kernel (256 | 384 | 448 | 512 | 640 | 768); round 2, on the 8-bit filter: 512 threads (128 registers) 86.6 us per dense iteration, 640 (96 registers) 95.2
#endif

Variant g_variants[V_COUNT] = {
    {reinterpret_cast<void*>(align_kernel<float, ESKF_ALIGN_U, 1, ESKF_ALIGN_MINB>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, ESKF_ALIGN_U, 1, ESKF_ALIGN_MINB>),
     kT, kT * ESKF_ALIGN_U, 1},
    {reinterpret_cast<void*>(align_kernel<float, 1, 7, 2>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 7, 2>), kT, kT, 1},
    {reinterpret_cast<void*>(align_kernel<double, 1, 1, 1>),
     reinterpret_cast<void*>(linearize_pass_kernel<double, 1, 1, 1>), kT, kT, 1},
    {reinterpret_cast<void*>(align_kernel<double, 1, 7, 1>),
     reinterpret_cast<void*>(linearize_pass_kernel<double, 1, 7, 1>), kT, kT, 1},
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 2, 384>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 2, 384>), 384, 384, 1},
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, kMaxT>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, kMaxT>), kMaxT, kMaxT, 1},
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, kMaxT, 4>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, kMaxT, 4>), kMaxT, kMaxT, 1},
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 640, 4>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 640, 4>), 640, 640, 1},
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 512, 4>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 512, 4>), 512, 512, 1},
    // depth 5 (SM-resident positions, probe filter, bulk-copy ring); the one-pass-per-launch kernels
    // of the NCCL-baseline mode cannot keep shared memory between launches and stay on depth 4
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 640, 5>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 640, 4>), 640, 640, 1, 1},
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 512, 5>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 512, 4>), 512, 512, 1, 1},
    // depth 6 (producer / consumer warps around a shared-memory hit queue)
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 640, 6>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 640, 4>), 640, 640, 1, 2},
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 512, 6>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 512, 4>), 512, 512, 1, 2},
    // depth 7 (phase-split passes around a per-CTA hit list)
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 640, 7>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 640, 4>), 640, 640, 1, 3},
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 512, 7>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 512, 4>), 512, 512, 1, 3},
    // depth 8 (depth 4's register pipeline, candidates parked per warp until 32 are there)
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 640, 8>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 640, 4>), 640, 640, 1, 4},
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 512, 8>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 512, 4>), 512, 512, 1, 4},
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 768, 8>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 768, 4>), 768, 768, 1, 4},
    // depth 9 (depth 8 with the positions streamed through a cp.async.bulk ring)
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 640, 9>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 640, 4>), 640, 640, 1, 5},
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 512, 9>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 512, 4>), 512, 512, 1, 5},
    // depth 4 with fewer, fatter threads (144 / 168 registers)
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 448, 4>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 448, 4>), 448, 448, 1},
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 384, 4>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 384, 4>), 384, 384, 1},
    // depth 11 (parked candidates, 2 tiles of lookups per trip; 3 per trip measured 157 us at 384 threads)
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 512, 11>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 512, 4>), 512, 512, 1, 4},
    {reinterpret_cast<void*>(align_kernel<float, 1, 1, 1, 384, 11>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, 1, 1, 1, 384, 4>), 384, 384, 1, 4},
    // the 3-stage loop on the 8-bit probe filter (large clouds; "align_block" 257)
    {reinterpret_cast<void*>(align_kernel<float, ESKF_ALIGN_U, 1, ESKF_ALIGN_MINB, kT, 2>),
     reinterpret_cast<void*>(linearize_pass_kernel<float, ESKF_ALIGN_U, 1, ESKF_ALIGN_MINB>), kT, kT * ESKF_ALIGN_U, 1},
};

// Large clouds (>= ctx->opt_align_fat_points, default 2^17) take one fat CTA per SM: fewer partial
// sums, tickets and pollers at the per-iteration hand-off, and with depth 5 the working positions
// stay in shared memory.  Frames (tens of thousands of points) keep 256-thread CTAs.
int variant_index(const eskf_ctx* ctx, const AlignArgs& a) {
  const int base = (a.fp64_math ? 2 : 0) + (a.neighbor_mode == 7 ? 1 : 0);
  if (base != V_F32_N1) return base;
  int threads = ctx->opt_align_block;  // 0 = choose by cloud size
  const bool fat = a.cloud && static_cast<int64_t>(a.cloud->n) >= ctx->opt_align_fat_points;
  // (a large cloud: the 4-deep 512-thread shape, unless this context's one-off timing of the two shapes
  // on its own device and data said the 3-stage 256-thread one is faster there: autotune_fat below)
  if (threads == 0) threads = fat ? (ctx->tuned_block != 0 && ctx->opt_align_depth == 0 ? ctx->tuned_block : ESKF_ALIGN_FAT_T) : kT;
  const int depth = ctx->opt_align_depth != 0 ? ctx->opt_align_depth : 4;
  switch (threads) {
    case 768: return depth == 8 ? V_F32_N1_T768P : depth == 4 ? V_F32_N1_T768D4 : V_F32_N1_T768;
    case 640: return depth == 9 ? V_F32_N1_T640B : depth == 8 ? V_F32_N1_T640P : depth == 7 ? V_F32_N1_T640S : depth == 6 ? V_F32_N1_T640Q : depth == 5 ? V_F32_N1_T640R : V_F32_N1_T640D4;
    case 512: return depth == 11 ? V_F32_N1_T512U2 : depth == 9 ? V_F32_N1_T512B : depth == 8 ? V_F32_N1_T512P : depth == 7 ? V_F32_N1_T512S : depth == 6 ? V_F32_N1_T512Q : depth == 5 ? V_F32_N1_T512R : V_F32_N1_T512D4;
    case 448: return V_F32_N1_T448D4;
    case 384: return depth == 11 ? V_F32_N1_T384U2 :
                     depth == 4 && ctx->opt_align_depth == 4 ? V_F32_N1_T384D4 : V_F32_N1_T384;
    case 257: return V_F32_N1_F;
    case 769: return V_F32_N1_T768;  // (the 3-stage loop at 768 threads whatever "align_depth" says)
    default: return V_F32_N1;
  }
}

struct TraceLayout {
  size_t o_state, o_sums, o_ll, o_H, o_b, o_nc, o_step, o_stamps, total;
};

TraceLayout trace_layout(int max_it) {
  TraceLayout L;
  L.o_state = 0;
  L.o_sums = 512;
  L.o_ll = L.o_sums + 32 * 8;  // flagged words of the pose broadcast (zeroed with the state)
  L.o_H = L.o_ll + 40 * 8;
  L.o_b = L.o_H + static_cast<size_t>(max_it) * 36 * 8;
  L.o_nc = L.o_b + static_cast<size_t>(max_it) * 6 * 8;
  L.o_step = L.o_nc + static_cast<size_t>(max_it) * 8;
  L.o_stamps = L.o_step + static_cast<size_t>(max_it) * 12 * 8;
  L.total = L.o_stamps + static_cast<size_t>(max_it) * kStampSlots * 8;
  return L;
}
static_assert(sizeof(AlignState) <= 512, "AlignState grew past its slot");

int fill_params(eskf_ctx* ctx, const AlignArgs& a, int max_it, AlignParams* P, TraceLayout* L, int* G,
                size_t* dyn_smem = nullptr) {
  const eskf_map* m = a.map;
  const eskf_cloud* c = a.cloud;
  ESKF_REQUIRE(m && c, "null map/cloud");
  ESKF_REQUIRE(m->ctx == ctx && c->ctx == ctx, "map/cloud belong to another context");
  ESKF_REQUIRE(c->has_cov && c->has_c32, "registration needs a cloud with covariances");
  ESKF_REQUIRE(a.neighbor_mode == 1 || a.neighbor_mode == 7, "neighbor_mode must be 1 or 7");
  ESKF_REQUIRE(max_it > 0, "max_iteration must be positive");
  ESKF_REQUIRE(c->n < (1ull << 31), "cloud too large");
  const unsigned n = static_cast<unsigned>(c->n);
  const size_t pitch = (static_cast<size_t>(n) + 63) / 64 * 64 + 64;
  ESKF_TRY(ctx->work.ensure(pitch * 3 * sizeof(double)));
  const Variant& var = g_variants[variant_index(ctx, a)];
  int g = static_cast<int>((n + var.pts_per_block - 1) / var.pts_per_block);
  const int g_max = var.per_sm * ctx->sm_count;
  if (g > g_max) g = g_max;
  if (g < 1) g = 1;
  *G = g;
  ESKF_TRY(ctx->partials.ensure(static_cast<size_t>(g) * kAcc * sizeof(double)));
  *L = trace_layout(max_it);
  ESKF_TRY(ctx->astate.ensure(L->total));
  char* base = ctx->astate.as<char>();
  std::memset(P, 0, sizeof *P);
  P->tags = m->tags;
  P->slots = m->slots;
  P->n_slots = static_cast<uint32_t>(m->n_slots);
  P->voxel = m->voxel;
  P->x0 = c->x();
  P->y0 = c->y();
  P->z0 = c->z();
  P->c4 = c->c4;
  P->c2 = c->c2;
  P->wx = ctx->work.as<double>();
  P->wy = P->wx + pitch;
  P->wz = P->wx + 2 * pitch;
  P->n = n;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) P->guess[3 * i + j] = a.guess[4 * i + j];
    P->guess[9 + i] = a.guess[4 * i + 3];
  }
  P->max_iteration = a.max_iteration;
  P->neighbor_mode = a.neighbor_mode;
  P->trans_sq_thr = a.trans_sq_thr;
  P->cos_thr = a.cos_thr;
  P->fixed_iterations = a.fixed_iterations;
  P->dynamic_tiles = ctx->opt_align_dynamic;
  P->ticket_chunk = ctx->opt_align_chunk;
  P->dyn16 = ctx->opt_align_dyn16;
  P->st = reinterpret_cast<AlignState*>(base + L->o_state);
  P->partials = ctx->partials.as<double>();
  P->sums = reinterpret_cast<double*>(base + L->o_sums);
  P->trace_H = reinterpret_cast<double*>(base + L->o_H);
  P->trace_b = reinterpret_cast<double*>(base + L->o_b);
  P->trace_ncorr = reinterpret_cast<unsigned long long*>(base + L->o_nc);
  P->trace_step = reinterpret_cast<double*>(base + L->o_step);
  P->hit = a.d_hit;
  P->world = 1;
  P->rank = 0;
  P->mail = nullptr;
  P->llbox = reinterpret_cast<unsigned long long*>(base + L->o_ll);
  P->stamps = ctx->opt_align_stamps ? reinterpret_cast<unsigned long long*>(base + L->o_stamps) : nullptr;
  P->ll = ctx->opt_align_ll;
  P->flags = static_cast<unsigned>(ctx->opt_align_flags);
  if (dyn_smem) *dyn_smem = 0;
  P->n_cons = ctx->opt_align_cons;
  P->xchg_ll = ctx->opt_align_xchg_ll;
  {
    const int vi = variant_index(ctx, a);
    // large clouds probe the 8-bit filter (frames keep the 16-bit tags: their map changes every frame and
    // the filter would have to follow it)
    const bool fat = static_cast<int64_t>(n) >= ctx->opt_align_fat_points;
    const bool d4 = vi == V_F32_N1_T768D4 || vi == V_F32_N1_T640D4 || vi == V_F32_N1_T512D4 || vi == V_F32_N1_T448D4 ||
                    vi == V_F32_N1_T384D4;
    (void)fat;
    if ((d4 || vi == V_F32_N1_F) && ctx->opt_align_filter && dyn_smem) ESKF_TRY(map_probe_filter(m, &P->filt));
    ESKF_REQUIRE(vi != V_F32_N1_F || P->filt != nullptr, "align_block 257 (3-stage loop on the probe filter) needs align_filter on");
  }
  if (var.depth5 == 5 && dyn_smem) {
    const size_t nw = static_cast<size_t>(var.threads / 32);
    *dyn_smem = nw * kWarpBytes9 + nw * kRing9 * sizeof(uint64_t);
    ESKF_TRY(map_probe_filter(m, &P->filt));
  }
  if (var.depth5 == 4 && dyn_smem) {
    *dyn_smem = static_cast<size_t>(var.threads / 32) * 10u * kPark * sizeof(uint32_t);
    ESKF_TRY(map_probe_filter(m, &P->filt));
  }
  if (var.depth5 == 3 && dyn_smem) {
    // hit list: as many 40 B entries as shared memory holds (a multiple of 32), the rest of a CTA's
    // worst case (every point of its tiles finds a voxel) in its spill region
    const unsigned n_tiles = (n + 31u) / 32u;
    const unsigned per_cta = (n_tiles + static_cast<unsigned>(g) - 1u) / static_cast<unsigned>(g) * 32u;  // points
    unsigned cap = static_cast<unsigned>(var.max_dyn_smem / 40u) / 32u * 32u;
    if (ctx->opt_align_resident >= 0 && static_cast<unsigned>(ctx->opt_align_resident) * 32u < cap)
      cap = static_cast<unsigned>(ctx->opt_align_resident) * 32u;  // (tests: force the spill path)
    if (cap > per_cta) cap = per_cta;
    if (cap < 32u) cap = 32u;
    const unsigned spill_cap = per_cta > cap ? per_cta - cap : 32u;
    ESKF_TRY(ctx->spill.ensure(static_cast<size_t>(g) * 10u * spill_cap * sizeof(uint32_t)));
    P->spill = ctx->spill.as<uint32_t>();
    P->spill_cap = spill_cap;
    P->k_res = static_cast<int>(cap);
    *dyn_smem = static_cast<size_t>(cap) * 40u;
    ESKF_TRY(map_probe_filter(m, &P->filt));
  }
  if (var.depth5 == 2 && dyn_smem) {
    // resident tiles per CTA: all of the CTA's tiles (b, b + G, ...) if shared memory holds them
    const unsigned n_tiles = (n + 31u) / 32u;
    const unsigned per_cta = (n_tiles + static_cast<unsigned>(g) - 1u) / static_cast<unsigned>(g);
    const size_t room = var.max_dyn_smem > sizeof(HitQueue) ? var.max_dyn_smem - sizeof(HitQueue) : 0;
    const unsigned k_max = static_cast<unsigned>(room / kTileBytes);
    unsigned k = per_cta < k_max ? per_cta : k_max;
    if (ctx->opt_align_resident >= 0 && static_cast<unsigned>(ctx->opt_align_resident) < k)
      k = static_cast<unsigned>(ctx->opt_align_resident);
    P->k_res = static_cast<int>(k);
    *dyn_smem = static_cast<size_t>(k) * kTileBytes + sizeof(HitQueue);
    ESKF_TRY(map_probe_filter(m, &P->filt));
  }
  if (var.depth5 == 1 && dyn_smem) {
    // resident tiles per warp: all of a warp's statically dealt tiles if shared memory holds them
    const unsigned nw = static_cast<unsigned>(var.threads / 32);
    const unsigned n_tiles = (n + 31u) / 32u;
    const unsigned warps = static_cast<unsigned>(g) * nw;
    unsigned per_warp = (n_tiles + warps - 1u) / warps;
    const size_t fixed = static_cast<size_t>(nw) * kRing * (kTileBytes + sizeof(uint64_t));
    const size_t room = var.max_dyn_smem > fixed ? var.max_dyn_smem - fixed : 0;
    const unsigned k_max = static_cast<unsigned>(room / (static_cast<size_t>(nw) * kTileBytes));
    unsigned k = per_warp < k_max ? per_warp : k_max;
    if (ctx->opt_align_resident >= 0 && static_cast<unsigned>(ctx->opt_align_resident) < k)
      k = static_cast<unsigned>(ctx->opt_align_resident);
    P->k_res = static_cast<int>(k);
    *dyn_smem = static_cast<size_t>(nw) * (k + kRing) * kTileBytes + static_cast<size_t>(nw) * kRing * sizeof(uint64_t);
    ESKF_TRY(map_probe_filter(m, &P->filt));
  }
  return ESKF_OK;
}

// copy state + traces back and fill the caller's outputs
int read_back(eskf_ctx* ctx, const AlignArgs& a, const TraceLayout& L, int max_it, double T_out[16],
              eskf_align_info* info, unsigned mail_seq = 0u) {
  const bool want_trace = info && (info->trace_H || info->trace_b || info->trace_ncorr || info->trace_step);
  if (mail_seq != 0u && !want_trace) {
    const int w = wait_mail(ctx, &ctx->mail_h->align_seq, mail_seq);
    if (w == ESKF_ERR_CUDA) return w;
    if (w == ESKF_OK) {
      const HostMail* m = ctx->mail_h;
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T_out[4 * i + j] = m->align_T[3 * i + j];
        T_out[4 * i + 3] = m->align_T[9 + i];
      }
      T_out[12] = 0.0; T_out[13] = 0.0; T_out[14] = 0.0; T_out[15] = 1.0;
      if (info) {
        info->iterations = m->align_iter;
        info->converged = m->align_converged;
        info->n_corr_last = m->align_ncorr;
      }
      return ESKF_OK;
    }
    // the stream went idle without a published result (error path): read the state the slow way
  }
  char* h = nullptr;
  ESKF_TRY(ctx_pinned(ctx, L.total, reinterpret_cast<void**>(&h)));
  const size_t bytes = want_trace ? L.total : L.o_sums;
  ESKF_CUDA(cudaMemcpyAsync(h, ctx->astate.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
  const AlignState* st = reinterpret_cast<const AlignState*>(h + L.o_state);
  if (ctx->opt_align_stamps && want_trace) {
    // ESKF_ALIGN_STAMPS=1 (debug aid): where an iteration's time goes, in us relative to the start of the pass on
    // CTA 0: its own pass end | last CTA's arrival | partials reduced | sums exchanged | solved | pose published |
    // next pose received by CTA 0
    const unsigned long long* s8 = reinterpret_cast<const unsigned long long*>(h + L.o_stamps);
    const int nit = st->iter < max_it ? st->iter : max_it;
    for (int k = 0; k < nit; ++k) {
      char line[256];
      int o = std::snprintf(line, sizeof line, "[eskf stamps] dev %d it %d:", ctx->device, k);
      for (int j = 1; j < kStampSlots; ++j)
        o += std::snprintf(line + o, sizeof line - static_cast<size_t>(o), "%s%.1f", j == 8 ? " | solve: " : " ",
                           (static_cast<double>(s8[kStampSlots * k + j]) - static_cast<double>(s8[kStampSlots * k])) * 1e-3);
      std::fprintf(stderr, "%s\n", line);
    }
  }
  if (st->error) {
    set_error(st->error == 2u ? "align kernel: a peer rank's H/b contribution did not arrive (NVLink mailbox timeout)"
                              : "align kernel: iteration hand-off timed out");
    return ESKF_ERR_INTERNAL;
  }
  const double* Tt = st->iter > 0 ? st->T_total : nullptr;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T_out[4 * i + j] = Tt ? Tt[3 * i + j] : a.guess[4 * i + j];
    T_out[4 * i + 3] = Tt ? Tt[9 + i] : a.guess[4 * i + 3];
  }
  T_out[12] = 0.0; T_out[13] = 0.0; T_out[14] = 0.0; T_out[15] = 1.0;
  if (info) {
    info->iterations = st->iter;
    info->converged = st->converged;
    info->n_corr_last = st->n_corr;
    const int it = st->iter < max_it ? st->iter : max_it;
    if (info->trace_H) std::memcpy(info->trace_H, h + L.o_H, static_cast<size_t>(it) * 36 * 8);
    if (info->trace_b) std::memcpy(info->trace_b, h + L.o_b, static_cast<size_t>(it) * 6 * 8);
    if (info->trace_ncorr) std::memcpy(info->trace_ncorr, h + L.o_nc, static_cast<size_t>(it) * 8);
    if (info->trace_step) {
      const double* s = reinterpret_cast<const double*>(h + L.o_step);
      for (int k = 0; k < it; ++k) {
        double* o = info->trace_step + 16 * k;
        for (int i = 0; i < 3; ++i) {
          for (int j = 0; j < 3; ++j) o[4 * i + j] = s[12 * k + 3 * i + j];
          o[4 * i + 3] = s[12 * k + 9 + i];
        }
        o[12] = 0.0; o[13] = 0.0; o[14] = 0.0; o[15] = 1.0;
      }
    }
  }
  return ESKF_OK;
}

// A registration in flight between align_begin and align_end (one per context)
struct PendingAlign {
  bool active = false;
  bool trivial = false;  // empty cloud: nothing was launched
  AlignArgs a;
  TraceLayout L;
  int max_it = 0;
  unsigned mail_seq = 0;
  // depth 10: the iterations are separate launches, queued a group at a time
  bool mk = false;
  AlignParams P;
  MkList list;
  int mk_next = 0;      // first iteration not queued yet
  int mk_lookup_grid = 0, mk_gather_grid = 0;
};

constexpr int kMkGroup = 8;  // iterations queued between two looks at the convergence word

// queue iterations [pd->mk_next, upto) of a depth-10 registration
int mk_enqueue(eskf_ctx* ctx, PendingAlign* pd, int upto) {
  for (int it = pd->mk_next; it < upto; ++it) {
    if (pd->mk_lookup_grid > 0) {
      mk_lookup_kernel<<<pd->mk_lookup_grid, kMkLookupT, 0, ctx->stream>>>(pd->P, pd->list, it);
      ESKF_CUDA(cudaGetLastError());
    }
    mk_gather_kernel<<<pd->mk_gather_grid, kMkGatherT, 0, ctx->stream>>>(pd->P, pd->list, it);
    ESKF_CUDA(cudaGetLastError());
    mk_solve_kernel<<<1, kMkGatherT, 0, ctx->stream>>>(pd->P, pd->list, it);
    ESKF_CUDA(cudaGetLastError());
    count_launch(ctx, pd->mk_lookup_grid > 0 ? 3 : 2);
  }
  pd->mk_next = upto;
  return ESKF_OK;
}
PendingAlign* pending(eskf_ctx* ctx) {
  if (!ctx->pending_align) ctx->pending_align = std::shared_ptr<void>(new PendingAlign(), [](void* p) {
    delete static_cast<PendingAlign*>(p);
  });
  return static_cast<PendingAlign*>(ctx->pending_align.get());
}

}  // namespace

// launch the persistent Gauss-Newton kernel and return; align_end collects the result
// keep the probed tag array (or filter) resident in L2 across iterations (the position /
// covariance streams would otherwise evict it every pass): a persisting access-policy window on the stream
int apply_l2_window(eskf_ctx* ctx, const AlignParams& P, const AlignArgs& a) {
  if (ctx->opt_l2_persist && ctx->l2_persist_bytes > 0) {
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof attr);
    // (depth 5 probes the 1 B/slot filter instead of the tags)
    const void* probed = P.filt ? static_cast<const void*>(P.filt) : static_cast<const void*>(a.map->tags);
    size_t bytes = static_cast<size_t>(a.map->n_slots) * (P.filt ? 1 : sizeof(tag_t));
    if (bytes > ctx->l2_window_max) bytes = ctx->l2_window_max;
    attr.accessPolicyWindow.base_ptr = const_cast<void*>(probed);
    attr.accessPolicyWindow.num_bytes = bytes;
    const double ratio = static_cast<double>(ctx->l2_persist_bytes) / static_cast<double>(bytes);
    attr.accessPolicyWindow.hitRatio = ratio > 1.0 ? 1.0f : static_cast<float>(ratio);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (ctx->l2_win_ptr != probed || ctx->l2_win_bytes != bytes) {  // the attribute sticks to the stream
      ESKF_CUDA(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
      ctx->l2_win_ptr = probed;
      ctx->l2_win_bytes = bytes;
    }
  } else if (ctx->l2_win_ptr != nullptr) {  // the option was switched off: drop the window the stream still carries
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof attr);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    ESKF_CUDA(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    ctx->l2_win_ptr = nullptr;
    ctx->l2_win_bytes = 0;
  }
  return ESKF_OK;
}

// The pool's B200s fall into two kinds on the dense pass (same clocks, same binary): on one the 4-deep
// 512-thread shape runs a 2 M-point iteration in ~86 us and the 3-stage 256-thread shape in ~88-98, on the
// other 103 and 90 (profiles/r2_align_experiments.md section 8).  Nothing the API reports tells them apart, so
// the first large-cloud registration of a context times both shapes on its own device, map and cloud (6 to
// 48 iterations per launch, results discarded, no exchange) and the context keeps the faster one; again when the
// cloud size changes by more than 2x.  Both shapes give the same correspondence sets and poses equal to
// rounding (tests).  Off: option "align_autotune" 0, or any explicit "align_block" / "align_depth".
int autotune_fat(eskf_ctx* ctx, const AlignArgs& a) {
  // 4-deep 512 on the filter; 3-stage 3 x 256 on the tags; 3-stage 1 x 768 on the tags; 3-stage 3 x 256 on the filter
  const int cand[4] = {ESKF_ALIGN_FAT_T, kT, 769, 257};
  AlignArgs ta = a;
  // (the first launch of a shape warms up; the fastest of the rest counts.  Enough iterations per launch
  // that the shapes' difference, a few per cent, stands clear of the launch overhead: ~1 ms of kernel)
  constexpr int kTuneReps = 4;
  const size_t per_it = a.cloud->n > 0 ? a.cloud->n : 1;
  const int kTuneIters = static_cast<int>(std::min<size_t>(48, std::max<size_t>(6, 12000000 / per_it)));
  ta.fixed_iterations = kTuneIters;
  ta.max_iteration = kTuneIters;
  ta.comm = nullptr;
  ta.d_hit = nullptr;
  cudaEvent_t e0, e1;
  ESKF_CUDA(cudaEventCreate(&e0));
  ESKF_CUDA(cudaEventCreate(&e1));
  float best = 0.f;
  int best_block = 0, rc = ESKF_OK;
  for (int c = 0; c < (ctx->opt_align_filter ? 4 : 3) && rc == ESKF_OK; ++c) {
    ctx->tuned_block = cand[c];
    float t_min = 0.f;
    for (int rep = 0; rep < kTuneReps && rc == ESKF_OK; ++rep) {
      AlignParams P;
      TraceLayout L;
      int G = 1;
      size_t dyn_smem = 0;
      rc = fill_params(ctx, ta, kTuneIters, &P, &L, &G, &dyn_smem);
      if (rc != ESKF_OK) break;
      rc = apply_l2_window(ctx, P, ta);
      if (rc != ESKF_OK) break;
      void* args[] = {&P};
      const Variant& var = g_variants[variant_index(ctx, ta)];
      cudaError_t e = cudaMemsetAsync(ctx->astate.p, 0, L.o_H, ctx->stream);
      if (e == cudaSuccess) e = cudaEventRecord(e0, ctx->stream);
      if (e == cudaSuccess) e = cudaLaunchCooperativeKernel(var.align, dim3(G), dim3(var.threads), args, dyn_smem, ctx->stream);
      if (e == cudaSuccess) e = cudaEventRecord(e1, ctx->stream);
      if (e == cudaSuccess) e = cudaEventSynchronize(e1);
      float ms = 0.f;
      if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
      if (e != cudaSuccess) {
        set_error("autotune: %s", cudaGetErrorString(e));
        rc = ESKF_ERR_CUDA;
        break;
      }
      count_launch(ctx);
      if (rep > 0 && (t_min == 0.f || ms < t_min)) t_min = ms;
    }
    if (rc == ESKF_OK && (best_block == 0 || t_min < best)) {
      best = t_min;
      best_block = cand[c];
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  ctx->tuned_block = rc == ESKF_OK ? best_block : 0;
  ctx->tuned_n = rc == ESKF_OK ? a.cloud->n : 0;
  if (ctx->opt_trace && rc == ESKF_OK)
    std::fprintf(stderr, "[eskf trace] align autotune: %d threads for %zu points\n", best_block, static_cast<size_t>(a.cloud->n));
  return rc;
}

int align_begin(eskf_ctx* ctx, const AlignArgs& a, const eskf_align_info* info) {
  ESKF_CUDA(cudaSetDevice(ctx->device));
  PendingAlign* pd = pending(ctx);
  ESKF_REQUIRE(!pd->active, "a registration is already in flight on this context (call eskf_align_end)");
  const int max_it = a.fixed_iterations > 0 ? a.fixed_iterations : a.max_iteration;
  const bool p2p = a.comm != nullptr && a.comm->world > 1;
  pd->a = a;
  pd->max_it = max_it;
  pd->mail_seq = 0u;
  pd->trivial = a.cloud && a.cloud->n == 0 && !p2p;
  if (pd->trivial) {
    pd->active = true;
    return ESKF_OK;
  }
  if (ctx->opt_align_autotune && ctx->opt_align_block == 0 && ctx->opt_align_depth == 0 && !a.fp64_math &&
      a.neighbor_mode == 1 && a.cloud && static_cast<int64_t>(a.cloud->n) >= ctx->opt_align_fat_points &&
      (ctx->tuned_block == 0 || a.cloud->n > 2 * ctx->tuned_n || 2 * a.cloud->n < ctx->tuned_n))
    ESKF_TRY(autotune_fat(ctx, a));
  AlignParams P;
  TraceLayout L;
  int G = 1;
  size_t dyn_smem = 0;
  ESKF_TRY(fill_params(ctx, a, max_it, &P, &L, &G, &dyn_smem));
  pd->L = L;
  if (p2p) {
    eskf_comm* c = a.comm;
    ESKF_REQUIRE(c->ctx == ctx, "communicator belongs to another context");
    ESKF_REQUIRE(c->connected, "communicator is not connected (eskf_comm_connect)");
    ESKF_REQUIRE(max_it < 65535, "too many iterations for the exchange flag encoding");
    P.world = c->world;
    P.rank = c->rank;
    P.seq = ++c->seq;
    for (int r = 0; r < c->world; ++r) P.peers[r] = c->peers[r];
  }
  const bool want_trace = info && (info->trace_H || info->trace_b || info->trace_ncorr || info->trace_step);
  if (ctx->opt_mapped_results && ctx->mail_h != nullptr && !want_trace && !ctx->opt_trace) {
    pd->mail_seq = ++ctx->align_seq;
    if (pd->mail_seq == 0u) pd->mail_seq = ++ctx->align_seq;
    P.mail = ctx->mail_d;
    P.mail_seq = pd->mail_seq;
  }
  ESKF_CUDA(cudaMemsetAsync(ctx->astate.p, 0, L.o_H, ctx->stream));
  ESKF_TRY(apply_l2_window(ctx, P, a));
  pd->mk = false;
  if (ctx->opt_align_depth == 10 && !a.fp64_math && a.neighbor_mode != 7 && a.cloud &&
      static_cast<int64_t>(a.cloud->n) >= ctx->opt_align_fat_points) {
    // depth 10: three launches per Gauss-Newton iteration (see mk_lookup_kernel)
    const unsigned n = static_cast<unsigned>(a.cloud->n);
    const unsigned n_tiles = (n + 31u) / 32u;
    const unsigned cap = (n + 63u) / 64u * 64u + 64u;
    ESKF_TRY(ctx->spill.ensure(static_cast<size_t>(10) * cap * sizeof(uint32_t)));
    ESKF_TRY(ctx->partials.ensure(static_cast<size_t>(ctx->sm_count) * kAcc * sizeof(double)));
    P.partials = ctx->partials.as<double>();
    ESKF_TRY(map_probe_filter(a.map, &P.filt));
    pd->mk = true;
    pd->P = P;
    pd->list.words = ctx->spill.as<uint32_t>();
    pd->list.count = &P.st->tile_counter;  // (zeroed with the state; reset by every solve)
    pd->list.cap = cap;
    const unsigned rounds = (n_tiles + kMkLookupT / 32 - 1u) / (kMkLookupT / 32);
    pd->mk_lookup_grid = static_cast<int>(rounds < static_cast<unsigned>(3 * ctx->sm_count) ? rounds : 3u * ctx->sm_count);
    pd->mk_gather_grid = ctx->sm_count;
    pd->mk_next = 0;
    trace_mark(ctx, "start");
    ESKF_TRY(mk_enqueue(ctx, pd, a.fixed_iterations > 0 ? max_it : (max_it < kMkGroup ? max_it : kMkGroup)));
    trace_mark(ctx, "align");
    pd->active = true;
    return ESKF_OK;
  }
  void* args[] = {&P};
  const Variant& var = g_variants[variant_index(ctx, a)];
  trace_mark(ctx, "start");
  ESKF_CUDA(cudaLaunchCooperativeKernel(var.align, dim3(G), dim3(var.threads), args, dyn_smem, ctx->stream));
  count_launch(ctx);
  trace_mark(ctx, "align");
  pd->active = true;
  return ESKF_OK;
}

int align_end(eskf_ctx* ctx, double T_out[16], eskf_align_info* info) {
  PendingAlign* pd = pending(ctx);
  ESKF_REQUIRE(pd->active, "no registration in flight (call eskf_align_cloud_begin first)");
  pd->active = false;
  if (pd->trivial) {
    // zero correspondences: zero step, "converged" after one iteration
    // (Eigen LDLT of a zero matrix solves to zero)
    for (int i = 0; i < 16; ++i) T_out[i] = pd->a.guess[i];
    if (info) {
      info->iterations = 1;
      info->converged = 1;
      info->n_corr_last = 0;
    }
    return ESKF_OK;
  }
  if (pd->mk) {
    // more iterations to queue?  (the kernels of iterations past convergence return at once)
    while (pd->mk_next < pd->max_it) {
      ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
      int* h = nullptr;
      ESKF_TRY(ctx_pinned(ctx, sizeof(AlignState), reinterpret_cast<void**>(&h)));
      ESKF_CUDA(cudaMemcpyAsync(h, ctx->astate.p, sizeof(AlignState), cudaMemcpyDeviceToHost, ctx->stream));
      ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
      const AlignState* hs = reinterpret_cast<const AlignState*>(h);
      if (hs->done || hs->error) break;
      const int upto = pd->mk_next + kMkGroup < pd->max_it ? pd->mk_next + kMkGroup : pd->max_it;
      ESKF_TRY(mk_enqueue(ctx, pd, upto));
    }
  }
  const int rc = read_back(ctx, pd->a, pd->L, pd->max_it, T_out, info, pd->mail_seq);
  trace_flush(ctx, "align");
  return rc;
}

int align_device(eskf_ctx* ctx, const AlignArgs& a, double T_out[16], eskf_align_info* info) {
  ESKF_TRY(align_begin(ctx, a, info));
  return align_end(ctx, T_out, info);
}

int align_max_blocks(int sm_count, int* out) {
  int best = 1;
  int dev = 0, optin = 0;
  ESKF_CUDA(cudaGetDevice(&dev));
  ESKF_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  for (int v = 0; v < V_COUNT; ++v) {
    int per_sm = 0;
    if (g_variants[v].depth5) {
      // one CTA per SM with all the opt-in shared memory its static allocations leave
      cudaFuncAttributes fa;
      ESKF_CUDA(cudaFuncGetAttributes(&fa, g_variants[v].align));
      const size_t room = static_cast<size_t>(optin) > fa.sharedSizeBytes + 1024 ? static_cast<size_t>(optin) - fa.sharedSizeBytes - 1024 : 0;
      ESKF_CUDA(cudaFuncSetAttribute(g_variants[v].align, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(room)));
      g_variants[v].max_dyn_smem = room;
      g_variants[v].per_sm = 1;
      continue;
    }
    ESKF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, g_variants[v].align,
                                                            g_variants[v].threads, 0));
    if (per_sm < 1) per_sm = 1;
    g_variants[v].per_sm = per_sm;
    if (per_sm > best) best = per_sm;
  }
  *out = best * sm_count;
  return ESKF_OK;
}

// sharded registration: per iteration
//   linearize_pass (local point range) -> caller's all-reduce of the 28 sums
//   -> solve (identical on every rank) -> host reads `done`
int align_sharded(eskf_ctx* ctx, const AlignArgs& a, eskf_allreduce_fn allreduce, void* user,
                  double T_out[16], eskf_align_info* info) {
  ESKF_CUDA(cudaSetDevice(ctx->device));
  const int max_it = a.fixed_iterations > 0 ? a.fixed_iterations : a.max_iteration;
  AlignParams P;
  TraceLayout L;
  int G = 1;
  ESKF_TRY(fill_params(ctx, a, max_it, &P, &L, &G));
  ESKF_CUDA(cudaMemsetAsync(ctx->astate.p, 0, L.o_H, ctx->stream));
  int* h_done = nullptr;
  ESKF_TRY(ctx_pinned(ctx, L.total + 64, reinterpret_cast<void**>(&h_done)));
  h_done = reinterpret_cast<int*>(reinterpret_cast<char*>(h_done) + L.total);
  for (int it = 0; it < max_it; ++it) {
    if (P.n > 0) {
      void* pargs[] = {&P, &it};
      const Variant& var = g_variants[variant_index(ctx, a)];
      ESKF_CUDA(cudaLaunchKernel(var.pass, dim3(G), dim3(var.threads), pargs, 0, ctx->stream));
      count_launch(ctx);
    } else {
      ESKF_CUDA(cudaMemsetAsync(P.sums, 0, kAcc * sizeof(double), ctx->stream));
    }
    if (allreduce) {
      const int rc = allreduce(user, P.sums, kAcc, ctx->stream);
      if (rc != 0) {
        set_error("allreduce callback failed with %d", rc);
        return ESKF_ERR_INTERNAL;
      }
    }
    solve_kernel<<<1, 32, 0, ctx->stream>>>(P, it);
    ESKF_CUDA(cudaGetLastError());
    count_launch(ctx);
    if (a.fixed_iterations <= 0) {
      ESKF_CUDA(cudaMemcpyAsync(h_done, &P.st->done, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      ESKF_CUDA(cudaStreamSynchronize(ctx->stream));
      if (*h_done) break;
    }
  }
  return read_back(ctx, a, L, max_it, T_out, info);
}

}  // namespace eskf
