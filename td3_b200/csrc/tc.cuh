// tc.cuh -- tcgen05 (5th-generation tensor core) tile for the dense contractions of the TD3 update.
//
// One tile = a 128 x NT block of C (NT in {16,32,64,128}) accumulated in tensor memory (TMEM) by
// tcgen05.mma.cta_group::1.kind::tf32 instructions of shape M=128, N=NT, K=8, issued by one thread.
// Operands are the fp32 activations / parameters / gradients exactly as they sit in HBM/L2: TF32 MMA
// reads the upper 19 bits of each fp32 word, so no conversion pass and no second copy of anything.
// They are staged chunk by chunk (32 reduction steps) into a ring of shared-memory slots by TMA
// (cp.async.bulk.tensor, one elected producer thread); the tensor maps' swizzle modes produce the canonical
// UMMA layouts directly:
//   reduction-contiguous operand ("rc", e.g. x[B,K] or W[N,K] in a forward layer)  -> K-major atoms
//   output-contiguous operand    ("oc", e.g. dZ[B,N] as the A of dW = dZ^T x)      -> MN-major atoms
// so forward, dX and dW GEMMs all run without a transpose.  A slot is refilled as soon as the
// tcgen05.commit of the MMAs that read it has arrived on the slot's mbarrier.  Operands the TMA unit cannot
// address (rows not 16-byte aligned: only the K = 23 / 17 first-layer weights) keep their problems on the fp32
// FFMA tile.
//
// Layout facts used below (cute/atom/mma_traits_sm100.hpp "make_umma_desc"):
//   K-major  SW128: row r of the tile at byte r*128, 16-byte chunk c stored at chunk (c ^ (r & 7));
//                   8-row groups 1024 B apart (SBO); an MMA K-step (8 tf32 = 32 B) advances the start
//                   address by 32 B inside the swizzled row.
//   MN-major, 32-bit elements: the only legal layout is SWIZZLE_128B_BASE32B ("TF32 transpose",
//                   Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> over 128 B x 4 reduction rows): reduction row k at
//                   byte k*128 holds 32 consecutive MN elements, its 32-byte granule q stored at (q ^ (k & 3));
//                   4-row K groups 512 B apart (SBO), 32-element MN groups LBO bytes apart; an MMA K-step
//                   (8 reduction rows) advances the start address by 1024 B.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "stage.cuh"

namespace td3 {

constexpr int kTcSub = 32768;                       // one 32-step chunk: A 128 x 32 fp32 (16 KB) + B up to 128 x 32 fp32 (16 KB)
constexpr int kTcGroup = 2;                         // chunks per pipeline stage: one barrier round trip per 8 MMAs
constexpr int kTcSlots = 3;                         // pipeline stages in flight
constexpr int kTcSlotBytes = kTcGroup * kTcSub;
constexpr int kTcRingBytes = kTcSlots * kTcSlotBytes;
constexpr int kTcMaxSlots = kTcSlots;
constexpr int kTcCols = 128;                        // TMEM columns allocated per CTA (fp32 accumulator columns)
constexpr int kTcAllocWarp = 2;                     // allocates / frees TMEM (warp 0 produces, warp 1 issues MMAs)

#ifdef TD3_TILE_PROF
#define TCP(k) do { tp_[k] = clock64(); } while (0)
#else
#define TCP(k) do { } while (0)
#endif

struct TcState {                                    // lives in shared memory, one per CTA
  unsigned long long full_bar[kTcMaxSlots];         // "this slot's operand bytes have landed" (TMA complete_tx)
  unsigned long long empty_bar[kTcMaxSlots];        // "the MMAs that read this slot have completed" (tcgen05.commit)
  unsigned long long done_bar;                      // "the tile's accumulator is complete"
  unsigned long long tmem_bar;                      // "tmem_base is valid" (the allocation runs beside the first TMA loads)
  unsigned int tmem_base;
  unsigned int tma_chunk_count;                     // chunks staged so far (full/empty barrier phases)
  unsigned int tile_count;                          // TC tiles finished so far (done_bar phase)
  int prof_stage;                                   // >= 0: record clock64 stamps of this CTA's tiles (debug builds)
  int group;                                        // chunks per pipeline stage of this launch: kTcGroup, or 1 (small ring: two CTAs per SM)
};

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
  const unsigned int a = smem_u32(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a), "r"(parity)
      : "memory");
}

__device__ __forceinline__ unsigned int cluster_ctarank() {
  unsigned int r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

__device__ __forceinline__ bool elect_one() {
  unsigned int pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_wait_u32(unsigned int bar, unsigned int parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_u32(unsigned int dst, const void* tmap, int c0, int c1, unsigned int bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
               "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc_u32(unsigned int dst, const void* tmap, int c0, int c1, unsigned int bar,
                                                   unsigned short mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;\n" ::"r"(dst),
      "l"(tmap), "r"(c0), "r"(c1), "r"(bar), "h"(mask)
      : "memory");
}

// called by all threads of the CTA once, before the first TC tile.  cluster = CTAs that share (multicast) the A panel:
// a slot is free again only when the MMAs of ALL of them have finished reading it.
__device__ __forceinline__ void tc_setup(TcState* st, int cluster = 1, int group = kTcGroup) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < kTcMaxSlots; ++i) {
      mbar_init(&st->full_bar[i], 1);
      mbar_init(&st->empty_bar[i], (unsigned)cluster);
    }
    mbar_init(&st->done_bar, 1);
    mbar_init(&st->tmem_bar, 1);
    st->tma_chunk_count = 0;
    st->prof_stage = -1;
    st->tile_count = 0;
    st->group = group;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();                                  // barriers exist: the producer warp may start loading operands
  // The TMEM allocation (~270 cycles, longer under contention) is only needed by the first MMA, a full L2 round trip
  // later: warp kTcAllocWarp performs it and publishes the base address through tmem_bar instead of a CTA-wide sync.
  if ((threadIdx.x >> 5) == kTcAllocWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&st->tmem_base)),
                 "r"(kTcCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncwarp();
    if ((threadIdx.x & 31) == 0)
      asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&st->tmem_bar)) : "memory");
  }
}

// every reader of tmem_base passes through here first (a completed phase: later calls return at once)
__device__ __forceinline__ unsigned int tc_tmem_base(TcState* st) {
  mbar_wait(&st->tmem_bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  return *reinterpret_cast<volatile unsigned int*>(&st->tmem_base);
}

__device__ __forceinline__ void tc_teardown(TcState* st) {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if ((threadIdx.x >> 5) == kTcAllocWarp) {
    const unsigned int base = tc_tmem_base(st);
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base), "r"(kTcCols) : "memory");
  }
}

// shared-memory matrix descriptor, 128-byte swizzle, version 1 (sm_100)
__device__ __forceinline__ unsigned long long tc_desc(unsigned int smem_addr, unsigned int lbo_bytes, unsigned int sbo_bytes,
                                                      unsigned int layout_type) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr >> 4) & 0x3FFF);
  d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;                                   // descriptor version (sm_100)
  d |= (unsigned long long)layout_type << 61;        // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}

// instruction descriptor: D fp32, A/B tf32, M = 128, N = nt, major bits per operand (0 = K-major, 1 = MN-major)
__device__ __forceinline__ unsigned int tc_idesc(int nt, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)a_mn_major << 15) | ((unsigned)b_mn_major << 16) |
         ((unsigned)(nt >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void tc_mma(unsigned int tmem_d, unsigned long long da, unsigned long long db, unsigned int idesc,
                                       unsigned int accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMA (cp.async.bulk.tensor) staging: one elected thread moves whole boxes, the hardware applies the swizzle ----
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned int dst, const void* tmap, int c0, int c1, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
               "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(unsigned int dst, const void* tmap, int c0, int c1, unsigned long long* bar,
                                               unsigned short mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;\n" ::"r"(dst),
      "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(unsigned long long* bar, unsigned short mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// The A chunk (128 rows x 32 k) of a cluster of `c` CTAs that share the same M tile: CTA `rank` fetches 1/c of it and
// multicasts it into the same slot of every CTA of the cluster (each CTA's full barrier sees all 16 KB arrive).
__device__ __forceinline__ void tma_issue_a_multicast(unsigned int dst, const void* tmap, int rc, int i0, int k0, int c, int rank,
                                                      unsigned long long* bar) {
  const unsigned short mask = (unsigned short)((1u << c) - 1u);
  if (rc) {
    const int rows = 128 / c;
    tma_load_2d_mc(dst + rank * rows * 128, tmap, k0, i0 + rank * rows, bar, mask);
  } else {
    for (int g = 0; g < 4; ++g)
      if (g % c == rank) tma_load_2d_mc(dst + g * 4096, tmap, i0 + g * 32, k0, bar, mask);
  }
}

// one operand chunk: rc -> one box [rows_mn x 32 k]; oc -> one box [32 k x 32 mn] per 32-wide MN group
__device__ __forceinline__ void tma_issue_operand(unsigned int dst, const void* tmap, int rc, int rows_mn, int o0, int k0,
                                                  unsigned long long* bar) {
  if (rc) {
    tma_load_2d(dst, tmap, k0, o0, bar);
  } else {
    const int groups = (rows_mn + 31) >> 5;
    for (int g = 0; g < groups; ++g) tma_load_2d(dst + g * 4096, tmap, o0 + g * 32, k0, bar);
  }
}

__device__ __forceinline__ float4 tc_load_aux(const Problem& P, const float* aux0, bool aux_read, bool x_vec, bool row_ok, int i,
                                              int j) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (aux_read && row_ok) {
    const float* ap = aux0 + (size_t)i * P.ldaux + j;
    if (x_vec && j + 3 < P.N) {
      v = *reinterpret_cast<const float4*>(ap);
    } else {
      if (j < P.N) v.x = ap[0];
      if (j + 1 < P.N) v.y = ap[1];
      if (j + 2 < P.N) v.z = ap[2];
      if (j + 3 < P.N) v.w = ap[3];
    }
  }
  return v;
}

// One group of W (16 or 8) accumulator columns of this thread's row: operands of the epilogue loaded up front (they are in
// flight while the tile's last MMAs finish: the caller waits for the accumulator between the two halves), one
// tcgen05.ld for all W columns, straight-line per-kind math, 16-byte stores.
template <int W>
__device__ __forceinline__ void tc_load_aux_group(const Problem& P, const float* aux0, bool aux_read, bool x_vec, bool row_ok, int i,
                                                  int j, float (&ax)[W]) {
#pragma unroll
  for (int q = 0; q < W / 4; ++q) {
    const float4 v = tc_load_aux(P, aux0, aux_read, x_vec, row_ok, i, j + 4 * q);
    ax[4 * q] = v.x; ax[4 * q + 1] = v.y; ax[4 * q + 2] = v.z; ax[4 * q + 3] = v.w;
  }
}

template <int EPI, int W>
__device__ __forceinline__ void tc_epilogue_group(const Problem& P, float* __restrict__ C, float* aux0, const float* bias_s,
                                                  bool has_bias, bool c_vec, bool row_ok, bool have_acc, int i, int j, int jl,
                                                  unsigned int taddr, const float (&ax)[W]) {
  unsigned int r[W];
  if (W == 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8 % W]), "=r"(r[9 % W]),
          "=r"(r[10 % W]), "=r"(r[11 % W]), "=r"(r[12 % W]), "=r"(r[13 % W]), "=r"(r[14 % W]), "=r"(r[15 % W])
        : "r"(taddr)
        : "memory");
  } else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  if (!row_ok) return;
#pragma unroll
  for (int q = 0; q < W / 4; ++q) {
    const int jq = j + 4 * q;
    if (jq >= P.N) break;
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float aux_out = 0.f;
      const float v = have_acc ? __uint_as_float(r[4 * q + e]) : 0.f;
      o[e] = epi_apply<EPI>(v, has_bias ? bias_s[jl + 4 * q + e] : 0.f, ax[4 * q + e], P.f0, P.f1, aux_out);
      if (EPI == EPI_BIAS_TANH && jq + e < P.N) aux0[(size_t)i * P.ldaux + jq + e] = aux_out;
      if (P.rn_out) o[e] = rn_tf32(o[e]);
    }
#pragma unroll 1
    for (int d = 0; d < P.c_dups; ++d) {
      float* cp = C + d * P.c_dup_stride + (size_t)i * P.ldc + jq;
      if (c_vec && jq + 3 < P.N) {
        *reinterpret_cast<float4*>(cp) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
        cp[0] = o[0];
        if (jq + 1 < P.N) cp[1] = o[1];
        if (jq + 2 < P.N) cp[2] = o[2];
        if (jq + 3 < P.N) cp[3] = o[3];
      }
    }
  }
}

// ---- coalesced form of a 16-column group -------------------------------------------------------------------------
// tcgen05.ld hands lane l the 16 columns of accumulator ROW l, so a store of 16 bytes per lane touches 32 different
// rows: 32 half-used sectors per instruction, and on the many-tile stages (populations, batch-8192) the L1 / L2 request
// rate of those accesses -- not the tensor pipe -- set the tile time.  The group is therefore transposed through 2 KB of
// shared memory per warp (16-byte chunks XOR-swizzled by row pair: conflict-free both ways): afterwards lane l owns
// columns 4 (l & 3) .. + 3 of rows 8 p + (l >> 2), p = 0..3, i.e. four lanes cover 64 contiguous bytes of a row and an
// instruction touches 8 rows with every sector fully used.  The auxiliary operand (ReLU mask / tanh / noise) is read in
// the same layout.  The arithmetic per element is unchanged.
__device__ __forceinline__ void tc_load_aux_group_co(const Problem& P, const float* aux0, bool aux_read, bool x_vec, int ib, int j,
                                                     int lane, float (&ax)[16]) {
  const int c = lane & 3, rr = lane >> 2;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int i = ib + 8 * p + rr;
    const float4 v = tc_load_aux(P, aux0, aux_read, x_vec, i < P.M, i, j + 4 * c);
    ax[4 * p] = v.x; ax[4 * p + 1] = v.y; ax[4 * p + 2] = v.z; ax[4 * p + 3] = v.w;
  }
}

template <int EPI>
__device__ __forceinline__ void tc_epilogue_group_co(const Problem& P, float* __restrict__ C, float* aux0, const float* bias_s,
                                                     bool has_bias, bool c_vec, bool have_acc, int ib, int j, int jl,
                                                     unsigned int taddr, const float (&ax)[16], float* stg, int lane) {
  unsigned int r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  {
    uint4* row = reinterpret_cast<uint4*>(stg + lane * 16);
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int q = 0; q < 4; ++q) row[q ^ sw] = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
  }
  __syncwarp();
  const int c = lane & 3, rr = lane >> 2;
  uint4 t[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int rl = 8 * p + rr;
    t[p] = reinterpret_cast<const uint4*>(stg + rl * 16)[c ^ ((rl >> 1) & 3)];
  }
  __syncwarp();                                      // the next group overwrites the staging rows
  const int jq = j + 4 * c;
  if (jq >= P.N) return;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int i = ib + 8 * p + rr;
    if (i >= P.M) continue;
    const unsigned int rv[4] = {t[p].x, t[p].y, t[p].z, t[p].w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float aux_out = 0.f;
      const float v = have_acc ? __uint_as_float(rv[e]) : 0.f;
      o[e] = epi_apply<EPI>(v, has_bias ? bias_s[jl + 4 * c + e] : 0.f, ax[4 * p + e], P.f0, P.f1, aux_out);
      if (EPI == EPI_BIAS_TANH && jq + e < P.N) aux0[(size_t)i * P.ldaux + jq + e] = aux_out;
      if (P.rn_out) o[e] = rn_tf32(o[e]);
    }
#pragma unroll 1
    for (int d = 0; d < P.c_dups; ++d) {
      float* cp = C + d * P.c_dup_stride + (size_t)i * P.ldc + jq;
      if (c_vec && jq + 3 < P.N) {
        *reinterpret_cast<float4*>(cp) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
        cp[0] = o[0];
        if (jq + 1 < P.N) cp[1] = o[1];
        if (jq + 2 < P.N) cp[2] = o[2];
        if (jq + 3 < P.N) cp[3] = o[3];
      }
    }
  }
}

// Epilogue of one warp: its 32 accumulator rows (TMEM lanes) x `half` columns starting at local column jw.
template <int EPI, bool kCo>
__device__ __forceinline__ void tc_epilogue_cols(const Problem& P, float* __restrict__ C, float* aux0, const float* bias_s,
                                                 bool has_bias, bool aux_read, bool x_vec, bool c_vec, bool row_ok, bool have_acc,
                                                 int i, int j0, int jw, int half, unsigned int tmem_lane, unsigned int done_bar,
                                                 unsigned int done_parity, float* stg) {
  if (half >= 16 && !kCo) {
    // latency form (a launch is one wave of tiles on the update's dependency chain): row-per-lane accesses, no staging --
    // the transposition below costs two shared-memory round trips per group, ~0.5 us per update at cfg2
    float ax[16];
    tc_load_aux_group<16>(P, aux0, aux_read, x_vec, row_ok, i, j0 + jw, ax);
    if (have_acc) mbar_wait_u32(done_bar, done_parity);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll 1
    for (int jc = 0; jc < half; jc += 16) {
      float ax_next[16];
      const bool more = jc + 16 < half;
      if (more) tc_load_aux_group<16>(P, aux0, aux_read, x_vec, row_ok, i, j0 + jw + jc + 16, ax_next);
      tc_epilogue_group<EPI, 16>(P, C, aux0, bias_s, has_bias, c_vec, row_ok, have_acc, i, j0 + jw + jc, jw + jc,
                                 tmem_lane + (unsigned)(jw + jc), ax);
      if (more) {
#pragma unroll
        for (int q = 0; q < 16; ++q) ax[q] = ax_next[q];
      }
    }
  } else if (kCo && half >= 16) {
    // throughput form (many tiles per SM): coalesced accesses through the per-warp staging
    const int lane = threadIdx.x & 31, ib = i - lane;    // first accumulator row of this warp
    float ax[16];
    tc_load_aux_group_co(P, aux0, aux_read, x_vec, ib, j0 + jw, lane, ax);
    if (have_acc) mbar_wait_u32(done_bar, done_parity);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // the auxiliary operand (ReLU mask / tanh / noise) of group jc + 1 is requested before group jc is processed: on the
    // many-tile stages it comes from HBM, and one DRAM round trip per 16 columns was most of a tile's time
#pragma unroll 1
    for (int jc = 0; jc < half; jc += 16) {
      float ax_next[16];
      const bool more = jc + 16 < half;
      if (more) tc_load_aux_group_co(P, aux0, aux_read, x_vec, ib, j0 + jw + jc + 16, lane, ax_next);
      tc_epilogue_group_co<EPI>(P, C, aux0, bias_s, has_bias, c_vec, have_acc, ib, j0 + jw + jc, jw + jc,
                                tmem_lane + (unsigned)(jw + jc), ax, stg, lane);
      if (more) {
#pragma unroll
        for (int q = 0; q < 16; ++q) ax[q] = ax_next[q];
      }
    }
  } else {   // NT = 16: eight columns per warp half
    float ax[8];
    tc_load_aux_group<8>(P, aux0, aux_read, x_vec, row_ok, i, j0 + jw, ax);
    if (have_acc) mbar_wait_u32(done_bar, done_parity);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    tc_epilogue_group<EPI, 8>(P, C, aux0, bias_s, has_bias, c_vec, row_ok, have_acc, i, j0 + jw, jw, tmem_lane + (unsigned)jw, ax);
  }
}

// ------------------------------------------------------------------------------------
// One 128 x NT tile.  `ring` is kTcRingBytes of 1024-byte-aligned shared memory.
// ------------------------------------------------------------------------------------
// kCo: coalesced (throughput) epilogue -- a template parameter, not a run-time switch: merely compiling the second epilogue
// into the latency kernel cost the single-agent update 2 us (profiles/r02g_code_size_ab.txt: it runs every launch cold)
template <bool kCo = false>
__device__ __forceinline__ void gemm_tile_tc(const Problem& P, int tile, unsigned char* ring, TcState* st,
                                             const TensorMapBlob* param_maps = nullptr) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NT = P.tc_nt;
#ifdef TD3_TILE_PROF
  long long tp_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
  TCP(0);

  int t = tile;
  const int g = t / P.tiles_per_group;
  t -= g * P.tiles_per_group;
  const int go = g / P.groups_inner, gi = g - go * P.groups_inner;
  int ks = 0;
  if (P.ksplit > 1) {
    const int per = P.tiles_m * P.tiles_n;
    ks = t / per;
    t -= ks * per;
  }
  const int tm = t / P.tiles_n, tn = t - tm * P.tiles_n;
  const int i0 = tm * 128, j0 = tn * NT;
  const int arc = P.a_rc, brc = P.b_rc;

  int k_begin = 0, k_end = P.K;
  if (P.ksplit > 1) {
    const int chunks = (P.K + 31) / 32;
    const int per = (chunks + P.ksplit - 1) / P.ksplit;
    k_begin = min(P.K, ks * per * 32);
    k_end = min(P.K, (ks + 1) * per * 32);
  }
  const int n_chunks = (k_end - k_begin + 31) / 32;

  const unsigned int idesc = tc_idesc(NT, !arc, !brc);
  // per-operand descriptor geometry
  const unsigned int a_lbo = arc ? 16 : 4096, b_lbo = brc ? 16 : 4096;
  const unsigned int a_sbo = arc ? 1024 : 512, b_sbo = brc ? 1024 : 512;
  const unsigned int a_lt = arc ? 2 : 1, b_lt = brc ? 2 : 1;
  const unsigned int a_kstep = arc ? 32 : 1024, b_kstep = brc ? 32 : 1024;

  const unsigned int tile_no = st->tile_count;
  const int grp = st->group;                                      // chunks per ring slot (barrier round trip)
  const int ring_bytes = kTcSlots * grp * kTcSub;
  // bias strip of this tile -> shared memory (read by every row's epilogue; a global load there would sit on the
  // critical path once per 4-column iteration)
  float* bias_s = reinterpret_cast<float*>(ring + ring_bytes);
  if (P.bias && tid >= 64 && tid < 64 + NT) {
    const float* bias = P.bias + go * P.bias_go + gi * P.bias_gi;
    const int j = j0 + tid - 64;
    bias_s[tid - 64] = j < P.N ? bias[j] : 0.f;
  }
  {
    // warp-specialised: one producer thread (TMA), one MMA-issuing thread; everybody else waits for the accumulator
    const unsigned int tchunk = st->tma_chunk_count;
    const bool pm = param_maps != nullptr && P.map_a >= 0 && P.map_b >= 0;
    const unsigned char* mapA = pm ? reinterpret_cast<const unsigned char*>(param_maps + P.map_a + g)
                                   : reinterpret_cast<const unsigned char*>(P.tmapA) + (size_t)g * 128;
    const unsigned char* mapB = pm ? reinterpret_cast<const unsigned char*>(param_maps + P.map_b + g)
                                   : reinterpret_cast<const unsigned char*>(P.tmapB) + (size_t)g * 128;
    const unsigned int bytes = 16384u + (arc ? 0u : 0u) + (brc ? (unsigned)NT * 128u : (unsigned)((NT + 31) >> 5) * 4096u);
    TCP(1);
    const unsigned int n_slots = kTcSlots, slot_bytes = (unsigned)(grp * kTcSub);
    const int n_stages = (n_chunks + grp - 1) / grp;
    if (tid == 0) asm volatile("prefetch.tensormap [%0];\n" ::"l"(mapA) : "memory");
    if (tid == 32) asm volatile("prefetch.tensormap [%0];\n" ::"l"(mapB) : "memory");
    const int csz = P.tc_cluster > 1 ? P.tc_cluster : 1;
    const int crank = csz > 1 ? (int)cluster_ctarank() : 0;
    // Producer (warp 0) and MMA issuer (warp 1).  The whole warp walks the loop with warp-uniform values (made
    // provably uniform with a shuffle) and one elected lane issues the TMA / tcgen05 instructions: UTMALDG / UTCHMMA
    // take their operands from UNIFORM registers, and with thread-divergent operands the compiler wraps every single
    // one in an ELECT + 4x R2UR.BROADCAST + branch "waterfall" loop (measured: 640-707 cycles per 4-MMA chunk,
    // independent of tile width, ring depth, operand orientation and cluster size).
#define TD3_UNI(x) __shfl_sync(0xffffffffu, (x), 0)
    if (warp == 0) {
      const unsigned int ring_u = TD3_UNI(smem_u32(ring)), fb0 = TD3_UNI(smem_u32(&st->full_bar[0])),
                         eb0 = TD3_UNI(smem_u32(&st->empty_bar[0]));
      const unsigned int nsl = TD3_UNI(n_slots), sbytes = TD3_UNI(slot_bytes), ubytes = TD3_UNI(bytes);
      const unsigned int tch = TD3_UNI(tchunk);
      const int nch = TD3_UNI(n_chunks), ui0 = TD3_UNI(i0), uj0 = TD3_UNI(j0), unt = TD3_UNI(NT);
      const int uarc = TD3_UNI(arc), ubrc = TD3_UNI(brc), ucsz = TD3_UNI(csz), ucrank = TD3_UNI(crank);
      const unsigned long long ma = ((unsigned long long)TD3_UNI((unsigned int)((unsigned long long)mapA >> 32)) << 32) |
                                    TD3_UNI((unsigned int)(unsigned long long)mapA);
      const unsigned long long mb = ((unsigned long long)TD3_UNI((unsigned int)((unsigned long long)mapB >> 32)) << 32) |
                                    TD3_UNI((unsigned int)(unsigned long long)mapB);
      int kc = TD3_UNI(k_begin);
      unsigned int slot = tch % nsl, use = tch / nsl;
      unsigned int sa0 = ring_u + slot * sbytes, fb = fb0 + slot * 8, eb = eb0 + slot * 8;
      const unsigned short mask = (unsigned short)((1u << ucsz) - 1u);
      const int a_rows = 128 / ucsz;
      const int nst = TD3_UNI(n_stages), ugrp = TD3_UNI(grp);
#pragma unroll 1
      for (int c = 0; c < nst; ++c) {
        if (use > 0) mbar_wait_u32(eb, (use - 1) & 1);                     // every CTA of the cluster is done with the slot
        const int nsub = min(ugrp, nch - c * ugrp);
        if (elect_one()) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(fb), "r"(ubytes * nsub) : "memory");
#pragma unroll 1
          for (int u = 0; u < nsub; ++u) {
          const unsigned int sa = sa0 + u * kTcSub;
          const int k0 = kc + u * 32;
          if (ucsz > 1) {
            if (uarc) {
              tma_load_2d_mc_u32(sa + ucrank * a_rows * 128, (const void*)ma, k0, ui0 + ucrank * a_rows, fb, mask);
            } else {
              for (int g4 = ucrank; g4 < 4; g4 += ucsz) tma_load_2d_mc_u32(sa + g4 * 4096, (const void*)ma, ui0 + g4 * 32, k0, fb, mask);
            }
          } else if (uarc) {
            tma_load_2d_u32(sa, (const void*)ma, k0, ui0, fb);
          } else {
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) tma_load_2d_u32(sa + g4 * 4096, (const void*)ma, ui0 + g4 * 32, k0, fb);
          }
          if (ubrc) {
            tma_load_2d_u32(sa + 16384, (const void*)mb, k0, uj0, fb);
          } else {
            for (int g4 = 0; g4 * 32 < unt; ++g4) tma_load_2d_u32(sa + 16384 + g4 * 4096, (const void*)mb, uj0 + g4 * 32, k0, fb);
          }
          }
        }
        __syncwarp();
        kc += ugrp * 32; sa0 += sbytes; fb += 8; eb += 8;
        if (++slot == nsl) { slot = 0; ++use; sa0 = ring_u; fb = fb0; eb = eb0; }
      }
      TCP(2);
    } else if (warp == 1) {
      const unsigned int tmem = tc_tmem_base(st);
      const unsigned int ring_u = TD3_UNI(smem_u32(ring)), fb0 = TD3_UNI(smem_u32(&st->full_bar[0])),
                         eb0 = TD3_UNI(smem_u32(&st->empty_bar[0])), done_u = TD3_UNI(smem_u32(&st->done_bar));
      const unsigned int nsl = TD3_UNI(n_slots), sbytes = TD3_UNI(slot_bytes), tch = TD3_UNI(tchunk);
      const unsigned int utmem = TD3_UNI(tmem), uidesc = TD3_UNI(idesc);
      const int nch = TD3_UNI(n_chunks), ucsz = TD3_UNI(csz);
      // descriptor words: hi = SBO | version | layout type (constant per operand); lo = start address | LBO
      const unsigned int a_hi = TD3_UNI((a_sbo >> 4) | (1u << 14) | (a_lt << 29)), b_hi = TD3_UNI((b_sbo >> 4) | (1u << 14) | (b_lt << 29));
      const unsigned int a_lo0 = TD3_UNI((a_lbo >> 4) << 16), b_lo0 = TD3_UNI((b_lbo >> 4) << 16);
      const unsigned int a_ks = TD3_UNI(a_kstep >> 4), b_ks = TD3_UNI(b_kstep >> 4);
      unsigned int slot = tch % nsl, use = tch / nsl;
      unsigned int sa = ring_u + slot * sbytes, fb = fb0 + slot * 8, eb = eb0 + slot * 8;
      const unsigned short mask = (unsigned short)((1u << ucsz) - 1u);
      const int nst = TD3_UNI(n_stages), ugrp = TD3_UNI(grp);
#pragma unroll 1
      for (int c = 0; c < nst; ++c) {
        mbar_wait_u32(fb, use & 1);
        if (c == 0) TCP(1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const int nsub = min(ugrp, nch - c * ugrp);
        if (elect_one()) {
          const unsigned int a_lo = a_lo0 | (sa >> 4), b_lo = b_lo0 | ((sa + 16384) >> 4);
          tc_mma(utmem, ((unsigned long long)a_hi << 32) | a_lo, ((unsigned long long)b_hi << 32) | b_lo, uidesc, c > 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 1; kk < 4; ++kk)
            tc_mma(utmem, ((unsigned long long)a_hi << 32) | (a_lo + kk * a_ks), ((unsigned long long)b_hi << 32) | (b_lo + kk * b_ks),
                   uidesc, 1u);
          if (nsub > 1) {
            const unsigned int a2 = a_lo + (kTcSub >> 4), b2 = b_lo + (kTcSub >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              tc_mma(utmem, ((unsigned long long)a_hi << 32) | (a2 + kk * a_ks), ((unsigned long long)b_hi << 32) | (b2 + kk * b_ks),
                     uidesc, 1u);
          }
          if (ucsz > 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(eb),
                         "h"(mask)
                         : "memory");
          else
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(eb) : "memory");
          if (c == nst - 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(done_u) : "memory");
        }
        __syncwarp();
        sa += sbytes; fb += 8; eb += 8;
        if (++slot == nsl) { slot = 0; ++use; sa = ring_u; fb = fb0; eb = eb0; }
      }
      TCP(2);
    }
#undef TD3_UNI
    __syncwarp();
  }

  // ---- epilogue: TMEM -> registers -> (bias / activation) -> global ----
  float* __restrict__ C = P.C + go * P.c_go + gi * P.c_gi + (long long)ks * P.c_split;
  const bool has_bias = P.bias != nullptr;
  float* aux0 = P.aux0 ? P.aux0 + go * P.aux0_go + gi * P.aux0_gi : nullptr;
  const int epi = P.epi;
  const bool aux_read = aux0 && (epi == EPI_BIAS_TANH_NOISE || epi == EPI_RELU_MASK || epi == EPI_TANH_GRAD);
  const int i = i0 + (warp & 3) * 32 + lane;         // TMEM lane == accumulator row
  const bool row_ok = i < P.M;
  const bool c_vec = P.c_vec, x_vec = P.aux_vec;
  // warps 0-3 own the lower half of the tile's columns, warps 4-7 the upper half; 4 columns per (rolled) iteration
  const int half = NT >> 1, jw = (warp >> 2) * half;
  __syncthreads();                                         // bias strip visible (the MMA pipeline is busy meanwhile)
  const unsigned int tmem = tc_tmem_base(st);
  TCP(3);
  // transposition staging of the coalesced epilogue: 2 KB per warp at the start of the ring (the staging is touched only
  // after done_bar has completed, i.e. after the tensor core has consumed every operand byte of this tile)
  float* stg = kCo ? reinterpret_cast<float*>(ring) + warp * 512 : nullptr;
  TD3_DISPATCH_EPI(epi, (tc_epilogue_cols<E, kCo>(P, C, aux0, bias_s, has_bias, aux_read, x_vec, c_vec, row_ok, n_chunks > 0, i, j0, jw,
                                             half, tmem + (((unsigned)(warp & 3) * 32u) << 16), smem_u32(&st->done_bar), tile_no & 1,
                                             stg)));
  if (kCo) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy staging accesses before the next tile's TMA writes
  TCP(4);
  TCP(5);
  // every warp is past its TMEM reads before the next tile's first MMA overwrites the accumulator
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  TCP(6);
#ifdef TD3_TILE_PROF
  if (blockIdx.x == 0 && st->prof_stage >= 0 && (tid == 0 || tid == 32)) {
    long long* o = g_tp + st->prof_stage * 16 + (tid == 32 ? 8 : 0);
    for (int k = 0; k < 8; ++k) o[k] = tp_[k];
  }
#endif
  if (tid == 0) {
    st->tma_chunk_count += (unsigned)max((n_chunks + grp - 1) / grp, 0);
    if (n_chunks > 0) st->tile_count = tile_no + 1;
  }
  __syncthreads();
}

}  // namespace td3
