// tc.cuh -- tcgen05 (5th-generation tensor core) tile for the dense contractions of the TD3 update.
//
// One tile = a 128 x NT block of C (NT in {16,32,64,128}) accumulated in tensor memory (TMEM) by
// tcgen05.mma.cta_group::1.kind::tf32 instructions of shape M=128, N=NT, K=8, issued by one thread.
// Operands are the fp32 activations / parameters / gradients exactly as they sit in HBM/L2: TF32 MMA
// reads the upper 19 bits of each fp32 word, so no conversion pass and no second copy of anything.
// They are staged chunk by chunk (32 reduction steps) into a ring of shared-memory slots with 16-byte
// cp.async copies that write the canonical 128-byte-swizzled UMMA layouts directly:
//   reduction-contiguous operand ("rc", e.g. x[B,K] or W[N,K] in a forward layer)  -> K-major atoms
//   output-contiguous operand    ("oc", e.g. dZ[B,N] as the A of dW = dZ^T x)      -> MN-major atoms
// so forward, dX and dW GEMMs all run without a transpose.  A slot is refilled as soon as the
// tcgen05.commit of the MMAs that read it has arrived on the slot's mbarrier.
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

constexpr int kTcSlots = 6;
constexpr int kTcSlotBytes = 32768;                 // A: 128 x 32 fp32 (16 KB) + B: up to 128 x 32 fp32 (16 KB)
constexpr int kTcRingBytes = kTcSlots * kTcSlotBytes;
constexpr int kTcCols = 128;                        // TMEM columns allocated per CTA (fp32 accumulator columns)

#ifdef TD3_TC_DEBUG
__device__ int g_tc_dbg[8] = {1, 512, 4096, 1024, 0, 0, 0, 0};   // MN-major: layout type, SBO, LBO, K-step, store swizzle, swap majors
#endif

struct TcState {                                    // lives in shared memory, one per CTA
  unsigned long long slot_bar[kTcSlots];            // "MMAs that read this slot have completed"
  unsigned long long done_bar;                      // "the tile's accumulator is complete"
  unsigned int tmem_base;
  unsigned int chunk_count;                         // chunks issued by this CTA so far (slot phase bookkeeping)
  unsigned int tile_count;                          // TC tiles finished so far (done_bar phase)
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

// called by all threads of the CTA once, before the first TC tile
__device__ __forceinline__ void tc_setup(TcState* st) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < kTcSlots; ++i) mbar_init(&st->slot_bar[i], 1);
    mbar_init(&st->done_bar, 1);
    st->chunk_count = 0;
    st->tile_count = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&st->tmem_base)),
                 "r"(kTcCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}

__device__ __forceinline__ void tc_teardown(TcState* st) {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if ((threadIdx.x >> 5) == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(st->tmem_base), "r"(kTcCols) : "memory");
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

// global -> shared copy of one operand chunk (rows_mn output rows x 32 reduction steps) in UMMA layout.
// rc (K-major): global row = output index, 8 chunks of 16 B along the reduction.
// oc (MN-major): global row = reduction index, rows_mn/4 chunks of 16 B along the output dimension.
__device__ __forceinline__ void tc_issue_operand(unsigned char* dst, const float* __restrict__ base, int ld, int rc,
                                                 int rows_mn, int o0, int O, int k0, int K, int tid) {
  const unsigned int d0 = smem_u32(dst);
  if (rc) {
    const int n16 = rows_mn * 8;
    for (int q = tid; q < n16; q += kStageThreads) {
      const int r = q >> 3, c = q & 7;
      const int oi = o0 + r, ki = k0 + c * 4;
      const bool ok = oi < O && ki < K;
      const float* src = ok ? base + (size_t)oi * ld + ki : base;
      const unsigned int d = d0 + r * 128 + ((c ^ (r & 7)) << 4);
      const int n = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(n) : "memory");
    }
  } else {
    const int cpr = rows_mn >> 2;                    // 16-byte chunks per reduction row (a power of two)
    const int sh = 31 - __clz(cpr);
    const int n16 = 32 * cpr;
    for (int q = tid; q < n16; q += kStageThreads) {
      const int kr = q >> sh, mc = q & (cpr - 1);    // reduction row, chunk along MN
      const int g = mc >> 3, c = mc & 7;             // 32-element MN group, chunk inside the 128-byte row
      const int oi = o0 + mc * 4, ki = k0 + kr;
      const bool ok = oi < O && ki < K;
      const float* src = ok ? base + (size_t)ki * ld + oi : base;
#ifdef TD3_TC_DEBUG
      unsigned int cs = (unsigned)c;
      if (g_tc_dbg[4] == 0) cs = ((((c >> 1) ^ (kr & 3)) << 1) | (c & 1));
      else if (g_tc_dbg[4] == 1) cs = c ^ (kr & 7);
      const unsigned int d = d0 + g * g_tc_dbg[2] + kr * 128 + (cs << 4);
#else
      const unsigned int d = d0 + g * 4096 + kr * 128 + (((((c >> 1) ^ (kr & 3)) << 1) | (c & 1)) << 4);
#endif
      const int n = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(n) : "memory");
    }
  }
}

// ------------------------------------------------------------------------------------
// One 128 x NT tile.  `ring` is kTcRingBytes of 1024-byte-aligned shared memory.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void gemm_tile_tc(const Problem& P, int tile, unsigned char* ring, TcState* st) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NT = P.tc_nt;

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
  const float* __restrict__ A = P.A + go * P.a_go + gi * P.a_gi;
  const float* __restrict__ B = P.B + go * P.b_go + gi * P.b_gi;

  int k_begin = 0, k_end = P.K;
  if (P.ksplit > 1) {
    const int chunks = (P.K + 31) / 32;
    const int per = (chunks + P.ksplit - 1) / P.ksplit;
    k_begin = min(P.K, ks * per * 32);
    k_end = min(P.K, (ks + 1) * per * 32);
  }
  const int n_chunks = (k_end - k_begin + 31) / 32;

  const unsigned int tmem = st->tmem_base;
  unsigned int gchunk = st->chunk_count;             // uniform across the CTA (every thread reads the same value)
  const unsigned int idesc = tc_idesc(NT, !arc, !brc);
  // per-operand descriptor geometry
  const unsigned int a_lbo = arc ? 16 : 4096, b_lbo = brc ? 16 : 4096;
  const unsigned int a_sbo = arc ? 1024 : 512, b_sbo = brc ? 1024 : 512;
  const unsigned int a_lt = arc ? 2 : 1, b_lt = brc ? 2 : 1;
  const unsigned int a_kstep = arc ? 32 : 1024, b_kstep = brc ? 32 : 1024;
#ifdef TD3_TC_DEBUG
  const unsigned int a_lbo_ = arc ? 16 : g_tc_dbg[2], b_lbo_ = brc ? 16 : g_tc_dbg[2];
  const unsigned int a_sbo_ = arc ? 1024 : g_tc_dbg[1], b_sbo_ = brc ? 1024 : g_tc_dbg[1];
  const unsigned int a_lt_ = arc ? 2 : g_tc_dbg[0], b_lt_ = brc ? 2 : g_tc_dbg[0];
  const unsigned int a_kstep_ = arc ? 32 : g_tc_dbg[3], b_kstep_ = brc ? 32 : g_tc_dbg[3];
  const unsigned int idesc_ = g_tc_dbg[5] ? tc_idesc(NT, !brc, !arc) : idesc;
#define a_lbo a_lbo_
#define b_lbo b_lbo_
#define a_sbo a_sbo_
#define b_sbo b_sbo_
#define a_lt a_lt_
#define b_lt b_lt_
#define a_kstep a_kstep_
#define b_kstep b_kstep_
#define idesc idesc_
#endif

#pragma unroll 1
  for (int c = -(kTcSlots - 1); c < n_chunks; ++c) {
    const int cn = c + kTcSlots - 1;                 // chunk to stage now
    if (cn < n_chunks) {
      const unsigned int gc = gchunk + cn, slot = gc % kTcSlots, use = gc / kTcSlots;
      if (use > 0) mbar_wait(&st->slot_bar[slot], (use - 1) & 1);     // the MMAs of the slot's previous chunk are done
      unsigned char* sa = ring + slot * kTcSlotBytes;
      tc_issue_operand(sa, A, P.lda, arc, 128, i0, P.M, k_begin + cn * 32, k_end, tid);
      tc_issue_operand(sa + 16384, B, P.ldb, brc, NT, j0, P.N, k_begin + cn * 32, k_end, tid);
    }
    cp_async_commit();
    if (c >= 0) {
      cp_async_wait<kTcSlots - 2>();                 // this thread's part of chunk c has landed ...
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // ... and is visible to the tensor core's proxy
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const unsigned int slot = (gchunk + c) % kTcSlots;
        const unsigned int sa = smem_u32(ring + slot * kTcSlotBytes), sb = sa + 16384;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const unsigned long long da = tc_desc(sa + kk * a_kstep, a_lbo, a_sbo, a_lt);
          const unsigned long long db = tc_desc(sb + kk * b_kstep, b_lbo, b_sbo, b_lt);
          tc_mma(tmem, da, db, idesc, (c > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(&st->slot_bar[slot]);
        if (c == n_chunks - 1) tc_commit(&st->done_bar);
      }
    }
  }
  cp_async_wait<0>();

  // ---- epilogue: TMEM -> registers -> (bias / activation) -> global ----
  float* __restrict__ C = P.C + go * P.c_go + gi * P.c_gi + (long long)ks * P.c_split;
  const float* bias = P.bias ? P.bias + go * P.bias_go + gi * P.bias_gi : nullptr;
  float* aux0 = P.aux0 ? P.aux0 + go * P.aux0_go + gi * P.aux0_gi : nullptr;
  const int epi = P.epi;
  const bool aux_read = aux0 && (epi == EPI_BIAS_TANH_NOISE || epi == EPI_RELU_MASK || epi == EPI_TANH_GRAD);
  const int i = i0 + (warp & 3) * 32 + lane;         // TMEM lane == accumulator row
  const bool row_ok = i < P.M;
  const unsigned int tile_no = st->tile_count;
  bool waited = false;
  const bool c_vec = P.c_vec, x_vec = P.aux_vec;
#pragma unroll 1
  for (int cg = (warp >> 2); cg < (NT >> 4); cg += 2) {   // 16-column groups; warps 0-3 take even, 4-7 odd groups
    const int jb = j0 + cg * 16;
    float bv[16], av[16];
#pragma unroll
    for (int e4 = 0; e4 < 4; ++e4) {                      // operands of the epilogue: in flight while the MMAs finish
      const int j = jb + e4 * 4;
      if (aux_read && row_ok && x_vec && j + 3 < P.N) {
        const float4 v = *reinterpret_cast<const float4*>(aux0 + (size_t)i * P.ldaux + j);
        av[e4 * 4 + 0] = v.x; av[e4 * 4 + 1] = v.y; av[e4 * 4 + 2] = v.z; av[e4 * 4 + 3] = v.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) av[e4 * 4 + e] = (aux_read && row_ok && j + e < P.N) ? aux0[(size_t)i * P.ldaux + j + e] : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) bv[e4 * 4 + e] = (bias && j + e < P.N) ? bias[j + e] : 0.f;
    }
    if (!waited) {
      if (n_chunks > 0) mbar_wait(&st->done_bar, tile_no & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      waited = true;
    }
    unsigned int r[16];
    const unsigned int taddr = tmem + (((unsigned)(warp & 3) * 32u) << 16) + (unsigned)(cg * 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    if (row_ok) {
#pragma unroll
      for (int e4 = 0; e4 < 4; ++e4) {
        const int j = jb + e4 * 4;
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = n_chunks > 0 ? __uint_as_float(r[e4 * 4 + e]) : 0.f;
          if (epi != EPI_STORE && j + e < P.N) {
            const float2 ev = apply_epilogue(epi, v, bv[e4 * 4 + e], av[e4 * 4 + e], P.f0, P.f1);
            v = ev.x;
            if (epi == EPI_BIAS_TANH) aux0[(size_t)i * P.ldaux + j + e] = ev.y;
          }
          o[e] = v;
        }
#pragma unroll 1
        for (int d = 0; d < P.c_dups; ++d) {
          float* cp = C + d * P.c_dup_stride + (size_t)i * P.ldc + j;
          if (c_vec && j + 3 < P.N) {
            *reinterpret_cast<float4*>(cp) = make_float4(o[0], o[1], o[2], o[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (j + e < P.N) cp[e] = o[e];
          }
        }
      }
    }
  }
  if (!waited && n_chunks > 0) mbar_wait(&st->done_bar, tile_no & 1);   // warps without a column group still track the phase
  // every warp is past its TMEM reads before the next tile's first MMA overwrites the accumulator
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    st->chunk_count = gchunk + (unsigned)max(n_chunks, 0);
    if (n_chunks > 0) st->tile_count = tile_no + 1;
  }
  __syncthreads();
}

#ifdef TD3_TC_DEBUG
#undef a_lbo
#undef b_lbo
#undef a_sbo
#undef b_sbo
#undef a_lt
#undef b_lt
#undef a_kstep
#undef b_kstep
#undef idesc
#endif
}  // namespace td3
