// tcpipe.cuh -- throughput form of the tensor-core stage launch (populations of agents, the batch-8192 update).
//
// The latency form (persist.cuh: stage_kernel) gives every 128 x NT tile its own CTA: barrier init, TMEM allocation, the
// first operand round trip, the K loop, the accumulator drain and the epilogue of a tile run back to back, and a CTA
// retires after ~4 us of which the tensor pipe is busy for a tenth.  That is the right shape for the single-agent update
// (a launch is one wave of tiles on a dependency chain) and the wrong one when a launch holds several tiles per SM: ncu
// shows the many-tile stages at 9 % tensor-pipe activity, 5-6 us of SM time per tile.
//
// Here one CTA per SM walks the tiles blockIdx.x, + gridDim.x, ... with its warps specialised across tiles:
//   warp 0      producer: TMA loads of tile n + 1 are issued while tile n is still being multiplied / drained (the operand
//               ring does not care about tile boundaries);
//   warp 1      MMA issuer: two TMEM accumulators (2 x 128 columns), tile n + 1 accumulates while tile n is drained;
//   warps 4-7   epilogue: tcgen05.ld -> bias / activation / mask -> global, one warp per 32 TMEM lanes, all NT columns;
//   warp 2      TMEM allocation.
// Barriers: full / empty per ring slot (as in tc.cuh), acc_full / acc_empty per accumulator.  The tiles of the stage that
// do not run on the tensor cores (column sums, FFMA problems) follow on the same CTAs once the pipeline has drained.
#pragma once

#include "persist.cuh"

namespace td3 {

constexpr int kPipeCols = 256;                      // two 128-column accumulators
constexpr int kPipeSmemBytes = kDynSmemBytes + 4 * 2048;   // + the epilogue warps' transposition staging

struct PipeBars {
  unsigned long long acc_full[2];                   // "accumulator b holds a finished tile" (tcgen05.commit)
  unsigned long long acc_empty[2];                  // "accumulator b has been read out" (one arrival per epilogue warp)
};

struct PipeTile {
  int pi, g, go, gi, i0, j0, n_chunks, NT, arc, brc;
};

__device__ __forceinline__ PipeTile pipe_tile(const StageParams& S, int tile_global) {
  PipeTile T;
  int pi = 0;
#pragma unroll
  for (int q = 1; q < kMaxProblemsPerStage; ++q)
    if (q < S.n_problems && tile_global >= S.p[q].tile_begin) pi = q;
  const Problem& P = S.p[pi];
  int t = tile_global - P.tile_begin;
  T.pi = pi;
  T.g = t / P.tiles_per_group;
  t -= T.g * P.tiles_per_group;
  T.go = T.g / P.groups_inner;
  T.gi = T.g - T.go * P.groups_inner;
  const int tm = t / P.tiles_n, tn = t - tm * P.tiles_n;
  T.NT = P.tc_nt;
  T.i0 = tm * 128;
  T.j0 = tn * T.NT;
  T.n_chunks = (P.K + 31) / 32;                     // ksplit == 1 (host: layout_stage)
  T.arc = P.a_rc;
  T.brc = P.b_rc;
  return T;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kStageThreads, 1) stage_pipe_kernel(const __grid_constant__ StageParams S) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ TcState tc;
  __shared__ PipeBars pb;
  unsigned char* ring = aligned_smem(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();
  if (tid == 0) {
    for (int i = 0; i < kTcMaxSlots; ++i) {
      mbar_init(&tc.full_bar[i], 1);
      mbar_init(&tc.empty_bar[i], 1);
    }
    mbar_init(&tc.done_bar, 1);
    mbar_init(&tc.tmem_bar, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&pb.acc_full[b], 1);
      mbar_init(&pb.acc_empty[b], 4);
    }
    tc.tma_chunk_count = 0;
    tc.prof_stage = -1;
    tc.tile_count = 0;
    tc.group = kTcGroup;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (warp == kTcAllocWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tc.tmem_base)), "r"(kPipeCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&tc.tmem_bar)) : "memory");
  }
  pdl_wait();                                       // the previous stage's outputs are complete and visible from here on

  const int tc_tiles = S.pipe_tiles;
  const int grp = kTcGroup;
  const unsigned int slot_bytes = (unsigned)(grp * kTcSub);
  float* bias_s = reinterpret_cast<float*>(ring + kTcSlots * grp * kTcSub);
#define TD3_UNI(x) __shfl_sync(0xffffffffu, (x), 0)
  if (warp == 0) {
    // ---- producer ----
    const unsigned int ring_u = TD3_UNI(smem_u32(ring)), fb0 = TD3_UNI(smem_u32(&tc.full_bar[0])), eb0 = TD3_UNI(smem_u32(&tc.empty_bar[0]));
    const unsigned int nsl = TD3_UNI((unsigned)kTcSlots), sbytes = TD3_UNI(slot_bytes);
    unsigned int slot = 0, use = 0;
    unsigned int sa0 = ring_u, fb = fb0, eb = eb0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < tc_tiles; tile += gridDim.x) {
      const PipeTile T = pipe_tile(S, tile);
      const Problem& P = S.p[T.pi];
      const bool pm = P.map_a >= 0 && P.map_b >= 0;
      const unsigned char* mapA = pm ? reinterpret_cast<const unsigned char*>(S.maps + P.map_a + T.g)
                                     : reinterpret_cast<const unsigned char*>(P.tmapA) + (size_t)T.g * 128;
      const unsigned char* mapB = pm ? reinterpret_cast<const unsigned char*>(S.maps + P.map_b + T.g)
                                     : reinterpret_cast<const unsigned char*>(P.tmapB) + (size_t)T.g * 128;
      if (lane == 0) asm volatile("prefetch.tensormap [%0];\n" ::"l"(mapA) : "memory");
      if (lane == 1) asm volatile("prefetch.tensormap [%0];\n" ::"l"(mapB) : "memory");
      const unsigned int bytes = 16384u + (T.brc ? (unsigned)T.NT * 128u : (unsigned)((T.NT + 31) >> 5) * 4096u);
      const unsigned int ubytes = TD3_UNI(bytes);
      const int nch = TD3_UNI(T.n_chunks), ui0 = TD3_UNI(T.i0), uj0 = TD3_UNI(T.j0), unt = TD3_UNI(T.NT);
      const int uarc = TD3_UNI(T.arc), ubrc = TD3_UNI(T.brc);
      const unsigned long long ma = ((unsigned long long)TD3_UNI((unsigned int)((unsigned long long)mapA >> 32)) << 32) |
                                    TD3_UNI((unsigned int)(unsigned long long)mapA);
      const unsigned long long mb = ((unsigned long long)TD3_UNI((unsigned int)((unsigned long long)mapB >> 32)) << 32) |
                                    TD3_UNI((unsigned int)(unsigned long long)mapB);
      const int nst = (nch + grp - 1) / grp;
      int kc = 0;
#pragma unroll 1
      for (int c = 0; c < nst; ++c) {
        if (use > 0) mbar_wait_u32(eb, (use - 1) & 1);
        const int nsub = min(grp, nch - c * grp);
        if (elect_one()) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(fb), "r"(ubytes * nsub) : "memory");
#pragma unroll 1
          for (int u = 0; u < nsub; ++u) {
            const unsigned int sa = sa0 + u * kTcSub;
            const int k0 = kc + u * 32;
            if (uarc) {
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
        kc += grp * 32; sa0 += sbytes; fb += 8; eb += 8;
        if (++slot == nsl) { slot = 0; ++use; sa0 = ring_u; fb = fb0; eb = eb0; }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    const unsigned int tmem = TD3_UNI(tc_tmem_base(&tc));
    const unsigned int ring_u = TD3_UNI(smem_u32(ring)), fb0 = TD3_UNI(smem_u32(&tc.full_bar[0])), eb0 = TD3_UNI(smem_u32(&tc.empty_bar[0]));
    const unsigned int af0 = TD3_UNI(smem_u32(&pb.acc_full[0])), ae0 = TD3_UNI(smem_u32(&pb.acc_empty[0]));
    const unsigned int nsl = TD3_UNI((unsigned)kTcSlots), sbytes = TD3_UNI(slot_bytes);
    unsigned int slot = 0, use = 0;
    unsigned int sa = ring_u, fb = fb0, eb = eb0;
    unsigned int n = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < tc_tiles; tile += gridDim.x, ++n) {
      const PipeTile T = pipe_tile(S, tile);
      const unsigned int buf = n & 1u;
      if (n >= 2) mbar_wait_u32(ae0 + buf * 8, ((n >> 1) - 1) & 1);     // the epilogue warps have read this accumulator out
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const int arc = T.arc, brc = T.brc;
      const unsigned int a_lbo = arc ? 16 : 4096, b_lbo = brc ? 16 : 4096;
      const unsigned int a_sbo = arc ? 1024 : 512, b_sbo = brc ? 1024 : 512;
      const unsigned int a_lt = arc ? 2 : 1, b_lt = brc ? 2 : 1;
      const unsigned int a_kstep = arc ? 32 : 1024, b_kstep = brc ? 32 : 1024;
      const unsigned int uidesc = TD3_UNI(tc_idesc(T.NT, !arc, !brc));
      const unsigned int a_hi = TD3_UNI((a_sbo >> 4) | (1u << 14) | (a_lt << 29)), b_hi = TD3_UNI((b_sbo >> 4) | (1u << 14) | (b_lt << 29));
      const unsigned int a_lo0 = TD3_UNI((a_lbo >> 4) << 16), b_lo0 = TD3_UNI((b_lbo >> 4) << 16);
      const unsigned int a_ks = TD3_UNI(a_kstep >> 4), b_ks = TD3_UNI(b_kstep >> 4);
      const unsigned int utmem = TD3_UNI(tmem + buf * 128u);
      const int nch = TD3_UNI(T.n_chunks);
      const int nst = (nch + grp - 1) / grp;
#pragma unroll 1
      for (int c = 0; c < nst; ++c) {
        mbar_wait_u32(fb, use & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const int nsub = min(grp, nch - c * grp);
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
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(eb) : "memory");
          if (c == nst - 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(af0 + buf * 8) : "memory");
        }
        __syncwarp();
        sa += sbytes; fb += 8; eb += 8;
        if (++slot == nsl) { slot = 0; ++use; sa = ring_u; fb = fb0; eb = eb0; }
      }
    }
  } else if (warp >= 4) {
    // ---- epilogue: warp w drains TMEM lanes [32 (w & 3), + 32) = accumulator rows, all NT columns ----
    const unsigned int tmem = tc_tmem_base(&tc);
    float* stg = bias_s + 128 + (warp & 3) * 512;   // transposition staging of the coalesced epilogue (tc.cuh), 2 KB per warp
    unsigned int n = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < tc_tiles; tile += gridDim.x, ++n) {
      const PipeTile T = pipe_tile(S, tile);
      const Problem& P = S.p[T.pi];
      const unsigned int buf = n & 1u;
      const int NT = T.NT, i0 = T.i0, j0 = T.j0;
      const long long go = T.go, gi = T.gi;
      named_bar_sync(1, 128);                       // the previous tile's bias strip has been read by all four warps
      if (P.bias && tid - 128 < NT) {
        const float* bias = P.bias + go * P.bias_go + gi * P.bias_gi;
        const int j = j0 + tid - 128;
        bias_s[tid - 128] = j < P.N ? bias[j] : 0.f;
      }
      named_bar_sync(1, 128);
      float* __restrict__ C = P.C + go * P.c_go + gi * P.c_gi;
      const bool has_bias = P.bias != nullptr;
      float* aux0 = P.aux0 ? P.aux0 + go * P.aux0_go + gi * P.aux0_gi : nullptr;
      const int epi = P.epi;
      const bool aux_read = aux0 && (epi == EPI_BIAS_TANH_NOISE || epi == EPI_RELU_MASK || epi == EPI_TANH_GRAD);
      const int i = i0 + (warp & 3) * 32 + lane;
      const bool row_ok = i < P.M;
      const bool c_vec = P.c_vec, x_vec = P.aux_vec;
      TD3_DISPATCH_EPI(epi, (tc_epilogue_cols<E, true>(P, C, aux0, bias_s, has_bias, aux_read, x_vec, c_vec, row_ok, true, i, j0, 0, NT,
                                                 tmem + buf * 128u + (((unsigned)(warp & 3) * 32u) << 16), smem_u32(&pb.acc_full[buf]),
                                                 (n >> 1) & 1, stg)));
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&pb.acc_empty[buf])) : "memory");
    }
  }
#undef TD3_UNI
  // pipeline drained: every operand byte has landed and been consumed, every accumulator read out
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  // the stage's other tiles (column sums, FFMA problems, reductions): the ring is free now
#pragma unroll 1
  for (int tile = tc_tiles + blockIdx.x; tile < S.total_tiles; tile += gridDim.x) run_stage_tile<false>(S, tile, ring, &tc, true);
  __syncthreads();
  if (warp == kTcAllocWarp) {
    const unsigned int base = tc_tmem_base(&tc);
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base), "r"(kPipeCols) : "memory");
  }
}

}  // namespace td3
