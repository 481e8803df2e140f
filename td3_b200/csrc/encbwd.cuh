// encbwd.cuh -- K6 backward: the particle-set encoder's backward pass (TD3_particles.py:53-58 under autograd) as two
// persistent tcgen05 kernels that read nothing but the particles, three ReLU bitmaps and d(loss)/d(pooled):
//
//     g[b, o]    = dpool[b, o] / N * (pooled[b, o] > 0)                       gradient of relu(mean_n h2[b, n, o])
//     dz2[p, o]  = g[b(p), o] * (h2[p, o] > 0)
//     dW2[o, c]  = sum_p dz2[p, o] h1[p, c]         db2[o] = sum_p dz2[p, o]                         enc_bwd_w2_kernel
//     dz1[p, c]  = (sum_o dz2[p, o] W2[o, c]) * (h1[p, c] > 0)
//     dW1[c, d]  = sum_p dz1[p, c] P[p, d]          db1[c] = sum_p dz1[p, c]                         enc_bwd_x_kernel
//
// As generic stages (pool-backward -> conv2 dX + dW split-K -> conv1 dW) the intermediates dz2 / dz1 ([B*N, 128] and
// [B*N, 256] floats) and the stored activations h1 / h2 cross HBM four times: 1.3 ms of a 1.9 ms critic update at
// B = 256, N = 1024.  Here the forward pass stores only the sign bits of h1 / h2 (enc.cuh: 64 bytes per particle
// instead of 1.5 KB), h1 is recomputed from the particles by one K = 8 MMA per tile, and every contraction runs on
// tcgen05 with hand-built K-major operands (128-byte swizzle) whose rows are written by the thread that owns the
// corresponding TMEM lane -- every product is computed TRANSPOSED (output channel = TMEM lane, particle = column), so
// an epilogue thread holds 32 consecutive particles of its channel: one 128-byte operand row per pass, its bitmap word
// in one register, bias gradients as a scalar running sum (a popcount for db2).
//
//   enc_bwd_w2_kernel   per 128-particle tile: z1^T = W1ext Pext^T (two M128 N128 K8 MMAs) -> relu, round -> h1^T rows
//                       [256 c][128 p]; dz2^T rows [128 o][128 p] = bit ? g : 0; dW2 += dz2^T (h1^T)^T as 16 M128 N256 K8
//                       MMAs into a TMEM accumulator that lives for the CTA's whole run of tiles
//   enc_bwd_x_kernel    a CTA owns one half (128) of the hidden channels: W2^T half resident (64 KB); per tile dz2 rows
//                       [128 p][128 o] -> dh1^T half = W2^T dz2^T (16 M128 N128 K8 MMAs, two TMEM accumulators) -> mask,
//                       round -> dz1^T rows [128 c][128 p] -> dW1 += dz1^T [P_hi | P_lo] (16 M128 N16 K8 MMAs: the
//                       particle coordinates enter as a two-term TF32 split, so they count at fp32 accuracy)
//
// Every CTA leaves one partial (dW2 / db2, or its half of dW1 / db1) in the split-K buffer the stage path already
// reduces in fixed order (PK_REDUCE_SPLITS): deterministic, no atomics.
// Shapes: 256 hidden / 128 output channels, D <= 7, N % 128 == 0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "enc.cuh"

namespace td3 {

constexpr int kEbThreads = 9 * 32;                                 // dW2 kernel: 8 worker warps + the MMA warp
constexpr int kEbXThreads = 13 * 32;                               // dX kernel: 4 builder + 8 epilogue warps + the MMA warp
constexpr int kEbChunk128 = 128 * 128, kEbChunk256 = 256 * 128, kEbChunk16 = 16 * 128;   // bytes of a [rows][32 fp32] K-major chunk
constexpr int kEbW2SmemBytes = 1024 + 4 * kEbChunk256 + 4 * kEbChunk128 + kEncW1Bytes + 2 * kEncPBytes + 256;
constexpr int kEbXSmemBytes = 1024 + 3 * 4 * kEbChunk128 + 2 * 4 * kEbChunk16 + 2 * 128 * 4 + 256;    // W2^T half + two dz2 buffers + particles + g

struct EncBwdParams {
  const float* P; long long p_go;                     // particles [rows, D] of outer group o at P + o * p_go
  const float* W1; const float* b1;                   // conv1 parameters (fp32 masters) of group (o, i) at + o * w_go + i * w_gi
  const float* W2;                                    // conv2 weight [128, 256] (the TF32-rounded shadow when there is one)
  long long w_go, w_gi;
  const float* dpool; int ld_dpool, pad0; long long dpool_go, dpool_gi;     // d(loss)/d(pooled) [B, >= 128]
  const float* pooled; int ld_pooled, pad1; long long pooled_go, pooled_gi; // relu(mean) [B, >= 128]
  const unsigned int* bits; long long bits_go, bits_gi;                     // EncBits block of a group (words)
  float* part; long long part_go, part_gi;            // this kernel's split-K partials: dW2 [ks][128*256], db2 [ks][128] / dW1 [ks][256*D], db1 [ks][256]
  int rows, D, n_particles, n_inner, n_groups, tiles_per_group, ks, pad2;
  float inv_n, pad3;
};

// one 128-byte row (32 consecutive K elements) of a K-major, 128-byte-swizzled operand chunk
__device__ __forceinline__ void eb_store_row(unsigned char* chunk, int row, const float (&v)[32]) {
  unsigned char* r = chunk + row * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(r + ((j ^ (row & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
// byte offset of element (row, k) inside such a chunk
__device__ __forceinline__ int eb_elem(int row, int k) { return row * 128 + ((((k & 31) >> 2) ^ (row & 7)) << 4) + (k & 3) * 4; }

// TMEM columns of the dX kernel: [0, 256) the two dh1^T accumulators, [256, 384) dz1^T (A operand of the dW1 MMAs, written
// by the epilogue warps with tcgen05.st: it never touches shared memory), [384, 400) the dW1 accumulator
constexpr unsigned int kEbXDz1Col = 256u, kEbXDw1Col = 384u;

// 32 consecutive columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void eb_tmem_st32(unsigned int taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]),
      "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]),
      "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
      : "memory");
}

// tcgen05.mma with the A operand in tensor memory (lane = row, column = reduction index) and B from a shared-memory descriptor
__device__ __forceinline__ void tc_mma_ts(unsigned int tmem_d, unsigned int tmem_a, unsigned long long db, unsigned int idesc,
                                          unsigned int accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void eb_tmem_ld16(unsigned int taddr, unsigned int (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// g[b, o] of this tile's sample, rounded to nearest TF32
__device__ __forceinline__ float eb_pool_grad(const EncBwdParams& E, int go, int gi, int b, int o) {
  const float dp = E.dpool[(long long)go * E.dpool_go + (long long)gi * E.dpool_gi + (long long)b * E.ld_dpool + o];
  const float pl = E.pooled[(long long)go * E.pooled_go + (long long)gi * E.pooled_gi + (long long)b * E.ld_pooled + o];
  return rn_tf32(pl > 0.f ? dp * E.inv_n : 0.f);
}

// ---------------------------------------------------------------------------------------------------------------------
// dW2 / db2
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEbThreads, 1) enc_bwd_w2_kernel(const __grid_constant__ EncBwdParams E) {
  extern __shared__ unsigned char eb_smem_raw[];
  __shared__ unsigned long long p_full[2], h1_full, h1_empty, ops_full, ops_empty, done_bar;
  __shared__ unsigned int tmem_base_s;
  unsigned char* base = eb_smem_raw + ((1024u - (smem_u32(eb_smem_raw) & 1023u)) & 1023u);
  unsigned char* H1T = base;                                        // 4 chunks x [256 c][32 p]
  unsigned char* DZ2T = H1T + 4 * kEbChunk256;                      // 4 chunks x [128 o][32 p]
  unsigned char* W1k = DZ2T + 4 * kEbChunk128;                      // [256 rows][8]: W1 row, 0 padding, bias in slot 7
  unsigned char* Pk = W1k + kEncW1Bytes;                            // 2 x [128 rows][8]: particle, 0 padding, 1 in slot 7
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&p_full[0], 4); mbar_init(&p_full[1], 4);
    mbar_init(&h1_full, 1); mbar_init(&h1_empty, 8);
    mbar_init(&ops_full, 8); mbar_init(&ops_empty, 1); mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned int tmem = *reinterpret_cast<volatile unsigned int*>(&tmem_base_s);

  const int g = (int)blockIdx.x / E.ks, k = (int)blockIdx.x - g * E.ks;
  const int go = g / E.n_inner, gi = g - go * E.n_inner;
  const int D = E.D;
  const int n_my = (E.tiles_per_group - k + E.ks - 1) / E.ks;       // tiles k, k + ks, ...  (>= 1: the host keeps ks <= tiles)
  const long long wofs = (long long)go * E.w_go + (long long)gi * E.w_gi;
  const unsigned int* bits2T = E.bits + (long long)go * E.bits_go + (long long)gi * E.bits_gi + (long long)E.rows * 4;

  if (warp < 8) {
    const int lq = warp & 3, half = warp >> 2;
    const int c = half * 128 + lq * 32 + lane;                      // this thread's hidden channel == its TMEM lane (+ accumulator half)
    {                                                               // layer-1 operand: [W1 | 0 | b1] rows, rounded to nearest TF32
      float w[8];
#pragma unroll
      for (int d = 0; d < 8; ++d) w[d] = d < D ? rn_tf32(E.W1[wofs + (long long)tid * D + d]) : (d == 7 ? rn_tf32(E.b1[wofs + tid]) : 0.f);
      *reinterpret_cast<float4*>(W1k + enc_k8_offset(tid, 0)) = make_float4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<float4*>(W1k + enc_k8_offset(tid, 1)) = make_float4(w[4], w[5], w[6], w[7]);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      enc_named_barrier(2, 256);
    }
    auto stage_particles = [&](int rt, unsigned int n) {
      if (tid < kEncTile) {
        const float* Pg = E.P + (long long)go * E.p_go + ((long long)rt * kEncTile + tid) * D;
        float p[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) p[d] = d < D ? rn_tf32(__ldg(Pg + d)) : (d == 7 ? 1.f : 0.f);
        unsigned char* dst = Pk + (n & 1u) * kEncPBytes;
        *reinterpret_cast<float4*>(dst + enc_k8_offset(tid, 0)) = make_float4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<float4*>(dst + enc_k8_offset(tid, 1)) = make_float4(p[4], p[5], p[6], p[7]);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) enc_arrive(&p_full[n & 1u]);
      }
    };
    stage_particles(k, 0);
    float db2 = 0.f;
    for (int t = 0; t < n_my; ++t) {
      const int rt = k + t * E.ks;
      if (t + 1 < n_my) stage_particles(rt + E.ks, (unsigned)t + 1u);
      uint4 bw = make_uint4(0, 0, 0, 0);
      float gv = 0.f;
      if (tid < kEncO) {                                            // output channel o = tid: its gradient scalar and bitmap row
        const int b = (int)(((long long)rt * kEncTile) / E.n_particles);
        gv = eb_pool_grad(E, go, gi, b, tid);
        bw = __ldg(reinterpret_cast<const uint4*>(bits2T + ((long long)rt * kEncO + tid) * 4));
      }
      if (t > 0) mbar_wait(&ops_empty, (unsigned)(t - 1) & 1u);     // the previous tile's MMAs have read both operand buffers
      if (tid < kEncO) {
        const unsigned int ww[4] = {bw.x, bw.y, bw.z, bw.w};
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = (ww[ch] >> i) & 1u ? gv : 0.f;
          eb_store_row(DZ2T + ch * kEbChunk128, tid, v);
        }
        db2 += gv * (float)(__popc(bw.x) + __popc(bw.y) + __popc(bw.z) + __popc(bw.w));
      }
      mbar_wait(&h1_full, (unsigned)t & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      // the four 32-particle passes, two at a time: the TMEM read of the second is in flight while the first is converted
      // and stored
      const unsigned int z1_addr = tmem + (unsigned)(half * 128) + (((unsigned)lq * 32u) << 16);
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
        unsigned int ra[32], rb[32];
        enc_tmem_ld32_issue(z1_addr + (unsigned)(pp * 64), ra);
        enc_tmem_ld32_issue(z1_addr + (unsigned)(pp * 64 + 32), rb);
        enc_tmem_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = rn_tf32_finite(fmaxf(__uint_as_float(ra[i]), 0.f));
        eb_store_row(H1T + (2 * pp) * kEbChunk256, c, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = rn_tf32_finite(fmaxf(__uint_as_float(rb[i]), 0.f));
        eb_store_row(H1T + (2 * pp + 1) * kEbChunk256, c, v);
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) { enc_arrive(&h1_empty); enc_arrive(&ops_full); }
    }
    // ---- the CTA's partial: dW2 rows o = TMEM lanes (warps 0..3), 256 columns ----
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (half == 0) {
      const int o = lq * 32 + lane;
      float* pw = E.part + (long long)go * E.part_go + (long long)gi * E.part_gi;
      float* dst = pw + (long long)k * (kEncO * kEncH) + (long long)o * kEncH;
#pragma unroll 1
      for (int pass = 0; pass < 8; ++pass) {
        unsigned int r[32];
        enc_tmem_ld32(tmem + 256u + (unsigned)(pass * 32) + (((unsigned)lq * 32u) << 16), r);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + pass * 32 + j) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
      pw[(long long)E.ks * (kEncO * kEncH) + (long long)k * kEncO + o] = db2;
    }
  } else {
    // ------------------------------------------------------------------ MMA warp
    const unsigned int idesc1 = tc_idesc(kEncTile, 0, 0), idesc2 = tc_idesc(kEncH, 0, 0);
    const unsigned int hi = (1024u >> 4) | (1u << 14) | (2u << 29);        // SBO 1024 B, descriptor version, SWIZZLE_128B
    const unsigned int lo0 = (16u >> 4) << 16;
    const unsigned int hi1 = (256u >> 4) | (1u << 14);                     // layer-1 operands: SBO 256 B, no swizzle
    const unsigned int lo1 = (128u >> 4) << 16;
    const unsigned int wa = __shfl_sync(0xffffffffu, smem_u32(W1k), 0);
    const unsigned int a_base = __shfl_sync(0xffffffffu, smem_u32(DZ2T), 0);
    const unsigned int b_base = __shfl_sync(0xffffffffu, smem_u32(H1T), 0);
    for (int t = 0; t < n_my; ++t) {
      const unsigned int ut = (unsigned)t;
      mbar_wait(&p_full[ut & 1u], (ut >> 1) & 1u);
      if (t > 0) mbar_wait(&h1_empty, (ut - 1u) & 1u);              // the workers have read the previous tile's z1
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const unsigned int pa = __shfl_sync(0xffffffffu, smem_u32(Pk + (ut & 1u) * kEncPBytes), 0);
      if (elect_one()) {
        const unsigned long long db = ((unsigned long long)hi1 << 32) | (lo1 | (pa >> 4));
        tc_mma(tmem, ((unsigned long long)hi1 << 32) | (lo1 | (wa >> 4)), db, idesc1, 0u);
        tc_mma(tmem + 128u, ((unsigned long long)hi1 << 32) | (lo1 | ((wa + 4096u) >> 4)), db, idesc1, 0u);
        tc_commit(&h1_full);
      }
      __syncwarp();
      mbar_wait(&ops_full, ut & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const unsigned int a_lo = lo0 | ((a_base + ch * kEbChunk128) >> 4), b_lo = lo0 | ((b_base + ch * kEbChunk256) >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tc_mma(tmem + 256u, ((unsigned long long)hi << 32) | (a_lo + kk * 2), ((unsigned long long)hi << 32) | (b_lo + kk * 2), idesc2,
                   (ut | (unsigned)ch | (unsigned)kk) != 0 ? 1u : 0u);
        }
        tc_commit(&ops_empty);
        if (t + 1 == n_my) tc_commit(&done_bar);
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// dz2 -> dh1 -> dz1 -> dW1 / db1, one half of the hidden channels per CTA
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEbXThreads, 1) enc_bwd_x_kernel(const __grid_constant__ EncBwdParams E) {
  extern __shared__ unsigned char eb_smem_raw[];
  __shared__ unsigned long long dz2_full[2], dz2_empty[2], dh1_full[2], dh1_empty[2], dz1_full, dz1_empty, pt_full[2], pt_empty[2], done_bar;
  __shared__ unsigned int tmem_base_s;
  unsigned char* base = eb_smem_raw + ((1024u - (smem_u32(eb_smem_raw) & 1023u)) & 1023u);
  unsigned char* W2T = base;                                        // 4 chunks x [128 c][32 o]   (this CTA's half of the channels)
  unsigned char* DZ2 = W2T + 4 * kEbChunk128;                       // 2 buffers x 4 chunks x [128 p][32 o]
  unsigned char* PT = DZ2 + 2 * 4 * kEbChunk128;                       // 2 x 4 chunks x [16 d][32 p]: rows d < D = P_hi, 8 + d = P_lo
  float* gs = reinterpret_cast<float*>(PT + 2 * 4 * kEbChunk16);    // 2 x [128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&dz1_full, 8); mbar_init(&dz1_empty, 1); mbar_init(&done_bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&dh1_full[i], 1); mbar_init(&dh1_empty[i], 8); mbar_init(&pt_empty[i], 1); mbar_init(&pt_full[i], 4);
                                  mbar_init(&dz2_full[i], 4); mbar_init(&dz2_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == kEbXThreads / 32 - 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  }
  const int g = (int)blockIdx.x / (2 * E.ks);
  const int h = ((int)blockIdx.x / E.ks) & 1, k = (int)blockIdx.x % E.ks;
  const int go = g / E.n_inner, gi = g - go * E.n_inner;
  const int D = E.D;
  const int n_my = (E.tiles_per_group - k + E.ks - 1) / E.ks;
  const long long wofs = (long long)go * E.w_go + (long long)gi * E.w_gi;
  const unsigned int* bits2 = E.bits + (long long)go * E.bits_go + (long long)gi * E.bits_gi;
  const unsigned int* bits1T = bits2 + (long long)E.rows * 8;

  if (tid < 256) {
    // W2^T half, K-major: element (c, o) = W2[o][h * 128 + c]  (reads coalesced over c; a one-off 64 KB per CTA)
    const int c = tid & 127, o0 = (tid >> 7) * 64;
    const float* W2 = E.W2 + wofs + h * 128 + c;
#pragma unroll 4
    for (int o = o0; o < o0 + 64; ++o)
      *reinterpret_cast<float*>(W2T + (o >> 5) * kEbChunk128 + eb_elem(c, o)) = rn_tf32(__ldg(W2 + (long long)o * kEncH));
    float4* z = reinterpret_cast<float4*>(PT);                      // rows the particle staging never writes stay zero
    for (int i = tid; i < 2 * 4 * kEbChunk16 / 16; i += 256) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned int tmem = *reinterpret_cast<volatile unsigned int*>(&tmem_base_s);

  if (warp < 4) {
    // ------------------------------------------------------------------ builders: thread = particle p of the tile
    // The builder is the head of the tile's dependency chain (dz2 -> dh1 MMAs -> epilogue -> dW1 MMAs): its global reads
    // are issued one tile ahead and consumed at the top of the next iteration, so their latency stays off the chain.
    const int p = tid;
    auto sample_of = [&](int t) { return (int)(((long long)(k + t * E.ks) * kEncTile) / E.n_particles); };
    const long long gofs_dp = (long long)go * E.dpool_go + (long long)gi * E.dpool_gi + p;
    const long long gofs_pl = (long long)go * E.pooled_go + (long long)gi * E.pooled_gi + p;
    float dp_n = E.dpool[gofs_dp + (long long)sample_of(0) * E.ld_dpool];
    float pl_n = E.pooled[gofs_pl + (long long)sample_of(0) * E.ld_pooled];
    uint4 bw_n = __ldg(reinterpret_cast<const uint4*>(bits2 + ((long long)k * kEncTile + p) * 4));
    for (int t = 0; t < n_my; ++t) {
      const unsigned int ut = (unsigned)t;
      float* gsb = gs + (ut & 1u) * 128;
      gsb[p] = rn_tf32(pl_n > 0.f ? dp_n * E.inv_n : 0.f);          // g[b, o = p]  (thread index doubles as the output channel here)
      const uint4 bw = bw_n;
      enc_named_barrier(3, 128);                                    // gs of this tile complete (the other buffer is two tiles old)
      if (t + 1 < n_my) {
        const int rt1 = k + (t + 1) * E.ks;
        dp_n = E.dpool[gofs_dp + (long long)sample_of(t + 1) * E.ld_dpool];
        pl_n = E.pooled[gofs_pl + (long long)sample_of(t + 1) * E.ld_pooled];
        bw_n = __ldg(reinterpret_cast<const uint4*>(bits2 + ((long long)rt1 * kEncTile + p) * 4));
      }
      if (t >= 2) mbar_wait(&dz2_empty[ut & 1u], ((ut >> 1) - 1u) & 1u);   // the dh1 MMAs of two tiles ago have read this DZ2 buffer
      unsigned char* dzb = DZ2 + (ut & 1u) * 4 * kEbChunk128;
      const unsigned int ww[4] = {bw.x, bw.y, bw.z, bw.w};
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float v[32];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 g4 = *reinterpret_cast<const float4*>(gsb + ch * 32 + 4 * i4);
          v[4 * i4] = (ww[ch] >> (4 * i4)) & 1u ? g4.x : 0.f;
          v[4 * i4 + 1] = (ww[ch] >> (4 * i4 + 1)) & 1u ? g4.y : 0.f;
          v[4 * i4 + 2] = (ww[ch] >> (4 * i4 + 2)) & 1u ? g4.z : 0.f;
          v[4 * i4 + 3] = (ww[ch] >> (4 * i4 + 3)) & 1u ? g4.w : 0.f;
        }
        eb_store_row(dzb + ch * kEbChunk128, p, v);
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) enc_arrive(&dz2_full[ut & 1u]);
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------ epilogue: thread = hidden channel h * 128 + c == TMEM lane c;
    // the two warps of a lane quarter take two of the four 32-particle passes each (one warp per scheduler cannot hide
    // the latency of its own instruction stream: two can)
    const int e = warp & 3, hi2 = (warp - 4) >> 2, c = e * 32 + lane;
    float db1 = 0.f;
    // the second warp of each lane quarter also stages the particle coordinates of the tile (B operand of the dW1 MMAs: rows
    // d = P_hi, 8 + d = P_lo; thread = particle c), read one tile ahead
    float pn[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) pn[d] = 0.f;
    if (hi2 == 1) {
      const float* Pg = E.P + (long long)go * E.p_go + ((long long)k * kEncTile + c) * D;
#pragma unroll
      for (int d = 0; d < 8; ++d)
        if (d < D) pn[d] = __ldg(Pg + d);
    }
    for (int t = 0; t < n_my; ++t) {
      const unsigned int ut = (unsigned)t, buf = ut & 1u;
      const int rt = k + t * E.ks;
      if (hi2 == 1) {
        if (t >= 2) mbar_wait(&pt_empty[ut & 1u], ((ut >> 1) - 1u) & 1u);   // the dW1 MMAs of two tiles ago have read this PT buffer
        unsigned char* pt = PT + (ut & 1u) * 4 * kEbChunk16 + (c >> 5) * kEbChunk16;
#pragma unroll
        for (int d = 0; d < 8; ++d)
          if (d < D) {
            const float ph = rn_tf32(pn[d]);
            *reinterpret_cast<float*>(pt + eb_elem(d, c)) = ph;
            *reinterpret_cast<float*>(pt + eb_elem(8 + d, c)) = rn_tf32(pn[d] - ph);
          }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) enc_arrive(&pt_full[ut & 1u]);
        if (t + 1 < n_my) {
          const float* Pg = E.P + (long long)go * E.p_go + ((long long)(rt + E.ks) * kEncTile + c) * D;
#pragma unroll
          for (int d = 0; d < 8; ++d)
            if (d < D) pn[d] = __ldg(Pg + d);
        }
      }
      const uint2 bw = __ldg(reinterpret_cast<const uint2*>(bits1T + ((long long)rt * kEncH + h * 128 + c) * 4 + 2 * hi2));
      mbar_wait(&dh1_full[buf], (ut >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (t > 0) mbar_wait(&dz1_empty, (ut - 1u) & 1u);             // the previous tile's dW1 MMAs have read dz1^T
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int pass = 2 * hi2 + cc;
        unsigned int r[32];
        enc_tmem_ld32(tmem + buf * 128u + (unsigned)(pass * 32) + (((unsigned)e * 32u) << 16), r);
        const unsigned int w = cc == 0 ? bw.x : bw.y;
        float v[32];
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] = (w >> i) & 1u ? rn_tf32_finite(__uint_as_float(r[i])) : 0.f;
          s4[i & 3] += v[i];
        }
        db1 += (s4[0] + s4[1]) + (s4[2] + s4[3]);
        eb_tmem_st32(tmem + kEbXDz1Col + (unsigned)(pass * 32) + (((unsigned)e * 32u) << 16), v);   // row c of dz1^T: the A operand, in TMEM
      }
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) { enc_arrive(&dh1_empty[buf]); enc_arrive(&dz1_full); }
    }
    // the two warps of a lane quarter hold disjoint particle ranges of the same channel: summed in fixed order
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (hi2 == 1) gs[c] = db1;                                      // (gs is free: every build has been consumed)
    enc_named_barrier(4, 256);
    if (hi2 == 0) {
      unsigned int r[16];
      eb_tmem_ld16(tmem + kEbXDw1Col + (((unsigned)e * 32u) << 16), r);
      float* pw = E.part + (long long)go * E.part_go + (long long)gi * E.part_gi;     // (the host passes this kernel's region)
      float* dst = pw + (long long)k * (kEncH * D) + (long long)(h * 128 + c) * D;
      for (int d = 0; d < D; ++d) dst[d] = __uint_as_float(r[d]) + __uint_as_float(r[8 + d]);
      pw[(long long)E.ks * (kEncH * D) + (long long)k * kEncH + h * 128 + c] = db1 + gs[c];
    }
  } else {
    // ------------------------------------------------------------------ MMA warp
    const unsigned int idesc_x = tc_idesc(kEncTile, 0, 0), idesc_w = tc_idesc(16, 0, 0);
    const unsigned int hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const unsigned int lo0 = (16u >> 4) << 16;
    const unsigned int w2t = __shfl_sync(0xffffffffu, smem_u32(W2T), 0);
    const unsigned int dz2 = __shfl_sync(0xffffffffu, smem_u32(DZ2), 0);
    const unsigned int ptb = __shfl_sync(0xffffffffu, smem_u32(PT), 0);
    auto issue_dx = [&](unsigned int u) {       // dh1^T[c, p] = sum_o W2^T[c, o] dz2[p, o] of tile u
      mbar_wait(&dz2_full[u & 1u], (u >> 1) & 1u);
      if (u >= 2) mbar_wait(&dh1_empty[u & 1u], ((u >> 1) - 1u) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const unsigned int a_lo = lo0 | ((w2t + ch * kEbChunk128) >> 4), b_lo = lo0 | ((dz2 + (u & 1u) * 4 * kEbChunk128 + ch * kEbChunk128) >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tc_mma(tmem + (u & 1u) * 128u, ((unsigned long long)hi << 32) | (a_lo + kk * 2), ((unsigned long long)hi << 32) | (b_lo + kk * 2),
                   idesc_x, (ch | kk) != 0 ? 1u : 0u);
        }
        tc_commit(&dz2_empty[u & 1u]);
        tc_commit(&dh1_full[u & 1u]);
      }
      __syncwarp();
    };
    issue_dx(0u);
    for (int t = 0; t < n_my; ++t) {
      const unsigned int ut = (unsigned)t;
      if (t + 1 < n_my) issue_dx(ut + 1u);
      mbar_wait(&dz1_full, ut & 1u);
      mbar_wait(&pt_full[ut & 1u], (ut >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const unsigned int b_lo = lo0 | ((ptb + (ut & 1u) * 4 * kEbChunk16 + ch * kEbChunk16) >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)      // A = 8 columns (particles) of dz1^T straight from tensor memory
            tc_mma_ts(tmem + kEbXDw1Col, tmem + kEbXDz1Col + (unsigned)(ch * 32 + kk * 8), ((unsigned long long)hi << 32) | (b_lo + kk * 2), idesc_w,
                      (ut | (unsigned)ch | (unsigned)kk) != 0 ? 1u : 0u);
        }
        tc_commit(&dz1_empty);
        tc_commit(&pt_empty[ut & 1u]);
        if (t + 1 == n_my) tc_commit(&done_bar);
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == kEbXThreads / 32 - 1) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace td3
