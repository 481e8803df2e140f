// enc.cuh -- K6: the fused particle-set encoder forward (TD3_particles.py:29-32,53-58 / :104-109):
//
//     h1 = relu(P W1^T + b1)          Conv2d(1, 256, (1, D))  == per-particle linear D -> 256        CUDA cores
//     h2 = relu(h1 W2^T + b2)         Conv1d(256, 128, 1)     == per-particle linear 256 -> 128      tcgen05 (TF32)
//     pooled[b] = mean_n h2[b, n]     AvgPool2d((1, N))                                              epilogue
//
// As separate stages the two activations ([B*N, 256] and [B*N, 128] floats: 268 + 134 MB at B = 256, N = 1024) are
// written to HBM and read back, and the layer-2 GEMM starts from a cold operand stream.  Here one persistent CTA per
// SM walks 128-particle tiles and nothing but the 3 KB of particles enters and 512 bytes of pooled partial sums leave:
//
//   layer 1            is ONE tcgen05.mma per tile (M128 N256 K8): the particle row padded to [p_0 .. p_{D-1}, 0, 1] against
//                      [W1 | 0 | b1], both rounded to nearest TF32, into a 256-column TMEM accumulator (first version: FFMA
//                      in the producer warps -- 16 k warp-instructions per tile, issue-bound at 15 k cycles per tile)
//   8 producer warps   tcgen05.ld that accumulator -> ReLU -> round to nearest TF32 -> the UMMA A-operand layout of layer 2 in
//                      shared memory (K-major, 128-byte swizzle), a quarter of the 256 reduction columns at a time
//   1 MMA warp         W2 (128 x 256 fp32 = 128 KB) is loaded ONCE per CTA by TMA and stays resident; an elected lane
//                      issues tcgen05.mma M128 N128 K8 over each quarter as it lands; two 128-column TMEM accumulators
//   4 epilogue warps   tcgen05.ld -> + b2, ReLU -> column sums over the tile's 128 rows (butterfly across the warp, then
//                      across the four warps in fixed order) while the next tile's layer 1 and MMAs are already running
//
// A sample's N particles are N / 128 consecutive tiles: the tile writes its partial MEAN (sum / 128) to
// part[tile][128]; the existing pooling tile (stage.cuh: pool_fwd_tile with K = N / 128 rows per sample) finishes
// relu(mean) in a fixed order -- deterministic, no atomics.  For the online networks (whose backward pass needs them)
// the kernel can also store h1 / h2; the target networks run with no activation traffic at all.
// Shapes: 256 hidden / 128 output channels (the reference's constants), D <= 8, rows % 128 == 0, N % 128 == 0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc.cuh"

namespace td3 {

constexpr int kEncH = 256, kEncO = 128, kEncTile = 128;
constexpr int kEncProducerWarps = 8, kEncEpiWarps = 4;
constexpr int kEncThreads = (kEncEpiWarps + kEncProducerWarps + 1) * 32;      // 416
constexpr int kEncMaxGroups = 4;
constexpr int kEncChunk = 16384;                                   // one 128 x 32 fp32 K-major operand chunk
constexpr int kEncW2Bytes = (kEncH / 32) * kEncChunk;              // 131072
constexpr int kEncQuarterBytes = 2 * kEncChunk;                    // 64 reduction columns of the A operand
constexpr int kEncW1Bytes = kEncH * 32, kEncPBytes = kEncTile * 32;   // K = 8 operands of layer 1: 32 bytes per row, no swizzle
constexpr int kEncSmemBytes = 1024 + kEncW2Bytes + 2 * kEncQuarterBytes + kEncW1Bytes + 2 * kEncPBytes +
                              kEncEpiWarps * kEncO * 4 + kEncO * 4 + 256;
constexpr unsigned int kEncTmemCols = 512;         // [0,128) [128,256): layer-2 accumulators; [256,512): layer-1 accumulator

// K-major operand without swizzle: 8-row x 16-byte core matrices; row r, 16-byte granule j of a K = 8 (32-byte) row
__device__ __forceinline__ int enc_k8_offset(int r, int j) { return (r >> 3) * 256 + j * 128 + (r & 7) * 16; }

struct EncParams {
  const float* P; long long p_go;                    // particles [rows, D] of outer group (agent) o at P + o * p_go
  const float* W1; const float* b1; const float* b2; // layer parameters of group 0; group (o, i) at + o * w_go + i * w_gi
  long long w_go, w_gi;
  float* h1; long long h1_go, h1_gi;                 // optional [rows, 256] (nullptr: not stored)
  float* h2; long long h2_go, h2_gi;                 // optional [rows, 128]
  float* part; long long part_go, part_gi;           // [rows / 128, 128] partial means
  // optional ReLU bitmaps for the fused backward (encbwd.cuh), one block of 16 * rows words per group:
  //   bits2  [rows][4]          bit o % 32 of word o / 32      = h2[p, o] > 0
  //   bits2T [rows / 128][128][4]   bit p % 32 of word p / 32  = h2[tile * 128 + p, o] > 0      (+ 4 * rows words)
  //   bits1T [rows / 128][256][4]   the same for h1[., c] > 0                                   (+ 8 * rows words)
  unsigned int* bits; long long bits_go, bits_gi;
  int rows, D, n_inner, n_groups, tiles_per_group, pad;
  TensorMapBlob w2_map[kEncMaxGroups];               // W2 [128, 256] of every group (box 32 k x 128 rows, 128-byte swizzle)
};

// column sums of a 32 (lanes = rows) x 32 (registers = columns) block: after the butterfly lane l holds the sum of column l
__device__ __forceinline__ float enc_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// 32 x 32 bit-matrix transpose across a warp: lane i holds row i (bit j = column j); returns column `lane` (bit i = row i).
// Five butterfly steps swap the off-diagonal blocks of size 16, 8, 4, 2, 1: 5 shuffles instead of 32 ballots.
__device__ __forceinline__ unsigned int enc_transpose32(unsigned int x, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const unsigned int m = s == 16 ? 0x0000ffffu : s == 8 ? 0x00ff00ffu : s == 4 ? 0x0f0f0f0fu : s == 2 ? 0x33333333u : 0x55555555u;
    const unsigned int y = __shfl_xor_sync(0xffffffffu, x, s);
    x = (lane & s) ? (((y & ~m) >> s) | (x & ~m)) : ((x & m) | ((y & m) << s));
  }
  return x;
}

__device__ __forceinline__ void enc_named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void enc_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// 32 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void enc_tmem_ld32(unsigned int taddr, unsigned int (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// the same load without the wait: the caller overlaps it with other work and calls enc_tmem_wait before reading r
__device__ __forceinline__ void enc_tmem_ld32_issue(unsigned int taddr, unsigned int (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void enc_tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

__global__ void __launch_bounds__(kEncThreads, 1) enc_fwd_kernel(const __grid_constant__ EncParams E) {
  extern __shared__ unsigned char enc_smem_raw[];
  __shared__ unsigned long long w2_bar, w2_free, a_full[2], a_empty[2], acc_full[2], acc_empty[2], p_full[2], h1_full, h1_empty;
  __shared__ unsigned int tmem_base_s;
  unsigned char* base = enc_smem_raw + ((1024u - (smem_u32(enc_smem_raw) & 1023u)) & 1023u);
  unsigned char* W2s = base;                                        // 8 chunks x 16 KB
  unsigned char* As = W2s + kEncW2Bytes;                            // 2 quarter buffers x 32 KB
  unsigned char* W1k = As + 2 * kEncQuarterBytes;                   // [256 rows][8]: W1 row, 0 padding, bias in slot 7
  unsigned char* Pk = W1k + kEncW1Bytes;                            // 2 x [128 rows][8]: particle, 0 padding, 1 in slot 7
  float* part_s = reinterpret_cast<float*>(Pk + 2 * kEncPBytes);    // [4 warps][128]
  float* b2s = part_s + kEncEpiWarps * kEncO;                       // [128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&w2_bar, 1);
    mbar_init(&w2_free, 1);
    mbar_init(&h1_full, 1);
    mbar_init(&h1_empty, kEncProducerWarps);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], kEncProducerWarps);
      mbar_init(&a_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEncEpiWarps);
      mbar_init(&p_full[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == kEncEpiWarps + kEncProducerWarps) {                   // the MMA warp owns the TMEM allocation (all 512 columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(kEncTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned int tmem = *reinterpret_cast<volatile unsigned int*>(&tmem_base_s);

  // this CTA's contiguous run of tiles (group-major order: at most one change of network per CTA)
  const int total = E.n_groups * E.tiles_per_group;
  const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int t_begin = (int)blockIdx.x * per, t_end = min(total, t_begin + per);
  const int D = E.D;

  if (warp < kEncEpiWarps) {
    // ------------------------------------------------------------------ epilogue warps
    const int e = warp;
    int cur_g = -1;
    unsigned int tcount = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++tcount) {
      const int g = tile / E.tiles_per_group, rt = tile - g * E.tiles_per_group;
      const int go = g / E.n_inner, gi = g - go * E.n_inner;
      if (g != cur_g) {                                             // this network's output bias
        enc_named_barrier(1, kEncEpiWarps * 32);
        b2s[tid] = E.b2[(long long)go * E.w_go + (long long)gi * E.w_gi + tid];
        enc_named_barrier(1, kEncEpiWarps * 32);
        cur_g = g;
      }
      const unsigned int acc = tcount & 1u, ause = tcount >> 1;
      mbar_wait(&acc_full[acc], ause & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const long long row = (long long)rt * kEncTile + e * 32 + lane;
      float* h2row = E.h2 ? E.h2 + (long long)go * E.h2_go + (long long)gi * E.h2_gi + row * kEncO : nullptr;
      unsigned int* bits2 = E.bits ? E.bits + (long long)go * E.bits_go + (long long)gi * E.bits_gi : nullptr;
#pragma unroll 1
      for (int pass = 0; pass < 4; ++pass) {
        unsigned int r[32];
        enc_tmem_ld32(tmem + acc * 128u + (unsigned)(pass * 32) + (((unsigned)e * 32u) << 16), r);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(__uint_as_float(r[j]) + b2s[pass * 32 + j], 0.f);
        if (h2row) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(h2row + pass * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (bits2) {                                                // sign bits of this thread's particle, and -- transposed across
          unsigned int nat = 0;                                     // the warp -- of its 32 particles for channel pass * 32 + lane
#pragma unroll
          for (int j = 0; j < 32; ++j) nat |= v[j] > 0.f ? (1u << j) : 0u;
          bits2[row * 4 + pass] = nat;
          bits2[(long long)E.rows * 4 + ((long long)rt * kEncO + pass * 32 + lane) * 4 + e] = enc_transpose32(nat, lane);
        }
        part_s[e * kEncO + pass * 32 + lane] = enc_colsum32(v, lane);
      }
      // the accumulator may be overwritten by the MMAs of the tile after next
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) enc_arrive(&acc_empty[acc]);
      enc_named_barrier(1, kEncEpiWarps * 32);
      {
        const float tot = ((part_s[tid] + part_s[kEncO + tid]) + part_s[2 * kEncO + tid]) + part_s[3 * kEncO + tid];
        E.part[(long long)go * E.part_go + (long long)gi * E.part_gi + (long long)rt * kEncO + tid] = tot * (1.f / (float)kEncTile);
      }
      enc_named_barrier(1, kEncEpiWarps * 32);
    }
  } else if (warp < kEncEpiWarps + kEncProducerWarps) {
    // ------------------------------------------------------------------ producer warps: stage P, move relu(h1) TMEM -> UMMA A operand
    const int pw = warp - kEncEpiWarps, ptid = tid - kEncEpiWarps * 32;     // 0..255
    const int lq = pw & 3, half = pw >> 2;                          // TMEM lane quarter (== warp % 4), which chunk of a quarter
    const int row = lq * 32 + lane;                                 // this thread's particle of the tile
    int cur_g = -1;
    unsigned int qcount = 0, tcount = 0;
    // the K = 8 operand row of particle `ptid` of tile t -> Pk[t & 1] (threads 0..127: the four warps with half == 0)
    auto stage_particles = [&](int tile, unsigned int n) {
      if (ptid < kEncTile) {
        const int g = tile / E.tiles_per_group, rt = tile - g * E.tiles_per_group;
        const int go = g / E.n_inner;
        const float* Pg = E.P + (long long)go * E.p_go + ((long long)rt * kEncTile + ptid) * D;
        float p[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) p[d] = d < D ? rn_tf32(__ldg(Pg + d)) : (d == 7 ? 1.f : 0.f);
        unsigned char* dst = Pk + (n & 1u) * kEncPBytes;
        *reinterpret_cast<float4*>(dst + enc_k8_offset(ptid, 0)) = make_float4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<float4*>(dst + enc_k8_offset(ptid, 1)) = make_float4(p[4], p[5], p[6], p[7]);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) enc_arrive(&p_full[n & 1u]);
      }
    };
    if (t_begin < t_end) {
      // first network's layer-1 operand must be in place before the first p_full arrival releases MMA 1
      const int g = t_begin / E.tiles_per_group, go = g / E.n_inner, gi = g - go * E.n_inner;
      const float* W1 = E.W1 + (long long)go * E.w_go + (long long)gi * E.w_gi;
      const float* b1 = E.b1 + (long long)go * E.w_go + (long long)gi * E.w_gi;
      float w[8];
#pragma unroll
      for (int d = 0; d < 8; ++d) w[d] = d < D ? rn_tf32(W1[(long long)ptid * D + d]) : (d == 7 ? rn_tf32(b1[ptid]) : 0.f);
      *reinterpret_cast<float4*>(W1k + enc_k8_offset(ptid, 0)) = make_float4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<float4*>(W1k + enc_k8_offset(ptid, 1)) = make_float4(w[4], w[5], w[6], w[7]);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      enc_named_barrier(2, kEncProducerWarps * 32);
      cur_g = g;
      stage_particles(t_begin, 0);
    }
    for (int tile = t_begin; tile < t_end; ++tile, ++tcount) {
      const int g = tile / E.tiles_per_group, rt = tile - g * E.tiles_per_group;
      const int go = g / E.n_inner, gi = g - go * E.n_inner;
      // the next tile's particles go in now, so that its layer-1 MMA can be issued the moment this tile's h1 has been read;
      // a change of network first waits for that point itself (W1k is rewritten)
      const bool next_same = tile + 1 < t_end && (tile + 1) / E.tiles_per_group == g;
      if (next_same) stage_particles(tile + 1, tcount + 1);
      mbar_wait(&h1_full, tcount & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      float* h1g = E.h1 ? E.h1 + (long long)go * E.h1_go + (long long)gi * E.h1_gi + ((long long)rt * kEncTile + row) * kEncH : nullptr;
      unsigned int* bits1T = E.bits ? E.bits + (long long)go * E.bits_go + (long long)gi * E.bits_gi + (long long)E.rows * 8 : nullptr;
#pragma unroll 1
      for (int q = 0; q < 4; ++q, ++qcount) {
        const unsigned int buf = qcount & 1u, use = qcount >> 1;
        const int chunk = 2 * q + half;                             // 32 hidden channels
        unsigned int r[32];
        enc_tmem_ld32(tmem + 256u + (unsigned)(chunk * 32) + (((unsigned)lq * 32u) << 16), r);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rn_tf32_finite(fmaxf(__uint_as_float(r[j]), 0.f));
        if (use > 0) mbar_wait(&a_empty[buf], (use - 1u) & 1u);     // the MMAs that read this buffer have completed
        // operand layout: chunk = 32 reduction columns; row r at r * 128 B, 16-byte granule j at (j ^ (r & 7))
        unsigned char* arow = As + buf * kEncQuarterBytes + half * kEncChunk + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(arow + ((j ^ (row & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // generic-proxy stores -> visible to the tensor core
        __syncwarp();
        if (lane == 0) enc_arrive(&a_full[buf]);
        if (bits1T) {                                               // h1 > 0 of the warp's 32 particles, per hidden channel
          unsigned int nat = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) nat |= v[j] > 0.f ? (1u << j) : 0u;
          bits1T[((long long)rt * kEncH + chunk * 32 + lane) * 4 + lq] = enc_transpose32(nat, lane);
        }
        if (h1g) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(h1g + chunk * 32 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      // this warp is done reading the layer-1 accumulator
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) enc_arrive(&h1_empty);
      if (!next_same && tile + 1 < t_end) {                         // next tile belongs to another network: new W1 operand, then its particles
        const int g2 = (tile + 1) / E.tiles_per_group, go2 = g2 / E.n_inner, gi2 = g2 - go2 * E.n_inner;
        // MMA 1 of this tile has completed (h1_full was observed), so W1k is free
        enc_named_barrier(2, kEncProducerWarps * 32);
        const float* W1 = E.W1 + (long long)go2 * E.w_go + (long long)gi2 * E.w_gi;
        const float* b1 = E.b1 + (long long)go2 * E.w_go + (long long)gi2 * E.w_gi;
        float w[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) w[d] = d < D ? rn_tf32(W1[(long long)ptid * D + d]) : (d == 7 ? rn_tf32(b1[ptid]) : 0.f);
        *reinterpret_cast<float4*>(W1k + enc_k8_offset(ptid, 0)) = make_float4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<float4*>(W1k + enc_k8_offset(ptid, 1)) = make_float4(w[4], w[5], w[6], w[7]);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        enc_named_barrier(2, kEncProducerWarps * 32);
        cur_g = g2;
        stage_particles(tile + 1, tcount + 1);
      }
    }
    (void)cur_g;
  } else {
    // ------------------------------------------------------------------ MMA warp (+ the one-off W2 load per network)
    const unsigned int idesc2 = tc_idesc(kEncO, 0, 0), idesc1 = tc_idesc(kEncH, 0, 0);
    const unsigned int hi = (1024u >> 4) | (1u << 14) | (2u << 29);        // SBO 1024 B, descriptor version, SWIZZLE_128B
    const unsigned int lo0 = (16u >> 4) << 16;
    const unsigned int hi1 = (256u >> 4) | (1u << 14);                     // layer-1 operands: SBO 256 B, no swizzle
    const unsigned int lo1 = (128u >> 4) << 16;                            //                   LBO 128 B
    int cur_g = -1;
    unsigned int qcount = 0, tcount = 0, w2_loads = 0, w2_frees = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++tcount) {
      const int g = tile / E.tiles_per_group;
      if (g != cur_g) {
        if (cur_g >= 0) {                                           // the previous network's MMAs must be done with W2
          if (elect_one()) tc_commit(&w2_free);
          __syncwarp();
          mbar_wait(&w2_free, w2_frees & 1u);
          ++w2_frees;
        }
        if (elect_one()) {
          mbar_expect_tx(&w2_bar, (unsigned)kEncW2Bytes);
          for (int ch = 0; ch < kEncH / 32; ++ch) tma_load_2d(smem_u32(W2s + ch * kEncChunk), &E.w2_map[g], ch * 32, 0, &w2_bar);
        }
        __syncwarp();
        mbar_wait(&w2_bar, w2_loads & 1u);
        ++w2_loads;
        cur_g = g;
      }
      // ---- layer 1: one MMA, [128 x 8] . [256 x 8]^T -> TMEM columns [256, 512) ----
      mbar_wait(&p_full[tcount & 1u], (tcount >> 1) & 1u);
      if (tcount > 0) mbar_wait(&h1_empty, (tcount - 1u) & 1u);    // the producers have read the previous tile's h1
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      {
        const unsigned int pa = __shfl_sync(0xffffffffu, smem_u32(Pk + (tcount & 1u) * kEncPBytes), 0);
        const unsigned int wa = __shfl_sync(0xffffffffu, smem_u32(W1k), 0);
        if (elect_one()) {
          tc_mma(tmem + 256u, ((unsigned long long)hi1 << 32) | (lo1 | (pa >> 4)), ((unsigned long long)hi1 << 32) | (lo1 | (wa >> 4)), idesc1, 0u);
          tc_commit(&h1_full);
        }
        __syncwarp();
      }
      const unsigned int acc = tcount & 1u, ause = tcount >> 1;
      if (ause > 0) mbar_wait(&acc_empty[acc], (ause - 1u) & 1u);   // the epilogue has read this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const unsigned int d_tmem = tmem + acc * 128u;
#pragma unroll 1
      for (int q = 0; q < 4; ++q, ++qcount) {
        const unsigned int buf = qcount & 1u, use = qcount >> 1;
        mbar_wait(&a_full[buf], use & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        // warp-uniform operand addresses (UTCHMMA takes uniform registers: see tc.cuh)
        const unsigned int a_base = __shfl_sync(0xffffffffu, smem_u32(As + buf * kEncQuarterBytes), 0);
        const unsigned int b_base = __shfl_sync(0xffffffffu, smem_u32(W2s + (q * 2) * kEncChunk), 0);
        const unsigned int uq = __shfl_sync(0xffffffffu, (unsigned)q, 0);
        if (elect_one()) {
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const unsigned int a_lo = lo0 | ((a_base + cc * kEncChunk) >> 4), b_lo = lo0 | ((b_base + cc * kEncChunk) >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              tc_mma(d_tmem, ((unsigned long long)hi << 32) | (a_lo + kk * 2), ((unsigned long long)hi << 32) | (b_lo + kk * 2), idesc2,
                     (uq | (unsigned)cc | (unsigned)kk) != 0 ? 1u : 0u);
          }
          tc_commit(&a_empty[buf]);
          if (uq == 3) tc_commit(&acc_full[acc]);
        }
        __syncwarp();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == kEncEpiWarps + kEncProducerWarps) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(kEncTmemCols) : "memory");
  }
}

}  // namespace td3
