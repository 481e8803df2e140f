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
//   8 producer warps   layer 1 on FFMA straight into the UMMA operand layout in shared memory (K-major, 128-byte
//                      swizzle, values rounded to nearest TF32), a quarter of the 256 reduction columns at a time
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
constexpr int kEncSmemBytes = 1024 + kEncW2Bytes + 2 * kEncQuarterBytes + kEncH * 8 * 4 + kEncProducerWarps * 16 * 8 * 4 +
                              kEncEpiWarps * kEncO * 4 + kEncO * 4 + 256;

struct EncParams {
  const float* P; long long p_go;                    // particles [rows, D] of outer group (agent) o at P + o * p_go
  const float* W1; const float* b1; const float* b2; // layer parameters of group 0; group (o, i) at + o * w_go + i * w_gi
  long long w_go, w_gi;
  float* h1; long long h1_go, h1_gi;                 // optional [rows, 256] (nullptr: not stored)
  float* h2; long long h2_go, h2_gi;                 // optional [rows, 128]
  float* part; long long part_go, part_gi;           // [rows / 128, 128] partial means
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

__device__ __forceinline__ void enc_named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void enc_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kEncThreads, 1) enc_fwd_kernel(const __grid_constant__ EncParams E) {
  extern __shared__ unsigned char enc_smem_raw[];
  __shared__ unsigned long long w2_bar, w2_free, a_full[2], a_empty[2], acc_full[2], acc_empty[2];
  __shared__ unsigned int tmem_base_s;
  unsigned char* base = enc_smem_raw + ((1024u - (smem_u32(enc_smem_raw) & 1023u)) & 1023u);
  unsigned char* W2s = base;                                        // 8 chunks x 16 KB
  unsigned char* As = W2s + kEncW2Bytes;                            // 2 quarter buffers x 32 KB
  float* W1s = reinterpret_cast<float*>(As + 2 * kEncQuarterBytes); // [256][8]: 6 (<= 7) weights, then the bias in slot 7
  float* Ps = W1s + kEncH * 8;                                      // [8 warps][16 rows][8]
  float* part_s = Ps + kEncProducerWarps * 16 * 8;                  // [4 warps][128]
  float* b2s = part_s + kEncEpiWarps * kEncO;                       // [128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&w2_bar, 1);
    mbar_init(&w2_free, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], kEncProducerWarps);
      mbar_init(&a_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEncEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == kEncEpiWarps + kEncProducerWarps) {                   // the MMA warp owns the TMEM allocation: 2 x 128 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(256) : "memory");
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
#pragma unroll 1
      for (int pass = 0; pass < 4; ++pass) {
        unsigned int r[32];
        const unsigned int taddr = tmem + acc * 128u + (unsigned)(pass * 32) + (((unsigned)e * 32u) << 16);
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
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(__uint_as_float(r[j]) + b2s[pass * 32 + j], 0.f);
        if (h2row) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(h2row + pass * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
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
    // ------------------------------------------------------------------ producer warps: layer 1 -> UMMA A operand
    const int pw = warp - kEncEpiWarps, ptid = tid - kEncEpiWarps * 32;     // 0..255
    const int r0 = pw * 16;
    float* myP = Ps + pw * 16 * 8;
    int cur_g = -1;
    unsigned int qcount = 0;
    for (int tile = t_begin; tile < t_end; ++tile) {
      const int g = tile / E.tiles_per_group, rt = tile - g * E.tiles_per_group;
      const int go = g / E.n_inner, gi = g - go * E.n_inner;
      if (g != cur_g) {                                             // this network's first-layer weights -> [c][8]
        enc_named_barrier(2, kEncProducerWarps * 32);
        const float* W1 = E.W1 + (long long)go * E.w_go + (long long)gi * E.w_gi;
        const float* b1 = E.b1 + (long long)go * E.w_go + (long long)gi * E.w_gi;
        for (int c = ptid; c < kEncH; c += kEncProducerWarps * 32) {
#pragma unroll
          for (int d = 0; d < 7; ++d) W1s[c * 8 + d] = d < D ? W1[(long long)c * D + d] : 0.f;
          W1s[c * 8 + 7] = b1[c];
        }
        enc_named_barrier(2, kEncProducerWarps * 32);
        cur_g = g;
      }
      // this warp's 16 particles: rows r0 .. r0 + 15 of the tile, padded to 8 floats (slot 7 = 1 multiplies the bias)
      {
        const float* Pg = E.P + (long long)go * E.p_go + ((long long)rt * kEncTile + r0) * D;
        __syncwarp();
        for (int i = lane; i < 16 * 8; i += 32) {
          const int rr = i >> 3, d = i & 7;
          myP[i] = d < D ? __ldg(Pg + rr * D + d) : (d == 7 ? 1.f : 0.f);
        }
        __syncwarp();
      }
      float* h1g = E.h1 ? E.h1 + (long long)go * E.h1_go + (long long)gi * E.h1_gi + ((long long)rt * kEncTile + r0) * kEncH : nullptr;
#pragma unroll 1
      for (int q = 0; q < 4; ++q, ++qcount) {
        const unsigned int buf = qcount & 1u, use = qcount >> 1;
        if (use > 0) mbar_wait(&a_empty[buf], (use - 1u) & 1u);     // the MMAs that read this buffer have completed
        const int c = q * 64 + 2 * lane;                            // this lane's two hidden channels
        const float4 wa0 = *reinterpret_cast<const float4*>(W1s + c * 8), wa1 = *reinterpret_cast<const float4*>(W1s + c * 8 + 4);
        const float4 wb0 = *reinterpret_cast<const float4*>(W1s + c * 8 + 8), wb1 = *reinterpret_cast<const float4*>(W1s + c * 8 + 12);
        // operand layout: chunk = 32 reduction columns; row r of a chunk at r * 128 B, 16-byte granule j at (j ^ (r & 7))
        unsigned char* abuf = As + buf * kEncQuarterBytes + (lane >> 4) * kEncChunk;
        const int gran = (lane & 15) >> 1, sub = (lane & 1) * 8;
#pragma unroll 4
        for (int rr = 0; rr < 16; ++rr) {
          const float4 p0 = *reinterpret_cast<const float4*>(myP + rr * 8), p1 = *reinterpret_cast<const float4*>(myP + rr * 8 + 4);
          float v0 = p0.x * wa0.x;
          v0 = fmaf(p0.y, wa0.y, v0); v0 = fmaf(p0.z, wa0.z, v0); v0 = fmaf(p0.w, wa0.w, v0);
          v0 = fmaf(p1.x, wa1.x, v0); v0 = fmaf(p1.y, wa1.y, v0); v0 = fmaf(p1.z, wa1.z, v0); v0 = fmaf(p1.w, wa1.w, v0);
          float v1 = p0.x * wb0.x;
          v1 = fmaf(p0.y, wb0.y, v1); v1 = fmaf(p0.z, wb0.z, v1); v1 = fmaf(p0.w, wb0.w, v1);
          v1 = fmaf(p1.x, wb1.x, v1); v1 = fmaf(p1.y, wb1.y, v1); v1 = fmaf(p1.z, wb1.z, v1); v1 = fmaf(p1.w, wb1.w, v1);
          v0 = rn_tf32(fmaxf(v0, 0.f));
          v1 = rn_tf32(fmaxf(v1, 0.f));
          const int row = r0 + rr;
          *reinterpret_cast<float2*>(abuf + row * 128 + ((gran ^ (row & 7)) << 4) + sub) = make_float2(v0, v1);
          if (h1g) *reinterpret_cast<float2*>(h1g + (long long)rr * kEncH + c) = make_float2(v0, v1);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // generic-proxy stores -> visible to the tensor core
        __syncwarp();
        if (lane == 0) enc_arrive(&a_full[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA warp (+ the one-off W2 load per network)
    const unsigned int idesc = tc_idesc(kEncO, 0, 0);
    const unsigned int hi = (1024u >> 4) | (1u << 14) | (2u << 29);        // SBO 1024 B, descriptor version, SWIZZLE_128B
    const unsigned int lo0 = (16u >> 4) << 16;
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
      const unsigned int acc = tcount & 1u, ause = tcount >> 1;
      if (ause > 0) mbar_wait(&acc_empty[acc], (ause - 1u) & 1u);   // the epilogue has read this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const unsigned int d_tmem = tmem + acc * 128u;
#pragma unroll 1
      for (int q = 0; q < 4; ++q, ++qcount) {
        const unsigned int buf = qcount & 1u, use = qcount >> 1;
        mbar_wait(&a_full[buf], use & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        if (elect_one()) {
          const unsigned int a_base = smem_u32(As + buf * kEncQuarterBytes), b_base = smem_u32(W2s + (q * 2) * kEncChunk);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const unsigned int a_lo = lo0 | ((a_base + cc * kEncChunk) >> 4), b_lo = lo0 | ((b_base + cc * kEncChunk) >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              tc_mma(d_tmem, ((unsigned long long)hi << 32) | (a_lo + kk * 2), ((unsigned long long)hi << 32) | (b_lo + kk * 2), idesc,
                     (q | cc | kk) != 0 ? 1u : 0u);
          }
          tc_commit(&a_empty[buf]);
          if (q == 3) tc_commit(&acc_full[acc]);
        }
        __syncwarp();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == kEncEpiWarps + kEncProducerWarps) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(256) : "memory");
  }
}

}  // namespace td3
