// td3_b200 -- fused "apply" launch: first-layer weight gradient + optimiser step in one kernel.
//
// The last link of a backward pass is dW_0 = dz_0^T x (and db_0 = sum_b dz_0): the reduction runs over the batch
// and the output is only [w_1, S + A] (400 x 23 for the 400-300 critic).  As a stage of its own it costs a full
// node of the update's dependency chain (~7 us on B200) for 2 MFLOP, and the Adam launch that follows waits for it.
// Here the tiles that compute dW_0 / db_0 apply torch's Adam (and the Polyak average on policy steps) to exactly the
// parameters they produced the gradient for, inside the same launch as the element-wise Adam/Polyak blocks that
// cover the rest of the family (whose gradients are complete one stage earlier):
//     critic_optimizer.step()   TD3_featured.py:151-153      actor_optimizer.step() + soft update   :162-171
// Arithmetic of the step is adam_element() / the Polyak expression of misc.cuh, unchanged; the gradient is also
// stored to the packed grad buffer so the observable state is the same as with the unfused sequence.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "misc.cuh"
#include "stage.cuh"

namespace td3 {

constexpr int kDwCols = 16;      // output channels (rows of W_0) per tile
constexpr int kDwRows = 256;     // batch rows staged per pass
constexpr int kDwRowsSmall = 128;
constexpr int kDwMaxK = 32;      // input width of the layer (<= kFrontMaxK)
constexpr int kDwThreads = kEwThreads;

struct DwParams {
  int n_tiles;                   // 0: plain Adam/Polyak launch
  int batch, n_agents, n_inner, col_blocks;
  int N, K, ld_dz, ldx, do_polyak;
  int rows_per_pass, pad_d;      // batch rows staged per pass: kDwRows (one agent: shortest chain) or kDwRowsSmall (populations:
                                 // half the shared memory per block doubles the resident element-wise blocks of the launch)
  const float* dz; long long dz_go, dz_gi;      // gradient w.r.t. the layer's pre-activation [B, N] (ReLU mask applied)
  const float* x; long long x_go, x_gi;         // the layer's input [B, ldx]
  float* p; float* g; float* m; float* v; float* tgt;    // packed buffers of the family
  float* p_sh; float* tgt_sh;                    // optional TF32-rounded shadows of p / tgt (tensor-core operands)
  long long p_go, p_gi, w_off, b_off;           // agent stride, twin stride, offsets of W_0 / b_0 inside a network
  const float* sc_ptr;                          // device-resident {step_size, sqrt(1 - beta2^t)} (adam_tick)
};

// 256 threads = 4 batch slices x 8 channel pairs x 8 groups of 4 input columns: a thread accumulates a 2 x 4 block of
// dW_0 (+ the two bias sums) over its slice, so one 8-byte and one 16-byte shared-memory load feed 8 FMAs.  The slices
// meet in shared memory (summed in slice order), and the optimiser step then walks the tile's parameters -- 16 whole
// rows of W_0, i.e. one contiguous run of 16 K floats, plus 16 biases -- with consecutive threads on consecutive
// addresses; their p / m / v / target values were requested before the first operand was staged.
constexpr int kDwSlices = 4, kDwPartLd = kDwMaxK + 4;      // partial tile [slice][16][36]: 32 dW columns + the bias sum + pad
static_assert(kDwSlices * kDwCols * kDwPartLd <= kDwRowsSmall * (kDwCols + kDwMaxK), "slice partials reuse the staging buffers");
constexpr int kDwElems = 2;                                  // optimiser elements per thread: 16 * (32 + 1) <= 2 * 256 + 16

__device__ __forceinline__ void dw_adam_body(const DwParams& D, const EwParams& E, int tile, float* smem) {
  const int RP = D.rows_per_pass;
  float* dzs = smem;                              // [RP][kDwCols]
  float* xs = smem + RP * kDwCols;                // [RP][kDwMaxK]
  const int tid = threadIdx.x;
  const int per_agent = D.n_inner * D.col_blocks;
  const int agent = tile / per_agent;
  const int rem = tile - agent * per_agent;
  const int inner = rem / D.col_blocks, c0 = (rem - inner * D.col_blocks) * kDwCols;
  const int slice = tid >> 6, u = tid & 63, cp = u >> 3, kq = u & 7;
  const int N = D.N, K = D.K;
  const int nch = min(kDwCols, N - c0);           // channels of this tile
  const long long net = (long long)agent * D.p_go + (long long)inner * D.p_gi;
  const int n_w = nch * K, n_all = n_w + nch;     // the tile's parameters: W_0 rows c0.. (contiguous), then biases
  // ---- optimiser operands: element e of the tile for e = tid, tid + 256 (+ 512 only when K > 30) ----
  float pv[kDwElems + 1], mv[kDwElems + 1], vv[kDwElems + 1], tv[kDwElems + 1];
  long long pi[kDwElems + 1];
#pragma unroll
  for (int j = 0; j <= kDwElems; ++j) {
    const int e = tid + j * kDwThreads;
    const bool ok = e < n_all;
    pi[j] = e < n_w ? net + D.w_off + (long long)c0 * K + e : net + D.b_off + c0 + (e - n_w);
    pv[j] = ok ? __ldcg(D.p + pi[j]) : 0.f;
    mv[j] = ok ? __ldcg(D.m + pi[j]) : 0.f;
    vv[j] = ok ? __ldcg(D.v + pi[j]) : 0.f;
    tv[j] = ok && D.do_polyak ? __ldcg(D.tgt + pi[j]) : 0.f;
  }
  const float step_size = __ldcg(D.sc_ptr), bc2s = __ldcg(D.sc_ptr + 1);
  const float* dz = D.dz + (long long)agent * D.dz_go + (long long)inner * D.dz_gi + c0;
  const float* x = D.x + (long long)agent * D.x_go + (long long)inner * D.x_gi;
  const int kg = (K + 3) >> 2;                    // 16-byte granules per input row
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, accb[2] = {0.f, 0.f};
#pragma unroll 1
  for (int b0 = 0; b0 < D.batch; b0 += RP) {
    const int rows = min(RP, D.batch - b0);
    if (b0 > 0) __syncthreads();
#pragma unroll 1
    for (int i = tid; i < rows * (kDwCols / 4); i += kDwThreads) {
      const int r = i >> 2, q = i & 3;
      cp_async16(dzs + r * kDwCols + 4 * q, dz + (long long)(b0 + r) * D.ld_dz + 4 * q, c0 + 4 * q < N);
    }
#pragma unroll 1
    for (int i = tid; i < rows * 8; i += kDwThreads) {
      const int r = i >> 3, q = i & 7;
      cp_async16(xs + r * kDwMaxK + 4 * q, x + (long long)(b0 + r) * D.ldx + 4 * q, q < kg);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int sr = (rows + kDwSlices - 1) / kDwSlices;
    const int rb = slice * sr, re = min(rows, rb + sr);
    const float4* xs4 = reinterpret_cast<const float4*>(xs);
    const float2* dz2 = reinterpret_cast<const float2*>(dzs);
#pragma unroll 4
    for (int r = rb; r < re; ++r) {
      const float2 d = dz2[r * (kDwCols / 2) + cp];
      const float4 xv = xs4[r * (kDwMaxK / 4) + kq];
      acc[0][0] = fmaf(d.x, xv.x, acc[0][0]); acc[0][1] = fmaf(d.x, xv.y, acc[0][1]);
      acc[0][2] = fmaf(d.x, xv.z, acc[0][2]); acc[0][3] = fmaf(d.x, xv.w, acc[0][3]);
      acc[1][0] = fmaf(d.y, xv.x, acc[1][0]); acc[1][1] = fmaf(d.y, xv.y, acc[1][1]);
      acc[1][2] = fmaf(d.y, xv.z, acc[1][2]); acc[1][3] = fmaf(d.y, xv.w, acc[1][3]);
      accb[0] += d.x; accb[1] += d.y;
    }
  }
  // ---- slice partials -> shared memory; summed in slice order by whoever owns the parameter ----
  __syncthreads();
  float* part = smem;                             // [kDwSlices][kDwCols][kDwPartLd]
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float* row = part + (slice * kDwCols + 2 * cp + h) * kDwPartLd;
    *reinterpret_cast<float4*>(row + 4 * kq) = make_float4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
    if (kq == 0) row[kDwMaxK] = accb[h];
  }
  __syncthreads();
  const float w1 = (float)(1.0 - E.beta1), b2 = (float)E.beta2, w2 = (float)(1.0 - E.beta2);
  const float eps = (float)E.eps, tau = (float)E.tau, omt = (float)(1.0 - E.tau);
#pragma unroll
  for (int j = 0; j <= kDwElems; ++j) {
    const int e = tid + j * kDwThreads;
    if (e >= n_all) break;
    int c, k;
    if (e < n_w) { c = e / K; k = e - c * K; }
    else { c = e - n_w; k = kDwMaxK; }
    float g = part[c * kDwPartLd + k];
#pragma unroll
    for (int sl = 1; sl < kDwSlices; ++sl) g += part[(sl * kDwCols + c) * kDwPartLd + k];
    const float pn = adam_element(pv[j], g, mv[j], vv[j], w1, b2, w2, bc2s, eps, -step_size);
    D.g[pi[j]] = g;
    D.m[pi[j]] = mv[j];
    D.v[pi[j]] = vv[j];
    D.p[pi[j]] = pn;
    if (D.p_sh) D.p_sh[pi[j]] = rn_tf32(pn);
    if (D.do_polyak) {
      const float tn = __fadd_rn(__fmul_rn(tau, pn), __fmul_rn(omt, tv[j]));
      D.tgt[pi[j]] = tn;
      if (D.tgt_sh) D.tgt_sh[pi[j]] = rn_tf32(tn);
    }
  }
}

constexpr int dw_smem_bytes(int rows_per_pass) { return rows_per_pass * (kDwCols + kDwMaxK) * 4; }

// blocks [0, n_tiles): first-layer gradient + step; the rest: element-wise Adam / Polyak over the family's other tensors
__global__ void __launch_bounds__(kDwThreads) apply_kernel(const __grid_constant__ EwParams E, const __grid_constant__ DwParams D) {
  extern __shared__ __align__(16) float apply_smem[];
  pdl_launch_dependents();
  pdl_wait();
  if ((int)blockIdx.x < D.n_tiles) dw_adam_body(D, E, blockIdx.x, apply_smem);
  else adam_polyak_body(E, (long long)blockIdx.x - D.n_tiles);
}

}  // namespace td3
