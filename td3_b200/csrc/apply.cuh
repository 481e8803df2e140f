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
constexpr int kDwMaxK = 32;      // input width of the layer (<= kFrontMaxK)
constexpr int kDwThreads = kEwThreads;

struct DwParams {
  int n_tiles;                   // 0: plain Adam/Polyak launch
  int batch, n_agents, n_inner, col_blocks;
  int N, K, ld_dz, ldx, do_polyak;
  const float* dz; long long dz_go, dz_gi;      // gradient w.r.t. the layer's pre-activation [B, N] (ReLU mask applied)
  const float* x; long long x_go, x_gi;         // the layer's input [B, ldx]
  float* p; float* g; float* m; float* v; float* tgt;    // packed buffers of the family
  long long p_go, p_gi, w_off, b_off;           // agent stride, twin stride, offsets of W_0 / b_0 inside a network
  const float* sc_ptr;                          // device-resident {step_size, sqrt(1 - beta2^t)} (adam_tick)
};

// 256 threads = 2 batch halves x 16 channels x 8 groups of 4 input columns.
__device__ __forceinline__ void dw_adam_body(const DwParams& D, const EwParams& E, int tile, float* smem) {
  float* dzs = smem;                              // [kDwRows][kDwCols]
  float* xs = smem + kDwRows * kDwCols;           // [kDwRows][kDwMaxK]
  const int tid = threadIdx.x;
  const int per_agent = D.n_inner * D.col_blocks;
  const int agent = tile / per_agent;
  const int rem = tile - agent * per_agent;
  const int inner = rem / D.col_blocks, c0 = (rem - inner * D.col_blocks) * kDwCols;
  const int half = tid >> 7, u = tid & 127, c = u >> 3, kq = u & 7;
  const int N = D.N, K = D.K;
  const bool c_ok = c0 + c < N;
  const long long net = (long long)agent * D.p_go + (long long)inner * D.p_gi;
  const long long wi = net + D.w_off + (long long)(c0 + c) * K + 4 * kq;
  const long long bi = net + D.b_off + c0 + c;
  const bool owner = half == 0 && c_ok;
  // ---- optimiser operands of this thread's parameters: issued first, consumed last ----
  float pv[5], mv[5], vv[5], tv[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const bool ok = owner && (j < 4 ? 4 * kq + j < K : kq == 0);
    const long long i = j < 4 ? wi + j : bi;
    pv[j] = ok ? __ldcg(D.p + i) : 0.f;
    mv[j] = ok ? __ldcg(D.m + i) : 0.f;
    vv[j] = ok ? __ldcg(D.v + i) : 0.f;
    tv[j] = ok && D.do_polyak ? __ldcg(D.tgt + i) : 0.f;
  }
  const float step_size = __ldcg(D.sc_ptr), bc2s = __ldcg(D.sc_ptr + 1);
  const float* dz = D.dz + (long long)agent * D.dz_go + (long long)inner * D.dz_gi + c0;
  const float* x = D.x + (long long)agent * D.x_go + (long long)inner * D.x_gi;
  const int kg = (K + 3) >> 2;                    // 16-byte granules per input row
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, accb = 0.f;
#pragma unroll 1
  for (int b0 = 0; b0 < D.batch; b0 += kDwRows) {
    const int rows = min(kDwRows, D.batch - b0);
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
    const int hr = (rows + 1) >> 1;
    const int rb = half * hr, re = min(rows, rb + hr);
    const float4* xs4 = reinterpret_cast<const float4*>(xs);
#pragma unroll 4
    for (int r = rb; r < re; ++r) {
      const float d = dzs[r * kDwCols + c];
      const float4 xv = xs4[r * (kDwMaxK / 4) + kq];
      acc[0] = fmaf(d, xv.x, acc[0]);
      acc[1] = fmaf(d, xv.y, acc[1]);
      acc[2] = fmaf(d, xv.z, acc[2]);
      acc[3] = fmaf(d, xv.w, acc[3]);
      accb += d;
    }
  }
  // ---- the two batch halves: second half hands over through shared memory, first half adds (fixed order) ----
  __syncthreads();
  float* hand = smem;                             // [128][5]
  if (half == 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) hand[u * 5 + j] = acc[j];
    hand[u * 5 + 4] = accb;
  }
  __syncthreads();
  if (!owner) return;
  float gv[5];
#pragma unroll
  for (int j = 0; j < 4; ++j) gv[j] = acc[j] + hand[u * 5 + j];
  gv[4] = accb + hand[u * 5 + 4];
  const float w1 = (float)(1.0 - E.beta1), b2 = (float)E.beta2, w2 = (float)(1.0 - E.beta2);
  const float eps = (float)E.eps, tau = (float)E.tau, omt = (float)(1.0 - E.tau);
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const bool ok = j < 4 ? 4 * kq + j < K : kq == 0;
    if (!ok) continue;
    const long long i = j < 4 ? wi + j : bi;
    const float pn = adam_element(pv[j], gv[j], mv[j], vv[j], w1, b2, w2, bc2s, eps, -step_size);
    D.g[i] = gv[j];
    D.m[i] = mv[j];
    D.v[i] = vv[j];
    D.p[i] = pn;
    if (D.do_polyak) D.tgt[i] = __fadd_rn(__fmul_rn(tau, pn), __fmul_rn(omt, tv[j]));
  }
}

constexpr int kDwSmemBytes = (kDwRows * kDwCols + kDwRows * kDwMaxK) * 4;

// blocks [0, n_tiles): first-layer gradient + step; the rest: element-wise Adam / Polyak over the family's other tensors
__global__ void __launch_bounds__(kDwThreads) apply_kernel(const __grid_constant__ EwParams E, const __grid_constant__ DwParams D) {
  __shared__ __align__(16) float apply_smem[kDwSmemBytes / 4];
  pdl_launch_dependents();
  pdl_wait();
  if ((int)blockIdx.x < D.n_tiles) dw_adam_body(D, E, blockIdx.x, apply_smem);
  else adam_polyak_body(E, (long long)blockIdx.x - D.n_tiles);
}

}  // namespace td3
