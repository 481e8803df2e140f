// td3_b200 -- row-local "front" kernel.
//
// Several links of the TD3 update's dependency chain touch only ONE batch row at a time and have a reduction
// that is either tiny (the S+A <= 32 input columns of a first layer) or produces only A <= 8 numbers per row (the
// actor's output layer).  As separate launches each of them costs a full node of the chain (about 7 us on B200)
// for a few hundred FMAs per row.  This kernel chains them inside one CTA:
//
//   [sample]  draw the replay index of each row and read the transition            (my_replay_buffer.py:58-69)
//   [head]    a[b, :A] = epi(h[b, :] . Wh^T + bh)                                   e.g. actor_target's last layer
//                                                                                   + smoothing noise + clamp
//                                                                                   (TD3_featured.py:131-138)
//   [layers]  out_n[b, :] = relu(x_n[b, :K] . W_n^T + b_n)  for up to 4 networks    e.g. both target critics'
//                                                                                   first layer on [s', a']
//
// and, in the actor's backward pass, the mirrored pair  da = (dz1 . W1[:, S:S+A]) * tanh'  ->  dz2 = (da . W3) * relu'.
// Weight element (j, k) of the head is Wh[j*hs_j + k*hs_k] and element (c, k) of a layer is W[c*ws_c + k*ws_k], so
// both orientations are the same code.
//
// Work split: a tile is (agent, block of kFrontRows rows, job); a job is a block of 128 output columns of one network.
// Every tile stages its rows (and recomputes the head for them: kFrontRows x A dot products, far cheaper than a
// launch); tile job 0 of a row block also writes what later stages read from global memory (the gathered batch,
// the action, tanh(y)).  Accumulation is fp32 FFMA in every precision mode.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "misc.cuh"
#include "stage.cuh"

namespace td3 {

constexpr int kFrontRows = 8, kFrontCols = 128, kFrontMaxK = 32, kFrontMaxA = 8, kFrontMaxNets = 4;
constexpr int kFrontRowsWide = 32;                  // rows per tile when there are thousands of rows (populations, large batches):
                                                    // the tile's weight block is staged once per 32 rows instead of once per 8
constexpr int kFrontMaxKh = 512;                    // head reduction length (hidden width)
constexpr int kFrontXs = 72;                        // staged row: a transition prefix [s | a | s'] (<= 2*32 + 8 floats)
constexpr int kFrontWs = kFrontMaxK + 1;            // odd stride: column-per-thread reads are conflict-free
constexpr int kFrontWh = kFrontMaxKh + 4;
// shared memory: staged rows, weight block, bias strip, head weights, head input rows
constexpr int front_smem_bytes(int rows, bool head) {
  return (rows * kFrontXs + kFrontCols * kFrontWs + kFrontCols + (head ? kFrontMaxA * kFrontWh + rows * kFrontWh : 0)) * 4;
}
constexpr int kFrontSmemBytes = front_smem_bytes(kFrontRowsWide, true);      // the largest configuration (kernel attribute)

struct FrontNet {
  const float* x;            // input rows [B, ldx] (nullptr: the sampled transition, or the head output alone)
  long long x_go;
  int ldx, x_off;            // x_off: first input column inside the staged row
  const float* W; const float* bias;          // bias may be nullptr
  long long w_go, w_gi;
  int ws_c, ws_k;
  float* out; const float* mask;              // mask (same shape as out): out = v * (mask > 0) instead of relu(v + bias)
  long long out_go, out_gi, mask_go, mask_gi;
  int ldo, K, N, n_inner;
  int act_col;               // >= 0: input columns [act_col, act_col + A) are the head's output
  int job_begin, col_blocks;
  int rn_out, pad_n;         // out stored rounded to nearest TF32 (operand of a tensor-core contraction)
};

struct FrontParams {
  int gather, head, n_nets, jobs;
  int batch, n_agents, row_blocks, A;
  int rows_per_tile, pad_r;                   // kFrontRows (latency: one agent, batch 256) or kFrontRowsWide
  GatherParams g;
  // head
  const float* h; long long h_go; int ldh, Kh;
  const float* Wh; const float* bh; long long wh_go; int hs_j, hs_k;
  int head_epi, pad0;                         // EPI_BIAS_TANH_NOISE | EPI_BIAS_TANH | EPI_TANH_GRAD
  const float* aux_in; long long aux_go;      // clipped noise (TANH_NOISE) / tanh(y) (TANH_GRAD)   [B, A]
  float* aux_out;                             // tanh(y) (BIAS_TANH)                                 [B, A]
  float* a_out; long long a_go; int a_ld, pad1;
  float f0, f1;
  FrontNet net[kFrontMaxNets];
};

// Code size is part of the cost here: a launch of the update chain starts with a cold instruction cache and a CTA
// runs this body once, so every unrolled copy of a loop is paid for in instruction-fetch latency (a fully unrolled
// version of this function was 9.5k instructions and took 38 us).  All global reads are therefore
// cp.async copies issued from ROLLED loops -- asynchronous, so they still overlap -- followed by one wait.

// n contiguous floats global -> shared, 16 bytes at a time when both sides allow it
__device__ __forceinline__ void front_copy(float* dst, const float* src, int n, int tid, int nthr) {
  const bool vec = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)__cvta_generic_to_shared(dst)) & 15) == 0;
  const int n4 = vec ? n >> 2 : 0;
#pragma unroll 1
  for (int i = tid; i < n4; i += nthr) cp_async16(dst + 4 * i, src + 4 * i, true);
#pragma unroll 1
  for (int i = 4 * n4 + tid; i < n; i += nthr) cp_async4(dst + i, src + i, true);
}

__device__ __forceinline__ void front_body(const FrontParams& P, int tile, float* smem) {
  const int R = P.rows_per_tile;
  float* xs = smem;                                   // [R][kFrontXs]
  float* ws = smem + R * kFrontXs;                    // [kFrontCols][ws_ld]
  float* bs = ws + kFrontCols * kFrontWs;             // [kFrontCols]
  float* whs = bs + kFrontCols;                       // [kFrontMaxA][h_ld]
  float* hs = whs + kFrontMaxA * kFrontWh;            // [R][h_ld]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x;
  const int per_agent = P.row_blocks * P.jobs;
  const int agent = tile / per_agent;
  int rem = tile - agent * per_agent;
  const int rb = rem / P.jobs, job = rem - rb * P.jobs;
  int ni = 0;
#pragma unroll 1
  for (int i = 1; i < P.n_nets; ++i)
    if (job >= P.net[i].job_begin) ni = i;
  const FrontNet& N = P.net[ni];
  const int lj = job - N.job_begin;
  const int inner = lj / N.col_blocks, c0 = (lj - inner * N.col_blocks) * kFrontCols;
  const int r0 = rb * R;
  const bool writer = job == 0;                       // net[0].x_off == 0 (host): the writer's rows are staged unshifted
  const int nc = min(kFrontCols, N.N - c0), K = N.K;
  const int A = P.A, Kh = P.Kh;
  const int nrows = min(R, P.batch - r0);
  const int h_ld = (Kh + 3) & ~3;                     // rows of hs / whs (a flat copy when the source is dense)
  const int ws_ld = K;                                // weight block [nc][K], flat
  const bool do_head = P.head && (writer || N.act_col >= 0);   // tiles of a network that does not consume the head skip it

  // ---- phase 1: every global read, as asynchronous copies ----
  if (P.gather) {
    const GatherParams& G = P.g;
    const int xlen = writer ? (int)G.row_floats : K;  // the writer scatters the whole transition afterwards
#pragma unroll 1
    for (int r = warp; r < nrows; r += (nthr >> 5)) {
      const long long idx = gather_index(G, agent, r0 + r);
      if (writer && lane == 0) G.idx_out[(long long)agent * G.batch + r0 + r] = idx;
      const float* src = G.rows + (long long)agent * G.rb_agent_stride + idx * G.row_stride + N.x_off;
      for (int i = lane; i < xlen; i += 32) cp_async4(xs + r * kFrontXs + i, src + i, true);
    }
  } else if (N.x) {
    const float* x = N.x + (long long)agent * N.x_go + (long long)r0 * N.ldx + N.x_off;
#pragma unroll 1
    for (int r = warp; r < nrows; r += (nthr >> 5))
      for (int i = lane; i < K; i += 32) cp_async4(xs + r * kFrontXs + i, x + (long long)r * N.ldx + i, true);
  }
  {
    const float* W = N.W + (long long)agent * N.w_go + (long long)inner * N.w_gi + (long long)c0 * N.ws_c;
    if (N.ws_k == 1 && N.ws_c == K) {
      front_copy(ws, W, nc * K, tid, nthr);
    } else {
#pragma unroll 1
      for (int k = 0; k < K; ++k)
        for (int c = tid; c < nc; c += nthr) cp_async4(ws + c * ws_ld + k, W + (long long)c * N.ws_c + (long long)k * N.ws_k, true);
    }
    if (N.bias && tid < nc) cp_async4(bs + tid, N.bias + (long long)agent * N.w_go + (long long)inner * N.w_gi + c0 + tid, true);
  }
  if (do_head) {
    const float* Wh = P.Wh + (long long)agent * P.wh_go;
    if (P.hs_k == 1 && P.hs_j == h_ld) {
      front_copy(whs, Wh, A * Kh, tid, nthr);
    } else {
#pragma unroll 1
      for (int j = 0; j < A; ++j)
        for (int k = tid; k < Kh; k += nthr) cp_async4(whs + j * h_ld + k, Wh + (long long)j * P.hs_j + (long long)k * P.hs_k, true);
    }
    const float* h = P.h + (long long)agent * P.h_go + (long long)r0 * P.ldh;
    if (P.ldh == h_ld) {
      front_copy(hs, h, nrows * Kh, tid, nthr);
    } else {
#pragma unroll 1
      for (int r = 0; r < nrows; ++r)
        for (int k = tid; k < Kh; k += nthr) cp_async4(hs + r * h_ld + k, h + (long long)r * P.ldh + k, true);
    }
  }
  // columns [K, 32) of the staged rows are multiplied by zero weights in phase 3: they must be finite.  The sampling
  // writer's rows hold the rest of the transition there; everything else is cleared.
  {
    const int filled = (P.gather && writer) ? (int)P.g.row_floats : K;
#pragma unroll 1
    for (int i = tid; i < nrows * kFrontMaxK; i += nthr) {
      const int r = i / kFrontMaxK, k = i - r * kFrontMaxK;
      if (k >= filled) xs[r * kFrontXs + k] = 0.f;
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  // the sampled batch for the later stages: scattered from the staged rows, plus the clipped smoothing noise
  if (P.gather && writer) {
#pragma unroll 1
    for (int r = warp; r < nrows; r += (nthr >> 5)) gather_scatter(P.g, agent, r0 + r, xs + r * kFrontXs, lane);
  }

  // ---- phase 2: head, A numbers per row; one warp per row, the reduction split over the lanes ----
  if (do_head) {
#pragma unroll 1
    for (int r = warp; r < nrows; r += (nthr >> 5)) {
      float acc[kFrontMaxA];
#pragma unroll
      for (int j = 0; j < kFrontMaxA; ++j) acc[j] = 0.f;
#pragma unroll 2
      for (int k = lane; k < Kh; k += 32) {
        const float hv = hs[r * h_ld + k];
#pragma unroll
        for (int j = 0; j < kFrontMaxA; ++j)
          if (j < A) acc[j] = fmaf(hv, whs[j * h_ld + k], acc[j]);
      }
      float v = 0.f;                                   // lane j keeps output j
#pragma unroll
      for (int j = 0; j < kFrontMaxA; ++j) {
        float t = acc[j];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (lane == j) v = t;
      }
      if (lane < A) {
        const int j = lane, b = r0 + r;
        const float bias = P.bh ? P.bh[(long long)agent * P.wh_go + j] : 0.f;
        const long long ai = (long long)agent * P.aux_go + (long long)b * A + j;
        float res;
        if (P.head_epi == EPI_TANH_GRAD) {
          const float y = P.aux_in[ai];
          res = v * P.f0 * (1.f - y * y);
        } else {
          const float y = tanhf(v + bias);
          res = P.f0 * y;
          if (P.head_epi == EPI_BIAS_TANH_NOISE) {
            res += P.aux_in[ai];
            if (P.f1 > 0.f) res = fminf(fmaxf(res, -P.f1), P.f1);
          } else if (writer) {
            P.aux_out[ai] = y;
          }
        }
        if (N.act_col >= 0) xs[r * kFrontXs + N.act_col + j] = res;
        if (writer) P.a_out[(long long)agent * P.a_go + (long long)b * P.a_ld + j] = res;
      }
    }
    __syncthreads();
  }

  // ---- phase 3: layer.  thread = output column, 8 rows each; weights in registers, rows broadcast from shared memory ----
  {
    const int c = tid & (kFrontCols - 1), half = tid >> 7;
    if (c < nc && half < 2) {
      float w[kFrontMaxK];
#pragma unroll
      for (int k = 0; k < kFrontMaxK; ++k) w[k] = k < K ? ws[c * ws_ld + k] : 0.f;
      const float bias = N.bias ? bs[c] : 0.f;
      float* out = N.out + (long long)agent * N.out_go + (long long)inner * N.out_gi;
      const float* mask = N.mask ? N.mask + (long long)agent * N.mask_go + (long long)inner * N.mask_gi : nullptr;
      const int rbeg = half * (R / 2), rend = min(nrows, rbeg + R / 2);
#pragma unroll 1
      for (int r = rbeg; r < rend; ++r) {
        const long long o = (long long)(r0 + r) * N.ldo + c0 + c;
        const float m = mask ? mask[o] : 1.f;
        const float4* xr = reinterpret_cast<const float4*>(xs + r * kFrontXs);
        float v = bias;
#pragma unroll
        for (int k4 = 0; k4 < kFrontMaxK / 4; ++k4)
          if (4 * k4 < K) {                            // w[k >= K] == 0 and the staged columns there are finite
            const float4 x4 = xr[k4];
            v = fmaf(x4.x, w[4 * k4], v);
            v = fmaf(x4.y, w[4 * k4 + 1], v);
            v = fmaf(x4.z, w[4 * k4 + 2], v);
            v = fmaf(x4.w, w[4 * k4 + 3], v);
          }
        const float ov = mask ? (m > 0.f ? v : 0.f) : fmaxf(v, 0.f);
        out[o] = N.rn_out ? rn_tf32(ov) : ov;
      }
    }
  }
  __syncthreads();        // shared memory is reused by the next tile (persistent kernel)
}

__global__ void __launch_bounds__(256) front_kernel(const __grid_constant__ FrontParams P) {
  extern __shared__ __align__(16) float front_smem[];
  pdl_launch_dependents();
  pdl_wait();
  front_body(P, blockIdx.x, front_smem);
}

}  // namespace td3
