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
  int w_window, units;                        // front_wide_body: floats of shared memory for weight matrices; (network, twin) units
  int rows_per_tile, job_groups;              // kFrontRows (latency: one agent, batch 256); job_groups > 0: the row-block-major
                                              // kernel for thousands of rows (front_wide_body), jobs split over that many tiles
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
    // lane k of a warp draws the index of the warp's k-th row (one Philox pass per warp instead of one per row)
    const int nw = nthr >> 5;
    const int my_r = warp + lane * nw;
    const long long my_idx = (lane < 8 && my_r < nrows) ? gather_index(G, agent, r0 + my_r) : 0;
    int kk = 0;
#pragma unroll 1
    for (int r = warp; r < nrows; r += nw, ++kk) {
      const long long idx = (long long)__shfl_sync(0xffffffffu, (unsigned long long)my_idx, kk & 31);
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

// ---- thousands of rows (a population of agents, a large data-parallel batch): row-block-major tiling ----
// The latency form above gives every (row block, 128-column job) its own CTA, so a row block is sampled, staged and
// its head recomputed once per job (12 times for the sampling launch of cfg2) and a CTA retires after ~100 FMAs per
// thread: at 8 agents x 256 rows that was 16 instructions per useful FMA and 30-40 us per launch.  Here a tile is
// (agent, block of R rows, group of units), a unit being one (network, twin): phase 1 / 2 run once and the units'
// whole weight matrices are staged with them (one wait for memory per tile), then every thread walks its output
// columns with the weights of a column in registers, four rows in flight, straight-line LDS.128 + FFMA.
constexpr int kFrontWideSmemBytes = 196608;         // everything a tile keeps in shared memory (fits the persistent kernel's ring)
__host__ __device__ constexpr int front_wide_smem_floats_fixed(int rows, bool head) {
  return rows * kFrontXs + kFrontMaxNets * rows * kFrontMaxK + (head ? kFrontMaxA * kFrontWh + rows * kFrontWh : 0);
}
// w_floats: the launch's weight window (FrontParams::w_window: the units of a tile all at once, or what is left of
// kFrontWideSmemBytes beside the row buffers)
constexpr int front_wide_smem_bytes(int rows, bool head, int w_floats) {
  return (front_wide_smem_floats_fixed(rows, head) + w_floats) * 4;
}
__host__ __device__ __forceinline__ int front_unit_floats(int N, int K) { return ((N * K + 3) & ~3) + ((N + 3) & ~3); }

struct FrontUnit { int ni, inner; };
__device__ __forceinline__ FrontUnit front_unit(const FrontParams& P, int u) {
  FrontUnit U{0, u};
#pragma unroll 1
  for (int i = 0; i + 1 < P.n_nets && U.inner >= P.net[i].n_inner; ++i) { U.inner -= P.net[i].n_inner; U.ni = i + 1; }
  return U;
}

// stages the weight matrices [N][K] + bias strips of units [lo, ...) that fit the launch's window, as asynchronous
// copies; returns the first unit left out
__device__ __forceinline__ int front_stage_units(const FrontParams& P, int lo, int hi, int agent, float* ws, int tid, int nthr) {
  int woff = 0, u = lo;
#pragma unroll 1
  for (; u < hi; ++u) {
    const FrontUnit U = front_unit(P, u);
    const FrontNet& N = P.net[U.ni];
    const int K = N.K, Nn = N.N;
    const int need = front_unit_floats(Nn, K);
    if (woff + need > P.w_window && u > lo) break;
    const long long po = (long long)agent * N.w_go + (long long)U.inner * N.w_gi;
    const float* W = N.W + po;
    float* dst = ws + woff;
    if (N.ws_k == 1 && N.ws_c == K) {
      front_copy(dst, W, Nn * K, tid, nthr);
    } else {
#pragma unroll 1
      for (int k = 0; k < K; ++k)
        for (int c = tid; c < Nn; c += nthr) cp_async4(dst + c * K + k, W + (long long)c * N.ws_c + (long long)k * N.ws_k, true);
    }
    if (N.bias) front_copy(dst + ((Nn * K + 3) & ~3), N.bias + po, Nn, tid, nthr);
    woff += need;
  }
  return u;
}

// fp32 -> nearest TF32 (ties away), the same value rn_tf32 returns, as two integer instructions
__device__ __forceinline__ float front_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

// every 128-th output column of one unit, starting at column c, over the tile's staged rows (stride kFrontMaxK; rows
// beyond `rows` hold finite-or-not garbage whose results are not stored); K4 = ceil(K / 4) quads of the reduction
template <int K4, bool kMask>
__device__ __forceinline__ void front_wide_unit(const float* __restrict__ wsb, const float* __restrict__ bsb, int K, int Nn, int c,
                                                const float* __restrict__ xr0, int rows, float* __restrict__ out0,
                                                const float* __restrict__ mask0, int ldo, bool rn) {
#pragma unroll 1
  for (int cc = c; cc < Nn; cc += kFrontCols) {
    const float* wc = wsb + cc * K;
    float w[4 * K4];
#pragma unroll
    for (int k = 0; k < 4 * K4; ++k) w[k] = (k < 4 * (K4 - 1) || k < K) ? wc[k] : 0.f;
    const float bias = bsb ? bsb[cc] : 0.f;
    const float4* xp = reinterpret_cast<const float4*>(xr0);
    float* op = out0 + cc;
    const float* mp = mask0 + cc;
#pragma unroll 1
    for (int rg = 0; rg < rows; rg += 4) {
      float v[4], m[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = bias;
        if (kMask) m[i] = rg + i < rows ? mp[(long long)i * ldo] : 0.f;
      }
#pragma unroll
      for (int k4 = 0; k4 < K4; ++k4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 x4 = xp[i * (kFrontMaxK / 4) + k4];
          v[i] = fmaf(x4.x, w[4 * k4], v[i]);
          v[i] = fmaf(x4.y, w[4 * k4 + 1], v[i]);
          v[i] = fmaf(x4.z, w[4 * k4 + 2], v[i]);
          v[i] = fmaf(x4.w, w[4 * k4 + 3], v[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (rg + i < rows) {
          const float ov = kMask ? (m[i] > 0.f ? v[i] : 0.f) : fmaxf(v[i], 0.f);
          op[(long long)i * ldo] = rn ? front_rn(ov) : ov;
        }
      xp += kFrontMaxK;                 // four rows of kFrontMaxK floats
      op += 4LL * ldo;
      if (kMask) mp += 4LL * ldo;
    }
  }
}

template <bool kMask>
__device__ __forceinline__ void front_wide_unit_k(int K4, const float* wsb, const float* bsb, int K, int Nn, int c, const float* xr0,
                                                  int rows, float* out0, const float* mask0, int ldo, bool rn) {
  switch (K4) {
    case 1: front_wide_unit<1, kMask>(wsb, bsb, K, Nn, c, xr0, rows, out0, mask0, ldo, rn); break;
    case 2: front_wide_unit<2, kMask>(wsb, bsb, K, Nn, c, xr0, rows, out0, mask0, ldo, rn); break;
    case 3: front_wide_unit<3, kMask>(wsb, bsb, K, Nn, c, xr0, rows, out0, mask0, ldo, rn); break;
    case 4: front_wide_unit<4, kMask>(wsb, bsb, K, Nn, c, xr0, rows, out0, mask0, ldo, rn); break;
    case 5: front_wide_unit<5, kMask>(wsb, bsb, K, Nn, c, xr0, rows, out0, mask0, ldo, rn); break;
    case 6: front_wide_unit<6, kMask>(wsb, bsb, K, Nn, c, xr0, rows, out0, mask0, ldo, rn); break;
    case 7: front_wide_unit<7, kMask>(wsb, bsb, K, Nn, c, xr0, rows, out0, mask0, ldo, rn); break;
    default: front_wide_unit<8, kMask>(wsb, bsb, K, Nn, c, xr0, rows, out0, mask0, ldo, rn); break;
  }
}

__device__ __forceinline__ void front_wide_body(const FrontParams& P, int tile, float* smem) {
  const int R = P.rows_per_tile;
  float* xfull = smem;                                        // [R][kFrontXs]       sampled transitions (scattered by the writer)
  float* xn = xfull + R * kFrontXs;                           // [nets][R][kFrontMaxK] input rows of every network, zero padded
  float* whs = xn + kFrontMaxNets * R * kFrontMaxK;           // [kFrontMaxA][h_ld]     (head launches only)
  float* hs = whs + kFrontMaxA * kFrontWh;                    // [R][h_ld]
  float* ws = smem + front_wide_smem_floats_fixed(R, P.head != 0);   // weight window: per unit [N][K] + [N] bias, 16-byte aligned
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x, nw = nthr >> 5;
  const int G = P.job_groups;
  const int per_agent = P.row_blocks * G;
  const int agent = tile / per_agent;
  const int rem = tile - agent * per_agent;
  const int rb = rem / G, jg = rem - rb * G;
  const int unit_lo = P.units * jg / G, unit_hi = P.units * (jg + 1) / G;
  const int r0 = rb * R;
  const bool writer = jg == 0;
  const int A = P.A, Kh = P.Kh;
  const int nrows = min(R, P.batch - r0);
  const int h_ld = (Kh + 3) & ~3;

  // ---- phase 1: every global read of the row block, and the tile's weights, as asynchronous copies ----
  if (P.gather) {
    const GatherParams& Gp = P.g;
    const int my_r = warp + lane * nw;                 // lane k draws the index of the warp's k-th row
    const long long my_idx = (my_r < nrows) ? gather_index(Gp, agent, r0 + my_r) : 0;
    int kk = 0;
#pragma unroll 1
    for (int r = warp; r < nrows; r += nw, ++kk) {
      const long long idx = (long long)__shfl_sync(0xffffffffu, (unsigned long long)my_idx, kk & 31);
      const float* src = Gp.rows + (long long)agent * Gp.rb_agent_stride + idx * Gp.row_stride;
      if (writer) {
        if (lane == 0) Gp.idx_out[(long long)agent * Gp.batch + r0 + r] = idx;
        for (int i = lane; i < (int)Gp.row_floats; i += 32) cp_async4(xfull + r * kFrontXs + i, src + i, true);
      }
#pragma unroll 1
      for (int ni = 0; ni < P.n_nets; ++ni)
        if (lane < P.net[ni].K) cp_async4(xn + (ni * R + r) * kFrontMaxK + lane, src + P.net[ni].x_off + lane, true);
    }
  } else {
#pragma unroll 1
    for (int ni = 0; ni < P.n_nets; ++ni) {
      const FrontNet& N = P.net[ni];
      if (!N.x) continue;
      const float* x = N.x + (long long)agent * N.x_go + (long long)r0 * N.ldx + N.x_off;
#pragma unroll 1
      for (int r = warp; r < nrows; r += nw)
        if (lane < N.K) cp_async4(xn + (ni * R + r) * kFrontMaxK + lane, x + (long long)r * N.ldx + lane, true);
    }
  }
  int chunk_hi = front_stage_units(P, unit_lo, unit_hi, agent, ws, tid, nthr);
  if (P.head) {
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
  // columns [K, 32) of every network's rows meet zero weights in phase 3, and rows past the batch are multiplied
  // without being stored: both must not hold anything that traps -- clear them (finite values are not required of
  // the unstored rows, zeros are simply the cheapest thing to write)
#pragma unroll 1
  for (int i = tid; i < P.n_nets * R * kFrontMaxK; i += nthr) {
    const int ni = i / (R * kFrontMaxK), r = (i / kFrontMaxK) % R, k = i & (kFrontMaxK - 1);
    if (k >= P.net[ni].K || r >= nrows) xn[i] = 0.f;
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  if (P.gather && writer) {
#pragma unroll 1
    for (int r = warp; r < nrows; r += nw) gather_scatter(P.g, agent, r0 + r, xfull + r * kFrontXs, lane);
  }

  // ---- phase 2: head, A numbers per row, once per row block ----
  if (P.head) {
#pragma unroll 1
    for (int r = warp; r < nrows; r += nw) {
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
      float v = 0.f;
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
#pragma unroll 1
        for (int ni = 0; ni < P.n_nets; ++ni)
          if (P.net[ni].act_col >= 0) xn[(ni * R + r) * kFrontMaxK + P.net[ni].act_col + j] = res;
        if (writer) P.a_out[(long long)agent * P.a_go + (long long)b * P.a_ld + j] = res;
      }
    }
    __syncthreads();
  }

  // ---- phase 3: the tile's units, a window of weight matrices at a time (normally all of them at once).
  // thread = output columns c, c + 128, ... of the rows of its 128-thread part ----
  const int c = tid & (kFrontCols - 1), part = tid >> 7, rpp = R / (nthr >> 7);      // rows per 128-thread part
  const int rbeg = part * rpp, rows = min(nrows - rbeg, rpp);
  int chunk_lo = unit_lo;
#pragma unroll 1
  while (true) {
    int woff = 0;
#pragma unroll 1
    for (int u = chunk_lo; u < chunk_hi; ++u) {
      const FrontUnit U = front_unit(P, u);
      const FrontNet& N = P.net[U.ni];
      const int K = N.K, Nn = N.N, ldo = N.ldo;
      const float* wsb = ws + woff;
      const float* bsb = N.bias ? wsb + ((Nn * K + 3) & ~3) : nullptr;
      woff += front_unit_floats(Nn, K);
      if (rows > 0) {
        const long long o0 = (long long)agent * N.out_go + (long long)U.inner * N.out_gi + (long long)(r0 + rbeg) * ldo;
        const float* xr0 = xn + (U.ni * R + rbeg) * kFrontMaxK;
        if (N.mask)
          front_wide_unit_k<true>((K + 3) >> 2, wsb, bsb, K, Nn, c, xr0, rows, N.out + o0,
                                  N.mask + (long long)agent * N.mask_go + (long long)U.inner * N.mask_gi + (long long)(r0 + rbeg) * ldo,
                                  ldo, N.rn_out != 0);
        else
          front_wide_unit_k<false>((K + 3) >> 2, wsb, bsb, K, Nn, c, xr0, rows, N.out + o0, nullptr, ldo, N.rn_out != 0);
      }
    }
    __syncthreads();            // the window is rewritten by the next chunk (and shared memory reused by the next tile)
    if (chunk_hi >= unit_hi) break;
    chunk_lo = chunk_hi;
    chunk_hi = front_stage_units(P, chunk_lo, unit_hi, agent, ws, tid, nthr);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
  }
}

constexpr int kFrontWideThreads = 512;
__global__ void __launch_bounds__(kFrontWideThreads, 1) front_wide_kernel(const __grid_constant__ FrontParams P) {
  extern __shared__ __align__(16) float front_smem[];
  pdl_launch_dependents();
  pdl_wait();
  front_wide_body(P, blockIdx.x, front_smem);
}

}  // namespace td3
