// stage.cuh -- the "stage" kernel: one launch = one dependency level of the TD3 update.
//
// A stage is a short table of tile problems (passed by value in the kernel parameter
// space).  Every CTA looks up the problem its blockIdx falls into and runs one tile of it.
// Problems of one stage are mutually independent (e.g. target-actor layer l and twin-critic
// layer l; or the dW and dX GEMMs of one backward layer), so a whole update is ~20 launches
// that are captured once into a CUDA graph (engine.cu).
//
// GEMM tiles are strict-fp32 FFMA (parity mode: results match the CPU oracle to ~1e-6); the
// large particle-encoder contractions go to the tcgen05 kernels in encoder_tc.cu instead.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace td3 {

enum ProblemKind : int {
  PK_GEMM = 0,          // C[i,j] = epi(sum_r A(i,r) B(r,j))
  PK_LN_FWD = 1,        // row LayerNorm (TD3_featured.py:44-46)
  PK_LN_BWD_ROWS = 2,   // d(input) of LayerNorm (+ optional ReLU mask)
  PK_LN_BWD_COLS = 3,   // d(gamma), d(beta)
  PK_POOL_FWD = 4,      // relu(mean over particles)  (TD3_particles.py:57)
  PK_POOL_BWD = 5,      // dH2 = dpool/N * (pool>0) * (h2>0)
  PK_REDUCE_SPLITS = 6, // C = sum_s partial[s]   (split-K second phase)
  PK_NEG_MEAN = 7       // scalar = -mean(A)      (actor loss read-back)
};

enum Epilogue : int {
  EPI_STORE = 0,
  EPI_BIAS = 1,             // C = acc + bias[j]
  EPI_BIAS_RELU = 2,        // C = relu(acc + bias[j])
  EPI_BIAS_TANH = 3,        // y = tanh(acc + bias[j]); aux0 <- y; C = f0 * y
  EPI_BIAS_TANH_NOISE = 4,  // C = clamp(f0 * tanh(acc + bias) + aux0[i,j], +-f1)   (f1 <= 0: no clamp)
  EPI_RELU_MASK = 5,        // C = acc * (aux0[i,j] > 0)
  EPI_TANH_GRAD = 6         // C = acc * f0 * (1 - aux0[i,j]^2)
};

struct Problem {
  int kind, epi;
  int M, N, K;                    // output M x N, reduction extent K
  int lda, ldb, ldc, ldaux;
  int a_rc, b_rc;                 // 1: reduction index is the contiguous one in memory
  int a_vec, b_vec;               // 1: 16-byte vector loads are legal for this operand
  int groups_inner;               // inner group count (twin critics); outer groups = agents
  int tiles_m, tiles_n, tiles_per_group;
  int tile_begin, tile_count;     // [tile_begin, tile_begin + tile_count) of the stage's grid
  int ksplit;                     // >1: reduction split over CTAs, partials at C + s*c_split
  int c_dups;                     // epilogue writes C to c_dups destinations c_dup_stride apart
  int reserved;
  long long c_split, c_dup_stride;
  const float* A; const float* B; float* C; const float* bias;
  float* aux0; float* aux1; float* aux2; float* aux3;
  // per-group pointer strides (floats): *_go outer group (agent), *_gi inner group (twin)
  long long a_go, a_gi, b_go, b_gi, c_go, c_gi, bias_go, bias_gi;
  long long aux0_go, aux0_gi, aux1_go, aux1_gi, aux2_go, aux2_gi, aux3_go, aux3_gi;
  float f0, f1;
};

constexpr int kMaxProblemsPerStage = 6;

struct StageParams {
  int n_problems;
  int total_tiles;
  Problem p[kMaxProblemsPerStage];
};

constexpr int kStageThreads = 256;
constexpr int kBM = 32, kBN = 32, kBK = 32;
constexpr int kLd = 36;                         // smem row stride (floats): 16B aligned, conflict-free
constexpr int kTileFloats = 32 * kLd;           // 1152
constexpr int kSmemFloats = 4 * kTileFloats;    // double-buffered A and B: 4608 floats = 18 KB

// ------------------------------------------------------------------------------------
// global -> register tile fragment.  Both layouts use the same thread map: row = tid/8,
// col4 = (tid%8)*4.  rc: rows index the output dim, cols the reduction; oc: the opposite.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float4 load_frag(const float* __restrict__ base, int ld, bool rc, bool vec,
                                            int o0, int k0, int O, int K, int tid) {
  const int row = tid >> 3, c4 = (tid & 7) << 2;
  const int row_idx = rc ? o0 + row : k0 + row;
  const int row_lim = rc ? O : K;
  const int col_idx = rc ? k0 + c4 : o0 + c4;
  const int col_lim = rc ? K : O;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row_idx < row_lim && col_idx < col_lim) {
    const float* p = base + (size_t)row_idx * ld + col_idx;
    if (vec) {
      v = __ldg(reinterpret_cast<const float4*>(p));
    } else {
      v.x = __ldg(p);
      if (col_idx + 1 < col_lim) v.y = __ldg(p + 1);
      if (col_idx + 2 < col_lim) v.z = __ldg(p + 2);
      if (col_idx + 3 < col_lim) v.w = __ldg(p + 3);
    }
  }
  return v;
}

__device__ __forceinline__ float apply_epilogue(const Problem& P, float v, int i, int j, const float* bias,
                                                float* aux0) {
  switch (P.epi) {
    case EPI_BIAS: return v + __ldg(bias + j);
    case EPI_BIAS_RELU: return fmaxf(v + __ldg(bias + j), 0.f);
    case EPI_BIAS_TANH: {
      float y = tanhf(v + __ldg(bias + j));
      aux0[(size_t)i * P.ldaux + j] = y;
      return P.f0 * y;
    }
    case EPI_BIAS_TANH_NOISE: {
      float a = P.f0 * tanhf(v + __ldg(bias + j)) + aux0[(size_t)i * P.ldaux + j];
      if (P.f1 > 0.f) a = fminf(fmaxf(a, -P.f1), P.f1);
      return a;
    }
    case EPI_RELU_MASK: return aux0[(size_t)i * P.ldaux + j] > 0.f ? v : 0.f;
    case EPI_TANH_GRAD: {
      float y = aux0[(size_t)i * P.ldaux + j];
      return v * P.f0 * (1.f - y * y);
    }
    default: return v;
  }
}

// ------------------------------------------------------------------------------------
// One 32x32 output tile; 256 threads = 4 k-groups x (8x8 threads x 4x4 micro-tile).
// ------------------------------------------------------------------------------------
template <bool ARC, bool BRC>
__device__ __forceinline__ void gemm_tile(const Problem& P, int tile, float* smem) {
  const int tid = threadIdx.x;
  const int kg = tid >> 6, t64 = tid & 63, ti = t64 >> 3, tj = t64 & 7;

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
  const int i0 = tm * kBM, j0 = tn * kBN;

  const float* __restrict__ A = P.A + go * P.a_go + gi * P.a_gi;
  const float* __restrict__ B = P.B + go * P.b_go + gi * P.b_gi;
  float* __restrict__ C = P.C + go * P.c_go + gi * P.c_gi + (long long)ks * P.c_split;
  const float* bias = P.bias ? P.bias + go * P.bias_go + gi * P.bias_gi : nullptr;
  float* aux0 = P.aux0 ? P.aux0 + go * P.aux0_go + gi * P.aux0_gi : nullptr;
  float* aux1 = P.aux1 ? P.aux1 + go * P.aux1_go + gi * P.aux1_gi : nullptr;

  // reduction range of this CTA
  int k_begin = 0, k_end = P.K;
  if (P.ksplit > 1) {
    const int chunks = (P.K + kBK - 1) / kBK;
    const int per = (chunks + P.ksplit - 1) / P.ksplit;
    k_begin = min(P.K, ks * per * kBK);
    k_end = min(P.K, (ks + 1) * per * kBK);
  }
  const int n_chunks = (k_end - k_begin + kBK - 1) / kBK;

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  float asum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool want_rowsum = (aux1 != nullptr) && (P.epi == EPI_STORE) && (tn == 0);

  const int srow = tid >> 3, sc4 = (tid & 7) << 2;
  float4 fa = make_float4(0.f, 0.f, 0.f, 0.f), fb = fa;
  if (n_chunks > 0) {
    fa = load_frag(A, P.lda, ARC, P.a_vec, i0, k_begin, P.M, k_end, tid);
    fb = load_frag(B, P.ldb, BRC, P.b_vec, j0, k_begin, P.N, k_end, tid);
  }
  for (int c = 0; c < n_chunks; ++c) {
    float* As = smem + (c & 1) * 2 * kTileFloats;
    float* Bs = As + kTileFloats;
    *reinterpret_cast<float4*>(As + srow * kLd + sc4) = fa;
    *reinterpret_cast<float4*>(Bs + srow * kLd + sc4) = fb;
    __syncthreads();
    if (c + 1 < n_chunks) {
      const int k0 = k_begin + (c + 1) * kBK;
      fa = load_frag(A, P.lda, ARC, P.a_vec, i0, k0, P.M, k_end, tid);
      fb = load_frag(B, P.ldb, BRC, P.b_vec, j0, k0, P.N, k_end, tid);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int r = kg * 8 + s * 4;
      float a[4][4], b[4][4];
      if (ARC) {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const float4 v = *reinterpret_cast<const float4*>(As + (ti + 8 * ci) * kLd + r);
          a[ci][0] = v.x; a[ci][1] = v.y; a[ci][2] = v.z; a[ci][3] = v.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(As + (r + q) * kLd + ti * 4);
          a[0][q] = v.x; a[1][q] = v.y; a[2][q] = v.z; a[3][q] = v.w;
        }
      }
      if (BRC) {
#pragma unroll
        for (int cj = 0; cj < 4; ++cj) {
          const float4 v = *reinterpret_cast<const float4*>(Bs + (tj + 8 * cj) * kLd + r);
          b[cj][0] = v.x; b[cj][1] = v.y; b[cj][2] = v.z; b[cj][3] = v.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(Bs + (r + q) * kLd + tj * 4);
          b[0][q] = v.x; b[1][q] = v.y; b[2][q] = v.z; b[3][q] = v.w;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
#pragma unroll
          for (int cj = 0; cj < 4; ++cj) acc[ci][cj] = fmaf(a[ci][q], b[cj][q], acc[ci][cj]);
      if (want_rowsum) {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) asum[ci] += (a[ci][0] + a[ci][1]) + (a[ci][2] + a[ci][3]);
      }
    }
    // the buffer written two iterations from now is this one: guard before it is overwritten
    if (c + 2 < n_chunks || true) __syncthreads();
  }

  // ---- reduce the 4 k-groups through shared memory ----
  float* red = smem;                       // [kg][e][t64] : 4*16*64 = 4096 floats
  float* sdb = smem + 4096;                // [kg][ti][c]  : 128 floats
#pragma unroll
  for (int ci = 0; ci < 4; ++ci)
#pragma unroll
    for (int cj = 0; cj < 4; ++cj) red[(kg * 16 + ci * 4 + cj) * 64 + t64] = acc[ci][cj];
  if (want_rowsum && tj == 0) {
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) sdb[(kg * 8 + ti) * 4 + ci] = asum[ci];
  }
  __syncthreads();

  const int ci = kg;
  const int i = i0 + (ARC ? ti + 8 * ci : ti * 4 + ci);
#pragma unroll
  for (int cj = 0; cj < 4; ++cj) {
    const int e = ci * 4 + cj;
    float v = (red[(0 * 16 + e) * 64 + t64] + red[(1 * 16 + e) * 64 + t64]) +
              (red[(2 * 16 + e) * 64 + t64] + red[(3 * 16 + e) * 64 + t64]);
    const int j = j0 + (BRC ? tj + 8 * cj : tj * 4 + cj);
    if (i < P.M && j < P.N) {
      v = apply_epilogue(P, v, i, j, bias, aux0);
      for (int d = 0; d < P.c_dups; ++d) C[d * P.c_dup_stride + (size_t)i * P.ldc + j] = v;
    }
  }
  if (want_rowsum && tid < 32) {
    const int rti = tid >> 2, rc = tid & 3;
    const float s = (sdb[(0 * 8 + rti) * 4 + rc] + sdb[(1 * 8 + rti) * 4 + rc]) +
                    (sdb[(2 * 8 + rti) * 4 + rc] + sdb[(3 * 8 + rti) * 4 + rc]);
    const int row = i0 + (ARC ? rti + 8 * rc : rti * 4 + rc);
    if (row < P.M) aux1[(long long)ks * P.M + row] = s;
  }
}

// ------------------------------------------------------------------------------------
// LayerNorm (eps = f0).  One warp per row, 8 rows per CTA.
//   fwd : A = input [M,lda], B = gamma, bias = beta, C = output, aux2 = mean[M], aux3 = rstd[M]
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void group_ptrs(const Problem& P, int g, long long& go, long long& gi) {
  go = g / P.groups_inner;
  gi = g - go * P.groups_inner;
}

__device__ __forceinline__ void ln_fwd_tile(const Problem& P, int tile) {
  const int g = tile / P.tiles_per_group;
  const int t = tile - g * P.tiles_per_group;
  long long go, gi;
  group_ptrs(P, g, go, gi);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = t * 8 + warp;
  if (row >= P.M) return;
  const float* x = P.A + go * P.a_go + gi * P.a_gi + (size_t)row * P.lda;
  const float* gamma = P.B + go * P.b_go + gi * P.b_gi;
  const float* beta = P.bias + go * P.bias_go + gi * P.bias_gi;
  float* y = P.C + go * P.c_go + gi * P.c_gi + (size_t)row * P.ldc;
  float s = 0.f;
  for (int j = lane; j < P.N; j += 32) s += x[j];
  const float mean = warp_sum(s) / (float)P.N;
  float ss = 0.f;
  for (int j = lane; j < P.N; j += 32) {
    const float d = x[j] - mean;
    ss = fmaf(d, d, ss);
  }
  const float var = warp_sum(ss) / (float)P.N;
  const float rstd = 1.f / sqrtf(var + P.f0);
  for (int j = lane; j < P.N; j += 32) y[j] = (x[j] - mean) * rstd * __ldg(gamma + j) + __ldg(beta + j);
  if (lane == 0) {
    (P.aux2 + go * P.aux2_go + gi * P.aux2_gi)[row] = mean;
    (P.aux3 + go * P.aux3_go + gi * P.aux3_gi)[row] = rstd;
  }
}

//   bwd rows : A = d(out) [M,lda], B = gamma, aux0 = LN input [M,ldaux], aux2/aux3 = mean/rstd,
//              C = d(in) [M,ldc]; epi == EPI_RELU_MASK multiplies by (input > 0) (ReLU precedes the LN)
__device__ __forceinline__ void ln_bwd_rows_tile(const Problem& P, int tile) {
  const int g = tile / P.tiles_per_group;
  const int t = tile - g * P.tiles_per_group;
  long long go, gi;
  group_ptrs(P, g, go, gi);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = t * 8 + warp;
  if (row >= P.M) return;
  const float* dy = P.A + go * P.a_go + gi * P.a_gi + (size_t)row * P.lda;
  const float* gamma = P.B + go * P.b_go + gi * P.b_gi;
  const float* x = P.aux0 + go * P.aux0_go + gi * P.aux0_gi + (size_t)row * P.ldaux;
  float* dx = P.C + go * P.c_go + gi * P.c_gi + (size_t)row * P.ldc;
  const float mean = (P.aux2 + go * P.aux2_go + gi * P.aux2_gi)[row];
  const float rstd = (P.aux3 + go * P.aux3_go + gi * P.aux3_gi)[row];
  float s1 = 0.f, s2 = 0.f;
  for (int j = lane; j < P.N; j += 32) {
    const float gdy = dy[j] * __ldg(gamma + j);
    const float xh = (x[j] - mean) * rstd;
    s1 += gdy;
    s2 = fmaf(gdy, xh, s2);
  }
  s1 = warp_sum(s1) / (float)P.N;
  s2 = warp_sum(s2) / (float)P.N;
  for (int j = lane; j < P.N; j += 32) {
    const float xv = x[j];
    const float gdy = dy[j] * __ldg(gamma + j);
    const float xh = (xv - mean) * rstd;
    float d = rstd * (gdy - s1 - xh * s2);
    if (P.epi == EPI_RELU_MASK && !(xv > 0.f)) d = 0.f;
    dx[j] = d;
  }
}

//   bwd cols : A = d(out), aux0 = LN input, aux2/aux3 = mean/rstd -> C = d(gamma)[N], aux1 = d(beta)[N]
__device__ __forceinline__ void ln_bwd_cols_tile(const Problem& P, int tile, float* smem) {
  const int g = tile / P.tiles_per_group;
  const int t = tile - g * P.tiles_per_group;
  long long go, gi;
  group_ptrs(P, g, go, gi);
  const int rl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = t * 32 + lane;
  const float* dy = P.A + go * P.a_go + gi * P.a_gi;
  const float* x = P.aux0 + go * P.aux0_go + gi * P.aux0_gi;
  const float* mean = P.aux2 + go * P.aux2_go + gi * P.aux2_gi;
  const float* rstd = P.aux3 + go * P.aux3_go + gi * P.aux3_gi;
  float dg = 0.f, db = 0.f;
  if (j < P.N) {
    for (int m = rl; m < P.M; m += 8) {
      const float d = dy[(size_t)m * P.lda + j];
      const float xh = (x[(size_t)m * P.ldaux + j] - mean[m]) * rstd[m];
      dg = fmaf(d, xh, dg);
      db += d;
    }
  }
  smem[rl * 32 + lane] = dg;
  smem[256 + rl * 32 + lane] = db;
  __syncthreads();
  if (rl == 0 && j < P.N) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      sg += smem[r * 32 + lane];
      sb += smem[256 + r * 32 + lane];
    }
    (P.C + go * P.c_go + gi * P.c_gi)[j] = sg;
    (P.aux1 + go * P.aux1_go + gi * P.aux1_gi)[j] = sb;
  }
}

// ------------------------------------------------------------------------------------
// Particle pooling (TD3_particles.py:57-58): C[b, c] = relu(mean_n A[(b*K + n), c]).
//   M = batch, N = channels, K = particles per sample; tile = 8 samples? -> one CTA per
//   (sample, 32-channel strip); 8 row-lanes walk the particles.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void pool_fwd_tile(const Problem& P, int tile, float* smem) {
  const int g = tile / P.tiles_per_group;
  int t = tile - g * P.tiles_per_group;
  long long go, gi;
  group_ptrs(P, g, go, gi);
  const int b = t / P.tiles_n, strip = t - b * P.tiles_n;
  const int rl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = strip * 32 + lane;
  const float* h = P.A + go * P.a_go + gi * P.a_gi + (size_t)b * P.K * P.lda;
  float s = 0.f;
  if (c < P.N)
    for (int n = rl; n < P.K; n += 8) s += h[(size_t)n * P.lda + c];
  smem[rl * 32 + lane] = s;
  __syncthreads();
  if (rl == 0 && c < P.N) {
    float tot = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) tot += smem[r * 32 + lane];
    const float v = fmaxf(tot / (float)P.K, 0.f);
    float* C = P.C + go * P.c_go + gi * P.c_gi;
    for (int d = 0; d < P.c_dups; ++d) C[d * P.c_dup_stride + (size_t)b * P.ldc + c] = v;
  }
}

// dH2[(b*K+n), c] = (A[b,c] / K) * (aux0[b,c] > 0) * (aux1[(b*K+n), c] > 0)
//   A = d(pool out) [M, lda], aux0 = pool out [M, ldaux], aux1 = h2 [M*K, N] (ld = ldb), C = dH2 [M*K, ldc]
//   tile = 8 particle rows x all channels
__device__ __forceinline__ void pool_bwd_tile(const Problem& P, int tile) {
  const int g = tile / P.tiles_per_group;
  const int t = tile - g * P.tiles_per_group;
  long long go, gi;
  group_ptrs(P, g, go, gi);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)t * 8 + warp;
  if (row >= (long long)P.M * P.K) return;
  const int b = (int)(row / P.K);
  const float* dp = P.A + go * P.a_go + gi * P.a_gi + (size_t)b * P.lda;
  const float* pool = P.aux0 + go * P.aux0_go + gi * P.aux0_gi + (size_t)b * P.ldaux;
  const float* h2 = P.aux1 + go * P.aux1_go + gi * P.aux1_gi + (size_t)row * P.ldb;
  float* out = P.C + go * P.c_go + gi * P.c_gi + (size_t)row * P.ldc;
  const float inv = 1.f / (float)P.K;
  for (int c = lane; c < P.N; c += 32) {
    const float gate = (pool[c] > 0.f && h2[c] > 0.f) ? 1.f : 0.f;
    out[c] = dp[c] * inv * gate;
  }
}

// C[e] = sum_s A[s*c_split + e] for e in tile of 1024 elements (M = element count, K = splits)
__device__ __forceinline__ void reduce_splits_tile(const Problem& P, int tile) {
  const int g = tile / P.tiles_per_group;
  const int t = tile - g * P.tiles_per_group;
  long long go, gi;
  group_ptrs(P, g, go, gi);
  const float* part = P.A + go * P.a_go + gi * P.a_gi;
  float* out = P.C + go * P.c_go + gi * P.c_gi;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int e = t * 1024 + u * 256 + threadIdx.x;
    if (e < P.M) {
      float s = 0.f;
      for (int k = 0; k < P.K; ++k) s += part[(long long)k * P.c_split + e];
      out[e] = s;
    }
  }
}

// scalar C[g] = f0 * mean(A[0..M*N)) with A [M, lda]; one CTA per group
__device__ __forceinline__ void neg_mean_tile(const Problem& P, int tile, float* smem) {
  long long go, gi;
  group_ptrs(P, tile, go, gi);
  const float* a = P.A + go * P.a_go + gi * P.a_gi;
  float s = 0.f;
  const int total = P.M * P.N;
  for (int e = threadIdx.x; e < total; e += kStageThreads) {
    const int i = e / P.N, j = e - i * P.N;
    s += a[(size_t)i * P.lda + j];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < kStageThreads / 32; ++w) tot += smem[w];
    (P.C + go * P.c_go + gi * P.c_gi)[0] = P.f0 * tot / (float)total;
  }
}

__global__ void __launch_bounds__(kStageThreads, 2) stage_kernel(const __grid_constant__ StageParams S) {
  __shared__ __align__(16) float smem[kSmemFloats];
  const int tile_global = blockIdx.x;
  int pi = 0;
#pragma unroll
  for (int q = 1; q < kMaxProblemsPerStage; ++q)
    if (q < S.n_problems && tile_global >= S.p[q].tile_begin) pi = q;
  const Problem& P = S.p[pi];
  const int tile = tile_global - P.tile_begin;
  switch (P.kind) {
    case PK_GEMM:
      if (P.a_rc && P.b_rc) gemm_tile<true, true>(P, tile, smem);
      else if (P.a_rc && !P.b_rc) gemm_tile<true, false>(P, tile, smem);
      else if (!P.a_rc && !P.b_rc) gemm_tile<false, false>(P, tile, smem);
      else gemm_tile<false, true>(P, tile, smem);
      break;
    case PK_LN_FWD: ln_fwd_tile(P, tile); break;
    case PK_LN_BWD_ROWS: ln_bwd_rows_tile(P, tile); break;
    case PK_LN_BWD_COLS: ln_bwd_cols_tile(P, tile, smem); break;
    case PK_POOL_FWD: pool_fwd_tile(P, tile, smem); break;
    case PK_POOL_BWD: pool_bwd_tile(P, tile); break;
    case PK_REDUCE_SPLITS: reduce_splits_tile(P, tile); break;
    case PK_NEG_MEAN: neg_mean_tile(P, tile, smem); break;
    default: break;
  }
}

}  // namespace td3
