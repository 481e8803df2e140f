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

#include "misc.cuh"

namespace td3 {

enum ProblemKind : int {
  PK_GEMM = 0,          // C[i,j] = epi(sum_r A(i,r) B(r,j))
  PK_LN_FWD = 1,        // row LayerNorm (TD3_featured.py:44-46)
  PK_LN_BWD_ROWS = 2,   // d(input) of LayerNorm (+ optional ReLU mask)
  PK_LN_BWD_COLS = 3,   // d(gamma), d(beta)
  PK_POOL_FWD = 4,      // relu(mean over particles)  (TD3_particles.py:57)
  PK_POOL_BWD = 5,      // dH2 = dpool/N * (pool>0) * (h2>0)
  PK_REDUCE_SPLITS = 6, // C = sum_s partial[s]   (split-K second phase)
  PK_NEG_MEAN = 7,      // scalar = -mean(A)      (actor loss read-back)
  PK_COLSUM = 8,        // C[j] = sum_i A[i,j]    (bias gradient next to a tensor-core dW)
  PK_SMALLK_FWD = 9,    // C[i,j] = relu(bias[j] + sum_{d<K} A[i,d] B[j,d]), K <= 8   (particle encoder layer 1: HBM-write bound)
  PK_SMALLK_DW = 10,    // C[j,d] = sum_i A[i,j] B[i,d], aux1[j] = sum_i A[i,j], N <= 8, reduction split over CTAs
  PK_ENC_FUSED = 11,    // host-side planning only: the fused set-encoder forward (enc.cuh), emitted as a launch of its own
  PK_ENC_BWD_W2 = 12,   // host-side planning only: fused set-encoder backward (encbwd.cuh), conv2 weight / bias gradient
  PK_ENC_BWD_X = 13     //                          ... and d(hidden) -> conv1 weight / bias gradient
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
  int use_tc;                     // 1: tcgen05 TF32 tile (128 x tc_nt) instead of the fp32 FFMA tile (32 x 32)
  int tc_nt, c_vec, aux_vec;      // TC tile width; 16-byte stores to C / loads from aux0 are legal
  int tc_cluster, rn_out;         // > 1: this many consecutive N tiles (one cluster) multicast their common A panel;
                                  // rn_out: the output is an operand of a later tensor-core contraction -> store it rounded to
                                  // nearest TF32 (the tensor core itself would truncate the low 13 mantissa bits)
  int map_a, map_b;               // >= 0: first tensor map of operand A / B inside StageParams::maps (kernel-parameter space:
                                  // the TMA unit reads the descriptor without a global-memory round trip); < 0: tmapA / tmapB
  long long c_split, c_dup_stride;
  const float* A; const float* B; float* C; const float* bias;
  const void* tmapA; const void* tmapB;   // device arrays of CUtensorMap (one per group) when the operand is TMA-loadable
  const float* B_master;                  // host-side planning only: B points at a TF32-rounded shadow of this buffer, which
                                          // only the tensor-core tile should read (finalize_problem puts it back otherwise)
  float* aux0; float* aux1; float* aux2; float* aux3;
  // per-group pointer strides (floats): *_go outer group (agent), *_gi inner group (twin)
  long long a_go, a_gi, b_go, b_gi, c_go, c_gi, bias_go, bias_gi;
  long long aux0_go, aux0_gi, aux1_go, aux1_gi, aux2_go, aux2_gi, aux3_go, aux3_gi;
  float f0, f1;
};

constexpr int kMaxProblemsPerStage = 6;
constexpr int kStageMaps = 16;    // tensor maps carried in the kernel parameters of a stage launch

struct alignas(64) TensorMapBlob { unsigned long long v[16]; };   // CUtensorMap (128 bytes, opaque)

struct StageParams {
  int n_problems;
  int total_tiles;
  int any_tc, cluster;            // some problem runs on the tensor cores (TMEM must be allocated); cluster size of the launch
  int small_ring, pipe_tiles;     // small_ring 1: many-tile launch -> one chunk per ring slot (96 KB instead of 192 KB): two CTAs per
                                  //    SM, so one tile's epilogue overlaps the other's operand loads and MMAs
                                  // pipe_tiles > 0: the launch's first pipe_tiles tiles are tensor-core tiles walked by one persistent
                                  //    CTA per SM with producer / MMA / epilogue warps pipelined ACROSS tiles (tcpipe.cuh)
  Problem p[kMaxProblemsPerStage];
  TensorMapBlob maps[kStageMaps];
};

constexpr int kStageThreads = 256;
constexpr int kBM = 32, kBN = 32, kBK = 32;
constexpr int kLd = 36;                         // smem row stride (floats): 16B aligned, conflict-free
constexpr int kTileFloats = 32 * kLd;           // 1152
constexpr int kRing = 8;                        // K-chunks resident in shared memory (kRing - 1 in flight)
constexpr int kRingFloats = kRing * 2 * kTileFloats;   // A and B rings: 18432 floats = 72 KB (dynamic)
constexpr int kSmemFloats = kRingFloats + kTileFloats + 32;   // + epilogue operand tile (aux0) + bias strip
constexpr int kSmemBytes = kSmemFloats * 4;

// ------------------------------------------------------------------------------------
// cp.async helpers.  Everything the update touches (activations, parameters) is also
// WRITTEN by other CTAs of the same persistent kernel, so no load may use the
// non-coherent path (ld.global.nc / __ldg): 16-byte copies bypass L1 (.cg), 4-byte
// copies and plain loads rely on the acquire fence of the grid barrier.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool pred) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = pred ? 16 : 0;                  // src-size 0 -> the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool pred) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// global -> shared copy of one 32x32 operand chunk.  Thread map: row = tid/8, col4 = (tid%8)*4.
// rc: rows index the output dim, cols the reduction; oc: the opposite.  The shared image keeps the
// global orientation (rows kLd floats apart).
#ifdef TD3_TILE_PROF
__device__ long long g_tp[128 * 16];
__device__ int g_tp_stage = -1;
#define TP(k) do { if (threadIdx.x == 0 && blockIdx.x == 0 && g_tp_stage >= 0) g_tp[g_tp_stage * 8 + (k)] = clock64(); } while (0)
#else
#define TP(k) do { } while (0)
#endif

struct Operand {
  const float* base;
  int ld, rc, vec, o0, O;
};

__device__ __forceinline__ void issue_chunk(float* dst, const Operand& op, int k0, int K, int tid) {
  const int row = tid >> 3, c4 = (tid & 7) << 2;
  const int row_idx = op.rc ? op.o0 + row : k0 + row;
  const int row_lim = op.rc ? op.O : K;
  const int col_idx = op.rc ? k0 + c4 : op.o0 + c4;
  const int col_lim = op.rc ? K : op.O;
  float* d = dst + row * kLd + c4;
  const bool row_ok = row_idx < row_lim;
  const float* p = op.base + (size_t)(row_ok ? row_idx : 0) * op.ld + col_idx;
  if (op.vec) {
    const bool ok = row_ok && col_idx < col_lim;
    cp_async16(d, ok ? p : op.base, ok);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool ok = row_ok && col_idx + e < col_lim;
      cp_async4(d + e, ok ? p + e : op.base, ok);
    }
  }
}

// The update kernel runs ~20 different stages back to back and every tile walks its prologue and epilogue
// exactly once, so the instruction footprint (unique 128-byte lines fetched, not instructions executed)
// decides how long the once-per-tile code takes: epilogues are ROLLED loops with one copy of this switch.
template <int EPI>
__device__ __forceinline__ float epi_apply(float v, float bias, float aux, float f0, float f1, float& aux_out) {
  if (EPI == EPI_BIAS) return v + bias;
  if (EPI == EPI_BIAS_RELU) return fmaxf(v + bias, 0.f);
  if (EPI == EPI_BIAS_TANH) {
    const float y = tanhf(v + bias);
    aux_out = y;                       // stored into aux0 by the caller
    return f0 * y;
  }
  if (EPI == EPI_BIAS_TANH_NOISE) {
    float a = f0 * tanhf(v + bias) + aux;
    if (f1 > 0.f) a = fminf(fmaxf(a, -f1), f1);
    return a;
  }
  if (EPI == EPI_RELU_MASK) return aux > 0.f ? v : 0.f;
  if (EPI == EPI_TANH_GRAD) return v * f0 * (1.f - aux * aux);
  return v;
}

// The epilogue kind is dispatched ONCE per tile (a switch per element cost ~400 cycles each on B200: an indirect
// branch over the two tanhf bodies); the per-kind code below is straight-line.
#define TD3_DISPATCH_EPI(epi, CALL)                                        \
  switch (epi) {                                                           \
    case EPI_BIAS: { constexpr int E = EPI_BIAS; CALL; } break;            \
    case EPI_BIAS_RELU: { constexpr int E = EPI_BIAS_RELU; CALL; } break;  \
    case EPI_BIAS_TANH: { constexpr int E = EPI_BIAS_TANH; CALL; } break;  \
    case EPI_BIAS_TANH_NOISE: { constexpr int E = EPI_BIAS_TANH_NOISE; CALL; } break; \
    case EPI_RELU_MASK: { constexpr int E = EPI_RELU_MASK; CALL; } break;  \
    case EPI_TANH_GRAD: { constexpr int E = EPI_TANH_GRAD; CALL; } break;  \
    default: { constexpr int E = EPI_STORE; CALL; } break;                 \
  }

// 8 reduction steps (this thread's k-group share of a 32-wide chunk) of the 4x4 micro-tile
template <bool ARC, bool BRC>
__device__ __forceinline__ void fma_block(const float* __restrict__ As, const float* __restrict__ Bs, int kg, int ti,
                                          int tj, float (&acc)[4][4], float (&asum)[4], bool want_rowsum) {
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
}

template <int EPI>
__device__ __forceinline__ void ffma_epilogue(const Problem& P, float* __restrict__ C, float* aux0, const float* red,
                                              const float* aux_s, const float* bias_s, bool aux_read, int i0, int j0, int il,
                                              int ci, int t64, int tj, int brc) {
  const bool has_bias = P.bias != nullptr;
  const int i = i0 + il;
#pragma unroll
  for (int cj = 0; cj < 4; ++cj) {
    const int e = ci * 4 + cj;
    float v = (red[(0 * 16 + e) * 64 + t64] + red[(1 * 16 + e) * 64 + t64]) +
              (red[(2 * 16 + e) * 64 + t64] + red[(3 * 16 + e) * 64 + t64]);
    const int jl = brc ? tj + 8 * cj : tj * 4 + cj;
    const int j = j0 + jl;
    if (i < P.M && j < P.N) {
      float aux_out = 0.f;
      v = epi_apply<EPI>(v, has_bias ? bias_s[jl] : 0.f, aux_read ? aux_s[il * kLd + jl] : 0.f, P.f0, P.f1, aux_out);
      if (EPI == EPI_BIAS_TANH) aux0[(size_t)i * P.ldaux + j] = aux_out;
      if (P.rn_out) v = rn_tf32(v);
#pragma unroll 1
      for (int d = 0; d < P.c_dups; ++d) C[d * P.c_dup_stride + (size_t)i * P.ldc + j] = v;
    }
  }
}

// ------------------------------------------------------------------------------------
// One 32x32 output tile; 256 threads = 4 k-groups x (8x8 threads x 4x4 micro-tile).
// Operand chunks stream through a kRing-deep cp.async ring so that up to kRing-1 chunks
// (the whole reduction for K <= 224, most of it for the 400/500-wide layers) are in flight
// at once: a tile costs about one L2 round trip plus its FFMA time instead of one round
// trip per chunk.  One copy of everything but the 64-FFMA block (4 operand-orientation
// variants, selected per chunk by a uniform branch).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void gemm_tile(const Problem& P, int tile, float* smem) {
  const int tid = threadIdx.x;
  const int kg = tid >> 6, t64 = tid & 63, ti = t64 >> 3, tj = t64 & 7;
  TP(0);

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
  const int arc = P.a_rc, brc = P.b_rc, variant = arc * 2 + brc;

  Operand opA{P.A + go * P.a_go + gi * P.a_gi, P.lda, arc, P.a_vec, i0, P.M};
  Operand opB{P.B + go * P.b_go + gi * P.b_gi, P.ldb, brc, P.b_vec, j0, P.N};
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

  // epilogue operands ride in the first commit group: the bias strip and the 32x32 tile of aux0 (ReLU mask /
  // noise / tanh output) are in shared memory long before the epilogue wants them
  const int epi = P.epi;
  float* aux_s = smem + kRingFloats;
  float* bias_s = aux_s + kTileFloats;
  float* aux0 = P.aux0 ? P.aux0 + go * P.aux0_go + gi * P.aux0_gi : nullptr;
  const bool aux_read = aux0 && (epi == EPI_BIAS_TANH_NOISE || epi == EPI_RELU_MASK || epi == EPI_TANH_GRAD);
  if (aux_read) {
    Operand opX{aux0, P.ldaux, 1, 0, i0, P.M};
    issue_chunk(aux_s, opX, j0, P.N, tid);
  }
  if (P.bias && tid < 32) {
    const float* bias = P.bias + go * P.bias_go + gi * P.bias_gi;
    const bool ok = j0 + tid < P.N;
    cp_async4(bias_s + tid, ok ? bias + j0 + tid : bias, ok);
  }
  TP(1);
  // c < 0: prologue (kRing-1 commit groups; empty ones keep the group arithmetic uniform)
#pragma unroll 1
  for (int c = -(kRing - 1); c < n_chunks; ++c) {
    if (c >= 0) {
      cp_async_wait<kRing - 2>();   // this thread's copies of chunk c have landed
      __syncthreads();              // ... everyone's have, and chunk c-1 has been consumed by all warps
      if (c == 0) TP(2);
    }
    const int cn = c + kRing - 1;   // refill the slot chunk c-1 occupied
    if (cn < n_chunks) {
      float* Ns = smem + (cn % kRing) * 2 * kTileFloats;
      issue_chunk(Ns, opA, k_begin + cn * kBK, k_end, tid);
      issue_chunk(Ns + kTileFloats, opB, k_begin + cn * kBK, k_end, tid);
    }
    cp_async_commit();
    if (c >= 0) {
      const float* As = smem + (c % kRing) * 2 * kTileFloats;
      const float* Bs = As + kTileFloats;
      switch (variant) {
        case 3: fma_block<true, true>(As, Bs, kg, ti, tj, acc, asum, want_rowsum); break;
        case 2: fma_block<true, false>(As, Bs, kg, ti, tj, acc, asum, want_rowsum); break;
        case 1: fma_block<false, true>(As, Bs, kg, ti, tj, acc, asum, want_rowsum); break;
        default: fma_block<false, false>(As, Bs, kg, ti, tj, acc, asum, want_rowsum); break;
      }
    }
  }
  TP(3);
  cp_async_wait<0>();
  __syncthreads();                  // all warps are done with the ring: reuse it for the reduction
  TP(4);

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

  {
    float* __restrict__ C = P.C + go * P.c_go + gi * P.c_gi + (long long)ks * P.c_split;
    const int ci = kg;
    const int il = arc ? ti + 8 * ci : ti * 4 + ci;      // row / column of this thread's outputs inside the tile
    TD3_DISPATCH_EPI(epi, (ffma_epilogue<E>(P, C, aux0, red, aux_s, bias_s, aux_read, i0, j0, il, ci, t64, tj, brc)));
  }
  if (want_rowsum && tid < 32) {
    const int rti = tid >> 2, rc = tid & 3;
    const float s = (sdb[(0 * 8 + rti) * 4 + rc] + sdb[(1 * 8 + rti) * 4 + rc]) +
                    (sdb[(2 * 8 + rti) * 4 + rc] + sdb[(3 * 8 + rti) * 4 + rc]);
    const int row = i0 + (arc ? rti + 8 * rc : rti * 4 + rc);
    if (row < P.M) aux1[(long long)ks * P.M + row] = s;
  }
  TP(5);
  __syncthreads();                  // the next tile of this CTA refills the ring
  TP(6);
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
  const int o = g / P.groups_inner;      // 32-bit division: the 64-bit one is a ~200-instruction routine
  go = o;
  gi = g - o * P.groups_inner;
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
  #pragma unroll 2
  for (int j = lane; j < P.N; j += 32) s += x[j];
  const float mean = warp_sum(s) / (float)P.N;
  float ss = 0.f;
  #pragma unroll 2
  for (int j = lane; j < P.N; j += 32) {
    const float d = x[j] - mean;
    ss = fmaf(d, d, ss);
  }
  const float var = warp_sum(ss) / (float)P.N;
  const float rstd = 1.f / sqrtf(var + P.f0);
  #pragma unroll 2
  for (int j = lane; j < P.N; j += 32) {
    const float v = (x[j] - mean) * rstd * gamma[j] + beta[j];
    y[j] = P.rn_out ? rn_tf32(v) : v;
  }
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
  #pragma unroll 2
  for (int j = lane; j < P.N; j += 32) {
    const float gdy = dy[j] * gamma[j];
    const float xh = (x[j] - mean) * rstd;
    s1 += gdy;
    s2 = fmaf(gdy, xh, s2);
  }
  s1 = warp_sum(s1) / (float)P.N;
  s2 = warp_sum(s2) / (float)P.N;
  #pragma unroll 2
  for (int j = lane; j < P.N; j += 32) {
    const float xv = x[j];
    const float gdy = dy[j] * gamma[j];
    const float xh = (xv - mean) * rstd;
    float d = rstd * (gdy - s1 - xh * s2);
    if (P.epi == EPI_RELU_MASK && !(xv > 0.f)) d = 0.f;
    dx[j] = P.rn_out ? rn_tf32(d) : d;
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
    #pragma unroll 4
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
  __syncthreads();                  // scratch is reused by this CTA's next tile
}

// ------------------------------------------------------------------------------------
// Particle pooling (TD3_particles.py:57-58): C[b, c] = relu(mean_n A[(b*K + n), c]).
//   M = batch, N = channels, K = particles per sample; tile = 8 samples? -> one CTA per
//   (sample, 32-channel strip); 8 row-lanes walk the particles.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void pool_fwd_tile(const Problem& P, int tile, float* smem) {
  // one CTA per sample; a thread owns 4 consecutive channels (one 16-byte load per particle row), 1024 / N row-lanes walk
  // the particles with 8 loads in flight each: the stage is bound by reading h2 once from HBM
  const int g = tile / P.tiles_per_group;
  const int b = tile - g * P.tiles_per_group;
  long long go, gi;
  group_ptrs(P, g, go, gi);
  const int N4 = P.N >> 2;                                 // channel quads (N <= 1024, N % 4 == 0)
  const int lanes = max(1, kStageThreads / N4);
  const int c4 = threadIdx.x % N4, rl = threadIdx.x / N4;
  const float* h = P.A + go * P.a_go + gi * P.a_gi + (size_t)b * P.K * P.lda + c4 * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < lanes) {
    int n = rl;
    for (; n + 7 * lanes < P.K; n += 8 * lanes) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4*>(h + (size_t)(n + u * lanes) * P.lda);
#pragma unroll
      for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    for (; n < P.K; n += lanes) {
      const float4 v = *reinterpret_cast<const float4*>(h + (size_t)n * P.lda);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  reinterpret_cast<float4*>(smem)[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < N4) {
    float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < lanes; ++r) {
      const float4 v = reinterpret_cast<float4*>(smem)[r * N4 + c4];
      tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
    }
    const float inv = 1.f / (float)P.K;
    float4 o = make_float4(fmaxf(tot.x * inv, 0.f), fmaxf(tot.y * inv, 0.f), fmaxf(tot.z * inv, 0.f), fmaxf(tot.w * inv, 0.f));
    if (P.rn_out) o = make_float4(rn_tf32(o.x), rn_tf32(o.y), rn_tf32(o.z), rn_tf32(o.w));
    float* C = P.C + go * P.c_go + gi * P.c_gi;
    for (int d = 0; d < P.c_dups; ++d) {
      float* cp = C + d * P.c_dup_stride + (size_t)b * P.ldc + c4 * 4;
      cp[0] = o.x; cp[1] = o.y; cp[2] = o.z; cp[3] = o.w;
    }
  }
  __syncthreads();
}

// dH2[(b*K+n), c] = (A[b,c] / K) * (aux0[b,c] > 0) * (aux1[(b*K+n), c] > 0)
//   A = d(pool out) [M, lda], aux0 = pool out [M, ldaux], aux1 = h2 [M*K, N] (ld = ldb), C = dH2 [M*K, ldc]
//   tile = 8 particle rows x all channels
constexpr int kPoolBwdRows = 256;    // rows per tile: 128 KB read + 128 KB written per CTA (64-row tiles spent their time being launched)
__device__ __forceinline__ void pool_bwd_tile(const Problem& P, int tile) {
  // tile = kPoolBwdRows particle rows; a thread owns 4 channels (16-byte accesses) and walks the rows
  const int g = tile / P.tiles_per_group;
  const int t = tile - g * P.tiles_per_group;
  long long go, gi;
  group_ptrs(P, g, go, gi);
  const int N4 = P.N >> 2;
  const int lanes = max(1, kStageThreads / N4);
  const int c4 = threadIdx.x % N4, rl = threadIdx.x / N4;
  const long long row0 = (long long)t * kPoolBwdRows;
  if (rl >= lanes) return;
  const float* dpb = P.A + go * P.a_go + gi * P.a_gi + c4 * 4;
  const float* poolb = P.aux0 + go * P.aux0_go + gi * P.aux0_gi + c4 * 4;
  const float inv = 1.f / (float)P.K;
  int b_cur = -1;
  float gsc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* h2 = P.aux1 + go * P.aux1_go + gi * P.aux1_gi + c4 * 4;
  float* out = P.C + go * P.c_go + gi * P.c_gi + c4 * 4;
#pragma unroll 8
  for (int r = rl; r < kPoolBwdRows; r += lanes) {
    const long long row = row0 + r;
    if (row >= (long long)P.M * P.K) break;
    const int b = (int)(row / P.K);                         // a tile may straddle samples when 64 does not divide N
    if (b != b_cur) {
      b_cur = b;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        gsc[e] = poolb[(size_t)b * P.ldaux + e] > 0.f ? dpb[(size_t)b * P.lda + e] * inv : 0.f;
        if (P.rn_out) gsc[e] = rn_tf32(gsc[e]);
      }
    }
    const float4 hv = *reinterpret_cast<const float4*>(h2 + (size_t)row * P.ldb);
    float4 o;
    o.x = hv.x > 0.f ? gsc[0] : 0.f; o.y = hv.y > 0.f ? gsc[1] : 0.f;
    o.z = hv.z > 0.f ? gsc[2] : 0.f; o.w = hv.w > 0.f ? gsc[3] : 0.f;
    *reinterpret_cast<float4*>(out + (size_t)row * P.ldc) = o;
  }
}

// C[e] = sum_s A[s*c_split + e] for e in tile of kReduceTile elements (M = element count, K = splits); a thread sums its
// elements' partials in split order (fixed: deterministic) with the loads of eight splits in flight
constexpr int kReduceTile = 512;
__device__ __forceinline__ void reduce_splits_tile(const Problem& P, int tile) {
  const int g = tile / P.tiles_per_group;
  const int t = tile - g * P.tiles_per_group;
  long long go, gi;
  group_ptrs(P, g, go, gi);
  const float* part = P.A + go * P.a_go + gi * P.a_gi;
  float* out = P.C + go * P.c_go + gi * P.c_gi;
  constexpr int kPer = kReduceTile / kStageThreads;
  const int e0 = t * kReduceTile + threadIdx.x;
  float s[kPer];
#pragma unroll
  for (int u = 0; u < kPer; ++u) s[u] = 0.f;
  const float* p = part + e0;
#pragma unroll 8
  for (int k = 0; k < P.K; ++k, p += P.c_split) {
#pragma unroll
    for (int u = 0; u < kPer; ++u)
      if (e0 + u * kStageThreads < P.M) s[u] += p[u * kStageThreads];
  }
#pragma unroll
  for (int u = 0; u < kPer; ++u)
    if (e0 + u * kStageThreads < P.M) out[e0 + u * kStageThreads] = s[u];
}

// Particle-encoder layer 1 (TD3_particles.py:29,54: Conv2d(1,256,(1,D)) == a D -> 256 linear layer per particle).  The
// reduction is D <= 8 long: a GEMM tile would spend its time on overhead, and the layer is bound by writing its
// [B*N, 256] output.  Tile = 128 particles; thread = output channel (weights in registers), particles from shared memory.
__device__ __forceinline__ void smallk_fwd_tile(const Problem& P, int tile, float* smem) {
  const int g = tile / P.tiles_per_group;
  const int t = tile - g * P.tiles_per_group;
  long long go, gi;
  group_ptrs(P, g, go, gi);
  const float* A = P.A + go * P.a_go + gi * P.a_gi;
  const float* W = P.B + go * P.b_go + gi * P.b_gi;
  const float* bias = P.bias + go * P.bias_go + gi * P.bias_gi;
  float* C = P.C + go * P.c_go + gi * P.c_gi;
  const int K = P.K, r0 = t * 128, nr = min(128, P.M - r0);
  for (int e = threadIdx.x; e < nr * K; e += kStageThreads) {
    const int r = e / K, d = e - r * K;
    smem[r * 8 + d] = A[(size_t)(r0 + r) * P.lda + d];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < P.N; c += kStageThreads) {
    float w[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) w[d] = d < K ? W[(size_t)c * P.ldb + d] : 0.f;
    const float b = bias[c];
#pragma unroll 8
    for (int r = 0; r < nr; ++r) {
      float v = b;
#pragma unroll
      for (int d = 0; d < 8; ++d)
        if (d < K) v = fmaf(smem[r * 8 + d], w[d], v);
      v = fmaxf(v, 0.f);
      C[(size_t)(r0 + r) * P.ldc + c] = P.rn_out ? rn_tf32(v) : v;
    }
  }
  __syncthreads();
}

// Weight gradient of the same layer: dW1[c, d] = sum_r dH1[r, c] P[r, d], db1[c] = sum_r dH1[r, c], reduction over the
// B*N particles split over `ksplit` CTAs (partials at C + ks*c_split and aux1 + ks*M, reduced by PK_REDUCE_SPLITS).
//   A = dH1 [K rows, lda] (M = channels), B = P [K rows, ldb] (N = D <= 8 columns)
constexpr int kSmallkDwRows = 1024;   // particle rows staged per pass (32 KB of shared memory, columns D..7 zero)
__device__ __forceinline__ void smallk_dw_tile(const Problem& P, int tile, float* smem) {
  const int g = tile / P.tiles_per_group;
  const int ks = tile - g * P.tiles_per_group;
  long long go, gi;
  group_ptrs(P, g, go, gi);
  const float* A = P.A + go * P.a_go + gi * P.a_gi;
  const float* Bm = P.B + go * P.b_go + gi * P.b_gi;
  float* C = P.C + go * P.c_go + gi * P.c_gi + (long long)ks * P.c_split;
  float* db = P.aux1 + go * P.aux1_go + gi * P.aux1_gi + (long long)ks * P.M;
  const int D = P.N;
  const int per = (P.K + P.ksplit - 1) / P.ksplit;
  const int r_begin = ks * per, r_end = min(P.K, r_begin + per);
  float acc[8], accb = 0.f;
#pragma unroll
  for (int d = 0; d < 8; ++d) acc[d] = 0.f;
  const int c = threadIdx.x;            // channel (M <= 256)
  for (int e = threadIdx.x; e < kSmallkDwRows * 8; e += kStageThreads) smem[e] = 0.f;      // columns D..7 stay zero
#pragma unroll 1
  for (int rb = r_begin; rb < r_end; rb += kSmallkDwRows) {
    const int nr = min(kSmallkDwRows, r_end - rb);
    __syncthreads();
    for (int e = threadIdx.x; e < nr * D; e += kStageThreads) {
      const int r = e / D, d = e - r * D;
      smem[r * 8 + d] = Bm[(size_t)(rb + r) * P.ldb + d];
    }
    __syncthreads();
    if (c < P.M) {
      const float* ap = A + (size_t)rb * P.lda + c;
#pragma unroll 1
      for (int r0 = 0; r0 < nr; r0 += 32) {
        float a[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) a[u] = r0 + u < nr ? ap[(size_t)(r0 + u) * P.lda] : 0.f;     // 32 rows in flight per thread
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          accb += a[u];
          const float4 p0 = *reinterpret_cast<const float4*>(smem + (r0 + u) * 8), p1 = *reinterpret_cast<const float4*>(smem + (r0 + u) * 8 + 4);
          acc[0] = fmaf(a[u], p0.x, acc[0]); acc[1] = fmaf(a[u], p0.y, acc[1]); acc[2] = fmaf(a[u], p0.z, acc[2]); acc[3] = fmaf(a[u], p0.w, acc[3]);
          acc[4] = fmaf(a[u], p1.x, acc[4]); acc[5] = fmaf(a[u], p1.y, acc[5]); acc[6] = fmaf(a[u], p1.z, acc[6]); acc[7] = fmaf(a[u], p1.w, acc[7]);
        }
      }
    }
  }
  if (c < P.M) {
#pragma unroll
    for (int d = 0; d < 8; ++d)
      if (d < D) C[(size_t)c * P.ldc + d] = acc[d];
    db[c] = accb;
  }
  __syncthreads();
}

// C[ks*c_split + j] = sum over this slice's rows i of A[i, j]  (A [K rows, lda], N columns; tile = 32-column strip)
__device__ __forceinline__ void colsum_tile(const Problem& P, int tile, float* smem) {
  const int g = tile / P.tiles_per_group;
  int t = tile - g * P.tiles_per_group;
  long long go, gi;
  group_ptrs(P, g, go, gi);
  const int ks = t / P.tiles_n, strip = t - ks * P.tiles_n;
  const int rl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = strip * 32 + lane;
  const int per = (P.K + P.ksplit - 1) / P.ksplit;
  const int r0 = ks * per, r1 = min(P.K, r0 + per);
  const float* a = P.A + go * P.a_go + gi * P.a_gi;
  // 16 rows in flight per thread, four independent partial sums (fixed order): with long reductions (the particle
  // encoder's bias gradients: 2048 rows per slice) a 4-deep dependent chain of loads was a fifth of the whole stage
  float s = 0.f;
  if (j < P.N) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* ap = a + (size_t)(r0 + rl) * P.lda + j;
    const size_t step = (size_t)8 * P.lda;
    int n = (r1 - r0 - rl + 7) / 8;
#pragma unroll 1
    for (; n >= 16; n -= 16, ap += 16 * step) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = ap[u * step];
#pragma unroll
      for (int u = 0; u < 16; ++u) acc[u & 3] += v[u];
    }
    for (int u = 0; u < n; ++u) acc[u & 3] += ap[u * step];
    s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  }
  smem[rl * 32 + lane] = s;
  __syncthreads();
  if (rl == 0 && j < P.N) {
    float tot = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) tot += smem[r * 32 + lane];
    (P.C + go * P.c_go + gi * P.c_gi)[(long long)ks * P.c_split + j] = tot;
  }
  __syncthreads();
}

// scalar C[g] = f0 * mean(A[0..M*N)) with A [M, lda]; one CTA per group
__device__ __forceinline__ void neg_mean_tile(const Problem& P, int tile, float* smem) {
  long long go, gi;
  group_ptrs(P, tile, go, gi);
  const float* a = P.A + go * P.a_go + gi * P.a_gi;
  float s = 0.f;
  const int total = P.M * P.N;
  #pragma unroll 1
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
  __syncthreads();
}

}  // namespace td3
