// chain.cuh -- layer-fused update chains for the plain-MLP TD3_featured networks (TD3_featured.py:123-171).
//
// Everything in a TD3 update that touches ONE batch row at a time -- sampling, the target actor, the smoothing noise,
// the twin target critics, the Bellman target, the online critics' forward pass, the loss gradient and the whole
// dX chain of the backward pass -- is independent across batch rows.  As stage-per-layer launches the update is a chain
// of 7 (critic-only) or 14 (policy) dependent launches of 6-8 us each whatever they contain.  Here
// one CTA owns a 64-row tile of the batch and walks a whole chain of layers with the activations never leaving the SM:
//
//   A operand   the tile's activations [64 x K], K-major / 128-byte swizzle in shared memory, written by the epilogue of
//               the previous layer (rounded to nearest TF32)
//   B operand   the layer's weights, streamed ONCE per CTA through a ring of shared-memory slots by TMA from the
//               round-to-nearest TF32 shadow of the packed parameters (3-D tensor maps: k, n, network); forward layers
//               read W[N][K] K-major, backward layers read the same matrix MN-major (dX = dZ . W, no transpose copy);
//               the K <= 32 first layer (92-byte rows: not TMA-addressable) is staged by the producer warp with cp.async
//   D           tcgen05.mma.cta_group::1.kind::tf32 M64 N<=256 K8 into a 512-column TMEM allocation (row m of the tile
//               lives in TMEM lane 32 (m / 16) + m % 16)
//   epilogue    8 warps: tcgen05.ld -> bias / ReLU / ReLU mask / tanh ... -> next layer's A operand (+ the copies the
//               weight-gradient stage needs in global memory: hidden activations and dZ of the online networks)
//
// Roles (one CTA = one role x agent x twin x 64-row tile; launch order = dependency order):
//   TARGET  sample s' -> target actor -> clipped noise, clamp -> target critic twin t -> Q_t'(s', a')   -> tq + flag
//   ACTOR   (policy steps) sample s -> online actor forward -> pi(s), tanh, hidden activations
//   CRITIC  sample (s, a, r, nd) -> online critic twin t forward -> Q_t; waits for both twins' tq flags of its tile ->
//           y = r + nd * discount * min(Q1', Q2'), loss term, dQ -> backward through the hidden layers -> dZ_l
//   POLICY  (second launch of a policy step, after the critic's optimiser) [s, pi(s)] -> Q1 forward with the stepped
//           critic -> -mean -> backward to the action -> through tanh -> backward through the actor -> dZ_l
// The reductions over the batch (dW = dZ^T . H, db) stay with the existing tensor-core stage and the fused
// first-layer-gradient + Adam launch (apply.cuh): a critic-only update is 3 launches, a policy update 6.
// The last CTA to finish sums the loss terms in a fixed order, mirrors the loss to the host, advances the sampling
// step and the optimiser's bias-correction scalars (what head_body did).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc.cuh"

namespace td3 {

constexpr int kChRows = 64;
constexpr int kChWorkers = 8;                          // epilogue / staging warps
constexpr int kChWorkerThreads = kChWorkers * 32;
constexpr int kChThreads = (kChWorkers + 2) * 32;      // + TMA producer warp + MMA warp
constexpr int kChMaxSteps = 20;
constexpr int kChMaxMaps = 20;
constexpr int kChMaxRoles = 4;
constexpr int kChMaxW = 512;                           // hidden width (TMEM columns)
constexpr int kChMaxSlots = 6;
constexpr int kChMaxMask = 3;                          // hidden layers per network
constexpr int kChX0Bytes = kChRows * 128;              // [64 x 32] input operand
constexpr int kChMaskBytes = kChRows * 16 * 4;         // one hidden layer's ReLU mask: 512 bits per row

enum ChainRole : int { CR_TARGET = 0, CR_ACTOR = 1, CR_CRITIC = 2, CR_POLICY = 3 };
enum ChainB : int { CB_TMA_K = 0, CB_TMA_MN = 1, CB_MANUAL_K = 2, CB_MANUAL_MN = 3 };
enum ChainEpi : int { CE_HIDDEN = 0, CE_ACTOR_OUT = 1, CE_BWD_MASK = 2, CE_DX_TANH = 3 };
enum ChainPost : int { CP_NONE = 0, CP_TQ = 1, CP_CRITIC = 2, CP_Q1 = 3 };

struct ChainStep {
  const float* P;               // packed master parameters of the step's network family (biases, head weights)
  const float* Psh;             // their TF32 shadow (manually staged weights)
  long long p_go, p_gi;         // strides: agent, twin
  long long w_off, bias_off;    // weight matrix / bias inside a network's block (bias_off < 0: none)
  long long head_w_off, head_b_off;   // post != CP_NONE: the Q head's weight row and bias
  float* out; long long out_go, out_gi;         // global copy of the epilogue's result [B][ld_out] (or nullptr)
  float* out2; long long out2_go, out2_gi;      // post CP_CRITIC: dZ of the last hidden layer [B][ld_out]
  const float* gmask; long long gmask_go, gmask_gi;   // CE_BWD_MASK: ReLU mask source in global memory (> 0), or nullptr
  int K, N;                     // reduction length, output columns
  int b_mode, map, z_o, z_i;    // operand staging; tensor map; network index = agent * z_o + twin * z_i
  int a_x0;                     // A operand: 1 = the chain input X0, 0 = the activation buffer
  int epi, post;
  int mask_w, mask_r;           // shared-memory mask slot written (CE_HIDDEN) / read (CE_BWD_MASK, post), -1: none
  int ld_out, ld_w, w_col0;     // manual modes: row stride of the weight matrix, first column (CB_MANUAL_MN)
  int nb, n_blk;                // output columns per MMA / number of column blocks
  int cps, blk_bytes;           // 32-step chunks per ring-slot use; bytes of one chunk of one column block
  int flavour;                  // CE_ACTOR_OUT: 0 = target (noise + clamp, into X0), 1 = online actor (global only)
};

struct ChainParams {
  GatherParams g;               // replay view, index / noise source (misc.cuh)
  int batch, n_agents, tiles, S, A, n_q, ld_q, ld_tanh;
  int n_roles, n_ctas;
  int role_kind[kChMaxRoles], role_first[kChMaxRoles + 1], role_step0[kChMaxRoles], role_nsteps[kChMaxRoles],
      role_inner[kChMaxRoles];
  int act_bytes, slot_bytes, n_slots, n_mask;
  float* xq; long long xq_go;           // [B][ld_q] online critic input [s | a]
  float* xq2; long long xq2_go;         // target critic input [s' | a']
  float* xpi; long long xpi_go;         // [s | pi(s)]
  float* r; float* nd; long long r_go;
  float* tanh_y; long long tanh_go;     // [B][A]
  float* q; float* tq; float* dq; long long q_go, q_gi;      // [B] per (agent, twin)
  float* y; long long y_go;
  float* q_pi; long long qpi_go;
  float* loss_part; long long lp_go;    // [n_agents][tiles * n_q]
  float* loss;                          // [n_agents] (critic) / + n_agents (actor)
  unsigned int* flags; long long flag_go;   // [n_agents][tiles][n_q] epoch of the last published tq
  unsigned int* epoch;                  // launches completed
  unsigned int* done;                   // CTAs finished in this launch
  unsigned long long* host_status; unsigned int* seq;
  AdamTick tick;
  long long* prof;                      // debug: clock stamps of tile 0 / agent 0 / twin 0 of every role, or nullptr
  int fin_mode, pad0;                   // 0: critic launch (loss, host mirror, sampling step, critic tick); 1: policy launch
  float discount, inv_norm, max_action, clamp_action;
  ChainStep steps[kChMaxSteps];
  TensorMapBlob maps[kChMaxMaps];
};

__device__ __forceinline__ void ch_named_barrier(int id, int threads) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void ch_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned int dst, const void* tmap, int c0, int c1, int c2, unsigned int bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
               "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void ch_tmem_ld32(unsigned int taddr, unsigned int (&r)[32]) {
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
// instruction descriptor: D fp32, A / B tf32, M = 64, N = n, A K-major, B K-major (0) or MN-major (1)
__device__ __forceinline__ unsigned int ch_idesc(int n, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)b_mn_major << 16) | ((unsigned)(n >> 3) << 17) | ((64u >> 4) << 24);
}
// element (row, col) of a [rows x 32] K-major / 128-byte-swizzle operand chunk (byte offset inside the chunk)
__device__ __forceinline__ int ch_kmajor_off(int row, int col) { return row * 128 + ((((col >> 2) ^ (row & 7))) << 4) + (col & 3) * 4; }

// 32 columns of one row -> chunk `g` of the activation operand (8 swizzled 16-byte granules)
__device__ __forceinline__ void ch_store_act(unsigned char* act, int g, int row, const float (&v)[32]) {
  unsigned char* arow = act + g * (kChRows * 128) + row * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(arow + ((j ^ (row & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void ch_store_global(float* dst, int col0, int n, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 4)
    if (col0 + j < n) *reinterpret_cast<float4*>(dst + col0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
}

__global__ void __launch_bounds__(kChThreads, 1) chain_kernel(const __grid_constant__ ChainParams C) {
  extern __shared__ unsigned char ch_smem_raw[];
  __shared__ unsigned long long full_bar[kChMaxSlots], empty_bar[kChMaxSlots], a_ready, acc_full;
  __shared__ unsigned int tmem_base_s, epoch_s;
  __shared__ int s_last;
  unsigned char* base = ch_smem_raw + ((1024u - (smem_u32(ch_smem_raw) & 1023u)) & 1023u);
  unsigned char* X0 = base;
  unsigned char* ACT = X0 + kChX0Bytes;
  unsigned char* RING = ACT + C.act_bytes;
  unsigned int* mask_s = reinterpret_cast<unsigned int*>(RING + C.n_slots * C.slot_bytes);    // [n_mask][64][16]
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(mask_s) + C.n_mask * kChMaskBytes);   // [512]
  float* headw_s = bias_s + kChMaxW;                                                         // [512]
  float* red_s = headw_s + kChMaxW;                                                          // [2][64]
  float* rnd_s = red_s + 2 * kChRows;                                                        // [2][64] reward, not_done
  float* lred_s = rnd_s + 2 * kChRows;                                                       // [64]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- which chain is this CTA ----
  int role = 0;
  while (role + 1 < C.n_roles && (int)blockIdx.x >= C.role_first[role + 1]) ++role;
  const int kind = C.role_kind[role], inner = C.role_inner[role];
  int local = (int)blockIdx.x - C.role_first[role];
  const int agent = local / (inner * C.tiles);
  local -= agent * inner * C.tiles;
  const int tile = local / inner, twin = local - tile * inner;
  const int s0 = C.role_step0[role], nsteps = C.role_nsteps[role];
  const int B = C.batch, S = C.S, A = C.A;

  pdl_launch_dependents();
  pdl_wait();
  if (tid == 0) {
    for (int i = 0; i < C.n_slots; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&a_ready, kChWorkers);
    mbar_init(&acc_full, 1);
    epoch_s = __ldcg(C.epoch) + 1u;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == kChWorkers + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned int tmem = *reinterpret_cast<volatile unsigned int*>(&tmem_base_s);
  const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(&epoch_s);

  if (warp < kChWorkers) {
    // =================================================================== workers: chain input, epilogues
    const int q = warp & 3, h = warp >> 2;
    const int row = 16 * q + (lane & 15);               // epilogue row of this thread (lanes 16-31: no accumulator row)
    const bool lane_ok = lane < 16;
    const int grow = tile * kChRows + row;
    const bool row_ok = lane_ok && grow < B;
    // ---- chain input: 4 threads per row, 8 columns each, into X0 (rounded) and the global network inputs ----
    {
      const int rl = tid >> 2, part = tid & 3;
      const int gr = tile * kChRows + rl;
      const bool ok = gr < B;
      float v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = 0.f;
      if (kind == CR_POLICY) {
        if (ok) {
          const float* src = C.xpi + (long long)agent * C.xpi_go + (long long)gr * C.ld_q;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int col = part * 8 + c;
            if (col < S + A) v[c] = __ldcg(src + col);
          }
        }
      } else if (ok) {
        const long long idx = gather_index(C.g, agent, gr);
        const float* src = C.g.rows + (long long)agent * C.g.rb_agent_stride + idx * C.g.row_stride;
        const int off = kind == CR_TARGET ? S + A : 0;
        const int ncol = kind == CR_CRITIC ? S + A : S;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int col = part * 8 + c;
          if (col < ncol) v[c] = __ldg(src + off + col);
        }
        if (twin == 0) {
          float* dst = kind == CR_TARGET ? C.xq2 + (long long)agent * C.xq2_go
                       : kind == CR_CRITIC ? C.xq + (long long)agent * C.xq_go
                                           : C.xpi + (long long)agent * C.xpi_go;
          dst += (long long)gr * C.ld_q;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int col = part * 8 + c;
            if (col < ncol) dst[col] = v[c];
          }
        }
        if (kind == CR_CRITIC && part == 0) {
          const float rv = __ldg(src + 2 * S + A), ndv = __ldg(src + 2 * S + A + 1);
          rnd_s[rl] = rv;
          rnd_s[kChRows + rl] = ndv;
          if (twin == 0) {
            C.r[(long long)agent * C.r_go + gr] = rv;
            C.nd[(long long)agent * C.r_go + gr] = ndv;
            C.g.idx_out[(long long)agent * B + gr] = idx;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = rn_tf32(v[c]);
      unsigned char* xr = X0 + rl * 128;
      *reinterpret_cast<float4*>(xr + (((2 * part) ^ (rl & 7)) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(xr + (((2 * part + 1) ^ (rl & 7)) << 4)) = make_float4(v[4], v[5], v[6], v[7]);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncwarp();
    if (lane == 0) ch_arrive(&a_ready);

    long long* prof = (C.prof && tid == 0 && agent == 0 && tile == 0 && twin == 0) ? C.prof + role * 48 : nullptr;
    if (prof) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(gt));
      prof[0] = (long long)gt;
      prof[1] = clock64();
    }
    float dqv = 0.f;                                     // dQ of this thread's row (CP_CRITIC / CP_Q1)
#pragma unroll 1
    for (int si = 0; si < nsteps; ++si) {
      const ChainStep& st = C.steps[s0 + si];
      const long long pofs = (long long)agent * st.p_go + (long long)twin * st.p_gi;
      const int N = st.N, ngroups = (N + 31) >> 5;
      // ---- this step's bias (and head weights) -> shared memory, while the MMAs run ----
      ch_named_barrier(1, kChWorkerThreads);             // the previous step's readers are done with bias_s / headw_s
      for (int j = tid; j < ngroups * 32; j += kChWorkerThreads) {
        bias_s[j] = (st.bias_off >= 0 && j < N) ? __ldg(st.P + pofs + st.bias_off + j) : 0.f;
        if (st.post != CP_NONE) headw_s[j] = j < N ? __ldg(st.P + pofs + st.head_w_off + j) : 0.f;
      }
      ch_named_barrier(1, kChWorkerThreads);
      mbar_wait(&acc_full, (unsigned)si & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (prof) prof[2 + 2 * si] = clock64();
      const unsigned int tlane = tmem + (((unsigned)q * 32u) << 16);

      if (st.epi == CE_HIDDEN) {
        float dot = 0.f;
        float* outr = (st.out && row_ok) ? st.out + (long long)agent * st.out_go + (long long)twin * st.out_gi + (long long)grow * st.ld_out : nullptr;
        unsigned int* mrow = st.mask_w >= 0 ? mask_s + st.mask_w * (kChRows * 16) + row * 16 : nullptr;
#pragma unroll 1
        const bool want_dot = st.post != CP_NONE;
        for (int g = h; g < ngroups; g += 2) {
          unsigned int r[32];
          ch_tmem_ld32(tlane + (unsigned)(g * 32), r);
          float v[32];
          const float4* b4 = reinterpret_cast<const float4*>(bias_s + g * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = b4[j];
            v[4 * j] = rn_tf32(fmaxf(__uint_as_float(r[4 * j]) + bb.x, 0.f));
            v[4 * j + 1] = rn_tf32(fmaxf(__uint_as_float(r[4 * j + 1]) + bb.y, 0.f));
            v[4 * j + 2] = rn_tf32(fmaxf(__uint_as_float(r[4 * j + 2]) + bb.z, 0.f));
            v[4 * j + 3] = rn_tf32(fmaxf(__uint_as_float(r[4 * j + 3]) + bb.w, 0.f));
          }
          if (g == ngroups - 1) {                        // columns past N: never written by this step's MMAs
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (g * 32 + j < N) ? v[j] : 0.f;
          }
          if (want_dot) {
            const float4* w4 = reinterpret_cast<const float4*>(headw_s + g * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 ww = w4[j];
              dot = fmaf(v[4 * j], ww.x, dot);
              dot = fmaf(v[4 * j + 1], ww.y, dot);
              dot = fmaf(v[4 * j + 2], ww.z, dot);
              dot = fmaf(v[4 * j + 3], ww.w, dot);
            }
          }
          if (lane_ok) {
            ch_store_act(ACT, g, row, v);
            if (mrow) {
              unsigned int bits = 0;
#pragma unroll
              for (int j = 0; j < 32; ++j) bits |= (v[j] > 0.f ? 1u : 0u) << j;
              mrow[g] = bits;
            }
            if (outr) ch_store_global(outr, g * 32, N, v);
          }
          __syncwarp();
        }
        if (st.post != CP_NONE) {
          // ---- Q head: the two column halves of a row meet in shared memory ----
          if (lane_ok) red_s[h * kChRows + row] = dot;
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          ch_named_barrier(1, kChWorkerThreads);
          const float qv = red_s[row] + red_s[kChRows + row] + __ldg(st.P + pofs + st.head_b_off);
          if (st.post == CP_TQ) {
            if (h == 0 && row_ok) C.tq[(long long)agent * C.q_go + (long long)twin * C.q_gi + grow] = qv;
            __threadfence();
            ch_named_barrier(1, kChWorkerThreads);
            if (tid == 0) {
              unsigned int* f = C.flags + (long long)agent * C.flag_go + tile * C.n_q + twin;
              asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(f), "r"(epoch) : "memory");
            }
          } else {
            float term = 0.f;
            if (st.post == CP_CRITIC) {
              if (tid == 0) {                            // both twins' target values of this tile have been published
                const unsigned int* f = C.flags + (long long)agent * C.flag_go + tile * C.n_q;
                for (int t = 0; t < C.n_q; ++t) {
                  unsigned int seen;
                  do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(f + t) : "memory");
                  } while (seen != epoch);
                }
              }
              ch_named_barrier(1, kChWorkerThreads);
              if (row_ok) {
                const long long qo = (long long)agent * C.q_go + grow;
                float tmin = __ldcg(C.tq + qo);
                if (C.n_q > 1) tmin = fminf(tmin, __ldcg(C.tq + qo + C.q_gi));
                const float yv = __fadd_rn(rnd_s[row], __fmul_rn(__fmul_rn(rnd_s[kChRows + row], C.discount), tmin));
                const float diff = qv - yv;
                dqv = 2.f * C.inv_norm * diff;
                term = diff * diff;
                if (h == 0) {
                  C.q[qo + (long long)twin * C.q_gi] = qv;
                  C.dq[qo + (long long)twin * C.q_gi] = rn_tf32(dqv);
                  if (twin == 0) C.y[(long long)agent * C.y_go + grow] = yv;
                }
              }
            } else {                                     // CP_Q1: actor loss = -mean Q1(s, pi(s))
              if (row_ok) {
                dqv = -C.inv_norm;
                term = qv;
                if (h == 0) C.q_pi[(long long)agent * C.qpi_go + grow] = qv;
              }
            }
            if (h == 0 && lane_ok) lred_s[row] = term;
            // ---- dZ of the last hidden layer: (dQ . w_head) gated by its ReLU -> the backward chain's first operand ----
            const unsigned int* mr = mask_s + st.mask_w * (kChRows * 16) + row * 16;
            float* o2 = (st.out2 && row_ok) ? st.out2 + (long long)agent * st.out2_go + (long long)twin * st.out2_gi + (long long)grow * st.ld_out : nullptr;
#pragma unroll 1
            for (int g = h; g < ngroups; g += 2) {
              if (lane_ok) {
                const unsigned int bits = mr[g];
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = ((bits >> j) & 1u) ? rn_tf32(dqv * headw_s[g * 32 + j]) : 0.f;
                ch_store_act(ACT, g, row, v);
                if (o2) ch_store_global(o2, g * 32, N, v);
              }
            }
            __syncwarp();
            ch_named_barrier(1, kChWorkerThreads);
            if (tid == 0) {                              // this CTA's loss term, rows in order
              float t = 0.f;
              for (int i = 0; i < kChRows; ++i) t += lred_s[i];
              C.loss_part[(long long)agent * C.lp_go + tile * inner + twin] = t;
            }
          }
        }
      } else if (st.epi == CE_BWD_MASK) {
        float* outr = (st.out && row_ok) ? st.out + (long long)agent * st.out_go + (long long)twin * st.out_gi + (long long)grow * st.ld_out : nullptr;
        const unsigned int* mr = st.mask_r >= 0 ? mask_s + st.mask_r * (kChRows * 16) + row * 16 : nullptr;
        const float* gm = (st.gmask && row_ok) ? st.gmask + (long long)agent * st.gmask_go + (long long)twin * st.gmask_gi + (long long)grow * N : nullptr;
#pragma unroll 1
        for (int g = h; g < ngroups; g += 2) {
          unsigned int r[32];
          ch_tmem_ld32(tlane + (unsigned)(g * 32), r);
          if (lane_ok) {
            unsigned int bits = mr ? mr[g] : 0u;
            if (gm) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (g * 32 + j < N) {
                  const float4 m4 = __ldcg(reinterpret_cast<const float4*>(gm + g * 32 + j));
                  bits |= (m4.x > 0.f ? 1u : 0u) << j | (m4.y > 0.f ? 1u : 0u) << (j + 1) | (m4.z > 0.f ? 1u : 0u) << (j + 2) |
                          (m4.w > 0.f ? 1u : 0u) << (j + 3);
                }
              }
            }
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ((bits >> j) & 1u) && (g * 32 + j < N) ? rn_tf32(__uint_as_float(r[j])) : 0.f;
            ch_store_act(ACT, g, row, v);
            if (outr) ch_store_global(outr, g * 32, N, v);
          }
          __syncwarp();
        }
      } else if (st.epi == CE_ACTOR_OUT) {
        if (h == 0) {
          unsigned int r[32];
          ch_tmem_ld32(tlane, r);
          if (row_ok) {
            const unsigned long long step = __ldcg(C.g.step_ptr);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (j >= A) break;
              const float ty = tanhf(__uint_as_float(r[j]) + bias_s[j]);
              if (st.flavour == 0) {
                const long long e = (long long)grow * A + j;
                float z = (C.g.rng_mode == 0)
                              ? philox_normal(C.g.seed + (unsigned long long)agent * 0x9E3779B97F4A7C15ull, PHILOX_NOISE, step,
                                              (uint32_t)(e + (long long)C.g.elem_offset * A))
                              : C.g.noise_in[(long long)agent * B * A + e];
                z = z * C.g.policy_noise;
                z = fminf(fmaxf(z, -C.g.noise_clip), C.g.noise_clip);
                float a = C.max_action * ty + z;
                if (C.clamp_action > 0.f) a = fminf(fmaxf(a, -C.clamp_action), C.clamp_action);
                *reinterpret_cast<float*>(X0 + ch_kmajor_off(row, S + j)) = rn_tf32(a);
                if (twin == 0) {
                  C.g.eps_out[(long long)agent * B * A + e] = z;
                  C.xq2[(long long)agent * C.xq2_go + (long long)grow * C.ld_q + S + j] = a;
                }
              } else {
                C.xpi[(long long)agent * C.xpi_go + (long long)grow * C.ld_q + S + j] = C.max_action * ty;
                C.tanh_y[(long long)agent * C.tanh_go + (long long)grow * C.ld_tanh + j] = ty;
              }
            }
          }
        }
      } else {   // CE_DX_TANH: d(action) through max_action * tanh -> dZ of the actor's output layer (K = 8 operand)
        if (h == 0) {
          unsigned int r8[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                       : "=r"(r8[0]), "=r"(r8[1]), "=r"(r8[2]), "=r"(r8[3]), "=r"(r8[4]), "=r"(r8[5]), "=r"(r8[6]), "=r"(r8[7])
                       : "r"(tlane + (unsigned)S)
                       : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          if (lane_ok) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (j < A) {
                  const float ty = __ldcg(C.tanh_y + (long long)agent * C.tanh_go + (long long)grow * C.ld_tanh + j);
                  v[j] = rn_tf32(__uint_as_float(r8[j]) * C.max_action * (1.f - ty * ty));
                }
              }
              if (st.out) {
                float* o = st.out + (long long)agent * st.out_go + (long long)grow * st.ld_out;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (j < A) o[j] = v[j];
              }
            }
            ch_store_act(ACT, 0, row, v);
          }
        }
      }
      if (prof) prof[3 + 2 * si] = clock64();
      // the next step's MMAs may read the operand / overwrite the accumulator
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0 && si + 1 < nsteps) ch_arrive(&a_ready);
    }
    if (prof) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(gt));
      prof[47] = (long long)gt;
    }
  } else if (warp == kChWorkers) {
    // =================================================================== producer: weights -> ring
    // one slot use = `cps` consecutive 32-step chunks of one column block (small blocks share a barrier round trip)
    unsigned int cnt = 0;
#pragma unroll 1
    for (int si = 0; si < nsteps; ++si) {
      const ChainStep& st = C.steps[s0 + si];
      const int nch = (st.K + 31) >> 5;
      const int z = agent * st.z_o + twin * st.z_i;
      const float* Wm = st.Psh + (long long)agent * st.p_go + (long long)twin * st.p_gi + st.w_off;
#pragma unroll 1
      for (int kc0 = 0; kc0 < nch; kc0 += st.cps) {
        const int kc1 = min(nch, kc0 + st.cps);
#pragma unroll 1
        for (int b = 0; b < st.n_blk; ++b, ++cnt) {
          const unsigned int slot = cnt % (unsigned)C.n_slots, use = cnt / (unsigned)C.n_slots;
          if (use > 0) mbar_wait(&empty_bar[slot], (use - 1u) & 1u);
          unsigned char* dst = RING + slot * C.slot_bytes;
          const unsigned int dst_u = __shfl_sync(0xffffffffu, smem_u32(dst), 0);
          const unsigned int fb = __shfl_sync(0xffffffffu, smem_u32(&full_bar[slot]), 0);
          if (st.b_mode == CB_TMA_K) {
            if (elect_one()) {
              asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(fb), "r"((unsigned)((kc1 - kc0) * st.blk_bytes)) : "memory");
              for (int kc = kc0; kc < kc1; ++kc) tma_load_3d(dst_u + (kc - kc0) * st.blk_bytes, &C.maps[st.map], kc * 32, b * st.nb, z, fb);
            }
            __syncwarp();
          } else if (st.b_mode == CB_TMA_MN) {
            const int ng = (st.nb + 31) >> 5;
            if (elect_one()) {
              asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(fb), "r"((unsigned)((kc1 - kc0) * st.blk_bytes)) : "memory");
              for (int kc = kc0; kc < kc1; ++kc)
                for (int g = 0; g < ng; ++g)
                  tma_load_3d(dst_u + (kc - kc0) * st.blk_bytes + g * 4096, &C.maps[st.map], b * st.nb + g * 32, kc * 32, z, fb);
            }
            __syncwarp();
          } else {
            // Manually staged first-layer weights W_0[N][K0] (K0 <= 32: 92-byte rows are not TMA-addressable).  The rows
            // of a block are one contiguous, 16-byte aligned run of floats: zero the destination, then 16-byte loads
            // (12 in flight per lane) scattered element by element into the operand layout.
            //   CB_MANUAL_K   rows [b nb, b nb + nb) -> [nb x 32] K-major block (forward: out = x . W_0^T)
            //   CB_MANUAL_MN  rows [32 kc, 32 kc + 32) -> one 32-wide MN group per chunk: reduction row k at k * 128 B, its
            //                 32-byte granule q at (q ^ (k & 3))                                   (backward: dx = dZ_0 . W_0)
            const bool kmaj = st.b_mode == CB_MANUAL_K;
            const int K0 = st.ld_w;
            const int zero16 = (kmaj ? st.blk_bytes : (kc1 - kc0) * 4096) >> 4;
            for (int i = lane; i < zero16; i += 32) *reinterpret_cast<float4*>(dst + i * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
            const int row0 = kmaj ? b * st.nb : kc0 * 32;
            const int row1 = kmaj ? min(st.N, row0 + st.nb) : min(st.K, kc1 * 32);
            const int nfl = max(0, row1 - row0) * K0;
            const float4* src4 = reinterpret_cast<const float4*>(Wm + (long long)row0 * K0);
            const int nvec = (nfl + 3) >> 2;
#pragma unroll 1
            for (int i0 = 0; i0 < nvec; i0 += 32 * 12) {
              float4 w[12];
#pragma unroll
              for (int u = 0; u < 12; ++u) {
                const int i = i0 + u * 32 + lane;
                w[u] = i < nvec ? __ldg(src4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int u = 0; u < 12; ++u) {
                const int f0 = (i0 + u * 32 + lane) * 4;
                int rl = f0 / K0, k = f0 - rl * K0;
                const float e[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  if (f0 + c < nfl) {
                    const int off = kmaj ? ch_kmajor_off(rl, k)
                                         : (rl >> 5) * 4096 + (rl & 31) * 128 + ((((k >> 3) ^ (rl & 3))) << 5) + (k & 7) * 4;
                    *reinterpret_cast<float*>(dst + off) = e[c];
                  }
                  if (++k == K0) { k = 0; ++rl; }
                }
              }
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) ch_arrive(&full_bar[slot]);
          }
        }
      }
    }
  } else {
    // =================================================================== MMA issuer
    unsigned int cnt = 0;
#pragma unroll 1
    for (int si = 0; si < nsteps; ++si) {
      const ChainStep& st = C.steps[s0 + si];
      const int nch = (st.K + 31) >> 5, ksteps_total = (st.K + 7) >> 3;
      const bool mn = st.b_mode == CB_TMA_MN || st.b_mode == CB_MANUAL_MN;
      const unsigned int idesc = __shfl_sync(0xffffffffu, ch_idesc(st.nb, mn ? 1 : 0), 0);
      const unsigned int a_hi = (1024u >> 4) | (1u << 14) | (2u << 29), a_lo0 = (16u >> 4) << 16;
      const unsigned int b_hi = mn ? ((512u >> 4) | (1u << 14) | (1u << 29)) : a_hi;
      const unsigned int b_lo0 = mn ? ((4096u >> 4) << 16) : a_lo0;
      const unsigned int b_ks = mn ? (1024u >> 4) : (32u >> 4);
      const unsigned int a_base = __shfl_sync(0xffffffffu, smem_u32(st.a_x0 ? X0 : ACT), 0);
      const unsigned int blk = __shfl_sync(0xffffffffu, (unsigned)st.blk_bytes, 0);
      mbar_wait(&a_ready, (unsigned)si & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll 1
      for (int kc0 = 0; kc0 < nch; kc0 += st.cps) {
        const int kc1 = min(nch, kc0 + st.cps);
#pragma unroll 1
        for (int b = 0; b < st.n_blk; ++b, ++cnt) {
          const unsigned int slot = cnt % (unsigned)C.n_slots, use = cnt / (unsigned)C.n_slots;
          mbar_wait(&full_bar[slot], use & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const unsigned int b_base = __shfl_sync(0xffffffffu, smem_u32(RING + slot * C.slot_bytes), 0);
          const unsigned int d = __shfl_sync(0xffffffffu, tmem + (unsigned)(b * st.nb), 0);
          const unsigned int ukc0 = __shfl_sync(0xffffffffu, (unsigned)kc0, 0), ukc1 = __shfl_sync(0xffffffffu, (unsigned)kc1, 0);
          const unsigned int ukt = __shfl_sync(0xffffffffu, (unsigned)ksteps_total, 0);
          const bool last = kc1 == nch && b == st.n_blk - 1;
          if (elect_one()) {
#pragma unroll 1
            for (unsigned int kc = ukc0; kc < ukc1; ++kc) {
              const unsigned int a_lo = a_lo0 | ((a_base + kc * (kChRows * 128)) >> 4), b_lo = b_lo0 | ((b_base + (kc - ukc0) * blk) >> 4);
              const unsigned int ks = min(4u, ukt - kc * 4u);
#pragma unroll 1
              for (unsigned int kk = 0; kk < ks; ++kk)
                tc_mma(d, ((unsigned long long)a_hi << 32) | (a_lo + kk * 2), ((unsigned long long)b_hi << 32) | (b_lo + kk * b_ks), idesc,
                       (kc | kk) != 0 ? 1u : 0u);
            }
            tc_commit(&empty_bar[slot]);
            if (last) tc_commit(&acc_full);
          }
          __syncwarp();
        }
      }
    }
  }

  // ---- teardown; the last CTA of the launch finishes the bookkeeping ----
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == kChWorkers + 1) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
  }
  if (tid == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(C.done, 1u);
    s_last = prev == gridDim.x - 1u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < 32) {
    const int nparts = C.tiles * (C.fin_mode == 0 ? C.n_q : 1);
    for (int ag = 0; ag < C.n_agents; ++ag) {
      float t = 0.f;
      for (int c2 = lane; c2 < ((nparts + 31) & ~31); c2 += 32) {
        float v = c2 < nparts ? __ldcg(C.loss_part + (long long)ag * C.lp_go + c2) : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        t += v;
      }
      if (lane == 0) {
        if (C.fin_mode == 0) {
          C.loss[ag] = t * C.inv_norm;
          if (C.host_status) {
            const unsigned int sq = C.seq[ag] + 1u;
            C.seq[ag] = sq;
            const unsigned long long word = ((unsigned long long)sq << 32) | (unsigned long long)__float_as_uint(t * C.inv_norm);
            *reinterpret_cast<volatile unsigned long long*>(C.host_status + ag) = word;
          }
        } else {
          C.loss[C.n_agents + ag] = -t * C.inv_norm;
        }
      }
    }
    if (lane == 0) {
      if (C.tick.state) {
        if (C.fin_mode == 0) C.tick.state[0] += 1;      // sampling step (Philox counter)
        adam_tick(C.tick);
      }
      *C.done = 0u;
      *C.epoch = epoch;
    }
  }
}

}  // namespace td3
