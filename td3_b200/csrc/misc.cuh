// misc.cuh -- bandwidth-bound kernels of the TD3 update: replay gather (+ Philox index and
// smoothing-noise generation), Bellman target / MSE gradient, fused Adam + Polyak.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace td3 {

// ------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Counter-based: every draw is a pure function of
// (seed, stream, step, element), so a captured CUDA graph replays correctly from a
// device-resident step counter and results do not depend on the launch geometry.
// ------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint64_t seed) {
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c, k);
}

enum PhiloxStream : uint32_t { PHILOX_INDICES = 0, PHILOX_NOISE = 1 };

// uniform index in [0, size): multiply-high of a 64-bit draw (bias < size * 2^-64)
__host__ __device__ __forceinline__ int64_t philox_index(uint64_t seed, uint32_t stream, uint64_t step, uint32_t elem,
                                                         int64_t size) {
  uint32_t c[4] = {elem, stream, (uint32_t)step, (uint32_t)(step >> 32)};
  philox4x32_10(c, seed);
  const uint64_t r = ((uint64_t)c[0] << 32) | c[1];
#ifdef __CUDA_ARCH__
  return (int64_t)__umul64hi(r, (uint64_t)size);
#else
  return (int64_t)(((unsigned __int128)r * (unsigned __int128)(uint64_t)size) >> 64);
#endif
}

// standard normal number `elem` of the (step, stream) sequence: Box-Muller on 24-bit uniforms
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t stream, uint64_t step, uint32_t elem) {
  uint32_t c[4] = {elem >> 2, stream, (uint32_t)step, (uint32_t)(step >> 32)};
  philox4x32_10(c, seed);
  const uint32_t pair = (elem >> 1) & 1u;
  const uint32_t a = pair ? c[2] : c[0], b = pair ? c[3] : c[1];
  const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float rad = sqrtf(-2.0f * logf(u1));
  float s, co;
  sincospif(2.0f * u2, &s, &co);
  return (elem & 1u) ? rad * s : rad * co;
}

// ------------------------------------------------------------------------------------
// K1: replay gather.  rows are array-of-rows fp32; row idx[b] is scattered to up to
// kMaxSeg destination segments (the concatenated network inputs are written directly, no
// torch.cat later).  Small rows: one warp per sampled row.  Large rows (particle sets):
// blockIdx.y splits the row so each CTA streams a contiguous slice with 16B accesses.
// The same launch draws the indices (Philox or injected) and the clipped smoothing noise
// (TD3_featured.py:131-133).
// ------------------------------------------------------------------------------------
constexpr int kMaxSeg = 16;

struct GatherParams {
  const float* rows;
  long long row_stride, size, rb_agent_stride;
  int batch, n_agents, n_seg, rng_mode;
  int seg_off[kMaxSeg], seg_len[kMaxSeg], dst_ld[kMaxSeg];
  float* dst[kMaxSeg];
  long long dst_agent_stride[kMaxSeg];
  const long long* idx_in;        // [n_agents][batch] (injected mode)
  long long* idx_out;             // [n_agents][batch] indices actually used
  const unsigned long long* step_ptr;
  const unsigned long long* size_ptr; // live buffer size in device memory (graph replay), or nullptr -> `size`
  unsigned long long seed;
  float* eps_out;                 // [n_agents][batch][action_dim]  clipped noise
  const float* noise_in;          // N(0,1) draws (injected mode), same shape
  int action_dim, slices;         // slices = gridDim.y
  float policy_noise, noise_clip;
};

__device__ __forceinline__ void gather_body(const GatherParams& G, int bx, int by) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long job = (long long)bx * warps_per_block + warp;   // (agent, b)
  if (job >= (long long)G.n_agents * G.batch) return;
  const int agent = (int)(job / G.batch), b = (int)(job - (long long)agent * G.batch);
  const unsigned long long step = __ldcg(G.step_ptr);
  const long long size = G.size_ptr ? (long long)__ldcg(G.size_ptr) : G.size;
  long long idx;
  if (G.rng_mode == 0) {
    idx = philox_index(G.seed + (unsigned long long)agent * 0x9E3779B97F4A7C15ull, PHILOX_INDICES, step,
                       (uint32_t)b, size);
  } else {
    idx = G.idx_in[job];
  }
  const int slice = by;
  if (slice == 0) {
    if (lane == 0) G.idx_out[job] = idx;
    for (int a = lane; a < G.action_dim; a += 32) {
      const long long e = (long long)b * G.action_dim + a;
      float z = (G.rng_mode == 0)
                    ? philox_normal(G.seed + (unsigned long long)agent * 0x9E3779B97F4A7C15ull, PHILOX_NOISE, step,
                                    (uint32_t)e)
                    : G.noise_in[(long long)agent * G.batch * G.action_dim + e];
      z = z * G.policy_noise;
      z = fminf(fmaxf(z, -G.noise_clip), G.noise_clip);
      G.eps_out[(long long)agent * G.batch * G.action_dim + e] = z;
    }
  }
  const float* __restrict__ src = G.rows + (long long)agent * G.rb_agent_stride + idx * G.row_stride;
  for (int s = 0; s < G.n_seg; ++s) {
    const int len = G.seg_len[s];
    float* __restrict__ d = G.dst[s] + (long long)agent * G.dst_agent_stride[s] + (long long)b * G.dst_ld[s];
    const float* __restrict__ sp = src + G.seg_off[s];
    // this slice's part of the segment, in units of 4 floats where alignment allows
    const bool vec = ((G.seg_off[s] | len | G.dst_ld[s]) & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(sp)) & 15) == 0;
    if (vec) {
      const int n4 = len >> 2;
      const int per = (n4 + G.slices - 1) / G.slices;
      const int lo = slice * per, hi = min(n4, lo + per);
      const float4* s4 = reinterpret_cast<const float4*>(sp);
      float4* d4 = reinterpret_cast<float4*>(d);
      for (int i = lo + lane; i < hi; i += 32) d4[i] = __ldg(s4 + i);
    } else {
      const int per = (len + G.slices - 1) / G.slices;
      const int lo = slice * per, hi = min(len, lo + per);
      for (int i = lo + lane; i < hi; i += 32) d[i] = __ldg(sp + i);
    }
  }
}

__global__ void __launch_bounds__(256) gather_kernel(const __grid_constant__ GatherParams G) {
  gather_body(G, blockIdx.x, blockIdx.y);
}

__global__ void philox_indices_kernel(long long* idx, long long batch, long long size, unsigned long long seed,
                                      unsigned int stream_id, unsigned long long step) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < batch) idx[b] = philox_index(seed, stream_id, step, (uint32_t)b, size);
}

// ------------------------------------------------------------------------------------
// Bellman target + MSE gradient (TD3_featured.py:141-148, TD3_particles.py:184-198).
//   y[b,j]   = r[b] + nd[b] * discount * min_g tq[g][b,j]
//   dq[g][b,j] = (2 / norm) * (q[g][b,j] - y[b,j]),  norm = global_batch * W   (mean over B*W)
//   loss     = sum_g mean((q_g - y)^2)
// One CTA per agent (deterministic block reduction).  Also advances the device step counters.
// ------------------------------------------------------------------------------------
struct LossParams {
  const float* q; const float* tq; const float* r; const float* nd;
  float* y; float* dq; float* loss;
  int batch, width, ldq, n_q;
  long long q_gi, q_go;            // strides between twins / agents in q, tq, dq
  long long y_go, r_go;
  float discount, inv_norm;        // inv_norm = 1 / (global_batch * width)
  unsigned long long* counters;    // [0] sample step, [1] critic Adam t  (both += 1 here)
};

__device__ __forceinline__ void loss_body(const LossParams& L, int agent, float* red) {
  const float* q = L.q + agent * L.q_go;
  const float* tq = L.tq + agent * L.q_go;
  float* dq = L.dq + agent * L.q_go;
  const float* r = L.r + agent * L.r_go;
  const float* nd = L.nd + agent * L.r_go;
  float* y = L.y + agent * L.y_go;
  float acc = 0.f;
  const int total = L.batch * L.width;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int b = e / L.width, j = e - b * L.width;
    const size_t o = (size_t)b * L.ldq + j;
    float t = __ldcg(tq + o);
    if (L.n_q > 1) t = fminf(t, __ldcg(tq + L.q_gi + o));
    const float yy = __fadd_rn(__ldcg(r + b), __fmul_rn(__fmul_rn(__ldcg(nd + b), L.discount), t));
    y[o] = yy;
    for (int g = 0; g < L.n_q; ++g) {
      const float d = __ldcg(q + g * L.q_gi + o) - yy;
      dq[g * L.q_gi + o] = 2.f * L.inv_norm * d;
      acc = fmaf(d, d, acc);
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    L.loss[agent] = tot * L.inv_norm;
    if (agent == 0 && L.counters) {
      atomicAdd(L.counters + 0, 1ull);
      atomicAdd(L.counters + 1, 1ull);
    }
  }
}

__global__ void __launch_bounds__(256) loss_kernel(const __grid_constant__ LossParams L) {
  __shared__ float red[8];
  loss_body(L, blockIdx.x, red);
}

__global__ void counter_add_kernel(unsigned long long* c, unsigned long long inc) {
  if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(c, inc);
}

// ------------------------------------------------------------------------------------
// K4/K5: Adam (+ Polyak) over packed fp32 buffers.  Arithmetic mirrors torch 2.11's
// single-tensor Adam (torch/optim/adam.py: lerp_, mul_/addcmul_, sqrt/div/add_, addcdiv_)
// so that, given identical gradients, parameters agree to <= 2 ulp:
//   m = m + (1-b1) * (g - m)                 (fma, as ATen's vectorised lerp)
//   v = v*b2 + (1-b2)*g*g
//   p = p + (-(lr/bc1) * m) / (sqrt(v)/sqrt(bc2) + eps)
// Polyak (TD3_featured.py:167-171): t = tau*p + (1-tau)*t with both products rounded first.
// Up to three ranges per launch (actor Adam+Polyak and critic Polyak fuse into one kernel).
// ------------------------------------------------------------------------------------
struct EwRange {
  float* p; const float* g; float* m; float* v; float* tgt;
  long long n, blk_begin;
  const unsigned long long* t_ptr;   // device-resident Adam step (already incremented), or
  long long t_val;                   // host-supplied step when t_ptr == nullptr
  double lr;
  int do_adam, do_polyak;
};

struct EwParams {
  int n_ranges;
  EwRange r[3];
  double beta1, beta2, eps, tau;
};

constexpr int kEwThreads = 256;
constexpr int kEwPerBlock = kEwThreads * 8;   // 2 float4 per thread

// smem2: two floats of shared scratch.  bx = block index within the launch / stage.
__device__ __forceinline__ void adam_polyak_body(const EwParams& E, long long bx, float* smem2) {
  int ri = 0;
  for (int q = 1; q < 3; ++q)
    if (q < E.n_ranges && bx >= E.r[q].blk_begin) ri = q;
  const EwRange& R = E.r[ri];
  __syncthreads();
  if (R.do_adam && threadIdx.x == 0) {
    const double t = (double)(R.t_ptr ? (long long)__ldcg(R.t_ptr) : R.t_val);
    const double bc1 = 1.0 - pow(E.beta1, t);
    const double bc2 = 1.0 - pow(E.beta2, t);
    smem2[0] = (float)(R.lr / bc1);
    smem2[1] = (float)sqrt(bc2);
  }
  __syncthreads();
  const float s_step_size = smem2[0], s_bc2_sqrt = smem2[1];
  const float w1 = (float)(1.0 - E.beta1), b2 = (float)E.beta2, w2 = (float)(1.0 - E.beta2);
  const float eps = (float)E.eps, tau = (float)E.tau, omt = (float)(1.0 - E.tau);
  const float neg_step = R.do_adam ? -s_step_size : 0.f;
  const float bc2s = R.do_adam ? s_bc2_sqrt : 1.f;
  const long long base = (bx - R.blk_begin) * kEwPerBlock;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const long long e = base + ((long long)u * kEwThreads + threadIdx.x) * 4;
    if (e >= R.n) continue;
    if (e + 4 <= R.n) {
      float4 p = __ldcg(reinterpret_cast<const float4*>(R.p + e));
      float pv[4] = {p.x, p.y, p.z, p.w};
      if (R.do_adam) {
        const float4 g4 = __ldcg(reinterpret_cast<const float4*>(R.g + e));
        float4 m4 = __ldcg(reinterpret_cast<const float4*>(R.m + e));
        float4 v4 = __ldcg(reinterpret_cast<const float4*>(R.v + e));
        float gv[4] = {g4.x, g4.y, g4.z, g4.w}, mv[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          mv[k] = fmaf(w1, __fsub_rn(gv[k], mv[k]), mv[k]);
          vv[k] = __fadd_rn(__fmul_rn(vv[k], b2), __fmul_rn(__fmul_rn(w2, gv[k]), gv[k]));
          const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv[k]), bc2s), eps);
          pv[k] = __fadd_rn(pv[k], __fdiv_rn(__fmul_rn(neg_step, mv[k]), denom));
        }
        *reinterpret_cast<float4*>(R.m + e) = make_float4(mv[0], mv[1], mv[2], mv[3]);
        *reinterpret_cast<float4*>(R.v + e) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        *reinterpret_cast<float4*>(R.p + e) = make_float4(pv[0], pv[1], pv[2], pv[3]);
      }
      if (R.do_polyak) {
        const float4 t4 = __ldcg(reinterpret_cast<const float4*>(R.tgt + e));
        float tv[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) tv[k] = __fadd_rn(__fmul_rn(tau, pv[k]), __fmul_rn(omt, tv[k]));
        *reinterpret_cast<float4*>(R.tgt + e) = make_float4(tv[0], tv[1], tv[2], tv[3]);
      }
    } else {
      for (long long k = e; k < R.n; ++k) {
        float pvk = __ldcg(R.p + k);
        if (R.do_adam) {
          const float g = __ldcg(R.g + k), m0 = __ldcg(R.m + k);
          const float m = fmaf(w1, __fsub_rn(g, m0), m0);
          const float v = __fadd_rn(__fmul_rn(__ldcg(R.v + k), b2), __fmul_rn(__fmul_rn(w2, g), g));
          const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2s), eps);
          pvk = __fadd_rn(pvk, __fdiv_rn(__fmul_rn(neg_step, m), denom));
          R.m[k] = m; R.v[k] = v; R.p[k] = pvk;
        }
        if (R.do_polyak) R.tgt[k] = __fadd_rn(__fmul_rn(tau, pvk), __fmul_rn(omt, __ldcg(R.tgt + k)));
      }
    }
  }
}

__global__ void __launch_bounds__(kEwThreads) adam_polyak_kernel(const __grid_constant__ EwParams E) {
  __shared__ float smem2[2];
  adam_polyak_body(E, blockIdx.x, smem2);
}

}  // namespace td3
