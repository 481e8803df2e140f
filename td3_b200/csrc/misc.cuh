// misc.cuh -- bandwidth-bound kernels of the TD3 update: replay gather (+ Philox index and
// smoothing-noise generation), Bellman target / MSE gradient, fused Adam + Polyak.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace td3 {

// Programmatic dependent launch: every kernel of the stage-per-launch / graph path lets its successor start early
// (launch latency, TMEM allocation, barrier initialisation overlap this kernel's tail) and itself waits for its
// predecessor's results only where it first touches global memory.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }

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
#pragma unroll 1
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
  // fast intrinsics: the draws only need to be N(0,1) to ~1e-6, and the accurate logf/sincospif paths are several
  // KB of code that every CTA of the update kernel would have to fetch once per update
  const float rad = sqrtf(-2.0f * __logf(u1));
  float s, co;
  __sincosf(6.28318530717958648f * u2, &s, &co);
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
  int elem_offset, pad;           // data-parallel shard: local row b is element b + elem_offset of the global batch
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
                       (uint32_t)(b + G.elem_offset), size);
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
                                    (uint32_t)(e + (long long)G.elem_offset * G.action_dim))
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
  pdl_launch_dependents();
  pdl_wait();
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
// ------------------------------------------------------------------------------------
// Adam step bookkeeping, kept on the device so that graph replays / the persistent kernel need no host
// involvement.  State block (u64 index): [1 + w] step count t, [4 + 2w], [5 + 2w] beta1^t, beta2^t as doubles,
// [10 + w] = {float step_size = lr / (1 - beta1^t), float sqrt(1 - beta2^t)}  (w = 0 critic, 1 actor).
// torch computes the same scalars in Python doubles (torch/optim/adam.py, _single_tensor_adam); the running
// products differ from pow() by < t * 2^-53 relative, far below the fp32 rounding of the two scalars.
// One thread calls this once per optimiser step.
// ------------------------------------------------------------------------------------
struct AdamTick {
  unsigned long long* state;       // base of the u64 state block, or nullptr
  int which, pad;
  double lr, beta1, beta2;
};

__device__ __forceinline__ void adam_tick(const AdamTick& T) {
  unsigned long long* st = T.state;
  const unsigned long long t0 = st[1 + T.which];
  double* pw = reinterpret_cast<double*>(st + 4 + 2 * T.which);
  const double p1 = (t0 == 0 ? 1.0 : pw[0]) * T.beta1, p2 = (t0 == 0 ? 1.0 : pw[1]) * T.beta2;
  pw[0] = p1;
  pw[1] = p2;
  float* sc = reinterpret_cast<float*>(st + 10 + T.which);
  sc[0] = (float)(T.lr / (1.0 - p1));
  sc[1] = (float)sqrt(1.0 - p2);
  st[1 + T.which] = t0 + 1;
}

struct LossParams {
  const float* q; const float* tq; const float* r; const float* nd;
  float* y; float* dq; float* loss;
  int batch, width, ldq, n_q;
  long long q_gi, q_go;            // strides between twins / agents in q, tq, dq
  long long y_go, r_go;
  float discount, inv_norm;        // inv_norm = 1 / (global_batch * width)
  AdamTick tick;                   // critic optimiser step (and [0] sample step += 1) done here by one thread
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
    if (agent == 0 && L.tick.state) {
      L.tick.state[0] += 1;          // sampling step (Philox counter); single writer, read after the next barrier
      adam_tick(L.tick);
    }
  }
}

__global__ void __launch_bounds__(256) loss_kernel(const __grid_constant__ LossParams L) {
  __shared__ float red[8];
  pdl_launch_dependents();
  pdl_wait();
  loss_body(L, blockIdx.x, red);
}

__global__ void adam_tick_kernel(const __grid_constant__ AdamTick T) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) adam_tick(T);
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
  const float* sc_ptr;               // device-resident {step_size, sqrt(1 - beta2^t)} written by adam_tick, or
  float step_size, bc2_sqrt;         // host-supplied scalars when sc_ptr == nullptr
  int do_adam, do_polyak;
};

struct EwParams {
  int n_ranges;
  EwRange r[3];
  double beta1, beta2, eps, tau;
};

constexpr int kEwThreads = 256;
constexpr int kEwPerBlock = kEwThreads * 8;   // 2 float4 per thread

__device__ __forceinline__ void ew_load4(const float* p, long long e, long long n, float (&v)[4]) {
  if (e + 4 <= n) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(p + e));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = e + k < n ? __ldcg(p + e + k) : 0.f;
  }
}
__device__ __forceinline__ void ew_store4(float* p, long long e, long long n, const float (&v)[4]) {
  if (e + 4 <= n) {
    *reinterpret_cast<float4*>(p + e) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (e + k < n) p[e + k] = v[k];
  }
}

// One element of torch's single-tensor Adam; out of line so the IEEE div/sqrt sequences exist once.
__device__ __noinline__ float adam_element(float p, float g, float& m, float& v, float w1, float b2, float w2, float bc2s,
                                          float eps, float neg_step) {
  m = fmaf(w1, __fsub_rn(g, m), m);
  v = __fadd_rn(__fmul_rn(v, b2), __fmul_rn(__fmul_rn(w2, g), g));
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2s), eps);
  return __fadd_rn(p, __fdiv_rn(__fmul_rn(neg_step, m), denom));
}

// bx = block index within the launch / stage.  Rolled on purpose (code size: see stage.cuh).
__device__ __forceinline__ void adam_polyak_body(const EwParams& E, long long bx) {
  int ri = 0;
  for (int q = 1; q < 3; ++q)
    if (q < E.n_ranges && bx >= E.r[q].blk_begin) ri = q;
  const EwRange& R = E.r[ri];
  const float s_step_size = R.do_adam ? (R.sc_ptr ? __ldcg(R.sc_ptr) : R.step_size) : 0.f;
  const float s_bc2_sqrt = R.do_adam ? (R.sc_ptr ? __ldcg(R.sc_ptr + 1) : R.bc2_sqrt) : 1.f;
  const float w1 = (float)(1.0 - E.beta1), b2 = (float)E.beta2, w2 = (float)(1.0 - E.beta2);
  const float eps = (float)E.eps, tau = (float)E.tau, omt = (float)(1.0 - E.tau);
  const float neg_step = -s_step_size;
  const long long base = (bx - R.blk_begin) * kEwPerBlock;
#pragma unroll 1
  for (int u = 0; u < 2; ++u) {
    const long long e = base + ((long long)u * kEwThreads + threadIdx.x) * 4;
    if (e >= R.n) continue;
    float pv[4];
    ew_load4(R.p, e, R.n, pv);
    if (R.do_adam) {
      float gv[4], mv[4], vv[4];
      ew_load4(R.g, e, R.n, gv);
      ew_load4(R.m, e, R.n, mv);
      ew_load4(R.v, e, R.n, vv);
#pragma unroll
      for (int k = 0; k < 4; ++k) pv[k] = adam_element(pv[k], gv[k], mv[k], vv[k], w1, b2, w2, s_bc2_sqrt, eps, neg_step);
      ew_store4(R.m, e, R.n, mv);
      ew_store4(R.v, e, R.n, vv);
      ew_store4(R.p, e, R.n, pv);
    }
    if (R.do_polyak) {
      float tv[4];
      ew_load4(R.tgt, e, R.n, tv);
#pragma unroll
      for (int k = 0; k < 4; ++k) tv[k] = __fadd_rn(__fmul_rn(tau, pv[k]), __fmul_rn(omt, tv[k]));
      ew_store4(R.tgt, e, R.n, tv);
    }
  }
}

__global__ void __launch_bounds__(kEwThreads) adam_polyak_kernel(const __grid_constant__ EwParams E) {
  pdl_launch_dependents();
  pdl_wait();
  adam_polyak_body(E, blockIdx.x);
}

}  // namespace td3
