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

// fp32 -> nearest TF32-representable fp32 (10 mantissa bits, ties away from zero).  tcgen05 kind::tf32 reads only the
// upper 19 bits of an operand word, i.e. it TRUNCATES; an operand stored through this function is read exactly, so
// the contraction sees round-to-nearest operands (unbiased, half the worst-case error) at no cost in the GEMM itself.
__device__ __forceinline__ float rn_tf32(float x) {
  unsigned int u;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// The same rounding as two integer instructions, for inner loops whose inputs are finite: add half an ulp of the 10-bit
// mantissa to the magnitude, drop the low 13 bits (ptxas expands cvt.rna.tf32 to exactly this plus an |x| < inf guard).
__device__ __forceinline__ float rn_tf32_finite(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

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
  int elem_offset, row_floats;    // data-parallel shard: local row b is element b + elem_offset of the global batch
  unsigned int seg_rn, pad_g;     // bit s: segment s feeds a tensor-core contraction -> stored rounded to nearest TF32
  float policy_noise, noise_clip;
};

// replay index of element b of agent's batch (Philox draw, or the injected one)
__device__ __forceinline__ long long gather_index(const GatherParams& G, int agent, int b) {
  if (G.rng_mode != 0) return G.idx_in[(long long)agent * G.batch + b];
  const unsigned long long step = __ldcg(G.step_ptr);
  const long long size = G.size_ptr ? (long long)__ldcg(G.size_ptr) : G.size;
  return philox_index(G.seed + (unsigned long long)agent * 0x9E3779B97F4A7C15ull, PHILOX_INDICES, step,
                      (uint32_t)(b + G.elem_offset), size);
}

// ---- bulk-copy (TMA engine, non-tensor form) staging of large row segments -----------------------------------------
// A particle set is 24 KB of contiguous floats per sampled row.  Instead of every lane issuing 16-byte loads and
// stores, one lane hands the warp's slice to the copy engine: cp.async.bulk global -> shared (completion on an
// mbarrier), then cp.async.bulk shared -> global.  No registers carry data, a warp has its whole slice in flight, and
// the SM's load/store pipes stay free.  Needs 16-byte aligned addresses and sizes (the row layout guarantees it).
constexpr int kGatherBulkBytes = 4096;            // staging bytes per warp (8 warps per CTA: 32 KB of dynamic shared memory)
constexpr int kGatherBulkMinFloats = 1024;        // segments at least this long take the bulk path

__device__ __forceinline__ void bulk_g2s(unsigned int dst_smem, const void* src, unsigned int bytes, unsigned int bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_smem), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned int src_smem, unsigned int bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}

// one warp scatters (its slice of) transition `idx` to every destination segment; slice 0 also records the index and
// draws the row's clipped smoothing noise.  stage / bar: this warp's bulk staging area and mbarrier in shared memory
// (nullptr: no bulk path), phase: the barrier's parity, advanced for every bulk load
__device__ __forceinline__ void gather_row(const GatherParams& G, int agent, int b, long long idx, int slice, int lane,
                                           unsigned char* stage = nullptr, unsigned long long* bar = nullptr,
                                           unsigned int* phase = nullptr) {
  const long long job = (long long)agent * G.batch + b;
  if (slice == 0) {
    if (lane == 0) G.idx_out[job] = idx;
    const unsigned long long step = __ldcg(G.step_ptr);
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
    if (vec && stage && len >= kGatherBulkMinFloats && !((G.seg_rn >> s) & 1u)) {
      // this slice's part of the segment, in 16-byte units, moved kGatherBulkBytes at a time by the copy engine
      const int n4 = len >> 2;
      const int per = (n4 + G.slices - 1) / G.slices;
      const int lo = slice * per, hi = min(n4, lo + per);
      const unsigned int st = (unsigned int)__cvta_generic_to_shared(stage), br = (unsigned int)__cvta_generic_to_shared(bar);
      for (int at = lo; at < hi; at += kGatherBulkBytes / 16) {
        const unsigned int bytes = (unsigned int)(min(hi - at, kGatherBulkBytes / 16) * 16);
        if (lane == 0) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(br), "r"(bytes) : "memory");
          bulk_g2s(st, reinterpret_cast<const float4*>(sp) + at, bytes, br);
          unsigned int done = 0;
          while (!done) {
            asm volatile(
                "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(done)
                : "r"(br), "r"(*phase & 1u)
                : "memory");
          }
          bulk_s2g(reinterpret_cast<float4*>(d) + at, st, bytes);
          asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");     // the staging area may be overwritten
        }
        *phase += 1;
      }
    } else if (vec) {
      const int n4 = len >> 2;
      const int per = (n4 + G.slices - 1) / G.slices;
      const int lo = slice * per, hi = min(n4, lo + per);
      const float4* s4 = reinterpret_cast<const float4*>(sp);
      float4* d4 = reinterpret_cast<float4*>(d);
      if ((G.seg_rn >> s) & 1u) {
        for (int i = lo + lane; i < hi; i += 32) {
          const float4 v = __ldg(s4 + i);
          d4[i] = make_float4(rn_tf32(v.x), rn_tf32(v.y), rn_tf32(v.z), rn_tf32(v.w));
        }
      } else {
        for (int i = lo + lane; i < hi; i += 32) d4[i] = __ldg(s4 + i);
      }
    } else {
      const int per = (len + G.slices - 1) / G.slices;
      const int lo = slice * per, hi = min(len, lo + per);
      const bool rn = (G.seg_rn >> s) & 1u;
      for (int i = lo + lane; i < hi; i += 32) {
        const float v = __ldg(sp + i);
        d[i] = rn ? rn_tf32(v) : v;
      }
    }
  }
}

// the same scatter from a copy of the transition in shared memory (front.cuh), plus the row's smoothing noise
__device__ __forceinline__ void gather_scatter(const GatherParams& G, int agent, int b, const float* row, int lane) {
  const unsigned long long step = __ldcg(G.step_ptr);
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
#pragma unroll 1
  for (int s = 0; s < G.n_seg; ++s) {
    float* d = G.dst[s] + (long long)agent * G.dst_agent_stride[s] + (long long)b * G.dst_ld[s];
    const float* sp = row + G.seg_off[s];
    for (int i = lane; i < G.seg_len[s]; i += 32) d[i] = sp[i];
  }
}

__device__ __forceinline__ void gather_body(const GatherParams& G, int bx, int by, unsigned char* stage = nullptr,
                                            unsigned long long* bars = nullptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long job = (long long)bx * warps_per_block + warp;   // (agent, b)
  if (job >= (long long)G.n_agents * G.batch) return;
  const int agent = (int)(job / G.batch), b = (int)(job - (long long)agent * G.batch);
  unsigned int phase = 0;
  gather_row(G, agent, b, gather_index(G, agent, b), by, lane, stage ? stage + warp * kGatherBulkBytes : nullptr,
             bars ? bars + warp : nullptr, &phase);
}

// bulk != 0: launched with 8 * kGatherBulkBytes of dynamic shared memory; large aligned segments go through the copy engine
__global__ void __launch_bounds__(256) gather_kernel(const __grid_constant__ GatherParams G, int bulk) {
  extern __shared__ __align__(128) unsigned char gather_stage[];
  __shared__ unsigned long long gather_bars[8];
  pdl_launch_dependents();
  if (bulk) {
    if (threadIdx.x < 8) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((unsigned int)__cvta_generic_to_shared(&gather_bars[threadIdx.x])) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    __syncthreads();
  }
  pdl_wait();
  gather_body(G, blockIdx.x, blockIdx.y, bulk ? gather_stage : nullptr, bulk ? gather_bars : nullptr);
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
  int rn_out, pad_l;               // dq stored rounded to nearest TF32 (operand of tensor-core contractions)
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
      const float dqv = 2.f * L.inv_norm * d;
      dq[g * L.q_gi + o] = L.rn_out ? rn_tf32(dqv) : dqv;
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

// ------------------------------------------------------------------------------------
// Fused critic head.  The last layer of a Q network has 1 (featured) or action_dim (particles) outputs: as GEMM
// stages, "online head forward", "target head forward", "loss" and "head backward" are four dependent launches
// of almost no work.  One kernel does them all for a block of batch rows:
//   tq_g = ht_g . Wt_g^T + bt_g            (target critics, TD3_featured.py:140)
//   y    = r + nd * discount * min_g tq_g  (:141-142)
//   q_g  = h_g . W_g^T + b_g               (:145)
//   loss += (q_g - y)^2, dq_g = 2 (q_g - y) / norm            (:148)
//   dh_g = (dq_g . W_g) * relu'(h_g)       (backward of the head through the preceding ReLU; no mask with LayerNorm)
//   dW_g = dq_g^T . h_g,  db_g = sum_b dq_g                   (deterministic two-level reduction, last CTA finishes)
// mode 1 (actor step, :159): q = Q1(s, pi(s)) head only, loss = -mean(q), dq = -1/norm, no weight gradients.
// One warp per row, lanes over the hidden width (<= 512), qw <= 4.
// ------------------------------------------------------------------------------------
constexpr int kHeadMaxQw = 4, kHeadMaxW = 512, kHeadRows = 8, kHeadThreads = 512;

struct HeadParams {
  const float* h; const float* ht;            // last hidden activations of the online / target critics [B, ldh]
  long long h_go, h_gi, ht_go, ht_gi;
  const float* W; const float* b; const float* Wt; const float* bt;   // head weight [qw, w] and bias [qw], online / target
  long long w_go, w_gi, wt_go, wt_gi;         // strides of the parameter buffers (agent, twin)
  const float* r; const float* nd; long long r_go;
  float* q; float* tq; float* dq; long long q_go, q_gi;               // [B, qw] each
  float* y; long long y_go;
  float* dz; long long dz_go, dz_gi;          // gradient w.r.t. h (ld = lddz)
  float* gW; float* gb; long long g_go, g_gi; // gradient of the head parameters
  float* part; long long part_go;             // scratch [n_cta][n_q][qw * w + qw] + [n_cta] loss partials, per agent
  unsigned int* counter;                      // per-agent arrival counter (zero between launches)
  float* loss;                                // loss[agent]
  int batch, w, qw, n_q, ldh, lddz, n_cta, mode, relu_mask;
  int skip_dw;                                // 1: dW/db of the head are a GEMM problem of the next stage (from dq and h)
  int defer_finish;                           // 1: the CTAs leave after storing their loss partial; head_finish_kernel -- off the update's
                                              //    dependency chain, beside the next stage -- sums them and does the bookkeeping
  int rn_out;                                 // dz / dq stored rounded to nearest TF32 (operands of tensor-core contractions)
  float discount, inv_norm;
  AdamTick tick;                              // critic optimiser tick + sampling step (mode 0), done by the finishing CTA
  // optional host mirror of the loss (mode 0): the finishing CTA of agent i bumps seq[i] and stores the 8-byte word
  // {loss bits, seq} to host_status[i] in mapped pinned host memory, so a host thread can pick the step's result up
  // as soon as it exists instead of draining the stream (td3_agent_bind_host_status)
  unsigned long long* host_status;
  unsigned int* seq;
};

// The launch's loss from the CTAs' partials (fixed order), the host mirror of it, the optimiser tick: one warp per agent.
// Run by the last CTA to arrive (head_body_t) or, with HeadParams::defer_finish, by head_finish_kernel.
__device__ __forceinline__ void head_finish_warp(const HeadParams& H, int agent, const float* lpart, int lane) {
  const bool critic = H.mode == 0;
  float t = 0.f;
  for (int c2 = lane; c2 < ((H.n_cta + 31) & ~31); c2 += 32) {
    float v = c2 < H.n_cta ? __ldcg(lpart + c2) : 0.f;
    // fixed-order tree over the 32 lanes, then accumulate the groups of 32 CTAs in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    t += v;
  }
  if (lane == 0) {
    H.loss[agent] = critic ? t * H.inv_norm : -t * H.inv_norm;
    if (critic && H.host_status) {
      const unsigned int sq = H.seq[agent] + 1u;
      H.seq[agent] = sq;
      const unsigned long long word = ((unsigned long long)sq << 32) | (unsigned long long)__float_as_uint(t * H.inv_norm);
      // plain posted store: no system-scope fence here (it would hold the chain for a PCIe round trip); the word
      // is a single aligned 8-byte write, and the end of the kernel flushes it at the latest
      *reinterpret_cast<volatile unsigned long long*>(H.host_status + agent) = word;
    }
    if (agent == 0 && H.tick.state) {
      if (critic) H.tick.state[0] += 1;          // sampling step (Philox counter)
      adam_tick(H.tick);
    }
  }
}

// One warp per (batch row, twin): every global load of the pair is issued before anything is consumed -- the kernel
// sits on a dependency chain, so the length of one warp's instruction stream, not throughput, is what the update
// pays for.  kHeadRows rows per CTA; the twins of a row meet in shared memory for min(Q1', Q2').
template <int QW, int WI>
__device__ __forceinline__ void head_body_t(const HeadParams& H, int tile, float* smem) {
  constexpr int NQ = 2;
  const int agent = tile / H.n_cta, cb = tile - agent * H.n_cta;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = H.w, qw = H.qw, nq = H.n_q, per_g = qw * w + qw;
  const bool critic = H.mode == 0;
  const bool want_dw = critic && !H.skip_dw;
  float* Ws = smem;                                     // [n_q][qw][w] online head weights, then [n_q][qw] biases
  float* bs = Ws + nq * qw * w;
  float* Wts = bs + NQ * kHeadMaxQw;                    // target
  float* bts = Wts + (critic ? nq * qw * w : 0);
  float* tqs = bts + NQ * kHeadMaxQw;                   // [kHeadRows][NQ][kHeadMaxQw] target-head outputs of the CTA's rows
  float* lred = tqs + kHeadRows * NQ * kHeadMaxQw;      // [kHeadRows * NQ] loss terms
  float* red = lred + kHeadRows * NQ;                   // [kHeadRows][n_q][per_g] weight-gradient terms of each row
  const int nwarps = blockDim.x >> 5;                   // 16 in head_kernel (one pass), 8 in the persistent kernel (two)
  const int pairs = kHeadRows * nq;
  bool staged = false;
#pragma unroll 1
  for (int base = 0; base < pairs; base += nwarps) {
    const int pi = base + warp;
    const bool active = pi < pairs;
    const int rl = pi / nq, g = pi - rl * nq;
    const int row = cb * kHeadRows + rl;
    const bool row_ok = active && row < H.batch;
    // ---- issue every global load of this (row, twin) ----
    float hv[WI], tv[WI];
    float rv = 0.f, ndv = 0.f;
    {
      const float* hrow = H.h + agent * H.h_go + g * H.h_gi + (size_t)row * H.ldh;
      const float* trow = critic ? H.ht + agent * H.ht_go + g * H.ht_gi + (size_t)row * H.ldh : nullptr;
#pragma unroll
      for (int i = 0; i < WI; ++i) {
        const int k = lane + 32 * i;
        const bool ok = row_ok && k < w;
        hv[i] = ok ? hrow[k] : 0.f;
        tv[i] = ok && critic ? trow[k] : 0.f;
      }
    }
    if (row_ok && critic) {
      rv = H.r[agent * H.r_go + row];
      ndv = H.nd[agent * H.r_go + row];
    }
    if (!staged) {                                        // head parameters -> shared memory (first pass only)
      for (int gg = 0; gg < nq; ++gg) {
        for (int e = threadIdx.x; e < qw * w; e += blockDim.x) {
          Ws[gg * qw * w + e] = H.W[agent * H.w_go + gg * H.w_gi + e];
          if (critic) Wts[gg * qw * w + e] = H.Wt[agent * H.wt_go + gg * H.wt_gi + e];
        }
        if ((int)threadIdx.x < qw) {
          bs[gg * kHeadMaxQw + threadIdx.x] = H.b[agent * H.w_go + gg * H.w_gi + threadIdx.x];
          if (critic) bts[gg * kHeadMaxQw + threadIdx.x] = H.bt[agent * H.wt_go + gg * H.wt_gi + threadIdx.x];
        }
      }
      __syncthreads();
      staged = true;
    }
    // ---- target head of this twin (TD3_featured.py:140); the twins of a row meet in shared memory ----
    if (critic) {
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < QW; ++j) {
          float d = 0.f;
          if (j < qw) {
#pragma unroll
            for (int i = 0; i < WI; ++i) {
              const int k = lane + 32 * i;
              if (k < w) d = fmaf(tv[i], Wts[(g * qw + j) * w + k], d);
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
          d += bts[g * kHeadMaxQw + j];
          if (lane == 0 && j < qw) {
            H.tq[agent * H.q_go + g * H.q_gi + (size_t)row * qw + j] = d;
            tqs[(rl * NQ + g) * kHeadMaxQw + j] = d;
          }
        }
      }
      __syncthreads();
    }
    float pair_loss = 0.f;
    if (row_ok) {
      // ---- Bellman target (:141-142) ----
      float yv[QW];
      if (critic) {
#pragma unroll
        for (int j = 0; j < QW; ++j) {
          float tmin = tqs[(rl * NQ) * kHeadMaxQw + j];
          if (nq > 1) tmin = fminf(tmin, tqs[(rl * NQ + 1) * kHeadMaxQw + j]);
          yv[j] = __fadd_rn(rv, __fmul_rn(__fmul_rn(ndv, H.discount), tmin));
          if (g == 0 && lane == 0 && j < qw) H.y[agent * H.y_go + (size_t)row * qw + j] = yv[j];
        }
      }
      // ---- online head, gradient w.r.t. the hidden activations, this row's weight-gradient terms ----
      float dqv[QW];
#pragma unroll
      for (int j = 0; j < QW; ++j) {
        float d = 0.f;
        if (j < qw) {
#pragma unroll
          for (int i = 0; i < WI; ++i) {
            const int k = lane + 32 * i;
            if (k < w) d = fmaf(hv[i], Ws[(g * qw + j) * w + k], d);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        d += bs[g * kHeadMaxQw + j];
        dqv[j] = 0.f;
        if (j < qw) {
          const size_t qo = agent * H.q_go + g * H.q_gi + (size_t)row * qw + j;
          if (lane == 0) H.q[qo] = d;
          if (critic) {
            const float diff = d - yv[j];
            dqv[j] = 2.f * H.inv_norm * diff;
            pair_loss = fmaf(diff, diff, pair_loss);
          } else {
            dqv[j] = -H.inv_norm;
            pair_loss += d;
          }
          if (lane == 0 && H.dq) H.dq[qo] = H.rn_out ? rn_tf32(dqv[j]) : dqv[j];
        }
      }
      float* dzr = H.dz + agent * H.dz_go + g * H.dz_gi + (size_t)row * H.lddz;
#pragma unroll
      for (int i = 0; i < WI; ++i) {
        const int k = lane + 32 * i;
        if (k < w) {
          float sacc = 0.f;
#pragma unroll
          for (int j = 0; j < QW; ++j)
            if (j < qw) {
              sacc = fmaf(dqv[j], Ws[(g * qw + j) * w + k], sacc);
              if (want_dw) red[(rl * nq + g) * per_g + j * w + k] = dqv[j] * hv[i];
            }
          const float dzv = (H.relu_mask && !(hv[i] > 0.f)) ? 0.f : sacc;
          dzr[k] = H.rn_out ? rn_tf32(dzv) : dzv;
        }
      }
      if (want_dw && lane < qw) {
        float dj = dqv[0];
#pragma unroll
        for (int j = 1; j < QW; ++j)
          if (lane == j) dj = dqv[j];
        red[(rl * nq + g) * per_g + qw * w + lane] = dj;
      }
    } else if (want_dw && active) {                // rows past the batch contribute zeros
      for (int e = lane; e < per_g; e += 32) red[(rl * nq + g) * per_g + e] = 0.f;
    }
    if (active && lane == 0) lred[pi] = pair_loss;
  }   // pass over this warp's (row, twin) pairs
  __syncthreads();
  // ---- CTA partials (fixed order over the rows), then the last CTA to arrive sums the CTA partials in order ----
  float* pbase = H.part + agent * H.part_go;
  if (want_dw)
    for (int e = threadIdx.x; e < nq * per_g; e += blockDim.x) {
      float sacc = 0.f;
#pragma unroll
      for (int wv = 0; wv < kHeadRows; ++wv) sacc += red[wv * nq * per_g + e];
      pbase[(size_t)cb * nq * per_g + e] = sacc;
    }
  __shared__ int s_last;
  float* lpart = pbase + (size_t)H.n_cta * nq * per_g;
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int wv = 0; wv < pairs; ++wv) t += lred[wv];
    lpart[cb] = t;
  }
  if (H.defer_finish) return;                 // (never together with want_dw: host)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(H.counter + agent, 1u);
    s_last = prev == (unsigned)H.n_cta - 1u;
    if (s_last) atomicExch(H.counter + agent, 0u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (want_dw) {
    const long long cstride = (long long)nq * per_g;
    for (int e = threadIdx.x; e < nq * per_g; e += blockDim.x) {
      float sacc = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < H.n_cta; c0 += 8) {       // eight independent loads in flight, summed in CTA order
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = c0 + u < H.n_cta ? __ldcg(pbase + (c0 + u) * cstride + e) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) sacc += v[u];
      }
      const int g = e / per_g, o = e - g * per_g;
      if (o < qw * w) H.gW[agent * H.g_go + g * H.g_gi + o] = sacc;
      else H.gb[agent * H.g_go + g * H.g_gi + (o - qw * w)] = sacc;
    }
  }
  if (threadIdx.x < 32) head_finish_warp(H, agent, lpart, lane);
}

__device__ __forceinline__ void head_body(const HeadParams& H, int tile, float* smem) {
  if (H.qw == 1) {
    if (H.w <= 320) head_body_t<1, 10>(H, tile, smem);
    else head_body_t<1, kHeadMaxW / 32>(H, tile, smem);
  } else {
    head_body_t<kHeadMaxQw, kHeadMaxW / 32>(H, tile, smem);
  }
}

__global__ void __launch_bounds__(kHeadThreads) head_kernel(const __grid_constant__ HeadParams H) {
  extern __shared__ __align__(16) float head_smem[];
  pdl_launch_dependents();
  pdl_wait();
  head_body(H, blockIdx.x, head_smem);
}

// the same body compiled for two CTAs per SM (64 registers): launches with more tiles than SMs (populations, the
// batch-8192 update) run one wave instead of two; the latency form above keeps its 96 registers
__global__ void __launch_bounds__(kHeadThreads, 2) head_kernel_wide(const __grid_constant__ HeadParams H) {
  extern __shared__ __align__(16) float head_smem[];
  pdl_launch_dependents();
  pdl_wait();
  head_body(H, blockIdx.x, head_smem);
}

// HeadParams::defer_finish: what the last CTA of the head launch would do, as a launch of its own that the graph runs
// BESIDE the backward stage (a fork in the captured graph, joined before the optimiser launch that reads the tick's
// scalars): the fence + atomic + partial read-back + bookkeeping tail (~2 us) leaves the update's dependency chain.
__global__ void __launch_bounds__(32) head_finish_kernel(const __grid_constant__ HeadParams H) {
  const int agent = blockIdx.x;
  const int per_g = H.qw * H.w + H.qw;
  const float* lpart = H.part + agent * H.part_go + (size_t)H.n_cta * H.n_q * per_g;
  head_finish_warp(H, agent, lpart, threadIdx.x & 31);
}

// ------------------------------------------------------------------------------------
// Weight normalisation (TD3_particles.py:48-50, torch.nn.utils.weight_norm on every `linears` module):
// the packed buffers hold (bias, weight_g [out], weight_v [out, in]) per layer and the layer computes with
//   W[n, :] = v[n, :] * (g[n] / ||v[n, :]||)                      (aten::_weight_norm, dim = 0).
// mode 0 materialises an "effective" copy of a family's packed buffer in which every weight_v slot holds W (all
//        other tensors are copied through), which is what every contraction then reads;
// mode 1 turns dL/dW, which the backward contractions leave in the gradient buffer's weight_v slot, into
//        dL/dg[n] = (dW[n]·v[n]) / ||v[n]||  and  dL/dv[n] = (g/||v||) * (dW[n] - (dL/dg[n] / ||v[n]||) v[n])  in place.
// One warp per output row; the row blocks of all layers, networks and jobs are flattened into one tile index.
// ------------------------------------------------------------------------------------
constexpr int kWnMaxLayers = 8, kWnMaxCopies = 12, kWnMaxJobs = 4, kWnCopyChunk = 2048;

struct WnLayout {
  int n_layers, n_copies, units, pad;                  // units = row blocks + copy chunks of one network
  long long w_off[kWnMaxLayers], g_off[kWnMaxLayers];
  int out[kWnMaxLayers], in[kWnMaxLayers], unit_begin[kWnMaxLayers];
  long long c_off[kWnMaxCopies];
  int c_len[kWnMaxCopies], c_unit_begin[kWnMaxCopies];
};

struct WnJob {
  const float* src;       // packed parameters (weight_g / weight_v)
  float* dst;             // mode 0: effective parameters; mode 1: gradient buffer (converted in place)
  long long net_stride;   // floats between consecutive networks of the buffer
  int n_nets, layout, tile_begin, pad;
};

struct WnParams {
  int mode, n_jobs, rows_per_tile, total_tiles;
  int rn_out, pad_w;      // mode 0: effective weights stored rounded to nearest TF32 (tensor-core operands)
  WnJob job[kWnMaxJobs];
  WnLayout lay[2];
};

__device__ __forceinline__ void wn_body(const WnParams& P, int tile) {
  int j = 0;
#pragma unroll
  for (int i = 1; i < kWnMaxJobs; ++i)
    if (i < P.n_jobs && tile >= P.job[i].tile_begin) j = i;
  const WnJob& J = P.job[j];
  const WnLayout& Y = P.lay[J.layout];
  const int local = tile - J.tile_begin;
  const int net = local / Y.units, unit = local - net * Y.units;
  const float* src = J.src + (long long)net * J.net_stride;
  float* dst = J.dst + (long long)net * J.net_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (Y.n_copies > 0 && unit >= Y.c_unit_begin[0]) {       // pass-through tensors (biases, encoder): mode 0 only
    int cseg = 0;
    for (int i = 1; i < Y.n_copies; ++i)
      if (unit >= Y.c_unit_begin[i]) cseg = i;
    const int beg = (unit - Y.c_unit_begin[cseg]) * kWnCopyChunk;
    const int end = min(beg + kWnCopyChunk, Y.c_len[cseg]);
    const long long off = Y.c_off[cseg];
    for (int i = beg + (int)threadIdx.x; i < end; i += blockDim.x) dst[off + i] = src[off + i];
    return;
  }
  int l = 0;
  for (int i = 1; i < Y.n_layers; ++i)
    if (unit >= Y.unit_begin[i]) l = i;
  const int row = (unit - Y.unit_begin[l]) * P.rows_per_tile + warp;
  if (warp >= P.rows_per_tile || row >= Y.out[l]) return;
  const int K = Y.in[l];
  const float* v = src + Y.w_off[l] + (long long)row * K;
  float* o = dst + Y.w_off[l] + (long long)row * K;
  const float g = src[Y.g_off[l] + row];
  float ss = 0.f, dot = 0.f;
  if (P.mode == 0) {
    for (int k = lane; k < K; k += 32) { const float x = v[k]; ss = fmaf(x, x, ss); }
  } else {
    for (int k = lane; k < K; k += 32) { const float x = v[k]; ss = fmaf(x, x, ss); dot = fmaf(o[k], x, dot); }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, d);
    dot += __shfl_xor_sync(0xffffffffu, dot, d);
  }
  const float nrm = sqrtf(ss);
  if (P.mode == 0) {
    const float sc = g / nrm;
    for (int k = lane; k < K; k += 32) {
      const float w = v[k] * sc;
      o[k] = P.rn_out ? rn_tf32(w) : w;
    }
  } else {
    const float dg = dot / nrm, sc = g / nrm, back = dg / nrm;
    for (int k = lane; k < K; k += 32) o[k] = sc * (o[k] - back * v[k]);
    if (lane == 0) dst[Y.g_off[l] + row] = dg;
  }
}

__global__ void __launch_bounds__(256) wn_kernel(const __grid_constant__ WnParams P) {
  pdl_launch_dependents();
  pdl_wait();
  wn_body(P, blockIdx.x);
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
constexpr int kMaxPeers = 8;       // one NVSwitch domain

// 16 bytes from a peer GPU's memory: system-scope relaxed load (never served from a stale local cache line)
__device__ __forceinline__ void peer_load4(const float* p, float (&v)[4]) {
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p) : "memory");
}

struct EwRange {
  float* p; const float* g; float* m; float* v; float* tgt;
  float* p_sh; float* tgt_sh;        // optional TF32-rounded shadows of p / tgt (what the tensor-core contractions read)
  // data-parallel update: n_peer > 0 -> the gradient is the sum over the ranks' buffers g_peer[0..n_peer), read over
  // NVLink peer mappings and added in rank order (every replica forms the bit-identical sum; no all-reduce pass)
  const float* g_peer[kMaxPeers];
  int n_peer, pad_e;
  long long n, blk_begin;
  const float* sc_ptr;               // device-resident {step_size, sqrt(1 - beta2^t)} written by adam_tick, or
  float step_size, bc2_sqrt;         // host-supplied scalars when sc_ptr == nullptr
  int do_adam, do_polyak;
  long long skip_period, skip_len;   // > 0: elements with (index % skip_period) < skip_len belong to the fused
                                     // first-layer tiles of the same launch (apply.cuh) and are left alone here
};

struct EwParams {
  int n_ranges;
  EwRange r[3];
  double beta1, beta2, eps, tau;
};

constexpr int kEwThreads = 256;
constexpr int kEwPerBlock = kEwThreads * 4;   // one float4 per thread

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
__device__ __forceinline__ float adam_element(float p, float g, float& m, float& v, float w1, float b2, float w2, float bc2s,
                                          float eps, float neg_step) {
  m = fmaf(w1, __fsub_rn(g, m), m);
  v = __fadd_rn(__fmul_rn(v, b2), __fmul_rn(__fmul_rn(w2, g), g));
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2s), eps);
  return __fadd_rn(p, __fdiv_rn(__fmul_rn(neg_step, m), denom));
}

// bx = block index within the launch / stage.  Each thread owns one float4 of every buffer: all of its loads are
// issued before anything is consumed (the kernel is a link of the update's dependency chain: one L2 round trip).
__device__ __forceinline__ void adam_polyak_body(const EwParams& E, long long bx) {
  int ri = 0;
  for (int q = 1; q < 3; ++q)
    if (q < E.n_ranges && bx >= E.r[q].blk_begin) ri = q;
  const EwRange& R = E.r[ri];
  const long long e = (bx - R.blk_begin) * kEwPerBlock + (long long)threadIdx.x * 4;
  if (e >= R.n) return;
  if (R.skip_period > 0 && (e % R.skip_period) < R.skip_len) return;    // both multiples of 4: a float4 never straddles
  float pv[4], gv[4], mv[4], vv[4], tv[4];
  ew_load4(R.p, e, R.n, pv);
  if (R.do_adam) {
    if (R.n_peer > 0) {              // sum of the ranks' gradients, in rank order (n is a multiple of 4: packed layout)
      float pg[kMaxPeers][4];
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r)
        if (r < R.n_peer) peer_load4(R.g_peer[r] + e, pg[r]);
#pragma unroll
      for (int k = 0; k < 4; ++k) gv[k] = pg[0][k];
#pragma unroll
      for (int r = 1; r < kMaxPeers; ++r)
        if (r < R.n_peer) {
#pragma unroll
          for (int k = 0; k < 4; ++k) gv[k] = __fadd_rn(gv[k], pg[r][k]);
        }
    } else {
      ew_load4(R.g, e, R.n, gv);
    }
    ew_load4(R.m, e, R.n, mv);
    ew_load4(R.v, e, R.n, vv);
  }
  if (R.do_polyak) ew_load4(R.tgt, e, R.n, tv);
  const float s_step_size = R.do_adam ? (R.sc_ptr ? __ldcg(R.sc_ptr) : R.step_size) : 0.f;
  const float s_bc2_sqrt = R.do_adam ? (R.sc_ptr ? __ldcg(R.sc_ptr + 1) : R.bc2_sqrt) : 1.f;
  const float w1 = (float)(1.0 - E.beta1), b2 = (float)E.beta2, w2 = (float)(1.0 - E.beta2);
  const float eps = (float)E.eps, tau = (float)E.tau, omt = (float)(1.0 - E.tau);
  const float neg_step = -s_step_size;
  if (R.do_adam) {
#pragma unroll
    for (int k = 0; k < 4; ++k) pv[k] = adam_element(pv[k], gv[k], mv[k], vv[k], w1, b2, w2, s_bc2_sqrt, eps, neg_step);
    ew_store4(R.m, e, R.n, mv);
    ew_store4(R.v, e, R.n, vv);
    ew_store4(R.p, e, R.n, pv);
    if (R.p_sh) {
      float sv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) sv[k] = rn_tf32(pv[k]);
      ew_store4(R.p_sh, e, R.n, sv);
    }
  }
  if (R.do_polyak) {
#pragma unroll
    for (int k = 0; k < 4; ++k) tv[k] = __fadd_rn(__fmul_rn(tau, pv[k]), __fmul_rn(omt, tv[k]));
    ew_store4(R.tgt, e, R.n, tv);
    if (R.tgt_sh) {
#pragma unroll
      for (int k = 0; k < 4; ++k) tv[k] = rn_tf32(tv[k]);
      ew_store4(R.tgt_sh, e, R.n, tv);
    }
  }
}

// ------------------------------------------------------------------------------------
// Data-parallel update, p2p mode: the ranks meet at flag words in each other's (symmetric) memory.
//   flags of rank r: [slot][sender] 32-bit counters, written by the senders with system-scope release stores.
//   signal(slot): my count for the slot += 1, stored to flags[slot][my rank] on every rank.
//   wait(slot):   spin until flags[slot][r] on MY device has reached my own count for every r (all ranks signal
//                 the same number of times, so "my count" is what everybody must have reached).
// Slots: "gradient complete" before the peer reads of an Adam kernel, "done reading" before the next backward pass
// overwrites the gradient (one pair for the critic family, one for the actor family).
// ------------------------------------------------------------------------------------
struct DpSyncParams {
  unsigned int* peer_flags[kMaxPeers];   // flag arrays of all ranks (mine included), peer-mapped
  unsigned int* counts;                  // my per-slot signal counts (device-resident: graph replays need no host)
  int world, rank, signal_slot, wait_slot;   // slot < 0: skip that half
};

__global__ void dp_signal_wait_kernel(const __grid_constant__ DpSyncParams P) {
  const int r = threadIdx.x;
  __shared__ unsigned int s_cnt[2];
  if (r == 0) {
    if (P.signal_slot >= 0) {
      s_cnt[0] = P.counts[P.signal_slot] + 1u;
      P.counts[P.signal_slot] = s_cnt[0];
    }
    if (P.wait_slot >= 0) s_cnt[1] = P.wait_slot == P.signal_slot ? s_cnt[0] : P.counts[P.wait_slot];
    __threadfence_system();              // everything earlier launches wrote is visible system-wide before the flag is
  }
  __syncthreads();
  if (r < P.world) {
    if (P.signal_slot >= 0) {
      unsigned int* f = P.peer_flags[r] + P.signal_slot * kMaxPeers + P.rank;
      asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(f), "r"(s_cnt[0]) : "memory");
    }
    if (P.wait_slot >= 0) {
      const unsigned int* f = P.peer_flags[P.rank] + P.wait_slot * kMaxPeers + r;
      unsigned int v;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(f) : "memory");
      } while ((int)(v - s_cnt[1]) < 0);
    }
  }
}

// out[e] = sum over r of peer[r][e], in rank order (the reduction the fused Adam kernel performs, on its own)
struct PeerSumParams {
  const float* peer[kMaxPeers];
  float* out;
  long long n;
  int world, pad;
};

__global__ void __launch_bounds__(kEwThreads) peer_sum_kernel(const __grid_constant__ PeerSumParams P) {
  const long long e = (long long)blockIdx.x * kEwPerBlock + (long long)threadIdx.x * 4;
  if (e >= P.n) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (e + 4 <= P.n) {
    for (int r = 0; r < P.world; ++r) {
      float v[4];
      peer_load4(P.peer[r] + e, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = r == 0 ? v[k] : __fadd_rn(acc[k], v[k]);
    }
  } else {
    for (int r = 0; r < P.world; ++r)
      for (int k = 0; k < 4; ++k)
        if (e + k < P.n) {
          const float v = *reinterpret_cast<const volatile float*>(P.peer[r] + e + k);
          acc[k] = r == 0 ? v : __fadd_rn(acc[k], v);
        }
  }
  ew_store4(P.out, e, P.n, acc);
}

// dst[e] = rn_tf32(src[e]): (re)builds the TF32 shadow of a packed parameter buffer after the caller changed it
struct RoundCopyParams {
  int n_ranges, pad;
  const float* src[4]; float* dst[4];
  long long n[4], blk_begin[4];
};

__global__ void __launch_bounds__(kEwThreads) round_copy_kernel(const __grid_constant__ RoundCopyParams R) {
  int ri = 0;
  for (int q = 1; q < 4; ++q)
    if (q < R.n_ranges && (long long)blockIdx.x >= R.blk_begin[q]) ri = q;
  const long long e = ((long long)blockIdx.x - R.blk_begin[ri]) * kEwPerBlock + (long long)threadIdx.x * 4;
  if (e >= R.n[ri]) return;
  float v[4];
  ew_load4(R.src[ri], e, R.n[ri], v);
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = rn_tf32(v[k]);
  ew_store4(R.dst[ri], e, R.n[ri], v);
}

__global__ void __launch_bounds__(kEwThreads) adam_polyak_kernel(const __grid_constant__ EwParams E) {
  pdl_launch_dependents();
  pdl_wait();
  adam_polyak_body(E, blockIdx.x);
}

}  // namespace td3
