// infer.cuh -- the B = 1 forward the reference's environment loop calls every step:
//   select_action(state)      = actor(state)                       TD3_featured.py:113-115   (main.py:44,250)
//   eval_q(state, action)     = [Q1(s, a), Q2(s, a)]               TD3_featured.py:117-121   (main.py:45)
// One row through a 3-4 layer MLP is a chain of matrix-vector products: as stage launches it is a launch per layer
// plus two copies, i.e. pure latency.  Here ONE kernel does the whole call: a cluster of 8 CTAs per network reads the
// input row straight from the caller's pinned host buffer (zero-copy over PCIe), splits every layer's output neurons
// (one warp per neuron, all loads of its weight row in flight at once), exchanges the activations through distributed
// shared memory with one cluster barrier per layer, and stores the result plus a
// sequence word back to pinned host memory, where the host picks it up by polling that word: no memcpy nodes, no stream
// synchronise, no temporary tensors.  Strict fp32 (the reference's arithmetic); LayerNorm after ReLU as TD3_featured.py:44-46.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "misc.cuh"

namespace td3 {

constexpr int kInferThreads = 512, kInferMaxWidth = 2048, kInferMaxLayers = 8;
constexpr int kInferInline = 64;
constexpr int kInferCluster = 8;       // CTAs per network: a thread-block cluster that splits every layer's output neurons and
                                       // exchanges the activations through distributed shared memory

struct InferParams {
  const float* x_host;               // [dims[0]] input row in mapped pinned host memory
  float* out_host;                   // [n_nets][dims[L]] results, then n_nets sequence words (uint32), mapped pinned host memory
  const float* W;                    // packed parameters of network 0 (the effective ones under weight normalisation)
  long long net_stride;              // floats between the twin networks
  int n_nets, n_linear, ln, final_tanh;
  int dims[kInferMaxLayers + 1];
  long long w_off[kInferMaxLayers], b_off[kInferMaxLayers], lng_off[kInferMaxLayers], lnb_off[kInferMaxLayers];
  float out_scale;
  unsigned int seq;
  int x_inline, pad_i;               // 1: the input row travels in x[] with the launch (rows of <= kInferInline floats): no PCIe read
  float x[kInferInline];
};

// one output neuron: the warp's lanes stride over the K weights of its row; every load of the row is issued before the
// first is consumed (K <= 512: 16 per lane), so a row costs ONE L2 round trip
__device__ __forceinline__ float infer_row_dot(const float* __restrict__ row, const float* __restrict__ in, int K, int lane) {
  float acc = 0.f;
#pragma unroll 1
  for (int k0 = 0; k0 < K; k0 += 512) {
    float w[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int k = k0 + lane + 32 * u;
      w[u] = k < K ? __ldg(row + k) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int k = k0 + lane + 32 * u;
      if (k < K) acc = fmaf(w[u], in[k], acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// two rows at once (K <= 512)
__device__ __forceinline__ void infer_row_dot2(const float* __restrict__ r0, const float* __restrict__ r1, const float* __restrict__ in,
                                               int K, int lane, float& s0, float& s1) {
  float w0[16], w1[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int k = lane + 32 * u;
    w0[u] = k < K ? __ldg(r0 + k) : 0.f;
    w1[u] = k < K ? __ldg(r1 + k) : 0.f;
  }
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int k = lane + 32 * u;
    if (k < K) {
      const float x = in[k];
      a0 = fmaf(w0[u], x, a0);
      a1 = fmaf(w1[u], x, a1);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  s0 = a0; s1 = a1;
}

__device__ __forceinline__ unsigned int infer_cluster_rank() {
  unsigned int r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void infer_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// store v at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void infer_store_remote(float* local_addr, unsigned int rank, float v) {
  const unsigned int a = (unsigned int)__cvta_generic_to_shared(local_addr);
  unsigned int ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;\n" ::"r"(ra), "f"(v) : "memory");
}

// grid = n_nets * kInferCluster CTAs, clusters of kInferCluster: cluster c = network c
__global__ void __launch_bounds__(kInferThreads) infer_b1_kernel(const __grid_constant__ InferParams P) {
  __shared__ float act[2][kInferMaxWidth];
  __shared__ float red[2][kInferThreads / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int net = blockIdx.x / kInferCluster;
  const unsigned int crank = infer_cluster_rank();
  const float* W = P.W + (long long)net * P.net_stride;
  for (int k = tid; k < P.dims[0]; k += kInferThreads) act[0][k] = P.x_inline ? P.x[k] : __ldcv(P.x_host + k);
  __syncthreads();
  infer_cluster_sync();                  // every CTA of the cluster is running: remote shared memory may be written
  int cur = 0;
  for (int l = 0; l < P.n_linear; ++l) {
    const int K = P.dims[l], N = P.dims[l + 1];
    const bool last = l == P.n_linear - 1;
    const float* Wl = W + P.w_off[l];
    const float* bl = W + P.b_off[l];
    const float* in = act[cur];
    float* out = act[cur ^ 1];
    // this CTA's share of the output neurons, one warp each; the result goes to every CTA of the cluster
    const int per = (N + kInferCluster - 1) / kInferCluster;
    const int j_begin = crank * per, j_end = min(N, j_begin + per);
    constexpr int kW = kInferThreads / 32;
    for (int j = j_begin + warp; j < j_end; j += 2 * kW) {       // two neurons per pass: both rows' loads in flight together
      const int j2 = j + kW;
      float s0, s1 = 0.f;
      if (j2 < j_end && K <= 512) {
        infer_row_dot2(Wl + (long long)j * K, Wl + (long long)j2 * K, in, K, lane, s0, s1);
        s1 += __ldg(bl + j2);
      } else {
        s0 = infer_row_dot(Wl + (long long)j * K, in, K, lane);
        if (j2 < j_end) s1 = infer_row_dot(Wl + (long long)j2 * K, in, K, lane) + __ldg(bl + j2);
      }
      s0 += __ldg(bl + j);
      s0 = last ? (P.final_tanh ? P.out_scale * tanhf(s0) : s0) : fmaxf(s0, 0.f);
      s1 = last ? (P.final_tanh ? P.out_scale * tanhf(s1) : s1) : fmaxf(s1, 0.f);
      if (lane < kInferCluster) {
        infer_store_remote(out + j, (unsigned int)lane, s0);
        if (j2 < j_end) infer_store_remote(out + j2, (unsigned int)lane, s1);
      }
    }
    infer_cluster_sync();                // all neurons of the layer have landed in every CTA
    if (!last && P.ln) {                 // LayerNorm over the N post-ReLU activations (eps 1e-5, biased variance); every CTA its own copy
      float s = 0.f;
      for (int j = tid; j < N; j += kInferThreads) s += out[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) red[0][warp] = s;
      __syncthreads();
      float mean = 0.f;
      for (int w2 = 0; w2 < kInferThreads / 32; ++w2) mean += red[0][w2];
      mean /= (float)N;
      float ss = 0.f;
      for (int j = tid; j < N; j += kInferThreads) {
        const float d = out[j] - mean;
        ss = fmaf(d, d, ss);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (lane == 0) red[1][warp] = ss;
      __syncthreads();
      float var = 0.f;
      for (int w2 = 0; w2 < kInferThreads / 32; ++w2) var += red[1][w2];
      const float rstd = 1.f / sqrtf(var / (float)N + 1e-5f);
      const float* g = W + P.lng_off[l];
      const float* b = W + P.lnb_off[l];
      for (int j = tid; j < N; j += kInferThreads) out[j] = (out[j] - mean) * rstd * __ldg(g + j) + __ldg(b + j);
      __syncthreads();
    }
    cur ^= 1;
  }
  if (crank == 0) {
    const int NO = P.dims[P.n_linear];
    for (int j = tid; j < NO; j += kInferThreads) P.out_host[net * NO + j] = act[cur][j];
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      unsigned int* flags = reinterpret_cast<unsigned int*>(P.out_host + P.n_nets * NO);
      *reinterpret_cast<volatile unsigned int*>(flags + net) = P.seq;
    }
  }
  infer_cluster_sync();                  // nobody exits while a peer may still write into its shared memory
}

}  // namespace td3
