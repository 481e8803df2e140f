// Cycles per tcgen05.mma as a function of the instruction shape, measured on one SM: a chain of `n` MMAs of shape
// M x N x K8 (kind::tf32, operands K-major SWIZZLE_128B in shared memory, SS form) or M x N x K16 (kind::f16, bf16
// operands), all accumulating into the same TMEM tile, then one commit; clock64 around issue -> accumulator complete.
// The operand bytes are whatever shared memory holds (the result is never read): only the timing matters.
// Question it answers: is the ~100 cycles per M128 x K8 MMA the update's K loops see a
// per-instruction cost, does it scale with M (A rows read from shared memory) or with N?
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long make_desc(unsigned smem_addr) {   // K-major, SW128: LBO 16 B, SBO 1024 B
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr >> 4) & 0x3FFF);
  d |= (unsigned long long)(16 >> 4) << 16;
  d |= (unsigned long long)(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

template <int KIND>   // 0 = tf32, 1 = bf16
__global__ void mma_chain(int M, int N, int n, long long* out, int slots) {   // slots: distinct 32 KB operand chunks walked (1 = same bytes every time)
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned tmem_base;
  __shared__ __align__(8) unsigned long long bar;
  const unsigned base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem + (base - smem_u32(smem)))[i] = 0.001f * (i & 255);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    // instruction descriptor: D fp32 (1 << 4); A/B format tf32 = 2, bf16 = 1 at bits 7 and 10; N >> 3 at 17; M >> 4 at 24
    const unsigned fmt = KIND == 0 ? 2u : 1u;
    const unsigned idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
    const unsigned long long da0 = make_desc(base), db0 = make_desc(base + 16384);
    const unsigned tm = tmem_base;
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      const unsigned chunk = (unsigned)((i >> 2) % slots) * (32768u >> 4);        // a different operand chunk every 4 MMAs, as a K loop does
      const unsigned long long da = da0 + chunk + (unsigned long long)((i & 3) * 2), db = db0 + chunk + (unsigned long long)((i & 3) * 2);   // +32 B per K step
      const unsigned acc = i > 0;
      if (KIND == 0)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile(
        "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
}

int main() {
  long long* dev; cudaMalloc(&dev, 16);
  const int smem_bytes = 200 * 1024;
  cudaFuncSetAttribute(mma_chain<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(mma_chain<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int n = 64;
  printf("chain of %d MMAs on one SM; cycles per MMA = (accumulator complete - first issue) / %d\n", n, n);
  for (int slots : {1, 6})
  for (int kind = 0; kind < 2; ++kind)
    for (int M : {128, 64})
      for (int N : {16, 32, 64, 128, 256}) {
        if (slots > 1 && N > 128) continue;            // B region of a chunk is 16 KB
        if (M == 128 && N % 16) continue;
        long long h[2] = {0, 0};
        for (int rep = 0; rep < 3; ++rep) {
          if (kind == 0) mma_chain<0><<<1, 128, smem_bytes>>>(M, N, n, dev, slots);
          else mma_chain<1><<<1, 128, smem_bytes>>>(M, N, n, dev, slots);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("%s M=%d N=%d: %s\n", kind ? "bf16" : "tf32", M, N, cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h, dev, 16, cudaMemcpyDeviceToHost);
        }
        printf("%d chunk(s) %s M=%3d N=%3d K=%2d : issue %6.1f cycles/MMA, complete %6.1f cycles/MMA\n", slots, kind ? "bf16" : "tf32", M, N, kind ? 16 : 8,
               (double)h[0] / n, (double)h[1] / n);
        (void)0;
      }
  cudaFree(dev);
  return 0;
}
