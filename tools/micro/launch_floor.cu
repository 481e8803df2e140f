// Floor of one node of a dependent kernel chain inside a CUDA graph on this GPU, as a function of what the nodes
// ask for: nothing, a large dynamic shared-memory carve-out, alternating carve-outs, a TMEM allocation.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/launch_floor tools/micro/launch_floor.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__global__ void k_empty(int* p) { if (p && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(p, 1); }
__global__ void k_smem(int* p) {
  extern __shared__ int sm[];
  if (threadIdx.x == 0) sm[0] = blockIdx.x;
  __syncthreads();
  if (p && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(p, sm[0] + 1);
}
__global__ void k_smem2(int* p) {   // a different function with its own attributes
  extern __shared__ int sm[];
  if (threadIdx.x == 0) sm[0] = blockIdx.x;
  __syncthreads();
  if (p && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(p, sm[0] + 1);
}
__global__ void k_tmem(int* p) {
  __shared__ unsigned slot;
  if (threadIdx.x < 32) {
    unsigned a = (unsigned)__cvta_generic_to_shared(&slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(a) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
  if (p && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(p, 1);
}

struct Node { void (*fn)(int*); int grid, block, smem; };

static float run(const char* name, std::vector<Node> pat, int chain, int* dev) {
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < chain; ++i) { const Node& n = pat[i % pat.size()]; n.fn<<<n.grid, n.block, n.smem, s>>>(dev); }
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  for (int i = 0; i < 20; ++i) cudaGraphLaunch(ge, s);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaStreamSynchronize(s);
  const int reps = 200;
  cudaEventRecord(e0, s);
  for (int i = 0; i < reps; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s);
  cudaStreamSynchronize(s);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const float us = ms * 1000.f / reps / chain;
  printf("%-58s %.2f us/node  (%s)\n", name, us, cudaGetErrorString(cudaGetLastError()));
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(s);
  return us;
}

int main() {
  int* dev; cudaMalloc(&dev, 4); cudaMemset(dev, 0, 4);
  const int big = 200 * 1024, mid = 72 * 1024;
  cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(k_smem2, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  run("empty 1x32", {{k_empty, 1, 32, 0}}, 12, dev);
  run("empty 148x256", {{k_empty, 148, 256, 0}}, 12, dev);
  run("empty 40x256", {{k_empty, 40, 256, 0}}, 12, dev);
  run("empty 296x256", {{k_empty, 296, 256, 0}}, 12, dev);
  run("smem 200K 148x256", {{k_smem, 148, 256, big}}, 12, dev);
  run("smem 72K 192x256", {{k_smem, 192, 256, mid}}, 12, dev);
  run("alternate smem 200K / empty", {{k_smem, 148, 256, big}, {k_empty, 255, 256, 0}}, 12, dev);
  run("alternate smem 200K / smem2 72K", {{k_smem, 148, 256, big}, {k_smem2, 192, 256, mid}}, 12, dev);
  run("alternate smem 200K / smem 72K (same function)", {{k_smem, 148, 256, big}, {k_smem, 192, 256, mid}}, 12, dev);
  cudaFuncSetAttribute(k_smem2, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(k_smem, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(k_empty, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  run("alternate 200K / 72K, carve-out 100 on both", {{k_smem, 148, 256, big}, {k_smem2, 192, 256, mid}}, 12, dev);
  run("alternate 200K / empty, carve-out 100 on both", {{k_smem, 148, 256, big}, {k_empty, 255, 256, 0}}, 12, dev);
  run("tmem alloc 512 cols 148x256", {{k_tmem, 148, 256, 0}}, 12, dev);
  run("tmem alloc 512 cols 40x256", {{k_tmem, 40, 256, 0}}, 12, dev);
  run("alternate tmem / smem 72K", {{k_tmem, 60, 256, 0}, {k_smem2, 192, 256, mid}}, 12, dev);
  cudaFree(dev);
  return 0;
}
