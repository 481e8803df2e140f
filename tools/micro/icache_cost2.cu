// Finer cold-code probe: per-node time of a graph chain when K distinct kernels (or K code regions of ONE kernel)
// of N unrolled FMAs each take turns.  Floor (empty node) is ~0.7 us.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

template <int N, int TAG>
__device__ __forceinline__ void body(float* out, float a) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
#pragma unroll
  for (int i = 0; i < N / 4; ++i) {
    x0 = fmaf(x0, a, 0.5f + TAG); x1 = fmaf(x1, a, 1.5f); x2 = fmaf(x2, a, 2.5f); x3 = fmaf(x3, a, 3.5f);
  }
  if (x0 + x1 + x2 + x3 == 12345.f) out[threadIdx.x] = x0;
}
template <int N, int TAG>
__global__ void k_one(float* out, float a, int) { body<N, TAG>(out, a); }
template <int N>
__global__ void k_uber(float* out, float a, int tag) {
  switch (tag) {
    case 0: body<N, 10>(out, a); break;
    case 1: body<N, 11>(out, a); break;
    case 2: body<N, 12>(out, a); break;
    default: body<N, 13>(out, a); break;
  }
}
typedef void (*Fn)(float*, float, int);
static void run(const char* name, std::vector<Fn> fns, bool uber_tags, float* dev, int grid = 148, int block = 32) {
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaGraph_t g; cudaGraphExec_t ge;
  const int chain = 12;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < chain; ++i) fns[i % fns.size()]<<<grid, block, 0, s>>>(dev, 1.0001f, uber_tags ? i % 4 : 0);
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
  printf("%-60s %.2f us/node (%s)\n", name, ms * 1000.f / reps / chain, cudaGetErrorString(cudaGetLastError()));
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(s);
}
#define SET(N) \
  run(#N " FMA: one kernel", {k_one<N, 0>}, false, dev); \
  run(#N " FMA: 2 kernels alternate", {k_one<N, 0>, k_one<N, 1>}, false, dev); \
  run(#N " FMA: 4 kernels rotate", {k_one<N, 0>, k_one<N, 1>, k_one<N, 2>, k_one<N, 3>}, false, dev); \
  run(#N " FMA: one uber kernel, 4 regions rotate", {k_uber<N>}, true, dev); \
  run(#N " FMA: one uber kernel, 1 region", {k_uber<N>}, false, dev);
int main() {
  float* dev; cudaMalloc(&dev, 4096);
  SET(256) SET(512) SET(1024) SET(2048) SET(4096)
  run("1024 FMA: 4 kernels rotate, 148x256", {k_one<1024, 0>, k_one<1024, 1>, k_one<1024, 2>, k_one<1024, 3>}, false, dev, 148, 256);
  run("1024 FMA: 4 kernels rotate, 16x512", {k_one<1024, 0>, k_one<1024, 1>, k_one<1024, 2>, k_one<1024, 3>}, false, dev, 16, 512);
  run("1024 FMA: one kernel, 16x512", {k_one<1024, 0>}, false, dev, 16, 512);
  cudaFree(dev);
  return 0;
}
