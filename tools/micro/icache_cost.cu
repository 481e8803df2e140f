// Cost of running cold code: a chain of graph nodes in which kernel A (N unrolled independent FMAs per thread,
// executed once) alternates with kernel B of the same shape, against the same arithmetic in a rolled loop.
// If straight-line code is fetched at F bytes/cycle when cold, time(unrolled) - time(rolled) ~ 16 N / F cycles.
#include <cuda_runtime.h>
#include <cstdio>

template <int N, int TAG>
__global__ void k_unrolled(float* out, float a) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
#pragma unroll
  for (int i = 0; i < N / 4; ++i) {
    x0 = fmaf(x0, a, 0.5f + TAG); x1 = fmaf(x1, a, 1.5f); x2 = fmaf(x2, a, 2.5f); x3 = fmaf(x3, a, 3.5f);
  }
  if (x0 + x1 + x2 + x3 == 12345.f) out[threadIdx.x] = x0;
}
template <int N>
__global__ void k_rolled(float* out, float a) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
#pragma unroll 1
  for (int i = 0; i < N / 4; ++i) {
    x0 = fmaf(x0, a, 0.5f); x1 = fmaf(x1, a, 1.5f); x2 = fmaf(x2, a, 2.5f); x3 = fmaf(x3, a, 3.5f);
  }
  if (x0 + x1 + x2 + x3 == 12345.f) out[threadIdx.x] = x0;
}

typedef void (*Fn)(float*, float);
static void run(const char* name, Fn f0, Fn f1, int grid, int block, float* dev) {
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaGraph_t g; cudaGraphExec_t ge;
  const int chain = 12;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < chain; ++i) ((i & 1) ? f1 : f0)<<<grid, block, 0, s>>>(dev, 1.0001f);
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
  printf("%-64s %.2f us/node (%s)\n", name, ms * 1000.f / reps / chain, cudaGetErrorString(cudaGetLastError()));
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(s);
}

int main() {
  float* dev; cudaMalloc(&dev, 4096);
  run("rolled 2048 FMA, 148x32", k_rolled<2048>, k_rolled<2048>, 148, 32, dev);
  run("unrolled 2048 (32 KB), same kernel every node, 148x32", k_unrolled<2048, 0>, k_unrolled<2048, 0>, 148, 32, dev);
  run("unrolled 2048 (32 KB), two kernels alternate, 148x32", k_unrolled<2048, 0>, k_unrolled<2048, 1>, 148, 32, dev);
  run("rolled 8192 FMA, 148x32", k_rolled<8192>, k_rolled<8192>, 148, 32, dev);
  run("unrolled 8192 (128 KB), same kernel every node, 148x32", k_unrolled<8192, 0>, k_unrolled<8192, 0>, 148, 32, dev);
  run("unrolled 8192 (128 KB), two kernels alternate, 148x32", k_unrolled<8192, 0>, k_unrolled<8192, 1>, 148, 32, dev);
  run("rolled 16384 FMA, 148x32", k_rolled<16384>, k_rolled<16384>, 148, 32, dev);
  run("unrolled 16384 (256 KB), same kernel every node, 148x32", k_unrolled<16384, 0>, k_unrolled<16384, 0>, 148, 32, dev);
  run("unrolled 16384 (256 KB), two kernels alternate, 148x32", k_unrolled<16384, 0>, k_unrolled<16384, 1>, 148, 32, dev);
  run("unrolled 8192, alternate, 148x256", k_unrolled<8192, 0>, k_unrolled<8192, 1>, 148, 256, dev);
  run("rolled 8192, 148x256", k_rolled<8192>, k_rolled<8192>, 148, 256, dev);
  run("unrolled 8192, alternate, 16x512", k_unrolled<8192, 0>, k_unrolled<8192, 1>, 16, 512, dev);
  run("unrolled 2048, alternate with unrolled 16384, 148x32", k_unrolled<2048, 0>, k_unrolled<16384, 1>, 148, 32, dev);
  cudaFree(dev);
  return 0;
}
