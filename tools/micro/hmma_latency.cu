// Microbenchmark: latency (dependent chain) and issue interval (independent chains) of mma.sync.m16n8k8 TF32 on sm_100a,
// one warp per SM sub-partition.  nvcc -gencode arch=compute_100a,code=sm_100a -o hmma_latency hmma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma8(float* c, const float* a, float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
        "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
template <int CHAINS>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[4] = {1.f + threadIdx.x, 2.f, 3.f, 4.f};
  float c[CHAINS][4];
  for (int j = 0; j < CHAINS; ++j) for (int i = 0; i < 4; ++i) c[j][i] = j + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) mma8(c[j], a, 0.5f, 0.25f);
  }
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < CHAINS; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
// A-operand dependency: result of one mma feeds the A fragment of the next (the MLP layer chaining)
__global__ void kdep_a(float* out, long long* cyc, int iters) {
  float a[4] = {1e-3f * threadIdx.x, 2e-3f, 3e-3f, 4e-3f};
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    mma8(c, a, 0.5f, 0.25f);
    a[0] = c[0]; a[1] = c[2]; a[2] = c[1]; a[3] = c[3];
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a[0] + a[1] + a[2] + a[3];
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void kffma(float* out, long long* cyc, int iters) {
  float x = threadIdx.x * 1e-3f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) x = fmaf(x, 1.0001f, 0.5f);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void kmufu(float* out, long long* cyc, int iters) {
  float x = threadIdx.x * 1e-3f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) { float e; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); x = e * 0.5f; }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  long long h;
#define RUN(CH) k<CH><<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("HMMA.1688.TF32 %d independent chain(s): %.2f cycles per mma, %.2f per round\n", CH, (double)h / iters / CH, (double)h / iters);
  RUN(1) RUN(2) RUN(3) RUN(4) RUN(6) RUN(8)
  k<4><<<1, 128>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("4 warps (one per sub-partition) x 4 chains: %.2f cycles per mma per warp\n", (double)h / iters / 4);
  k<4><<<1, 256>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("8 warps (two per sub-partition) x 4 chains: %.2f cycles per mma per warp\n", (double)h / iters / 4);
  kdep_a<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("mma -> 4 movs -> mma (A-operand dependency): %.2f cycles per round\n", (double)h / iters);
  kffma<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dependent FFMA: %.2f cycles\n", (double)h / iters);
  kmufu<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dependent MUFU.EX2 + FMUL: %.2f cycles\n", (double)h / iters);
  return 0;
}
