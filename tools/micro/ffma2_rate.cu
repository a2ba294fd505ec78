// Issue rate and dependent latency of the packed fp32 FMA (fma.rn.f32x2 -> FFMA2) against the scalar FFMA on sm_100a.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
// Per variant: `warps` warps per SM sub-partition (one CTA of 4*warps warps per SM), each running CHAINS independent
// dependency chains of ITER FMAs; cycles per warp-instruction per sub-partition = elapsed / (warps * CHAINS * ITER).
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, bool PACKED>
__global__ void fma_kernel(float* out, long long* cyc, int iters, float a0, float b0) {
  unsigned long long v[CHAINS];
  float s[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) {
    s[c] = threadIdx.x * 1e-3f + c;
    const float lo = s[c], hi = s[c] + 0.5f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v[c]) : "f"(lo), "f"(hi));
  }
  unsigned long long a, b;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(a0));
  asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(b0));
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (PACKED)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[c]) : "l"(a), "l"(b));
      else
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[c]) : "f"(a0), "f"(b0));
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[c]));
    acc += lo + hi + s[c];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CHAINS, bool PACKED>
void run(int warps_per_sp, const char* name) {
  const int sms = 148, threads = 128 * warps_per_sp, iters = 4096;
  float* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * threads);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  for (int r = 0; r < 2; ++r) fma_kernel<CHAINS, PACKED><<<sms, threads>>>(out, cyc, iters, 0.999f, 1e-3f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double m = 0;
  for (int i = 0; i < sms; ++i) m += h[i];
  m /= sms;
  const double per = m / ((double)warps_per_sp * CHAINS * iters);
  printf("%-8s chains %2d warps/subpartition %d: %.2f cycles per warp-instruction and sub-partition (%s)\n", name, CHAINS,
         warps_per_sp, per, CHAINS == 1 && warps_per_sp == 1 ? "dependent latency" : "issue interval");
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<1, false>(1, "FFMA");
  run<1, true>(1, "FFMA2");
  run<8, false>(1, "FFMA");
  run<8, true>(1, "FFMA2");
  run<8, false>(4, "FFMA");
  run<8, true>(4, "FFMA2");
  run<8, false>(8, "FFMA");
  run<8, true>(8, "FFMA2");
  return 0;
}
