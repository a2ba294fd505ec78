// Where do the warps of small CTAs land?  Records (%smid, %warpid) per warp and times an HMMA-bound loop per warp, for
// the launch shapes of the latency-form kernels (many 1- or 2-warp CTAs, ~3 per SM).
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma8(float* c, const float* a, float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
        "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__global__ void k(int* info, long long* cyc, float* out, int iters, int heavy_mask) {
  const int warp = threadIdx.x >> 5, gw = blockIdx.x * (blockDim.x >> 5) + warp;
  unsigned smid, warpid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
  float a[4] = {1.f + threadIdx.x, 2.f, 3.f, 4.f};
  float c[4][4];
  for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) c[j][i] = j + i;
  const bool heavy = (heavy_mask >> warp) & 1;
  long long t0 = clock64();
  if (heavy) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 4; ++j) mma8(c[j], a, 0.5f, 0.25f);
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) { info[2 * gw] = smid; info[2 * gw + 1] = warpid; cyc[gw] = t1 - t0; }
}
int main() {
  int* info; long long* cyc; float* out;
  const int maxw = 4096;
  cudaMalloc(&info, maxw * 8); cudaMalloc(&cyc, maxw * 8); cudaMalloc(&out, maxw * 32 * 4);
  static int hi[maxw * 2]; static long long hc[maxw];
  const int iters = 2000;
  struct { int grid, block, mask; const char* name; } cfg[] = {
    {450, 32, 1, "450 CTAs x 1 warp (fwd latency form)"},
    {450, 64, 1, "450 CTAs x 2 warps, warp 0 heavy (bwd latency form)"},
    {450, 64, 3, "450 CTAs x 2 warps, both heavy"},
    {225, 128, 5, "225 CTAs x 4 warps, warps 0 and 2 heavy"},
    {148, 128, 15, "148 CTAs x 4 warps, all heavy (one CTA per SM)"},
    {592, 32, 1, "592 CTAs x 1 warp"},
  };
  for (auto& c : cfg) {
    k<<<c.grid, c.block>>>(info, cyc, out, iters, c.mask);
    cudaDeviceSynchronize();
    const int nw = c.grid * (c.block / 32);
    cudaMemcpy(hi, info, nw * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, nw * 8, cudaMemcpyDeviceToHost);
    // per-SM sub-partition occupancy by heavy warps, assuming sub-partition = warpid % 4
    static int cnt[256][4]; memset(cnt, 0, sizeof(cnt));
    double worst = 0, sum = 0; int nh = 0;
    for (int w = 0; w < nw; ++w) {
      const int lw = w % (c.block / 32);
      if (!((c.mask >> lw) & 1)) continue;
      cnt[hi[2 * w]][hi[2 * w + 1] & 3]++;
      const double per = (double)hc[w] / iters / 4;
      worst = per > worst ? per : worst; sum += per; ++nh;
    }
    int hist[8] = {0};
    for (int s = 0; s < 256; ++s) for (int q = 0; q < 4; ++q) hist[cnt[s][q] > 7 ? 7 : cnt[s][q]]++;
    printf("%-52s cycles/mma: mean %.2f worst %.2f | sub-partitions holding 1/2/3/4 heavy warps: %d/%d/%d/%d | first warps (sm,warpid):", c.name, sum / nh, worst, hist[1], hist[2], hist[3], hist[4]);
    for (int w = 0; w < 8 && w < nw; ++w) printf(" (%d,%d)", hi[2 * w], hi[2 * w + 1]);
    printf("\n");
  }
  return 0;
}
