// Multi-GPU exchange step of the path, fused with the optimiser: ONE kernel that sums the ranks' flat gradient vectors
// over NVLink peer memory and applies Adam (vihds/training.py:82, :336 on every rank; the reference itself is
// single-device).  It replaces the pair ncclAllReduce (176 KB message: ~25-30 us of latency at any rank count) +
// adam_dev_kernel of the training step.
//
// Protocol ("push", one epoch per call, every rank runs the same sequence of calls):
//   A. every rank stores its gradient into slot [epoch parity][its rank] of EVERY rank's inbox (remote stores through
//      the CUDA-IPC mapping of the peers' buffers) and clears its own gradient vector;
//   B. the last thread block to finish A publishes `epoch + 1` in flag [parity][its rank] of every rank
//      (fence + system-scope release store) and advances the local epoch / step counters;
//   C. every block waits until all `world` flags of its own rank show `epoch + 1`, then sums the `world` inbox slots
//      in rank order -- the same order on every rank, so all ranks apply bit-identical updates -- and applies Adam.
// Inbox slots alternate with the epoch parity: a peer can only push epoch e + 2 after it has seen this rank's flag
// of epoch e + 1, which this rank publishes after it has finished reading epoch e (stream order).
// The grid is sized so that every block is resident (blocks spin in C while others may still be in A).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "../../include/vihds_b200.h"
#include "vh_math.cuh"

namespace vh {
void set_error(const char* fmt, ...);

// buffer layout (bytes): [0, 4096): uint64 [4][VH_PEER_MAX_WORLD] = flags of parity 0, 1, "bad" words of parity 0, 1
// |  [4096, ...): inbox R [2][world][n_pad]
constexpr int PEER_FLAG_BYTES = 4096;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// state: int64[4] = {epoch, ticket, timed_out (sticky), skipped steps}; step: int64[4] as for vh_adam_step_dev.
// guard: this rank's cost of the step (or NULL).  A NaN cost anywhere means a NaN gradient sum everywhere: every rank
// publishes a "bad" word next to its flag, every rank ORs all of them and skips the update together (the reference
// stops before optimizer.step(), vihds/training.py:331-333); state[3] / step[2] count the skipped calls.
// A wait that exceeds timeout_ns marks the exchange as timed out (sticky) and the CTA skips its part of the update; once
// the flag is set every later call returns at once, so the peers time out as well and every host finds the flag.
template <typename R>
__global__ void __launch_bounds__(256) adam_allreduce_kernel(size_t n, size_t n_pad, R* __restrict__ p, R* __restrict__ g,
                                                             R* __restrict__ m, R* __restrict__ v,
                                                             const double* __restrict__ hyper, long long* step,
                                                             long long* state, int rank, int world,
                                                             unsigned char* const* __restrict__ peers,
                                                             const R* __restrict__ guard, unsigned long long timeout_ns,
                                                             const R* __restrict__ wg_dpre, const R* __restrict__ wg_pooled,
                                                             int wg_B, int wg_H, int wg_NLIN, long long wg_off) {
  if (*(volatile long long*)(state + 2) != 0) return;  // the exchange has failed before: nothing may be applied
  const unsigned long long epoch = (unsigned long long)*(volatile long long*)state;
  const double t = (double)(*(volatile long long*)step + 1);
  const int par = (int)(epoch & 1);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ int s_skip, s_last;
  if (threadIdx.x == 0) s_skip = s_last = 0;
  // A. push
  for (size_t i = i0; i < n; i += stride) {
    R val = g[i];
    g[i] = R(0);
    // optional: the encoder's hidden-layer weight gradient dW[o][c] = sum_b d_pre[b][o] pooled[b][c] is formed here, on its
    // way to the peers, instead of by a launch of its own (small batches; vh_encoder.cu enc_lin_wgrad_small_kernel)
    const long long e = (long long)i - wg_off;
    if (wg_dpre && e >= 0 && e < (long long)wg_H * wg_NLIN) {
      const int o = (int)(e / wg_NLIN), c = (int)(e % wg_NLIN);
      R a0 = R(0), a1 = R(0);
      int b = 0;
#pragma unroll 6
      for (; b + 1 < wg_B; b += 2) {
        a0 += wg_dpre[(size_t)b * wg_H + o] * wg_pooled[(size_t)b * wg_NLIN + c];
        a1 += wg_dpre[(size_t)(b + 1) * wg_H + o] * wg_pooled[(size_t)(b + 1) * wg_NLIN + c];
      }
      if (b < wg_B) a0 += wg_dpre[(size_t)b * wg_H + o] * wg_pooled[(size_t)b * wg_NLIN + c];
      val += a0 + a1;
    }
    for (int r = 0; r < world; ++r) {
      R* inbox = reinterpret_cast<R*>(peers[r] + PEER_FLAG_BYTES);
      inbox[((size_t)par * world + rank) * n_pad + i] = val;
    }
  }
  __syncthreads();  // every thread of this CTA has pushed and has read epoch / step
  // B. last block publishes.  One system-scope fence per CTA, by the thread that takes the ticket: the CTA barrier
  // orders the other threads' stores before it and fences are cumulative (the cooperative-groups grid-sync pattern).
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned long long ticket = atomicAdd((unsigned long long*)(state + 1), 1ULL);
    if (ticket == (unsigned long long)gridDim.x - 1) {
      s_last = 1;
      const R c = guard ? *guard : R(0);
      // sticky, like vh_adam_step_dev: state[3] advances on every rank together, so the verdict stays collective
      const unsigned long long bad = ((c != c) || (guard && *(volatile long long*)(state + 3) != 0)) ? 1ULL : 0ULL;
      for (int r = 0; r < world; ++r) {
        unsigned long long* flags = reinterpret_cast<unsigned long long*>(peers[r]);
        flags[(2 + par) * VH_PEER_MAX_WORLD + rank] = bad;  // ordered before the flag by the release below
      }
      __threadfence_system();
      for (int r = 0; r < world; ++r) {
        unsigned long long* flags = reinterpret_cast<unsigned long long*>(peers[r]);
        st_release_sys(flags + par * VH_PEER_MAX_WORLD + rank, epoch + 1);
      }
      state[1] = 0;
      state[0] = (long long)(epoch + 1);
    }
  }
  // C. wait for every rank's push of this epoch (bounded: a lost peer must not wedge the GPU)
  if (threadIdx.x < world) {
    const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(peers[rank]);
    const unsigned long long t0 = global_ns();
    bool ok = true;
    while (ld_acquire_sys(flags + par * VH_PEER_MAX_WORLD + threadIdx.x) < epoch + 1) {
      if (global_ns() - t0 > timeout_ns) {
        ok = false;
        break;
      }
      __nanosleep(64);
    }
    if (!ok) {
      state[2] = 1;
      atomicOr(&s_skip, 2);
    } else if (ld_acquire_sys(flags + (2 + par) * VH_PEER_MAX_WORLD + threadIdx.x) != 0) {
      atomicOr(&s_skip, 1);
    }
  }
  __syncthreads();
  const int skip = s_skip;
  if (threadIdx.x == 0 && s_last) {  // the step / skip counters advance once per call, after the verdict
    if (skip == 0)
      step[0] += 1;
    else if (skip == 1) {
      step[2] += 1;
      state[3] += 1;
    }
  }
  if (skip) return;
  const double lr = hyper[0], b1d = hyper[1], b2d = hyper[2];
  const R b1 = (R)b1d, b2 = (R)b2d, eps = (R)hyper[3];
  const R bc1 = (R)(1.0 - pow(b1d, t));
  const R bc2_sqrt = (R)sqrt(1.0 - pow(b2d, t));
  const R* inbox = reinterpret_cast<const R*>(peers[rank] + PEER_FLAG_BYTES) + (size_t)par * world * n_pad;
  for (size_t i = i0; i < n; i += stride) {
    R gi = R(0);
    for (int r = 0; r < world; ++r) gi += __ldcv(inbox + (size_t)r * n_pad + i);  // written by peers: bypass L1
    const R mi = m[i] + (gi - m[i]) * (R(1) - b1);
    const R vi = b2 * v[i] + (R(1) - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const R denom = vsqrt(vi) / bc2_sqrt + eps;
    p[i] -= ((R)lr / bc1) * (mi / denom);
  }
}

}  // namespace vh

using namespace vh;

extern "C" {

size_t vh_peer_buffer_bytes(int dtype, size_t n, int world) {
  const size_t es = dtype == VH_F64 ? 8 : 4;
  const size_t n_pad = (n + 63) & ~(size_t)63;
  return PEER_FLAG_BYTES + 2 * (size_t)world * n_pad * es;
}

int vh_peer_buffer_create(size_t bytes, void** dev_ptr, void* handle64) {
  if (!dev_ptr || !handle64 || bytes == 0) {
    set_error("vh_peer_buffer_create: bad arguments");
    return VH_ERR_INVALID;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* ptr = nullptr;
  cudaError_t e = cudaMalloc(&ptr, bytes);
  if (e == cudaSuccess) e = cudaMemset(ptr, 0, bytes);
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr);
  if (e != cudaSuccess) {
    set_error("vh_peer_buffer_create(%zu bytes): %s", bytes, cudaGetErrorString(e));
    if (ptr) cudaFree(ptr);
    return VH_ERR_CUDA;
  }
  *dev_ptr = ptr;
  return VH_OK;
}

int vh_peer_buffer_open(const void* handle64, void** dev_ptr) {
  if (!dev_ptr || !handle64) {
    set_error("vh_peer_buffer_open: bad arguments");
    return VH_ERR_INVALID;
  }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("vh_peer_buffer_open: %s", cudaGetErrorString(e));
    return VH_ERR_CUDA;
  }
  return VH_OK;
}

int vh_peer_buffer_close(void* dev_ptr) { return cudaIpcCloseMemHandle(dev_ptr) == cudaSuccess ? VH_OK : VH_ERR_CUDA; }
int vh_peer_buffer_destroy(void* dev_ptr) { return cudaFree(dev_ptr) == cudaSuccess ? VH_OK : VH_ERR_CUDA; }

static int allreduce_step(int dtype, size_t n, void* param, void* grad, void* exp_avg, void* exp_avg_sq, const void* hyper,
                          void* step, void* state, int rank, int world, const void* peers, const void* guard, double timeout_s,
                          const vh_lin_wgrad* wg, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !hyper || !step || !state || !peers || n == 0 || world < 1 ||
      world > VH_PEER_MAX_WORLD || rank < 0 || rank >= world) {
    set_error("vh_adam_allreduce_step: bad arguments");
    return VH_ERR_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int block = 256;
  const size_t n_pad = (n + 63) & ~(size_t)63;
  // every block must be resident (blocks wait for peers while others may still be pushing): at most 2 per SM
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  size_t want = (n + block - 1) / block;
  const size_t cap = (size_t)(sms > 0 ? sms : 1) * 2;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  unsigned char* const* pp = (unsigned char* const*)peers;
  const unsigned long long tns = (unsigned long long)((timeout_s > 0 ? timeout_s : 10.0) * 1e9);
  if (dtype == VH_F32)
    adam_allreduce_kernel<float><<<grid, block, 0, s>>>(n, n_pad, (float*)param, (float*)grad, (float*)exp_avg,
                                                         (float*)exp_avg_sq, (const double*)hyper, (long long*)step,
                                                         (long long*)state, rank, world, pp, (const float*)guard, tns,
                                                         (const float*)(wg ? wg->d_pre : nullptr), (const float*)(wg ? wg->pooled : nullptr),
                                                         wg ? wg->B : 0, wg ? wg->H : 0, wg ? wg->NLIN : 0, wg ? wg->offset : 0);
  else if (dtype == VH_F64)
    adam_allreduce_kernel<double><<<grid, block, 0, s>>>(n, n_pad, (double*)param, (double*)grad, (double*)exp_avg,
                                                          (double*)exp_avg_sq, (const double*)hyper, (long long*)step,
                                                          (long long*)state, rank, world, pp, (const double*)guard, tns,
                                                          (const double*)(wg ? wg->d_pre : nullptr),
                                                          (const double*)(wg ? wg->pooled : nullptr), wg ? wg->B : 0,
                                                          wg ? wg->H : 0, wg ? wg->NLIN : 0, wg ? wg->offset : 0);
  else {
    set_error("unknown dtype %d", dtype);
    return VH_ERR_INVALID;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("adam_allreduce_kernel launch failed: %s", cudaGetErrorString(e));
    return VH_ERR_CUDA;
  }
  return VH_OK;
}

int vh_adam_allreduce_step(int dtype, size_t n, void* param, void* grad, void* exp_avg, void* exp_avg_sq, const void* hyper,
                           void* step, void* state, int rank, int world, const void* peers, const void* guard,
                           double timeout_s, void* stream) {
  return allreduce_step(dtype, n, param, grad, exp_avg, exp_avg_sq, hyper, step, state, rank, world, peers, guard, timeout_s,
                        nullptr, stream);
}

int vh_adam_allreduce_step_wgrad(int dtype, size_t n, void* param, void* grad, void* exp_avg, void* exp_avg_sq,
                                 const void* hyper, void* step, void* state, int rank, int world, const void* peers,
                                 const void* guard, double timeout_s, const vh_lin_wgrad* wg, void* stream) {
  if (wg && (!wg->d_pre || !wg->pooled || wg->B <= 0 || wg->B > 128 || wg->H <= 0 || wg->NLIN <= 0 || wg->offset < 0 ||
             (size_t)wg->offset + (size_t)wg->H * wg->NLIN > n)) {
    set_error("vh_adam_allreduce_step_wgrad: bad vh_lin_wgrad (B must be 1..128, the weight view must lie inside the flat vector)");
    return VH_ERR_INVALID;
  }
  return allreduce_step(dtype, n, param, grad, exp_avg, exp_avg_sq, hyper, step, state, rank, world, peers, guard, timeout_s, wg,
                        stream);
}

}  // extern "C"
