// Multi-GPU exchange step of the path, fused with the optimiser: ONE kernel that sums the ranks' flat gradient vectors
// over NVLink peer memory and applies Adam (vihds/training.py:82, :336 on every rank; the reference itself is
// single-device).  It replaces the pair ncclAllReduce (176 KB message: ~25-30 us of latency at any rank count) +
// adam_dev_kernel of the training step.
//
// Protocol ("push", one epoch per call, every rank runs the same sequence of calls):
//   A. every rank stores its gradient into slot [epoch parity][its rank] of EVERY rank's inbox (remote stores through
//      the CUDA-IPC mapping of the peers' buffers) and clears its own gradient vector;
//   B. every thread block then publishes `epoch + 1` (and its rank's NaN verdict) in flag [parity][block][its rank] of
//      every rank: system-scope fence + release store, one thread per destination rank;
//   C. and waits until the `world` flags of the SAME block index in its own buffer show `epoch + 1` -- block c of every
//      rank owns the same slice of the vector --, then sums the `world` inbox slots of its slice in rank order (the same
//      order on every rank, so all ranks apply bit-identical updates) and applies Adam;
//   D. the last block to finish advances the local epoch / step counters.
// Inbox slots alternate with the epoch parity: block c of a peer can only push epoch e + 2 after it has seen this rank's
// block-c flag of epoch e + 1, which is published in the NEXT launch, i.e. after this launch has finished reading epoch e.
// The grid is sized so that every block is resident (blocks spin in C while others may still be in A).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "../../include/vihds_b200.h"
#include "vh_math.cuh"
#include "vh_pdl.cuh"

namespace vh {
void set_error(const char* fmt, ...);

// buffer layout (bytes): [0, PEER_FLAG_BYTES): uint64 flags [2 parities][PEER_MAX_CTAS][VH_PEER_MAX_WORLD]
// |  [PEER_FLAG_BYTES, ...): inbox R [2][world][n_pad]
// One flag per (parity, CTA, source rank): CTA c of every rank owns the same slice of the flat vector (the grids are
// identical), so it only has to wait for CTA c of its peers.  A flag holds ((epoch + 1) << 1) | bad.
constexpr int PEER_MAX_CTAS = 512;
constexpr int PEER_FLAG_BYTES = 2 * PEER_MAX_CTAS * VH_PEER_MAX_WORLD * 8;

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
// publishes a "bad" bit with its flags, every CTA ORs the bits of all ranks and all of them skip the update together (the
// reference stops before optimizer.step(), vihds/training.py:331-333); state[3] / step[2] count the skipped calls.
// A wait that exceeds timeout_ns marks the exchange as timed out (sticky) and NO block of this rank applies the update (the
// blocks vote before they touch the parameters); once the flag is set every later call returns at once, so the peers time
// out as well and every host finds the flag.
//
// Per CTA: push the slice to every rank's inbox -> system fence -> one flag per peer -> wait for the same CTA of every peer
// -> sum the inboxes + Adam on the slice.  (Round 1 had ONE flag per rank, published by the last CTA to finish pushing: a
// grid-wide ticket, two system fences and 2 x world remote stores by a single thread sat on every step's critical path.)
template <typename R>
__global__ void __launch_bounds__(256) adam_allreduce_kernel(size_t n, size_t n_pad, R* __restrict__ p, R* __restrict__ g,
                                                             R* __restrict__ m, R* __restrict__ v,
                                                             const double* __restrict__ hyper, long long* step,
                                                             long long* state, int rank, int world,
                                                             unsigned char* const* __restrict__ peers,
                                                             const R* __restrict__ guard, unsigned long long timeout_ns,
                                                             const R* __restrict__ wg_dpre, const R* __restrict__ wg_pooled,
                                                             int wg_B, int wg_H, int wg_NLIN, long long wg_off) {
  const bool failed = *(volatile long long*)(state + 2) != 0;  // the exchange has failed before: nothing may be applied
  const unsigned long long epoch = (unsigned long long)*(volatile long long*)state;
  const double t = (double)(*(volatile long long*)step + 1);
  const int par = (int)(epoch & 1);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t fslot = ((size_t)par * PEER_MAX_CTAS + blockIdx.x) * VH_PEER_MAX_WORLD;
  __shared__ int s_skip;
  if (threadIdx.x == 0) s_skip = 0;
  // Launched with programmatic stream serialization behind the encoder backward (vh_pdl.cuh): the optimiser state of this
  // thread's first element is fetched while that launch is still running; the gradient is read after pdl_wait().
  R m0 = R(0), v0 = R(0), p0 = R(0);
  if (i0 < n) {
    m0 = m[i0];
    v0 = v[i0];
    p0 = p[i0];
  }
  pdl_wait();  // also on the failure path: a launch that left without it could complete before its predecessor
  if (failed) return;
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
  // B. publish: thread r tells rank r that this CTA's slice has landed.  The CTA barrier orders the other threads' stores
  // before this thread's system-scope fence (fences are cumulative), the release store follows the fence.
  if (threadIdx.x < world) {
    const R c = guard ? *guard : R(0);
    // sticky, like vh_adam_step_dev: state[3] advances on every rank together, so the verdict stays collective
    const unsigned long long bad = ((c != c) || (guard && *(volatile long long*)(state + 3) != 0)) ? 1ULL : 0ULL;
    __threadfence_system();
    unsigned long long* flags = reinterpret_cast<unsigned long long*>(peers[threadIdx.x]);
    st_release_sys(flags + fslot + rank, ((epoch + 1) << 1) | bad);
    // C. wait for the same CTA of rank r (bounded: a lost peer must not wedge the GPU)
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peers[rank]) + fslot + threadIdx.x;
    const unsigned long long t0 = global_ns();
    unsigned long long seen;
    bool ok = true;
    while ((seen = ld_acquire_sys(mine)) < ((epoch + 1) << 1)) {
      if (global_ns() - t0 > timeout_ns) {
        ok = false;
        break;
      }
      __nanosleep(32);
    }
    if (!ok) {
      state[2] = 1;
      atomicOr(&s_skip, 2);
    } else if (seen & 1ULL) {
      atomicOr(&s_skip, 1);
    }
  }
  __syncthreads();
  // All or nothing on this rank: a block applies its slice only when EVERY block of the grid has received its peers'
  // flags (local vote on the ticket counter; the grid is resident).  A block that timed out has set the sticky flag
  // instead of voting, which releases the others without an update.
  if (threadIdx.x == 0) {
    if (!(s_skip & 2)) {
      __threadfence();
      atomicAdd((unsigned long long*)(state + 1), 1ULL);
      const unsigned long long t0 = global_ns();
      while (*(volatile long long*)(state + 1) < (long long)gridDim.x) {
        if (*(volatile long long*)(state + 2) != 0 || global_ns() - t0 > timeout_ns) {
          state[2] = 1;
          s_skip |= 2;
          break;
        }
      }
    }
  }
  __syncthreads();
  const int skip = s_skip;
  if (!skip) {
    const double lr = hyper[0], b1d = hyper[1], b2d = hyper[2];
    const R b1 = (R)b1d, b2 = (R)b2d, eps = (R)hyper[3];
    const R bc1 = (R)(1.0 - pow(b1d, t));
    const R bc2_sqrt = (R)sqrt(1.0 - pow(b2d, t));
    const R* inbox = reinterpret_cast<const R*>(peers[rank] + PEER_FLAG_BYTES) + (size_t)par * world * n_pad;
    for (size_t i = i0; i < n; i += stride) {
      R gi = R(0);
      for (int r = 0; r < world; ++r) gi += __ldcv(inbox + (size_t)r * n_pad + i);  // written by peers: bypass L1
      const bool first = i == i0;
      const R mo = first ? m0 : m[i], vo = first ? v0 : v[i], po = first ? p0 : p[i];
      const R mi = mo + (gi - mo) * (R(1) - b1);
      const R vi = b2 * vo + (R(1) - b2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      const R denom = vsqrt(vi) / bc2_sqrt + eps;
      p[i] = po - ((R)lr / bc1) * (mi / denom);
    }
  }
  // D. the last CTA to finish advances the epoch and the step / skip counters, once per call.  Every CTA has read them by
  // then; the bad bits are the same for every CTA (one guard per rank), a time-out (skip & 2) is sticky in state[2].
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd((unsigned long long*)(state + 1), 1ULL);
    if (ticket == 2ULL * gridDim.x - 1) {  // votes + finishers (never reached after a time-out: the exchange is dead then)
      state[1] = 0;
      state[0] = (long long)(epoch + 1);
      if (skip == 0)
        step[0] += 1;
      else if (skip == 1) {
        step[2] += 1;
        state[3] += 1;
      }
    }
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
  size_t cap = (size_t)(sms > 0 ? sms : 1) * 2;
  if (cap > (size_t)PEER_MAX_CTAS) cap = PEER_MAX_CTAS;  // one flag row per CTA in the exchange buffer
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  unsigned char* const* pp = (unsigned char* const*)peers;
  const unsigned long long tns = (unsigned long long)((timeout_s > 0 ? timeout_s : 10.0) * 1e9);
  if (dtype == VH_F32)
    launch_maybe_pdl(adam_allreduce_kernel<float>, dim3(grid), dim3(block), 0, s, true, n, n_pad, (float*)param, (float*)grad,
                     (float*)exp_avg, (float*)exp_avg_sq, (const double*)hyper, (long long*)step, (long long*)state, rank, world,
                     pp, (const float*)guard, tns, (const float*)(wg ? wg->d_pre : nullptr),
                     (const float*)(wg ? wg->pooled : nullptr), wg ? wg->B : 0, wg ? wg->H : 0, wg ? wg->NLIN : 0,
                     (long long)(wg ? wg->offset : 0));
  else if (dtype == VH_F64)
    launch_maybe_pdl(adam_allreduce_kernel<double>, dim3(grid), dim3(block), 0, s, true, n, n_pad, (double*)param, (double*)grad,
                     (double*)exp_avg, (double*)exp_avg_sq, (const double*)hyper, (long long*)step, (long long*)state, rank, world,
                     pp, (const double*)guard, tns, (const double*)(wg ? wg->d_pre : nullptr),
                     (const double*)(wg ? wg->pooled : nullptr), wg ? wg->B : 0, wg ? wg->H : 0, wg ? wg->NLIN : 0,
                     (long long)(wg ? wg->offset : 0));
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
