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

// phase timestamps of block 0 of the last exchange launch (ns, %globaltimer): start, after griddepcontrol.wait, pushed,
// published, peers' flags seen, vote passed, update applied.  Read with vh_peer_debug_times (tools/peer_timing.py).
__device__ unsigned long long vh_peer_dbg[8];

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
// V elements per thread (1, or 16 bytes' worth when the launcher finds everything 16-byte aligned): 16-byte remote stores,
// one 16-byte load of the pooled features per V multiply-adds of the weight gradient.
template <typename R, int V>
struct PeerVec {
  struct alignas(sizeof(R) * V) T {
    R v[V];
  };
};
__device__ __forceinline__ void ldcv_vec(const float* p, PeerVec<float, 4>::T& o) {
  const float4 t = __ldcv(reinterpret_cast<const float4*>(p));
  o.v[0] = t.x; o.v[1] = t.y; o.v[2] = t.z; o.v[3] = t.w;
}
__device__ __forceinline__ void ldcv_vec(const double* p, PeerVec<double, 2>::T& o) {
  const double2 t = __ldcv(reinterpret_cast<const double2*>(p));
  o.v[0] = t.x; o.v[1] = t.y;
}
__device__ __forceinline__ void ldcv_vec(const float* p, PeerVec<float, 1>::T& o) { o.v[0] = __ldcv(p); }
__device__ __forceinline__ void ldcv_vec(const double* p, PeerVec<double, 1>::T& o) { o.v[0] = __ldcv(p); }

template <typename R, int V>
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
  if (blockIdx.x == 0 && threadIdx.x == 0) vh_peer_dbg[0] = global_ns();
  typedef typename PeerVec<R, V>::T Vec;
  const size_t stride = (size_t)gridDim.x * blockDim.x * V;
  const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  const size_t fslot = ((size_t)par * PEER_MAX_CTAS + blockIdx.x) * VH_PEER_MAX_WORLD;
  __shared__ int s_skip;
  if (threadIdx.x == 0) s_skip = 0;
  // Launched with programmatic stream serialization behind the encoder backward (vh_pdl.cuh): the optimiser state of this
  // thread's first elements is fetched while that launch is still running; the gradient is read after pdl_wait().
  Vec m0 = {}, v0 = {}, p0 = {};
  if (i0 < n) {
    m0 = *reinterpret_cast<const Vec*>(m + i0);
    v0 = *reinterpret_cast<const Vec*>(v + i0);
    p0 = *reinterpret_cast<const Vec*>(p + i0);
  }
  pdl_wait();  // also on the failure path: a launch that left without it could complete before its predecessor
  if (failed) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) vh_peer_dbg[1] = global_ns();
  // A. push
  for (size_t i = i0; i < n; i += stride) {
    Vec val = *reinterpret_cast<const Vec*>(g + i);
    *reinterpret_cast<Vec*>(g + i) = Vec{};
    // optional: the encoder's hidden-layer weight gradient dW[o][c] = sum_b d_pre[b][o] pooled[b][c] is formed here, on its
    // way to the peers, instead of by a launch of its own (small batches; vh_encoder.cu enc_lin_wgrad_small_kernel).
    // V > 1: the launcher guarantees that a thread's elements share their row o (offset and NLIN multiples of V)
    const long long e = (long long)i - wg_off;
    if (wg_dpre && e >= 0 && e < (long long)wg_H * wg_NLIN) {
      const int o = (int)(e / wg_NLIN), c = (int)(e % wg_NLIN);
      Vec a0 = {}, a1 = {};
      const R* dp = wg_dpre + o;
      const R* pl = wg_pooled + c;
      int b = 0;
#pragma unroll 6
      for (; b + 1 < wg_B; b += 2) {
        const R w0 = dp[(size_t)b * wg_H], w1 = dp[(size_t)(b + 1) * wg_H];
        const Vec x0 = *reinterpret_cast<const Vec*>(pl + (size_t)b * wg_NLIN);
        const Vec x1 = *reinterpret_cast<const Vec*>(pl + (size_t)(b + 1) * wg_NLIN);
#pragma unroll
        for (int k = 0; k < V; ++k) {
          a0.v[k] += w0 * x0.v[k];
          a1.v[k] += w1 * x1.v[k];
        }
      }
      if (b < wg_B) {
        const R w0 = dp[(size_t)b * wg_H];
        const Vec x0 = *reinterpret_cast<const Vec*>(pl + (size_t)b * wg_NLIN);
#pragma unroll
        for (int k = 0; k < V; ++k) a0.v[k] += w0 * x0.v[k];
      }
#pragma unroll
      for (int k = 0; k < V; ++k) val.v[k] += a0.v[k] + a1.v[k];
    }
    for (int r = 0; r < world; ++r) {
      R* inbox = reinterpret_cast<R*>(peers[r] + PEER_FLAG_BYTES);
      *reinterpret_cast<Vec*>(inbox + ((size_t)par * world + rank) * n_pad + i) = val;
    }
  }
  __syncthreads();  // every thread of this CTA has pushed and has read epoch / step
  if (blockIdx.x == 0 && threadIdx.x == 0) vh_peer_dbg[2] = global_ns();
  // B. publish: thread r tells rank r that this CTA's slice has landed (system-scope release store).
  if (threadIdx.x < world) {
    const R c = guard ? *guard : R(0);
    // sticky, like vh_adam_step_dev: state[3] advances on every rank together, so the verdict stays collective
    const unsigned long long bad = ((c != c) || (guard && *(volatile long long*)(state + 3) != 0)) ? 1ULL : 0ULL;
    // st.release.sys IS the fence: it orders this thread's earlier accesses and, cumulatively, the other threads' pushes that
    // the CTA barrier above put before it.  (An explicit __threadfence_system() in front -- fence.sc.sys -- cost 10.5 us per
    // step with 2 ranks: tools/peer_timing.py.)
    unsigned long long* flags = reinterpret_cast<unsigned long long*>(peers[threadIdx.x]);
    st_release_sys(flags + fslot + rank, ((epoch + 1) << 1) | bad);
    if (blockIdx.x == 0 && threadIdx.x == 0) vh_peer_dbg[3] = global_ns();
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
  if (blockIdx.x == 0 && threadIdx.x == 0) vh_peer_dbg[4] = global_ns();
  // All or nothing on this rank: a block applies its slice only when EVERY block of the grid has received its peers'
  // flags (local vote on the ticket counter; the grid is resident).  A block that timed out has set the sticky flag
  // instead of voting, which releases the others without an update.
  if (threadIdx.x == 0) {
    if (!(s_skip & 2)) {
      atomicAdd((unsigned long long*)(state + 1), 1ULL);  // a count of verdicts: nothing is published through it
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
  if (blockIdx.x == 0 && threadIdx.x == 0) vh_peer_dbg[5] = global_ns();
  const int skip = s_skip;
  if (!skip) {
    const double lr = hyper[0], b1d = hyper[1], b2d = hyper[2];
    const R b1 = (R)b1d, b2 = (R)b2d, eps = (R)hyper[3];
    const R bc1 = (R)(1.0 - pow(b1d, t));
    const R bc2_sqrt = (R)sqrt(1.0 - pow(b2d, t));
    const R* inbox = reinterpret_cast<const R*>(peers[rank] + PEER_FLAG_BYTES) + (size_t)par * world * n_pad;
    for (size_t i = i0; i < n; i += stride) {
      Vec gi = {};
      for (int r = 0; r < world; ++r) {  // written by peers: bypass L1
        Vec t;
        ldcv_vec(inbox + (size_t)r * n_pad + i, t);
#pragma unroll
        for (int k = 0; k < V; ++k) gi.v[k] += t.v[k];
      }
      const bool first = i == i0;
      Vec mo = first ? m0 : *reinterpret_cast<const Vec*>(m + i), vo = first ? v0 : *reinterpret_cast<const Vec*>(v + i);
      Vec po = first ? p0 : *reinterpret_cast<const Vec*>(p + i);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const R mi = mo.v[k] + (gi.v[k] - mo.v[k]) * (R(1) - b1);
        const R vi = b2 * vo.v[k] + (R(1) - b2) * gi.v[k] * gi.v[k];
        mo.v[k] = mi;
        vo.v[k] = vi;
        const R denom = vsqrt(vi) / bc2_sqrt + eps;
        po.v[k] = po.v[k] - ((R)lr / bc1) * (mi / denom);
      }
      *reinterpret_cast<Vec*>(m + i) = mo;
      *reinterpret_cast<Vec*>(v + i) = vo;
      *reinterpret_cast<Vec*>(p + i) = po;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) vh_peer_dbg[6] = global_ns();
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

// ---------------------------------------------------------------------------------------------------------------
// Low-latency form of the same step, opt-in (fp32, one 16-byte group of parameters per thread, the whole vector in one grid pass):
// every 4-byte gradient value travels WITH its epoch tag in one 8-byte word -- {value, ((epoch + 1) << 1) | bad} -- so the
// receiver validates each word by itself and no fence, flag or second message is needed: the sender's part is two 16-byte
// remote stores per destination, the receiver polls the words of its own elements.  (With flags the system-scope release
// that orders 2 x world remote pushes before the flag cost 7 us per step even with two ranks: tools/peer_timing.py.)
// 8-byte words are single-copy atomic; a 16-byte vector access is two of them.  The inbox holds 8 bytes per element:
// [2 parities][world][n_pad] words behind the (then unused) flag area; a slot is reused two epochs later, when its words
// carry a different tag.  Everything else -- NaN verdict, time-out, vote, counters -- is as in adam_allreduce_kernel.
// ---------------------------------------------------------------------------------------------------------------
struct alignas(16) LLPair {
  unsigned long long w[2];
};
__device__ __forceinline__ void st_ll(LLPair* p, const LLPair& v) {
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.w[0]), "l"(v.w[1]) : "memory");
}
__device__ __forceinline__ LLPair ld_ll(const LLPair* p) {
  LLPair v;
  asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.w[0]), "=l"(v.w[1]) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) adam_allreduce_ll_kernel(size_t n, size_t n_pad, float* __restrict__ p, float* __restrict__ g,
                                                                float* __restrict__ m, float* __restrict__ v,
                                                                const double* __restrict__ hyper, long long* step,
                                                                long long* state, int rank, int world,
                                                                unsigned char* const* __restrict__ peers,
                                                                const float* __restrict__ guard, unsigned long long timeout_ns,
                                                                const float* __restrict__ wg_dpre,
                                                                const float* __restrict__ wg_pooled, int wg_B, int wg_H,
                                                                int wg_NLIN, long long wg_off) {
  if (blockIdx.x == 0 && threadIdx.x == 0) vh_peer_dbg[0] = global_ns();
  typedef PeerVec<float, 4>::T Vec;
  const bool failed = *(volatile long long*)(state + 2) != 0;
  const unsigned long long epoch = (unsigned long long)*(volatile long long*)state;
  const double t = (double)(*(volatile long long*)step + 1);
  const int par = (int)(epoch & 1);
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;  // this thread's only group (launcher: n / 4 <= threads)
  const bool on = i < n;
  __shared__ int s_skip;
  if (threadIdx.x == 0) s_skip = 0;
  Vec m0 = {}, v0 = {}, p0 = {};
  if (on) {
    m0 = *reinterpret_cast<const Vec*>(m + i);
    v0 = *reinterpret_cast<const Vec*>(v + i);
    p0 = *reinterpret_cast<const Vec*>(p + i);
  }
  pdl_wait();
  if (failed) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) vh_peer_dbg[1] = global_ns();
  const float c = guard ? *guard : 0.f;
  const unsigned bad = ((c != c) || (guard && *(volatile long long*)(state + 3) != 0)) ? 1u : 0u;
  const unsigned want = (unsigned)(((epoch + 1) << 1) & 0xfffffffeu);
  const unsigned long long tag = (unsigned long long)(want | bad) << 32;
  // A. push
  if (on) {
    Vec val = *reinterpret_cast<const Vec*>(g + i);
    *reinterpret_cast<Vec*>(g + i) = Vec{};
    const long long e = (long long)i - wg_off;
    if (wg_dpre && e >= 0 && e < (long long)wg_H * wg_NLIN) {
      const int o = (int)(e / wg_NLIN), cc = (int)(e % wg_NLIN);
      Vec a0 = {}, a1 = {};
      const float* dp = wg_dpre + o;
      const float* pl = wg_pooled + cc;
      int b = 0;
#pragma unroll 6
      for (; b + 1 < wg_B; b += 2) {
        const float w0 = dp[(size_t)b * wg_H], w1 = dp[(size_t)(b + 1) * wg_H];
        const Vec x0 = *reinterpret_cast<const Vec*>(pl + (size_t)b * wg_NLIN);
        const Vec x1 = *reinterpret_cast<const Vec*>(pl + (size_t)(b + 1) * wg_NLIN);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          a0.v[k] += w0 * x0.v[k];
          a1.v[k] += w1 * x1.v[k];
        }
      }
      if (b < wg_B) {
        const float w0 = dp[(size_t)b * wg_H];
        const Vec x0 = *reinterpret_cast<const Vec*>(pl + (size_t)b * wg_NLIN);
#pragma unroll
        for (int k = 0; k < 4; ++k) a0.v[k] += w0 * x0.v[k];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) val.v[k] += a0.v[k] + a1.v[k];
    }
    LLPair lo, hi;
    lo.w[0] = tag | __float_as_uint(val.v[0]);
    lo.w[1] = tag | __float_as_uint(val.v[1]);
    hi.w[0] = tag | __float_as_uint(val.v[2]);
    hi.w[1] = tag | __float_as_uint(val.v[3]);
    for (int r = 0; r < world; ++r) {
      LLPair* inbox = reinterpret_cast<LLPair*>(peers[r] + PEER_FLAG_BYTES) + (((size_t)par * world + rank) * n_pad + i) / 2;
      st_ll(inbox, lo);
      st_ll(inbox + 1, hi);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) vh_peer_dbg[2] = vh_peer_dbg[3] = global_ns();
  // C. every thread collects its own elements from every rank, in rank order (bit-identical sums everywhere)
  Vec gi = {};
  if (on) {
    unsigned seen_bad = 0;
    bool ok = true;
    const unsigned long long t0 = global_ns();
    for (int r = 0; r < world && ok; ++r) {
      const LLPair* inbox =
          reinterpret_cast<const LLPair*>(peers[rank] + PEER_FLAG_BYTES) + (((size_t)par * world + r) * n_pad + i) / 2;
      LLPair lo, hi;
      unsigned spins = 0;
      for (;;) {
        lo = ld_ll(inbox);
        hi = ld_ll(inbox + 1);
        const unsigned t0w = (unsigned)(lo.w[0] >> 32), t1w = (unsigned)(lo.w[1] >> 32), t2w = (unsigned)(hi.w[0] >> 32),
                       t3w = (unsigned)(hi.w[1] >> 32);
        if ((t0w & ~1u) == want && (t1w & ~1u) == want && (t2w & ~1u) == want && (t3w & ~1u) == want) {
          seen_bad |= (t0w | t1w | t2w | t3w) & 1u;
          break;
        }
        if ((++spins & 63u) == 0 && global_ns() - t0 > timeout_ns) {
          ok = false;
          break;
        }
      }
      gi.v[0] += __uint_as_float((unsigned)lo.w[0]);
      gi.v[1] += __uint_as_float((unsigned)lo.w[1]);
      gi.v[2] += __uint_as_float((unsigned)hi.w[0]);
      gi.v[3] += __uint_as_float((unsigned)hi.w[1]);
    }
    if (!ok) {
      state[2] = 1;
      atomicOr(&s_skip, 2);
    } else if (seen_bad) {
      atomicOr(&s_skip, 1);
    }
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) vh_peer_dbg[4] = global_ns();
  if (threadIdx.x == 0) {
    if (!(s_skip & 2)) {
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
  if (blockIdx.x == 0 && threadIdx.x == 0) vh_peer_dbg[5] = global_ns();
  const int skip = s_skip;
  if (!skip && on) {
    const double lr = hyper[0], b1d = hyper[1], b2d = hyper[2];
    const float b1 = (float)b1d, b2 = (float)b2d, eps = (float)hyper[3];
    const float bc1 = (float)(1.0 - pow(b1d, t));
    const float bc2_sqrt = (float)sqrt(1.0 - pow(b2d, t));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float mi = m0.v[k] + (gi.v[k] - m0.v[k]) * (1.f - b1);
      const float vi = b2 * v0.v[k] + (1.f - b2) * gi.v[k] * gi.v[k];
      m0.v[k] = mi;
      v0.v[k] = vi;
      const float denom = vsqrt(vi) / bc2_sqrt + eps;
      p0.v[k] = p0.v[k] - ((float)lr / bc1) * (mi / denom);
    }
    *reinterpret_cast<Vec*>(m + i) = m0;
    *reinterpret_cast<Vec*>(v + i) = v0;
    *reinterpret_cast<Vec*>(p + i) = p0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) vh_peer_dbg[6] = global_ns();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd((unsigned long long*)(state + 1), 1ULL);
    if (ticket == 2ULL * gridDim.x - 1) {
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
  (void)dtype;  // 8 bytes per element for both: fp64 values, or fp32 value + epoch tag (low-latency form)
  const size_t n_pad = (n + 63) & ~(size_t)63;
  return PEER_FLAG_BYTES + 2 * (size_t)world * n_pad * 8;
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

int vh_peer_debug_times(unsigned long long* out8) {
  return cudaMemcpyFromSymbol(out8, vh_peer_dbg, sizeof(unsigned long long) * 8) == cudaSuccess ? VH_OK : VH_ERR_CUDA;
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
  const size_t es = dtype == VH_F64 ? 8 : 4;
  const int vmax = (int)(16 / es);
  // 16 bytes per thread when the vector, its views and the weight-gradient operands allow it
  const size_t al = (size_t)param | (size_t)grad | (size_t)exp_avg | (size_t)exp_avg_sq | (wg ? (size_t)wg->pooled : 0);
  const bool vec = n % vmax == 0 && (al & 15) == 0 && (!wg || (wg->offset % vmax == 0 && wg->NLIN % vmax == 0));
  const size_t per_thread = vec ? vmax : 1;
  // every block must be resident (blocks wait for peers while others may still be pushing): at most 2 per SM
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  size_t want = (n / per_thread + block - 1) / block;
  size_t cap = (size_t)(sms > 0 ? sms : 1) * 2;
  if (cap > (size_t)PEER_MAX_CTAS) cap = PEER_MAX_CTAS;  // one flag row per CTA in the exchange buffer
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  unsigned char* const* pp = (unsigned char* const*)peers;
  const unsigned long long tns = (unsigned long long)((timeout_s > 0 ? timeout_s : 10.0) * 1e9);
#define VH_PEER_LAUNCH(R, V)                                                                                                   \
  launch_maybe_pdl(adam_allreduce_kernel<R, V>, dim3(grid), dim3(block), 0, s, true, n, n_pad, (R*)param, (R*)grad,             \
                   (R*)exp_avg, (R*)exp_avg_sq, (const double*)hyper, (long long*)step, (long long*)state, rank, world, pp,     \
                   (const R*)guard, tns, (const R*)(wg ? wg->d_pre : nullptr), (const R*)(wg ? wg->pooled : nullptr),           \
                   wg ? wg->B : 0, wg ? wg->H : 0, wg ? wg->NLIN : 0, (long long)(wg ? wg->offset : 0))
  // low-latency form (VIHDS_PEER_LL=1; fp32, 16 bytes per thread, the whole vector in one pass of a resident grid).  Measured
  // against the flag protocol inside real steps: 2 GPUs 0.1287 vs 0.1293 ms per step, 8 GPUs 0.1407 vs 0.1357 ms -- twice the
  // bytes and world x as many polling loads cost more than the release fence they replace, so the flag protocol is the
  // default.  Every rank takes the same branch (same n, same alignment classes, same environment).
  const char* llenv = getenv("VIHDS_PEER_LL");
  // 64-thread blocks: the remote stores of a step leave through as many SMs as possible
  const int ll_block = 64;
  const size_t ll_want = (n / 4 + ll_block - 1) / ll_block;
  const bool ll = dtype == VH_F32 && vec && ll_want <= cap && llenv && *llenv == '1';
  if (ll) {
    launch_maybe_pdl(adam_allreduce_ll_kernel, dim3((unsigned)ll_want), dim3(ll_block), 0, s, true, n, n_pad, (float*)param, (float*)grad,
                     (float*)exp_avg, (float*)exp_avg_sq, (const double*)hyper, (long long*)step, (long long*)state, rank, world,
                     pp, (const float*)guard, tns, (const float*)(wg ? wg->d_pre : nullptr),
                     (const float*)(wg ? wg->pooled : nullptr), wg ? wg->B : 0, wg ? wg->H : 0, wg ? wg->NLIN : 0,
                     (long long)(wg ? wg->offset : 0));
  } else if (dtype == VH_F32) {
    if (vec) VH_PEER_LAUNCH(float, 4); else VH_PEER_LAUNCH(float, 1);
  } else if (dtype == VH_F64) {
    if (vec) VH_PEER_LAUNCH(double, 2); else VH_PEER_LAUNCH(double, 1);
  }
#undef VH_PEER_LAUNCH
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
