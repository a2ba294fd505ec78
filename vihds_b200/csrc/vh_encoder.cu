// Fused amortised encoder q(theta | x, d): forward in ONE launch, backward in two (vihds/encoders.py:16-55, :126-253,
// :383-404).  The reference runs Conv1d -> AvgPool1d -> Linear -> tanh and then 2 x n_param Linear(., 1) heads one by
// one; under PyTorch that is ~15 forward and ~35 backward launches of 2-7 us each, which after the ODE kernels were
// fused had become ~40 % of the ELBO-gradient step at the icml size.  Everything here is per individual and tiny
// (30 k + 36 k multiply-adds), so one CTA per individual does the whole chain out of shared memory:
//
//   enc_fwd_kernel   delta = diff(obs) -> conv (F x 4 x K) -> mean-pool (P) -> [saved] -> Linear + tanh [saved] ->
//                    packed heads (local: [enc, treatments?, devices?]; global-conditioned: [treatments?, devices?]) ->
//                    q_mu / q_prec rows (global and constant columns filled in the same pass)
//   enc_bwd_kernel   per individual: head / hidden-layer / pool / conv backward; small parameter gradients by atomics
//   enc_lin_wgrad    dW_lin[o][i] = sum_b d_pre[b][o] * pooled[b][i]   (one thread per weight, coalesced over i)
//
// Gradients are ACCUMULATED into the caller's buffers (views of the flat gradient vector, zeroed by the optimizer).
#include <cuda_runtime.h>

#include "../../include/vihds_b200.h"
#include "vh_math.cuh"
#include "vh_pdl.cuh"

namespace vh {
void set_error(const char* fmt, ...);

struct EncDims {
  int B, T, NS, F, K, PL, H, C, D;
  int nl, ng, nglob, nconst, P;
  int lt, ld, gt, gd;       // conditioning flags (treatments / devices) of the local and global-conditioned groups
  int L1, NCV, NP, NLIN;    // T-1, conv outputs per filter, pooled outputs per filter, F*NP
  int nin_l, nin_g;
  int stage;                // kernels copy their weights into shared memory up front (cp.async) -- see stage_async
  int phase;                // 0: whole chain in one launch; large batches run the hidden layer as GEMMs (enc_gemm_kernel)
                            // between two launches of these kernels: 1 = the part before it, 2 = the part after it
};

template <typename R>
struct EncPtrs {
  const R *obs, *inputs, *dev, *conv_w, *conv_b, *lin_w, *lin_b, *local_w, *local_b, *gcond_w, *global_free, *const_values;
  R *q_mu, *q_prec, *pooled, *enc;
  const R *d_q_mu, *d_q_prec;
  R *g_conv_w, *g_conv_b, *g_lin_w, *g_lin_b, *g_local_w, *g_local_b, *g_gcond_w, *g_global_free, *d_pre;
  R* dpool_g;  // large batches: cotangent of the pooled features [B][NLIN], produced by the GEMM between the two phases
};

template <typename R>
__device__ R warp_sum(R v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Cold-cache latency is what these kernels pay for (the step runs them once, with the weights in HBM): every phase
// that reads a weight array from global memory costs one or several dependent DRAM round trips (~0.8 us each; the
// hidden layer alone needed ~11).  So each kernel starts by issuing cp.async copies of EVERYTHING it will read later
// into shared memory -- one round trip, overlapped with the first phases -- and computes from there.
__device__ __forceinline__ void cp_async_bytes(void* dst, const void* src, int bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  if (bytes == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  else if (bytes == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
}
// n contiguous elements; 16-byte copies when both ends are 16-byte aligned (dst always is: carve-up in multiples of 4)
template <typename R>
__device__ void stage_async(R* dst, const R* src, int n, int tid, int nt) {
  if (n <= 0 || !src) return;
  constexpr int V = 16 / sizeof(R);
  if ((((size_t)src | (size_t)dst) & 15) == 0) {
    const int nv = n / V;
    for (int i = tid; i < nv; i += nt) cp_async_bytes(dst + i * V, src + i * V, 16);
    for (int i = nv * V + tid; i < n; i += nt) cp_async_bytes(dst + i, src + i, sizeof(R));
  } else {
    for (int i = tid; i < n; i += nt) cp_async_bytes(dst + i, src + i, sizeof(R));
  }
}
__host__ __device__ inline int up4(int n) { return (n + 3) & ~3; }

// shared-memory carve-up (elements): delta [NS][L1] | conv scratch [F][NCV] | pooled [NLIN] | xloc [nin_l] | freev | conv w+b
// Everything here is latency-bound (36 CTAs for the icml batch), so the loops that read weights from global memory
// keep several independent loads in flight: the hidden layer accumulates up to ENC_OPW outputs per warp at once, the
// heads use one warp per row.
// ---------------------------------------------------------------------------------------------------------------
// The hidden layer of the encoder at LARGE batches is three GEMMs (B individuals x NLIN pooled features x H units):
//   forward   pre[b][o]   = sum_i pooled[b][i] W[o][i]          (M, N, K) = (B, H, NLIN)
//   backward  dpool[b][i] = sum_o d_pre[b][o] W[o][i]           (B, NLIN, H)
//   weights   dW[o][i]   += sum_b d_pre[b][o] pooled[b][i]      (H, NLIN, B)
// Done per individual inside the monolithic kernels, every CTA streams the whole weight matrix from L2 (1 MB at T = 500)
// and the weight gradient is a B-deep sum per thread: 156 + 218 + 99 us at B = 1,024.  One shared-memory tiled SIMT GEMM
// with general strides serves all three: 64 x 64 tile per CTA, 16-deep k chunks, 4 x 4 outputs per thread, split-K over
// blockIdx.z with atomicAdd (C must be zeroed, or hold the value to accumulate into).  fp32 / fp64 FMAs: the problem is
// 250 MFLOP, a latency / bandwidth matter, not one for the tensor cores.
// ---------------------------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(256) enc_gemm_kernel(int M, int N, int K, const R* __restrict__ A, long long sam, long long sak,
                                                       const R* __restrict__ Bm, long long sbk, long long sbn, R* __restrict__ C,
                                                       long long ldc, int k_per_split) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ R As[TK][TM + 4];
  __shared__ R Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int k_lo = blockIdx.z * k_per_split, k_hi = min(K, k_lo + k_per_split);
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each a 4 x 4 block: rows ty*4.., columns tx*4..
  R acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = R(0);
  for (int k0 = k_lo; k0 < k_hi; k0 += TK) {
    // tile loads: consecutive threads along whichever index is contiguous in memory
#pragma unroll
    for (int e = tid; e < TM * TK; e += 256) {
      int mm, kk;
      if (sak == 1) { kk = e % TK; mm = e / TK; } else { mm = e % TM; kk = e / TM; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < k_hi) ? A[(long long)m * sam + (long long)k * sak] : R(0);
    }
#pragma unroll
    for (int e = tid; e < TN * TK; e += 256) {
      int nn, kk;
      if (sbk == 1) { kk = e % TK; nn = e / TK; } else { nn = e % TN; kk = e / TN; }
      const int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < N && k < k_hi) ? Bm[(long long)k * sbk + (long long)n * sbn] : R(0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      R a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) atomicAdd(C + (long long)m * ldc + n, acc[i][j]);
    }
  }
}

// C += A B with split-K sized so that the grid covers the machine about twice
template <typename R>
static void enc_gemm(int M, int N, int K, const R* A, long long sam, long long sak, const R* Bm, long long sbk, long long sbn, R* C,
                     long long ldc, cudaStream_t s) {
  const int gm = (M + 63) / 64, gn = (N + 63) / 64;
  int splits = (2 * 148 + gm * gn - 1) / (gm * gn);
  const int max_splits = (K + 63) / 64;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int k_per_split = (((K + splits - 1) / splits) + 15) / 16 * 16;
  dim3 grid(gn, gm, (K + k_per_split - 1) / k_per_split);
  enc_gemm_kernel<R><<<grid, 256, 0, s>>>(M, N, K, A, sam, sak, Bm, sbk, sbn, C, ldc, k_per_split);
}

#define ENC_THREADS 512
#define ENC_OPW 4
// Hidden layer of the forward kernel: enc[g][o] = tanh(b[o] + sum_i W[o][i] pooled[g][i]).  Each warp owns outputs o0,
// o0 + nw, ... and accumulates ENC_OPW of them for all G individuals per pass over the inputs; a lane takes 16 bytes of
// consecutive inputs at a time (one 16-byte load per operand and four (two) multiply-adds -- as scalar loads the loop
// spent ~12 instructions per multiply-add).  Rows that are not 16-byte aligned take the scalar loop.
template <typename R, int G>
__device__ __forceinline__ void enc_hidden_layer(const EncDims& d, const R* __restrict__ lin_w, const R* __restrict__ lin_b,
                                                 const R* __restrict__ pooled, R* __restrict__ xloc, R* __restrict__ enc_out,
                                                 int b0, int ngr, int warp, int lane, int nw) {
  constexpr int V = 16 / sizeof(R);
  struct alignas(16) Vec { R v[V]; };
  const bool vec = d.NLIN % V == 0 && ((size_t)lin_w & 15) == 0 && ((size_t)pooled & 15) == 0;
  for (int o0 = warp; o0 < d.H; o0 += nw * ENC_OPW) {
    R acc[ENC_OPW][G];
#pragma unroll
    for (int q = 0; q < ENC_OPW; ++q)
#pragma unroll
      for (int g = 0; g < G; ++g) acc[q][g] = R(0);
    if (vec) {
      const int nv = d.NLIN / V;
#pragma unroll 2
      for (int i = lane; i < nv; i += 32) {
        Vec x[G];
#pragma unroll
        for (int g = 0; g < G; ++g) x[g] = reinterpret_cast<const Vec*>(pooled + (size_t)g * d.NLIN)[i];
#pragma unroll
        for (int q = 0; q < ENC_OPW; ++q) {
          const int o = min(o0 + q * nw, d.H - 1);  // surplus slots redo the last row (discarded below): no branch here
          const Vec w = reinterpret_cast<const Vec*>(lin_w + (size_t)o * d.NLIN)[i];
#pragma unroll
          for (int g = 0; g < G; ++g)
#pragma unroll
            for (int e = 0; e < V; ++e) acc[q][g] += w.v[e] * x[g].v[e];
        }
      }
    } else {
#pragma unroll 2
      for (int i = lane; i < d.NLIN; i += 32) {
        R x[G];
#pragma unroll
        for (int g = 0; g < G; ++g) x[g] = pooled[(size_t)g * d.NLIN + i];
#pragma unroll
        for (int q = 0; q < ENC_OPW; ++q) {
          const int o = o0 + q * nw;
          if (o < d.H) {
            const R w = lin_w[(size_t)o * d.NLIN + i];
#pragma unroll
            for (int g = 0; g < G; ++g) acc[q][g] += w * x[g];
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < ENC_OPW; ++q) {
      const int o = o0 + q * nw;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const R a = warp_sum(acc[q][g]);
        if (lane == 0 && o < d.H && g < ngr) {
          const R e = vtanh(a + lin_b[o]);
          xloc[g * d.nin_l + o] = e;
          enc_out[(size_t)(b0 + g) * d.H + o] = e;
        }
      }
    }
  }
}
// G individuals per CTA: the hidden-layer weight matrix (H x NLIN, 1 MB at T = 500) is the only large operand and
// every CTA streams all of it from L2; with G > 1 each weight is loaded once and used G times (large batches).
template <typename R, int G>
__global__ void __launch_bounds__(ENC_THREADS) enc_fwd_kernel(const EncDims d, const EncPtrs<R> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* delta = reinterpret_cast<R*>(smem_raw);
  R* conv = delta + d.NS * d.L1;
  R* pooled = delta + up4(d.NS * d.L1 + d.F * d.NCV);  // [G][NLIN], 16-byte aligned (vector loads in the hidden layer)
  R* xloc = pooled + (size_t)G * d.NLIN;          // [G][nin_l]
  R* freev = xloc + G * d.nin_l;                  // [G][2*(nl+ng)]
  R* cw = freev + G * 2 * (d.nl + d.ng);          // conv weights [F][NS][K] + bias [F]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const int b0 = blockIdx.x * G;
  const int ngr = min(G, d.B - b0);
  pdl_trigger();  // the ODE forward launch may become resident; it waits for this grid before it reads the q tables
  const int ncw = d.F * d.NS * d.K;
  const int nfree = 2 * (d.nl + d.ng);
  // staged copies (d.stage): hidden-layer weights + bias, head weights + biases
  const R *lin_w = p.lin_w, *lin_b = p.lin_b, *local_w = p.local_w, *local_b = p.local_b, *gcond_w = p.gcond_w;
  if (d.stage) {
    R* st = delta + up4((int)(cw - delta) + ncw + d.F);  // 16-byte aligned
    R* s_lin_w = st;
    R* s_lin_b = s_lin_w + up4(d.H * d.NLIN);
    R* s_local_w = s_lin_b + up4(d.H);
    R* s_local_b = s_local_w + up4(2 * d.nl * d.nin_l);
    R* s_gcond_w = s_local_b + up4(2 * d.nl);
    stage_async(s_local_w, p.local_w, 2 * d.nl * d.nin_l, tid, nt);
    stage_async(s_local_b, p.local_b, 2 * d.nl, tid, nt);
    stage_async(s_gcond_w, p.gcond_w, 2 * d.ng * d.nin_g, tid, nt);
    stage_async(s_lin_b, p.lin_b, d.H, tid, nt);
    stage_async(s_lin_w, p.lin_w, d.H * d.NLIN, tid, nt);
    lin_w = s_lin_w; lin_b = s_lin_b; local_w = s_local_w; local_b = s_local_b; gcond_w = s_gcond_w;
  }
  for (int i = tid; i < ncw + d.F; i += nt) cw[i] = i < ncw ? p.conv_w[i] : p.conv_b[i - ncw];
  const R inv_pool = R(1) / R(d.PL);
  if (d.phase == 2) {
    // after the GEMM: p.enc holds the hidden layer's pre-activations (without bias); tanh in place, then the heads
    for (int e = tid; e < ngr * d.nin_l; e += nt) {
      const int g = e / d.nin_l, i = e % d.nin_l, b = b0 + g;
      R v;
      if (i < d.H) {
        v = vtanh(p.enc[(size_t)b * d.H + i] + lin_b[i]);
        p.enc[(size_t)b * d.H + i] = v;
      } else {
        const int j = i - d.H;
        v = (d.lt && j < d.C) ? p.inputs[(size_t)b * d.C + j] : p.dev[(size_t)b * d.D + (j - (d.lt ? d.C : 0))];
      }
      xloc[e] = v;
    }
  }
  for (int g = 0; g < ngr && d.phase != 2; ++g) {
    const int b = b0 + g;
    const R* obs = p.obs + (size_t)b * d.NS * d.T;
    __syncthreads();  // delta / conv scratch of the previous individual fully consumed
    for (int i = tid; i < d.NS * d.L1; i += nt) {
      const int c = i / d.L1, j = i % d.L1;
      delta[i] = obs[c * d.T + j + 1] - obs[c * d.T + j];
    }
    for (int i = tid; i < d.nin_l - d.H; i += nt)  // conditioning inputs of the local heads
      xloc[g * d.nin_l + d.H + i] =
          (d.lt && i < d.C) ? p.inputs[(size_t)b * d.C + i] : p.dev[(size_t)b * d.D + (i - (d.lt ? d.C : 0))];
    __syncthreads();
    // conv: a thread per (position j, pair of filters) -- every input sample is loaded once for two multiply-adds and
    // the filter taps are warp-wide broadcasts
    const int nfp = (d.F + 1) >> 1;
    for (int i = tid; i < nfp * d.NCV; i += nt) {
      const int fp = i / d.NCV, j = i - fp * d.NCV;
      const int f0 = 2 * fp, f1 = min(f0 + 1, d.F - 1);
      R a0 = cw[ncw + f0], a1 = cw[ncw + f1];
      for (int c = 0; c < d.NS; ++c) {
        const R* w0 = cw + (f0 * d.NS + c) * d.K;
        const R* w1 = cw + (f1 * d.NS + c) * d.K;
        const R* x = delta + c * d.L1 + j;
#pragma unroll 5
        for (int k = 0; k < d.K; ++k) {
          const R xv = x[k];
          a0 += w0[k] * xv;
          a1 += w1[k] * xv;
        }
      }
      conv[f0 * d.NCV + j] = a0;
      if (f1 != f0) conv[f1 * d.NCV + j] = a1;
    }
    __syncthreads();
    for (int i = tid; i < d.NLIN; i += nt) {
      const int f = i / d.NP, j = i % d.NP;
      R a = R(0);
      for (int k = 0; k < d.PL; ++k) a += conv[f * d.NCV + j + k];
      a *= inv_pool;
      pooled[(size_t)g * d.NLIN + i] = a;
      p.pooled[(size_t)b * d.NLIN + i] = a;
    }
  }
  if (d.stage) cp_async_commit_wait_all();
  __syncthreads();
  if (d.phase == 1) return;  // pooled features are in global memory: the hidden layer runs as a GEMM
  if (d.phase == 0) {
    if (d.stage) {  // two call sites: the staged copy is read with shared-memory loads, not generic ones
      R* st = delta + up4((int)(cw - delta) + ncw + d.F);
      enc_hidden_layer<R, G>(d, st, st + up4(d.H * d.NLIN), pooled, xloc, p.enc, b0, ngr, warp, lane, nw);
    } else {
      enc_hidden_layer<R, G>(d, p.lin_w, p.lin_b, pooled, xloc, p.enc, b0, ngr, warp, lane, nw);
    }
  }
  __syncthreads();
  // packed heads: one warp per (individual, row), lanes over the inputs
  for (int e = warp; e < ngr * nfree; e += nw) {
    const int g = e / nfree, r = e % nfree, b = b0 + g;
    R a = R(0);
    if (r < 2 * d.nl) {
      const R* w = local_w + (size_t)r * d.nin_l;
      for (int i = lane; i < d.nin_l; i += 32) a += w[i] * xloc[g * d.nin_l + i];
    } else {
      const R* w = gcond_w + (size_t)(r - 2 * d.nl) * d.nin_g;
      for (int i = lane; i < d.nin_g; i += 32) {
        const R x = (d.gt && i < d.C) ? p.inputs[(size_t)b * d.C + i] : p.dev[(size_t)b * d.D + (i - (d.gt ? d.C : 0))];
        a += w[i] * x;
      }
    }
    a = warp_sum(a);
    if (lane == 0) freev[g * nfree + r] = a + (r < 2 * d.nl ? local_b[r] : R(0));
  }
  __syncthreads();
  for (int e = tid; e < ngr * d.P; e += nt) {
    const int g = e / d.P, k = e % d.P;
    R* qm = p.q_mu + (size_t)(b0 + g) * d.P;
    R* qp = p.q_prec + (size_t)(b0 + g) * d.P;
    const int ncond = d.nl + d.ng;
    if (k < ncond) {
      qm[k] = freev[g * nfree + 2 * k];
      qp[k] = vexp(freev[g * nfree + 2 * k + 1]);
    } else if (k < ncond + d.nglob) {
      qm[k] = p.global_free[2 * (k - ncond)];
      qp[k] = vexp(p.global_free[2 * (k - ncond) + 1]);
    } else {
      qm[k] = p.const_values[k - ncond - d.nglob];
      qp[k] = R(1);
    }
  }
}

// backward, G individuals per CTA.  shared: delta | dconv [F][NCV] | dpool [G][NLIN] | xloc [G][nin_l] | dfree [G][..] | dpre [G][H]
template <typename R, int G>
__global__ void __launch_bounds__(ENC_THREADS) enc_bwd_kernel(const EncDims d, const EncPtrs<R> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* delta = reinterpret_cast<R*>(smem_raw);
  R* dconv = delta + d.NS * d.L1;
  R* dpool = dconv + d.F * d.NCV;
  R* xloc = dpool + (size_t)G * d.NLIN;
  const int ncond = d.nl + d.ng;
  const int nfr = 2 * (ncond + d.nglob);
  R* dfree = xloc + G * d.nin_l;
  R* dpre = dfree + G * nfr;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const int b0 = blockIdx.x * G;
  const int ngr = min(G, d.B - b0);
  // Launched with programmatic stream serialization (vh_pdl.cuh): the weight staging and the reads of this step's encoder
  // forward outputs below run under the tail of the preceding launch (the reverse sweep that produces d_q_mu / d_q_prec);
  // pdl_wait() stands in front of the first read of those.  The launch after this one (hidden-layer weight gradient +
  // Adam) may become resident right away: it prefetches the optimiser state and waits for this grid.
  pdl_trigger();
  // this CTA's filters: f = fy, fy + nfy, ...;  local filter index lf = f / nfy
  const int fy = blockIdx.y, nfy = gridDim.y;
  const int nfl = (d.F - fy + nfy - 1) / nfy;  // number of filters handled here
  const int ncol = nfl * d.NP;                 // pooled columns handled here
  // staged copies (d.stage, G == 1): this CTA's columns of the hidden-layer weights [H][ncol] and the local head weights
  const R* local_w = p.local_w;
  R* wl = delta + up4((int)(dpre - delta) + G * d.H);
  if (d.stage) {
    R* s_local_w = wl + up4(d.H * ncol);
    // per (output o, local filter lf) one contiguous run of NP weights: a warp per run, 16-byte copies when the runs are
    // 16-byte aligned on both sides (as element-wise 4-byte copies with two divisions each this loop was 39 % of the
    // kernel's instructions at the icml size)
    {
      constexpr int V = 16 / sizeof(R);
      const bool v16 = d.NP % V == 0 && d.NLIN % V == 0 && ((size_t)p.lin_w & 15) == 0;
      const int cpr = v16 ? d.NP / V : d.NP, step = v16 ? V : 1, bytes = v16 ? 16 : (int)sizeof(R);
      for (int run = warp; run < d.H * nfl; run += nw) {
        const int o = run / nfl, lf = run - o * nfl;
        const R* src = p.lin_w + (size_t)o * d.NLIN + (fy + lf * nfy) * d.NP;
        R* dst = wl + o * ncol + lf * d.NP;
        for (int c = lane; c < cpr; c += 32) cp_async_bytes(dst + c * step, src + c * step, bytes);
      }
    }
    stage_async(s_local_w, p.local_w, 2 * d.nl * d.nin_l, tid, nt);
    local_w = s_local_w;
    const R* obs = p.obs + (size_t)b0 * d.NS * d.T;  // the only individual of this CTA
    for (int i = tid; i < d.NS * d.L1; i += nt) {
      const int c = i / d.L1, j = i % d.L1;
      delta[i] = obs[c * d.T + j + 1] - obs[c * d.T + j];
    }
  }
  for (int e = tid; e < ngr * d.nin_l && d.phase != 2; e += nt) {
    const int g = e / d.nin_l, i = e % d.nin_l, b = b0 + g;
    R v;
    if (i < d.H)
      v = p.enc[(size_t)b * d.H + i];
    else {
      const int j = i - d.H;
      v = (d.lt && j < d.C) ? p.inputs[(size_t)b * d.C + j] : p.dev[(size_t)b * d.D + (j - (d.lt ? d.C : 0))];
    }
    xloc[e] = v;
  }
  pdl_wait();
  for (int e = tid; e < ngr * (ncond + d.nglob) && d.phase != 2; e += nt) {
    const int g = e / (ncond + d.nglob), k = e % (ncond + d.nglob), b = b0 + g;
    dfree[g * nfr + 2 * k] = p.d_q_mu[(size_t)b * d.P + k];
    dfree[g * nfr + 2 * k + 1] = p.d_q_prec[(size_t)b * d.P + k] * p.q_prec[(size_t)b * d.P + k];  // d exp(log_prec)
  }
  if (d.stage) cp_async_commit_wait_all();
  __syncthreads();
  // Small batches: gridDim.y CTAs share one individual, each taking the conv filters f = blockIdx.y (mod gridDim.y) and
  // their pooled columns -- the heavy phases below (dpool, dconv, conv weight gradients) need no exchange between
  // them; the cheap head part above is recomputed by each and only CTA y = 0 publishes its gradients.
  const bool lead = blockIdx.y == 0 && d.phase != 2;  // phase 2 (after the GEMM) has no head part: phase 1 did it
  // parameter gradients of the heads and the global free parameters: summed over the CTA's individuals, then one atomic
  for (int j = tid; lead && j < 2 * d.nglob; j += nt) {
    R a = R(0);
    for (int g = 0; g < ngr; ++g) a += dfree[g * nfr + 2 * ncond + j];
    atomicAdd(p.g_global_free + j, a);
  }
  for (int e = tid; lead && e < 2 * d.nl * d.nin_l; e += nt) {
    R a = R(0);
    for (int g = 0; g < ngr; ++g) a += dfree[g * nfr + e / d.nin_l] * xloc[g * d.nin_l + e % d.nin_l];
    atomicAdd(p.g_local_w + e, a);
  }
  for (int r = tid; lead && r < 2 * d.nl; r += nt) {
    R a = R(0);
    for (int g = 0; g < ngr; ++g) a += dfree[g * nfr + r];
    atomicAdd(p.g_local_b + r, a);
  }
  for (int e = tid; lead && e < 2 * d.ng * d.nin_g; e += nt) {
    const int r = e / d.nin_g, i = e % d.nin_g;
    R a = R(0);
    for (int g = 0; g < ngr; ++g) {
      const int b = b0 + g;
      const R x = (d.gt && i < d.C) ? p.inputs[(size_t)b * d.C + i] : p.dev[(size_t)b * d.D + (i - (d.gt ? d.C : 0))];
      a += dfree[g * nfr + 2 * d.nl + r] * x;
    }
    atomicAdd(p.g_gcond_w + e, a);
  }
  for (int e = tid; e < G * d.H && d.phase != 2; e += nt) {
    const int g = e / d.H, o = e % d.H;
    R gp = R(0);
    if (g < ngr) {
      R gg = R(0);
      for (int r = 0; r < 2 * d.nl; ++r) gg += local_w[(size_t)r * d.nin_l + o] * dfree[g * nfr + r];
      const R en = xloc[g * d.nin_l + o];
      gp = gg * (R(1) - en * en);  // tanh'
      if (lead) p.d_pre[(size_t)(b0 + g) * d.H + o] = gp;
    }
    dpre[e] = gp;
  }
  __syncthreads();
  for (int o = tid; lead && o < d.H; o += nt) {
    R a = R(0);
    for (int g = 0; g < ngr; ++g) a += dpre[g * d.H + o];
    atomicAdd(p.g_lin_b + o, a);
  }
  if (d.phase == 1) return;  // d_pre is in global memory: dpool and the weight gradient are GEMMs
  if (d.phase == 2) {        // ... and here dpool comes back
    for (int e = tid; e < ngr * d.NLIN; e += nt) dpool[e] = p.dpool_g[(size_t)b0 * d.NLIN + e];
  }
  // cotangent of the pooled features: dpool[g][i] = sum_o W[o][i] dpre[g][o]   (each weight loaded once for G individuals)
  for (int li = tid; li < nfl * d.NP && d.phase == 0; li += nt) {
    const int i = (fy + (li / d.NP) * nfy) * d.NP + li % d.NP;
    R acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = R(0);
#pragma unroll 8
    for (int o = 0; o < d.H; ++o) {
      const R w = d.stage ? wl[o * ncol + li] : p.lin_w[(size_t)o * d.NLIN + i];
#pragma unroll
      for (int g = 0; g < G; ++g) acc[g] += w * dpre[g * d.H + o];
    }
#pragma unroll
    for (int g = 0; g < G; ++g) dpool[(size_t)g * d.NLIN + i] = acc[g];
  }
  const R inv_pool = R(1) / R(d.PL);
  for (int g = 0; g < ngr; ++g) {
    const R* obs = p.obs + (size_t)(b0 + g) * d.NS * d.T;
    __syncthreads();  // dpool complete / previous individual's delta + dconv consumed
    for (int i = tid; !d.stage && i < d.NS * d.L1; i += nt) {
      const int c = i / d.L1, j = i % d.L1;
      delta[i] = obs[c * d.T + j + 1] - obs[c * d.T + j];
    }
    for (int li = tid; li < nfl * d.NCV; li += nt) {
      const int f = fy + (li / d.NCV) * nfy, j = li % d.NCV;
      R a = R(0);
      for (int k = 0; k < d.PL; ++k) {
        const int jp = j - k;
        if (jp >= 0 && jp < d.NP) a += dpool[(size_t)g * d.NLIN + f * d.NP + jp];
      }
      dconv[f * d.NCV + j] = a * inv_pool;
    }
    __syncthreads();
    // conv weight / bias gradients of this CTA's filters: four threads per weight, each a quarter of the NCV positions
    // (a warp per weight spent 128 instructions per weight, most of them on the 5-level reduction of a 76-term sum)
    const int wpf = d.NS * d.K + 1;  // weights + bias per filter
    for (int le0 = 0; le0 < nfl * wpf; le0 += nt >> 2) {  // trip count uniform over the CTA (shuffles below)
      const int le = le0 + (tid >> 2), q = tid & 3;
      const bool on = le < nfl * wpf;
      const int lf = on ? le / wpf : 0, r = on ? le - lf * wpf : 0, f = fy + lf * nfy;
      const bool bias = r >= d.NS * d.K;
      const int c = bias ? 0 : r / d.K, k = bias ? 0 : r - c * d.K;
      const R* dc = dconv + f * d.NCV;
      const R* dl = delta + c * d.L1 + k;
      R a = R(0);
      if (on) {
        if (bias) {
          for (int j = q; j < d.NCV; j += 4) a += dc[j];
        } else {
#pragma unroll 4
          for (int j = q; j < d.NCV; j += 4) a += dc[j] * dl[j];
        }
      }
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      if (on && q == 0) atomicAdd(bias ? p.g_conv_b + f : p.g_conv_w + f * d.NS * d.K + r, a);
    }
  }
}

// dW_lin[o][i] += sum_b d_pre[b][o] * pooled[b][i].  One thread per input column i (coalesced reads of pooled), all H
// outputs accumulated in registers, d_pre rows broadcast from shared memory; the individuals are split over
// blockIdx.y so that large batches fill the machine (partial sums meet in the atomics).  HMAX bounds the register
// tile; wider hidden layers take several passes.
#define ENC_WG_HMAX 64
#define ENC_WG_BCHUNK 32
template <typename R>
__global__ void __launch_bounds__(128) enc_lin_wgrad_kernel(const EncDims d, const EncPtrs<R> p, int b_per_block) {
  __shared__ __align__(16) R sh[ENC_WG_BCHUNK * ENC_WG_HMAX];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b0 = blockIdx.y * b_per_block, b1 = min(d.B, b0 + b_per_block);
  for (int o0 = 0; o0 < d.H; o0 += ENC_WG_HMAX) {
    const int nh = min(ENC_WG_HMAX, d.H - o0);
    R acc[ENC_WG_HMAX];
#pragma unroll
    for (int o = 0; o < ENC_WG_HMAX; ++o) acc[o] = R(0);
    for (int bb = b0; bb < b1; bb += ENC_WG_BCHUNK) {
      const int nb = min(ENC_WG_BCHUNK, b1 - bb);
      __syncthreads();
      for (int e = threadIdx.x; e < nb * ENC_WG_HMAX; e += blockDim.x) {
        const int r = e / ENC_WG_HMAX, o = e % ENC_WG_HMAX;
        sh[e] = o < nh ? p.d_pre[(size_t)(bb + r) * d.H + o0 + o] : R(0);
      }
      __syncthreads();
      if (i < d.NLIN) {
        // the d_pre row is a broadcast operand: read it 16 bytes at a time (one shared-memory load per 4 (2) FMAs
        // instead of one per FMA -- the loop was bound by the load/store unit, not by the FMA pipe)
        constexpr int V = 16 / sizeof(R);
        struct alignas(16) Vec { R v[V]; };
        for (int r = 0; r < nb; ++r) {
          const R x = p.pooled[(size_t)(bb + r) * d.NLIN + i];
          const Vec* row = reinterpret_cast<const Vec*>(sh + r * ENC_WG_HMAX);
#pragma unroll
          for (int o = 0; o < ENC_WG_HMAX / V; ++o) {
            const Vec q = row[o];
#pragma unroll
            for (int e = 0; e < V; ++e) acc[o * V + e] += q.v[e] * x;
          }
        }
      }
    }
    if (i < d.NLIN) {
#pragma unroll
      for (int o = 0; o < ENC_WG_HMAX; ++o)
        if (o < nh) atomicAdd(p.g_lin_w + (size_t)(o0 + o) * d.NLIN + i, acc[o]);
    }
  }
}

// small batches: one thread per weight (H * NLIN threads fill the machine; the B-deep sum is short)
template <typename R>
__global__ void __launch_bounds__(256) enc_lin_wgrad_small_kernel(const EncDims d, const EncPtrs<R> p) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.H * d.NLIN) return;
  const int o = e / d.NLIN, i = e % d.NLIN;
  R a0 = R(0), a1 = R(0);
  int b = 0;
#pragma unroll 6  // 12 independent load pairs in flight: the B-deep sum is latency-, not throughput-bound
  for (; b + 1 < d.B; b += 2) {
    a0 += p.d_pre[(size_t)b * d.H + o] * p.pooled[(size_t)b * d.NLIN + i];
    a1 += p.d_pre[(size_t)(b + 1) * d.H + o] * p.pooled[(size_t)(b + 1) * d.NLIN + i];
  }
  if (b < d.B) a0 += p.d_pre[(size_t)b * d.H + o] * p.pooled[(size_t)b * d.NLIN + i];
  p.g_lin_w[e] += a0 + a1;
}

// Small batches, single GPU: the hidden-layer weight gradient AND the Adam update of the whole flat parameter vector in one
// launch (vh_encoder_bwd_adam).  One thread per parameter of the flat vector; threads inside the view of lin_w first form
// their gradient entry (the B-deep sum above) -- it never travels through memory -- everybody then applies Adam exactly
// as adam_dev_kernel does (vh_api.cu: device-side step counter and hyper-parameters, NaN guard, gradient cleared).
template <typename R>
__global__ void __launch_bounds__(256) enc_lin_wgrad_adam_kernel(const EncDims d, const EncPtrs<R> p, size_t n, long long lin_off,
                                                                 R* __restrict__ prm, R* __restrict__ g, R* __restrict__ m,
                                                                 R* __restrict__ v, const double* __restrict__ hyper,
                                                                 long long* step, const R* __restrict__ guard) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ R s_bc[2];
  __shared__ int s_skip;
  if (threadIdx.x == 0) {
    const double t = (double)(*(volatile long long*)step + 1);
    s_bc[0] = (R)(1.0 - pow(hyper[1], t));
    s_bc[1] = (R)sqrt(1.0 - pow(hyper[2], t));
    const R c = guard ? *guard : R(0);
    s_skip = (c != c) || (guard && *(volatile long long*)(step + 2) != 0);
  }
  __syncthreads();
  const bool skip = s_skip != 0;
  if (i < n) {
    R gi = g[i];
    const long long e = (long long)i - lin_off;
    if (!skip && e >= 0 && e < (long long)d.H * d.NLIN) {
      const int o = (int)(e / d.NLIN), c = (int)(e % d.NLIN);
      R a0 = R(0), a1 = R(0);
      int b = 0;
#pragma unroll 6
      for (; b + 1 < d.B; b += 2) {
        a0 += p.d_pre[(size_t)b * d.H + o] * p.pooled[(size_t)b * d.NLIN + c];
        a1 += p.d_pre[(size_t)(b + 1) * d.H + o] * p.pooled[(size_t)(b + 1) * d.NLIN + c];
      }
      if (b < d.B) a0 += p.d_pre[(size_t)b * d.H + o] * p.pooled[(size_t)b * d.NLIN + c];
      gi += a0 + a1;
    }
    if (!skip) {
      const double lr = hyper[0], b1d = hyper[1], b2d = hyper[2];
      const R b1 = (R)b1d, b2 = (R)b2d, eps = (R)hyper[3];
      const R mi = m[i] + (gi - m[i]) * (R(1) - b1);
      const R vi = b2 * v[i] + (R(1) - b2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      const R denom = vsqrt(vi) / s_bc[1] + eps;
      prm[i] -= ((R)lr / s_bc[0]) * (mi / denom);
    }
    g[i] = R(0);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd((unsigned long long*)(step + 1), 1ULL);
    if (ticket == (unsigned long long)gridDim.x - 1) {
      step[1] = 0;
      step[skip ? 2 : 0] += 1;
    }
  }
}

// The same launch with 16 bytes of consecutive parameters per thread (4 fp32 / 2 fp64): used when the view of lin_w starts
// on a 16-byte boundary of the flat vector and NLIN is a multiple of the vector width, so that a thread's parameters
// share their row o and their pooled columns are one 16-byte load.  The optimiser state (parameter, gradient, both
// moments) is fetched BEFORE the B-deep sum: its (cold, DRAM) round trip overlaps the sum's L2 round trips instead of
// following them.
template <typename R>
__global__ void __launch_bounds__(64) enc_lin_wgrad_adam_vec_kernel(const EncDims d, const EncPtrs<R> p, size_t n, long long lin_off,
                                                                    R* __restrict__ prm, R* __restrict__ g, R* __restrict__ m,
                                                                    R* __restrict__ v, const double* __restrict__ hyper,
                                                                    long long* step, const R* __restrict__ guard) {
  constexpr int V = 16 / sizeof(R);
  struct alignas(16) Vec { R v[V]; };
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * V;  // n is a multiple of V (checked by the launcher)
  __shared__ R s_bc[2];
  __shared__ int s_skip;
  const bool on = i < n;
  Vec pv = {}, gv = {}, mv = {}, vv = {};
  if (on) {  // nobody writes these before this kernel does: fetched under the preceding launch's tail (vh_pdl.cuh)
    pv = *reinterpret_cast<const Vec*>(prm + i);
    mv = *reinterpret_cast<const Vec*>(m + i);
    vv = *reinterpret_cast<const Vec*>(v + i);
  }
  pdl_wait();  // the gradient vector, d_pre and the guard come from the launches before
  if (on) gv = *reinterpret_cast<const Vec*>(g + i);
  if (threadIdx.x == 0) {
    const double t = (double)(*(volatile long long*)step + 1);
    s_bc[0] = (R)(1.0 - pow(hyper[1], t));
    s_bc[1] = (R)sqrt(1.0 - pow(hyper[2], t));
    const R c = guard ? *guard : R(0);
    s_skip = (c != c) || (guard && *(volatile long long*)(step + 2) != 0);
  }
  __syncthreads();
  const bool skip = s_skip != 0;
  if (on) {
    const long long e = (long long)i - lin_off;
    if (!skip && e >= 0 && e < (long long)d.H * d.NLIN) {
      const int o = (int)(e / d.NLIN), c = (int)(e - (long long)o * d.NLIN);
      Vec a0 = {}, a1 = {};
      const R* dp = p.d_pre + o;
      const R* pl = p.pooled + c;
      int b = 0;
#pragma unroll 6
      for (; b + 1 < d.B; b += 2) {
        const R w0 = dp[(size_t)b * d.H], w1 = dp[(size_t)(b + 1) * d.H];
        const Vec x0 = *reinterpret_cast<const Vec*>(pl + (size_t)b * d.NLIN);
        const Vec x1 = *reinterpret_cast<const Vec*>(pl + (size_t)(b + 1) * d.NLIN);
#pragma unroll
        for (int k = 0; k < V; ++k) {
          a0.v[k] += w0 * x0.v[k];
          a1.v[k] += w1 * x1.v[k];
        }
      }
      if (b < d.B) {
        const R w0 = dp[(size_t)b * d.H];
        const Vec x0 = *reinterpret_cast<const Vec*>(pl + (size_t)b * d.NLIN);
#pragma unroll
        for (int k = 0; k < V; ++k) a0.v[k] += w0 * x0.v[k];
      }
#pragma unroll
      for (int k = 0; k < V; ++k) gv.v[k] += a0.v[k] + a1.v[k];
    }
    if (!skip) {
      const double lr = hyper[0], b1d = hyper[1], b2d = hyper[2];
      const R b1 = (R)b1d, b2 = (R)b2d, eps = (R)hyper[3];
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const R gi = gv.v[k];
        const R mi = mv.v[k] + (gi - mv.v[k]) * (R(1) - b1);
        const R vi = b2 * vv.v[k] + (R(1) - b2) * gi * gi;
        mv.v[k] = mi;
        vv.v[k] = vi;
        const R denom = vsqrt(vi) / s_bc[1] + eps;
        pv.v[k] -= ((R)lr / s_bc[0]) * (mi / denom);
      }
      *reinterpret_cast<Vec*>(m + i) = mv;
      *reinterpret_cast<Vec*>(v + i) = vv;
      *reinterpret_cast<Vec*>(prm + i) = pv;
    }
    *reinterpret_cast<Vec*>(g + i) = Vec{};
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd((unsigned long long*)(step + 1), 1ULL);
    if (ticket == (unsigned long long)gridDim.x - 1) {
      step[1] = 0;
      step[skip ? 2 : 0] += 1;
    }
  }
}

// device conditioner (vihds/ode.py:43-58, :99-116) including the reference's repeat/reshape quirk: sample n = b*IW + i
// of the GLOBAL batch receives the conditioner output of individual n % B_global.  A rank that holds the slab of
// individuals [b_offset, b_offset + B) passes the global one-hot table, so the result does not depend on the rank count.
template <typename R>
__global__ void conditioner_kernel(int B_global, long long n_offset, int N, int D, int n_cond, const R* __restrict__ dev,
                                   const R* __restrict__ rel, const R* __restrict__ w, const int* __restrict__ plus_one,
                                   R* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const R* d = dev + (size_t)((n_offset + n) % B_global) * D;
  for (int k = 0; k < n_cond; ++k) {
    R a = R(0);
    for (int j = 0; j < D; ++j) a += w[k * D + j] * (d[j] * rel[k * D + j]);
    a = a > R(0) ? a : R(0);
    out[(size_t)k * N + n] = plus_one[k] ? R(1) + a : a;
  }
}

static const char* fill_dims(const vh_encoder_desc* e, EncDims& d) {
  if (!e) return "null descriptor";
  d.B = e->B; d.T = e->T; d.NS = e->n_signals; d.F = e->n_filters; d.K = e->filter_size; d.PL = e->pool_size; d.H = e->n_hidden;
  d.C = e->C; d.D = e->D; d.nl = e->n_local; d.ng = e->n_gcond; d.nglob = e->n_global; d.nconst = e->n_const;
  d.P = d.nl + d.ng + d.nglob + d.nconst;
  d.lt = e->local_cond_treatments; d.ld = e->local_cond_devices; d.gt = e->gcond_cond_treatments; d.gd = e->gcond_cond_devices;
  d.L1 = d.T - 1;
  d.NCV = d.L1 - (d.K - 1);
  d.NP = d.NCV - (d.PL - 1);
  d.NLIN = d.F * d.NP;
  d.nin_l = d.H + (d.lt ? d.C : 0) + (d.ld ? d.D : 0);
  d.nin_g = (d.gt ? d.C : 0) + (d.gd ? d.D : 0);
  d.stage = 0;
  d.phase = 0;
  if (d.B <= 0 || d.NP <= 0 || d.H <= 0) return "encoder: B, n_hidden must be positive and T long enough for the conv + pool";
  if (d.ng > 0 && d.nin_g == 0) return "encoder: global-conditioned parameters need a conditioning input";
  return nullptr;
}

template <typename R>
static void fill_ptrs(const vh_encoder_io* io, const vh_encoder_grads* g, EncPtrs<R>& p) {
  p.obs = (const R*)io->observations; p.inputs = (const R*)io->inputs; p.dev = (const R*)io->dev_1hot;
  p.conv_w = (const R*)io->conv_w; p.conv_b = (const R*)io->conv_b; p.lin_w = (const R*)io->lin_w; p.lin_b = (const R*)io->lin_b;
  p.local_w = (const R*)io->local_w; p.local_b = (const R*)io->local_b; p.gcond_w = (const R*)io->gcond_w;
  p.global_free = (const R*)io->global_free; p.const_values = (const R*)io->const_values;
  p.q_mu = (R*)io->q_mu; p.q_prec = (R*)io->q_prec; p.pooled = (R*)io->pooled; p.enc = (R*)io->enc;
  if (g) {
    p.d_q_mu = (const R*)g->d_q_mu; p.d_q_prec = (const R*)g->d_q_prec;
    p.g_conv_w = (R*)g->g_conv_w; p.g_conv_b = (R*)g->g_conv_b; p.g_lin_w = (R*)g->g_lin_w; p.g_lin_b = (R*)g->g_lin_b;
    p.g_local_w = (R*)g->g_local_w; p.g_local_b = (R*)g->g_local_b; p.g_gcond_w = (R*)g->g_gcond_w;
    p.g_global_free = (R*)g->g_global_free; p.d_pre = (R*)g->d_pre;
    p.dpool_g = (R*)g->dpool;
  }
}

template <typename R, int G>
static void enc_fwd_g(const EncDims& d, const EncPtrs<R>& p, cudaStream_t s) {
  size_t smem = sizeof(R) * ((size_t)d.NS * d.L1 + (size_t)d.F * d.NCV + (size_t)G * d.NLIN + G * d.nin_l +
                             G * 2 * (d.nl + d.ng) + (size_t)up4(d.F * d.NS * d.K + d.F) + 8);
  // stage the weights in shared memory when they fit (fp32 at the icml size: 144 KB of hidden-layer weights)
  const size_t staged = sizeof(R) * ((size_t)up4(d.H * d.NLIN) + up4(d.H) + up4(2 * d.nl * d.nin_l) + up4(2 * d.nl) +
                                     up4(2 * d.ng * d.nin_g) + 8);
  EncDims dd = d;
  dd.stage = (G == 1 && smem + staged <= 220 * 1024) ? 1 : 0;
  if (dd.stage) smem += staged;
  if (smem > 48 * 1024) cudaFuncSetAttribute(enc_fwd_kernel<R, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  enc_fwd_kernel<R, G><<<(d.B + G - 1) / G, ENC_THREADS, smem, s>>>(dd, p);
}
template <typename R, int G>
static void enc_bwd_g(const EncDims& d, const EncPtrs<R>& p, cudaStream_t s) {
  size_t smem = sizeof(R) * ((size_t)d.NS * d.L1 + (size_t)d.F * d.NCV + (size_t)G * d.NLIN + G * d.nin_l +
                             G * 2 * (d.nl + d.ng + d.nglob) + G * d.H + 8);
  // small batches: split every individual over up to 5 CTAs by conv filter so that the latency-bound phases spread over
  // more SMs (36 individuals -> 180 CTAs)
  int fsplit = 1;
  if (G == 1) {
    fsplit = (2 * 148) / (d.B > 0 ? d.B : 1);
    if (fsplit > 5) fsplit = 5;
    if (fsplit > d.F) fsplit = d.F;
    if (fsplit < 1) fsplit = 1;
  }
  // stage this CTA's share of the weights in shared memory when it fits
  const int nfl_max = (d.F + fsplit - 1) / fsplit;
  const size_t staged = sizeof(R) * ((size_t)up4(d.H * nfl_max * d.NP) + up4(2 * d.nl * d.nin_l) + 8);
  EncDims dd = d;
  dd.stage = (G == 1 && smem + staged <= 200 * 1024) ? 1 : 0;
  if (dd.stage) smem += staged;
  if (smem > 48 * 1024) cudaFuncSetAttribute(enc_bwd_kernel<R, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((d.B + G - 1) / G, fsplit);
  launch_maybe_pdl(enc_bwd_kernel<R, G>, grid, dim3(ENC_THREADS), smem, s, true, dd, p);
}
// individuals per CTA: 1 while the batch does not fill the machine anyway, 4 for large batches (if it fits shared memory)
template <typename R>
static int enc_group(const EncDims& d) {
  const size_t smem4 = sizeof(R) * ((size_t)d.NS * d.L1 + (size_t)d.F * d.NCV + (size_t)4 * d.NLIN + 4 * (d.nin_l + 2 * d.P + d.H) + 1024);
  return (d.B >= 4 * 148 && smem4 <= 200 * 1024) ? 4 : 1;
}

// large batches: conv + pool | GEMM | tanh + heads
static bool enc_gemm_path(const EncDims& d) { return d.B >= 256; }

template <typename R>
static int enc_fwd_t(const EncDims& d, const vh_encoder_io* io, cudaStream_t s) {
  EncPtrs<R> p = {};
  fill_ptrs<R>(io, nullptr, p);
  if (enc_gemm_path(d)) {
    EncDims d1 = d, d2 = d;
    d1.phase = 1;
    d2.phase = 2;
    cudaMemsetAsync(p.enc, 0, sizeof(R) * (size_t)d.B * d.H, s);
    if (enc_group<R>(d) == 4) enc_fwd_g<R, 4>(d1, p, s); else enc_fwd_g<R, 1>(d1, p, s);
    enc_gemm<R>(d.B, d.H, d.NLIN, p.pooled, d.NLIN, 1, p.lin_w, 1, d.NLIN, p.enc, d.H, s);  // pre = pooled W^T
    if (enc_group<R>(d) == 4) enc_fwd_g<R, 4>(d2, p, s); else enc_fwd_g<R, 1>(d2, p, s);
    return 0;
  }
  if (enc_group<R>(d) == 4)
    enc_fwd_g<R, 4>(d, p, s);
  else
    enc_fwd_g<R, 1>(d, p, s);
  return 0;
}

struct AdamArgs {
  size_t n;
  void *param, *grad, *exp_avg, *exp_avg_sq;
  const void *hyper, *guard;
  void* step;
};

template <typename R>
static int enc_bwd_t(const EncDims& d, const vh_encoder_io* io, const vh_encoder_grads* g, cudaStream_t s, const AdamArgs* ad = nullptr) {
  EncPtrs<R> p = {};
  fill_ptrs<R>(io, g, p);
  if (enc_gemm_path(d) && p.dpool_g) {
    EncDims d1 = d, d2 = d;
    d1.phase = 1;
    d2.phase = 2;
    cudaMemsetAsync(p.dpool_g, 0, sizeof(R) * (size_t)d.B * d.NLIN, s);
    if (enc_group<R>(d) == 4) enc_bwd_g<R, 4>(d1, p, s); else enc_bwd_g<R, 1>(d1, p, s);
    enc_gemm<R>(d.B, d.NLIN, d.H, p.d_pre, d.H, 1, p.lin_w, d.NLIN, 1, p.dpool_g, d.NLIN, s);       // dpool = d_pre W
    enc_gemm<R>(d.H, d.NLIN, d.B, p.d_pre, 1, d.H, p.pooled, d.NLIN, 1, p.g_lin_w, d.NLIN, s);      // dW += d_pre^T pooled
    if (enc_group<R>(d) == 4) enc_bwd_g<R, 4>(d2, p, s); else enc_bwd_g<R, 1>(d2, p, s);
    return 0;
  }
  if (enc_group<R>(d) == 4)
    enc_bwd_g<R, 4>(d, p, s);
  else
    enc_bwd_g<R, 1>(d, p, s);
  if (g && g->skip_lin_wgrad) return 0;  // formed inside the exchange launch (vh_adam_allreduce_step_wgrad)
  if (ad) {  // B <= 128 (checked by the caller): weight gradient of the hidden layer + Adam over the flat vector, one launch
    const long long lin_off = (long long)(p.g_lin_w - (R*)ad->grad);
    constexpr int V = 16 / sizeof(R);
    const size_t al = (size_t)ad->param | (size_t)ad->grad | (size_t)ad->exp_avg | (size_t)ad->exp_avg_sq | (size_t)p.pooled;
    if (ad->n % V == 0 && lin_off % V == 0 && d.NLIN % V == 0 && (al & 15) == 0) {
      const size_t nthr = ad->n / V;
      launch_maybe_pdl(enc_lin_wgrad_adam_vec_kernel<R>, dim3((unsigned)((nthr + 63) / 64)), dim3(64), 0, s, true, d, p, ad->n,
                       lin_off, (R*)ad->param, (R*)ad->grad, (R*)ad->exp_avg, (R*)ad->exp_avg_sq, (const double*)ad->hyper,
                       (long long*)ad->step, (const R*)ad->guard);
      return 0;
    }
    enc_lin_wgrad_adam_kernel<R><<<(unsigned)((ad->n + 255) / 256), 256, 0, s>>>(
        d, p, ad->n, lin_off, (R*)ad->param, (R*)ad->grad, (R*)ad->exp_avg, (R*)ad->exp_avg_sq, (const double*)ad->hyper,
        (long long*)ad->step, (const R*)ad->guard);
    return 0;
  }
  if (d.B <= 128) {
    const int n = d.H * d.NLIN;
    enc_lin_wgrad_small_kernel<R><<<(n + 255) / 256, 256, 0, s>>>(d, p);
    return 0;
  }
  // enough CTAs for ~2 per SM: columns x splits of the individuals
  const int col_blocks = (d.NLIN + 127) / 128;
  int splits = (2 * 148 + col_blocks - 1) / col_blocks;
  const int max_splits = (d.B + ENC_WG_BCHUNK - 1) / ENC_WG_BCHUNK;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int b_per_block = (((d.B + splits - 1) / splits) + ENC_WG_BCHUNK - 1) / ENC_WG_BCHUNK * ENC_WG_BCHUNK;
  dim3 grid(col_blocks, (d.B + b_per_block - 1) / b_per_block);
  enc_lin_wgrad_kernel<R><<<grid, 128, 0, s>>>(d, p, b_per_block);
  return 0;
}

}  // namespace vh

using namespace vh;

extern "C" {

int vh_encoder_fwd(const vh_encoder_desc* e, const vh_encoder_io* io, void* stream) {
  EncDims d;
  if (const char* err = fill_dims(e, d)) {
    set_error("vh_encoder_fwd: %s", err);
    return VH_ERR_INVALID;
  }
  if (!io || !io->observations || !io->conv_w || !io->lin_w || !io->q_mu || !io->q_prec || !io->pooled || !io->enc ||
      (d.nl > 0 && (!io->local_w || !io->local_b)) || (d.ng > 0 && !io->gcond_w) || (d.nglob > 0 && !io->global_free) ||
      (d.nconst > 0 && !io->const_values)) {
    set_error("vh_encoder_fwd: missing buffer");
    return VH_ERR_INVALID;
  }
  if (e->dtype == VH_F32) enc_fwd_t<float>(d, io, (cudaStream_t)stream);
  else if (e->dtype == VH_F64) enc_fwd_t<double>(d, io, (cudaStream_t)stream);
  else {
    set_error("unknown dtype %d", e->dtype);
    return VH_ERR_INVALID;
  }
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    set_error("enc_fwd_kernel launch failed: %s", cudaGetErrorString(ce));
    return VH_ERR_CUDA;
  }
  return VH_OK;
}

int vh_device_conditioner(int dtype, int B, int IW, int D, int n_cond, int B_global, int b_offset, const void* dev_1hot,
                          const void* rel, const void* w, const int* plus_one, void* out, void* stream) {
  if (B <= 0 || IW <= 0 || D <= 0 || n_cond <= 0 || !dev_1hot || !rel || !w || !plus_one || !out || B_global < B ||
      b_offset < 0 || b_offset + B > B_global) {
    set_error("vh_device_conditioner: bad arguments");
    return VH_ERR_INVALID;
  }
  const long long n_off = (long long)b_offset * IW;
  const int N = B * IW, block = 128, grid = (N + block - 1) / block;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == VH_F32)
    conditioner_kernel<float><<<grid, block, 0, s>>>(B_global, n_off, N, D, n_cond, (const float*)dev_1hot, (const float*)rel, (const float*)w, plus_one, (float*)out);
  else if (dtype == VH_F64)
    conditioner_kernel<double><<<grid, block, 0, s>>>(B_global, n_off, N, D, n_cond, (const double*)dev_1hot, (const double*)rel, (const double*)w, plus_one, (double*)out);
  else {
    set_error("unknown dtype %d", dtype);
    return VH_ERR_INVALID;
  }
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    set_error("conditioner_kernel launch failed: %s", cudaGetErrorString(ce));
    return VH_ERR_CUDA;
  }
  return VH_OK;
}

int vh_encoder_bwd_adam(const vh_encoder_desc* e, const vh_encoder_io* io, const vh_encoder_grads* g, size_t n, void* param,
                        void* grad, void* exp_avg, void* exp_avg_sq, const void* hyper, void* step, const void* guard,
                        void* stream) {
  EncDims d;
  if (const char* err = fill_dims(e, d)) {
    set_error("vh_encoder_bwd_adam: %s", err);
    return VH_ERR_INVALID;
  }
  if (!io || !g || !g->d_q_mu || !g->d_q_prec || !g->g_conv_w || !g->g_conv_b || !g->g_lin_w || !g->g_lin_b || !g->d_pre ||
      !io->q_prec || !io->pooled || !io->enc || (d.nl > 0 && (!g->g_local_w || !g->g_local_b)) || (d.ng > 0 && !g->g_gcond_w) ||
      (d.nglob > 0 && !g->g_global_free) || !param || !grad || !exp_avg || !exp_avg_sq || !hyper || !step || n == 0) {
    set_error("vh_encoder_bwd_adam: missing buffer");
    return VH_ERR_INVALID;
  }
  const size_t es = e->dtype == VH_F64 ? 8 : 4;
  const char *gl = (const char*)g->g_lin_w, *g0 = (const char*)grad;
  if (d.B > 128 || gl < g0 || gl + es * (size_t)d.H * d.NLIN > g0 + es * n) {
    set_error("vh_encoder_bwd_adam: needs B <= 128 and g_lin_w inside the flat gradient vector (use vh_encoder_bwd + vh_adam_step_dev)");
    return VH_ERR_UNSUPPORTED;
  }
  AdamArgs ad = {n, param, grad, exp_avg, exp_avg_sq, hyper, guard, step};
  if (e->dtype == VH_F32) enc_bwd_t<float>(d, io, g, (cudaStream_t)stream, &ad);
  else if (e->dtype == VH_F64) enc_bwd_t<double>(d, io, g, (cudaStream_t)stream, &ad);
  else {
    set_error("unknown dtype %d", e->dtype);
    return VH_ERR_INVALID;
  }
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    set_error("enc_bwd / enc_lin_wgrad_adam launch failed: %s", cudaGetErrorString(ce));
    return VH_ERR_CUDA;
  }
  return VH_OK;
}

int vh_encoder_bwd(const vh_encoder_desc* e, const vh_encoder_io* io, const vh_encoder_grads* g, void* stream) {
  EncDims d;
  if (const char* err = fill_dims(e, d)) {
    set_error("vh_encoder_bwd: %s", err);
    return VH_ERR_INVALID;
  }
  if (!io || !g || !g->d_q_mu || !g->d_q_prec || !g->g_conv_w || !g->g_conv_b || !g->g_lin_w || !g->g_lin_b || !g->d_pre ||
      !io->q_prec || !io->pooled || !io->enc || (d.nl > 0 && (!g->g_local_w || !g->g_local_b)) || (d.ng > 0 && !g->g_gcond_w) ||
      (d.nglob > 0 && !g->g_global_free)) {
    set_error("vh_encoder_bwd: missing buffer");
    return VH_ERR_INVALID;
  }
  if (e->dtype == VH_F32) enc_bwd_t<float>(d, io, g, (cudaStream_t)stream);
  else if (e->dtype == VH_F64) enc_bwd_t<double>(d, io, g, (cudaStream_t)stream);
  else {
    set_error("unknown dtype %d", e->dtype);
    return VH_ERR_INVALID;
  }
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    set_error("enc_bwd_kernel launch failed: %s", cudaGetErrorString(ce));
    return VH_ERR_CUDA;
  }
  return VH_OK;
}

}  // extern "C"
