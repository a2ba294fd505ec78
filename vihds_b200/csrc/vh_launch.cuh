// __global__ wrappers around the per-trajectory code (vh_traj.cuh) and their launchers.
// Mapping: ONE THREAD PER TRAJECTORY (individual x importance sample).  The species state, the RHS constants and (in
// the reverse sweep) the adjoint state live in registers for all T steps; every global access is coalesced across
// the 32 trajectories of a warp because traces are laid out [T][S][N].  See DESIGN.md for why this mapping (and not
// a warp per trajectory) is the one that fills the machine: the white-box RHS is ~25 scalars of state with a long
// dependent chain, so the parallel axis is trajectories, not species.
#pragma once
#include <cuda_runtime.h>

#include "vh_dispatch.cuh"

namespace vh {

void set_error(const char* fmt, ...);

// segmented (by individual) warp reduction + one atomic per segment: folds per-trajectory (d mu, d prec) into [B][P]
template <typename R>
struct WarpSegRed {
  R* d_mu;
  R* d_prec;
  int P;
  unsigned same;  // bit o set: lane + (1<<o) exists and belongs to the same individual
  bool head;      // first lane of its segment
  __device__ WarpSegRed(R* dm, R* dp, int P_, int b, bool active) : d_mu(dm), d_prec(dp), P(P_) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int bl = active ? b : -1;
    same = 0;
#pragma unroll
    for (int o = 0; o < 5; ++o) {
      const int other = __shfl_down_sync(full, bl, 1 << o);
      if (lane + (1 << o) < 32 && other == bl) same |= 1u << o;
    }
    const int prev = __shfl_up_sync(full, bl, 1);
    head = active && (lane == 0 || prev != bl);
  }
  __device__ void operator()(int b, int k, R dmu, R dprec, bool) const {
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 0; o < 5; ++o) {
      const R m = __shfl_down_sync(full, dmu, 1 << o);
      const R p = __shfl_down_sync(full, dprec, 1 << o);
      if (same & (1u << o)) {
        dmu += m;
        dprec += p;
      }
    }
    if (head) {
      atomicAdd(d_mu + (size_t)b * P + k, dmu);
      atomicAdd(d_prec + (size_t)b * P + k, dprec);
    }
  }
};

template <class M>
struct NetInfo {
  static constexpr int NW = M::DYN ? LinPrecNet<typename M::real, M::NIN>::NW : 0;
};

template <class M, class TB>
__global__ void __launch_bounds__(128) elbo_fwd_kernel(const Call<typename M::real> a) {
  typedef typename M::real R;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* w = reinterpret_cast<R*>(smem_raw);
  if (M::DYN) {
    for (int i = threadIdx.x; i < NetInfo<M>::NW; i += blockDim.x) w[i] = a.weights[i];
    __syncthreads();
  }
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < a.N) traj_forward<M, TB>(a, n, w);
}

// 3 resident CTAs of 128 threads per SM (<= 168 registers) for the fp32 8-species models: measured faster than 2 CTAs
// at 190 registers and than 4 CTAs with spills; the wider models keep the full register file.
template <class M>
struct BwdBounds {
  static constexpr int min_blocks = (sizeof(typename M::real) == 4 && !M::DYN && !M::RELAY) ? 3 : 1;
};
template <class M, class TB>
__global__ void __launch_bounds__(128, BwdBounds<M>::min_blocks) elbo_bwd_kernel(const Call<typename M::real> a) {
  typedef typename M::real R;
  constexpr int NW = NetInfo<M>::NW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* w = reinterpret_cast<R*>(smem_raw);
  R* gw = w + NW;  // [NW][blockDim.x] per-thread accumulators, conflict-free (consecutive threads, consecutive banks)
  if (M::DYN) {
    for (int i = threadIdx.x; i < NW; i += blockDim.x) w[i] = a.weights[i];
    for (int i = threadIdx.x; i < NW * (int)blockDim.x; i += blockDim.x) gw[i] = R(0);
    __syncthreads();
  }
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = n < a.N;
  const int nn = active ? n : a.N - 1;
  WarpSegRed<R> red(a.d_q_mu, a.d_q_prec, a.P, nn / a.IW, active);
  if (M::DYN) {
    StridedGW<R> h{gw + threadIdx.x, (int)blockDim.x};
    traj_backward<M, TB>(a, nn, active, w, h, red);
    __syncthreads();
    for (int k = threadIdx.x; k < NW; k += blockDim.x) {
      R s = R(0);
      for (int t = 0; t < (int)blockDim.x; ++t) s += gw[k * blockDim.x + ((t + threadIdx.x) % blockDim.x)];
      atomicAdd(a.d_weights + k, s);
    }
  } else {
    NoGW<R> nogw;
    traj_backward<M, TB>(a, nn, active, w, nogw, red);
  }
}

inline int pick_block(int N) {
  // small batches are latency-bound: spread warps over as many SMs as possible (148 SMs x 4 schedulers)
  if (N <= 148 * 4 * 32) return 32;
  if (N <= 148 * 8 * 64) return 64;
  return 128;
}

template <typename R>
struct FwdLauncher {
  Call<R> a;
  cudaStream_t stream;
  template <class M, class TB>
  int run() {
    const int block = pick_block(a.N);
    const int grid = (a.N + block - 1) / block;
    const size_t smem = sizeof(R) * NetInfo<M>::NW;
    elbo_fwd_kernel<M, TB><<<grid, block, smem, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("elbo_fwd_kernel launch failed: %s", cudaGetErrorString(e));
      return VH_ERR_CUDA;
    }
    return VH_OK;
  }
};

template <typename R>
struct BwdLauncher {
  Call<R> a;
  cudaStream_t stream;
  template <class M, class TB>
  int run() {
    constexpr int NW = NetInfo<M>::NW;
    int block = pick_block(a.N);
    if (NW > 0 && sizeof(R) * NW * (block + 1) > 200 * 1024) block = 64;
    const int grid = (a.N + block - 1) / block;
    const size_t smem = sizeof(R) * NW * (block + 1);
    cudaError_t e;
    if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(elbo_bwd_kernel<M, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(smem=%zu) failed: %s", smem, cudaGetErrorString(e));
        return VH_ERR_CUDA;
      }
    }
    if (a.d_q_mu && a.P > 0) {
      cudaMemsetAsync(a.d_q_mu, 0, sizeof(R) * (size_t)a.B * a.P, stream);
      cudaMemsetAsync(a.d_q_prec, 0, sizeof(R) * (size_t)a.B * a.P, stream);
    }
    if (NW > 0) cudaMemsetAsync(a.d_weights, 0, sizeof(R) * NW, stream);
    elbo_bwd_kernel<M, TB><<<grid, block, smem, stream>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("elbo_bwd_kernel launch failed: %s", cudaGetErrorString(e));
      return VH_ERR_CUDA;
    }
    return VH_OK;
  }
};

template <class M>
int launch_fwd_model(const vh_problem* p, const vh_fwd_io* io, cudaStream_t stream) {
  typedef typename M::real R;
  FwdLauncher<R> f;
  if (const char* err = build_call<R>(p, io, nullptr, f.a)) {
    set_error("vh_elbo_terms_fwd: %s", err);
    return VH_ERR_INVALID;
  }
  f.stream = stream;
  if (M::DYN && p->n_hidden != 0) {
    set_error("NeuralPrecisions with a hidden layer (n_hidden=%d) is not implemented for the white-box models yet", p->n_hidden);
    return VH_ERR_UNSUPPORTED;
  }
  return dispatch_solver<M>(p->solver, f);
}

template <class M>
int launch_bwd_model(const vh_problem* p, const vh_bwd_io* io, cudaStream_t stream) {
  typedef typename M::real R;
  BwdLauncher<R> f;
  if (const char* err = build_call<R>(p, &io->fwd, io, f.a)) {
    set_error("vh_elbo_terms_bwd: %s", err);
    return VH_ERR_INVALID;
  }
  f.stream = stream;
  if (M::DYN && p->n_hidden != 0) {
    set_error("NeuralPrecisions with a hidden layer (n_hidden=%d) is not implemented for the white-box models yet", p->n_hidden);
    return VH_ERR_UNSUPPORTED;
  }
  return dispatch_solver<M>(p->solver, f);
}

}  // namespace vh
