// dr_blackbox kernels (models/dr_blackbox.py): forward and reverse launches + the warp-level weight-gradient GEMM.
//
// One thread per trajectory, like the white-box kernels; the flat MLP weights sit in shared memory (every lane reads
// the same word: broadcast), and every thread owns a scratch row in shared memory (odd row stride: conflict-free)
// that holds its hidden-layer activations / cotangents (vh_bb.cuh, BbRow).  Weight gradients are
// dW = sum over trajectories and RHS evaluations of (pre-activation cotangent) (x) (activation): after each reverse
// evaluation the 32 rows of a warp ARE the two operand matrices, and the warp runs the 32-deep outer-product
// accumulation with the OUTPUT distributed over lanes -- lane h owns row/column h of each matrix and keeps its slice
// in registers for the whole kernel (37 accumulators) -- so the per-step cost is shared-memory reads + FMAs, no
// atomics; one atomicAdd per matrix element per warp at the very end.
#include <cuda_runtime.h>

#include "vh_bb.cuh"
#include "vh_launch.cuh"

namespace vh {

// Warp-cooperative weight-gradient sink.  rows: this warp's 32 scratch rows (row of lane t at rows + t * ROW).
template <class F>
struct BbWarpWgrad {
  typedef typename F::real R;
  typedef typename F::L L;
  typedef typename F::ROWL RW;
  static constexpr int NST = F::NST, H = F::H, HP = F::HP, NC = F::NC, ROW = RW::ROW;

  const R* rows;
  int lane;
  R aW1[NST], aWp[NST], aWd[NST], ab[2];
  R aQ1[1 + NST], aQp[4], aQd[4], aqb[2];
  R* d;  // flat weight gradient in global memory

  __device__ void init(const R* warp_rows, R* d_weights) {
    rows = warp_rows;
    d = d_weights;
    lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NST; ++i) aW1[i] = aWp[i] = aWd[i] = R(0);
#pragma unroll
    for (int i = 0; i < 1 + NST; ++i) aQ1[i] = R(0);
#pragma unroll
    for (int i = 0; i < 4; ++i) aQp[i] = aQd[i] = R(0);
    ab[0] = ab[1] = aqb[0] = aqb[1] = R(0);
  }

  __device__ void begin() const { __syncwarp(); }

  __device__ void states(const R*) {
    __syncwarp();
    const int hl = lane < H ? lane : 0;
    const int ol = lane < NST ? lane : 0;
    const R mh = lane < H ? R(1) : R(0), mo = lane < NST ? R(1) : R(0);
#pragma unroll 4
    for (int t = 0; t < 32; ++t) {
      const R* r = rows + t * ROW;
      const R myhid = r[RW::sHID + hl] * mh, mygp = r[RW::sGPRE + hl] * mh;
#pragma unroll
      for (int s = 0; s < NST; ++s) aW1[s] += mygp * r[RW::sX + s];
#pragma unroll
      for (int o = 0; o < NST; ++o) {
        aWp[o] += r[RW::sGZP + o] * myhid;
        aWd[o] += r[RW::sGZD + o] * myhid;
      }
      ab[0] += r[RW::sGZP + ol] * mo;
      ab[1] += r[RW::sGZD + ol] * mo;
    }
  }

  __device__ void precisions(const R*) {
    __syncwarp();
    const int hl = lane < HP ? lane : 0;
    const int ol = lane < 4 ? lane : 0;
    const R mh = lane < HP ? R(1) : R(0), mo = lane < 4 ? R(1) : R(0);
#pragma unroll 4
    for (int t = 0; t < 32; ++t) {
      const R* r = rows + t * ROW;
      const R myhp = r[RW::pHP + hl] * mh, mygp = r[RW::pGPRE + hl] * mh;
#pragma unroll
      for (int i = 0; i < 1 + NST; ++i) aQ1[i] += mygp * r[i];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        aQp[o] += r[RW::pGZP + o] * myhp;
        aQd[o] += r[RW::pGZD + o] * myhp;
      }
      aqb[0] += r[RW::pGZP + ol] * mo;
      aqb[1] += r[RW::pGZD + ol] * mo;
    }
  }

  // once per trajectory, after the time loop: columns of W1 / Q1 that multiply the constants, and the hidden biases
  __device__ void consts(const R*) {
    __syncwarp();
    constexpr int total = (H + HP) * (NC + 1);
    for (int e = lane; e < total; e += 32) {
      const int h = e / (NC + 1), j = e % (NC + 1);  // j == NC: the bias
      const int gofs = h < H ? RW::GHC + h : RW::GHPC + (h - H);
      R acc = R(0);
      for (int t = 0; t < 32; ++t) {
        const R* r = rows + t * ROW;
        acc += r[gofs] * (j < NC ? r[j] : R(1));
      }
      int idx;
      if (h < H)
        idx = j < NC ? L::W1 + h * L::nin + NST + j : L::b1 + h;
      else
        idx = j < NC ? L::Q1 + (h - H) * (L::nin + 1) + 1 + NST + j : L::qb1 + (h - H);
      atomicAdd(d + idx, acc);
    }
    __syncwarp();
  }

  __device__ void flush() {
    if (lane < H) {
#pragma unroll
      for (int s = 0; s < NST; ++s) atomicAdd(d + L::W1 + lane * L::nin + s, aW1[s]);
#pragma unroll
      for (int o = 0; o < NST; ++o) {
        atomicAdd(d + L::Wp + o * H + lane, aWp[o]);
        atomicAdd(d + L::Wd + o * H + lane, aWd[o]);
      }
    }
    if (lane < NST) {
      atomicAdd(d + L::bp + lane, ab[0]);
      atomicAdd(d + L::bd + lane, ab[1]);
    }
    if (lane < HP) {
#pragma unroll
      for (int i = 0; i < 1 + NST; ++i) atomicAdd(d + L::Q1 + lane * (L::nin + 1) + i, aQ1[i]);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        atomicAdd(d + L::Qp + o * HP + lane, aQp[o]);
        atomicAdd(d + L::Qd + o * HP + lane, aQd[o]);
      }
    }
    if (lane < 4) {
      atomicAdd(d + L::qbp + lane, aqb[0]);
      atomicAdd(d + L::qbd + lane, aqb[1]);
    }
  }
};

#ifndef VH_BB_DIR
#define VH_BB_DIR 2  // 0: forward kernels only, 1: reverse kernels only, 2: both (single translation unit)
#endif

template <class F, class TB>
__global__ void __launch_bounds__(64) bb_fwd_kernel(const Call<typename F::real> a) {
  typedef typename F::real R;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* w = reinterpret_cast<R*>(smem_raw);
  R* rows = w + ((F::L::total + 3) & ~3);
  for (int i = threadIdx.x; i < F::L::total; i += blockDim.x) w[i] = a.weights[i];
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < a.N) bb_traj_forward<F, TB>(a, n, w, rows + threadIdx.x * F::ROWL::ROW);
}

template <class F, class TB>
__global__ void __launch_bounds__(64) bb_bwd_kernel(const Call<typename F::real> a) {
  typedef typename F::real R;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* w = reinterpret_cast<R*>(smem_raw);
  R* rows = w + ((F::L::total + 3) & ~3);
  for (int i = threadIdx.x; i < F::L::total; i += blockDim.x) w[i] = a.weights[i];
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = n < a.N;
  const int nn = active ? n : a.N - 1;
  WarpSegRed<R> red(a.d_q_mu, a.d_q_prec, a.P, nn / a.IW, active);
  BbWarpWgrad<F> sink;
  sink.init(rows + (threadIdx.x & ~31) * F::ROWL::ROW, a.d_weights);
  bb_traj_backward<F, TB>(a, nn, active, w, rows + threadIdx.x * F::ROWL::ROW, sink, red);
  sink.flush();
}

template <typename R, bool BWD>
struct BbLauncher {
  Call<R> a;
  cudaStream_t stream;
  template <class F, class TB>
  int run() {
    const int block = a.N <= 148 * 4 * 32 ? 32 : 64;
    const int grid = (a.N + block - 1) / block;
    const size_t smem = sizeof(R) * (((F::L::total + 3) & ~3) + (size_t)block * F::ROWL::ROW);
    cudaError_t e = cudaSuccess;
    if (BWD) {
#if VH_BB_DIR != 0
      if (smem > 48 * 1024) e = cudaFuncSetAttribute(bb_bwd_kernel<F, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(smem=%zu) failed: %s", smem, cudaGetErrorString(e));
        return VH_ERR_CUDA;
      }
      if (a.d_q_mu && a.P > 0) {
        cudaMemsetAsync(a.d_q_mu, 0, sizeof(R) * (size_t)a.B * a.P, stream);
        cudaMemsetAsync(a.d_q_prec, 0, sizeof(R) * (size_t)a.B * a.P, stream);
      }
      cudaMemsetAsync(a.d_weights, 0, sizeof(R) * F::L::total, stream);
      bb_bwd_kernel<F, TB><<<grid, block, smem, stream>>>(a);
#endif
    } else {
#if VH_BB_DIR != 1
      if (smem > 48 * 1024) e = cudaFuncSetAttribute(bb_fwd_kernel<F, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(smem=%zu) failed: %s", smem, cudaGetErrorString(e));
        return VH_ERR_CUDA;
      }
      bb_fwd_kernel<F, TB><<<grid, block, smem, stream>>>(a);
#endif
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("%s launch failed: %s", BWD ? "bb_bwd_kernel" : "bb_fwd_kernel", cudaGetErrorString(e));
      return VH_ERR_CUDA;
    }
    return VH_OK;
  }
};

template <class F, class L>
static int bb_dispatch_solver(int solver, L& f) {
  typedef typename F::real R;
  switch (solver) {
    case VH_SOLVER_EULER: return f.template run<F, TabEuler<R> >();
    case VH_SOLVER_MIDPOINT: return f.template run<F, TabMidpoint<R> >();
    case VH_SOLVER_RK4: return f.template run<F, TabRK4_38<R> >();
    case VH_SOLVER_MODEULER: return f.template run<F, TabHeun<R, true> >();
    case VH_SOLVER_MODEULERWHILE: return f.template run<F, TabHeun<R, false> >();
    default: return VH_ERR_UNSUPPORTED;
  }
}

// compiled network shape: (n_latent_species, n_hidden_decoder, n_hidden_decoder_precisions, n_z+n_x+n_y+C+D) =
// (2, 25, 20, 21), the only dr_blackbox spec the reference ships (specs/dr_blackbox_icml.yaml:20-25)
#define VH_BB_SHAPE 2, 25, 20, 21
static bool bb_shape_ok(const vh_problem* p) {
  const int nc = p->n_z + p->n_x + p->n_y + p->C + p->D;
  if (p->n_latent == 2 && p->n_hidden_states == 25 && p->n_hidden == 20 && nc == 21) return true;
  set_error("dr_blackbox kernels are compiled for (n_latent_species, n_hidden_decoder, n_hidden_decoder_precisions, "
            "n_z+n_x+n_y+C+D) = (2, 25, 20, 21); got (%d, %d, %d, %d)", p->n_latent, p->n_hidden_states, p->n_hidden, nc);
  return false;
}

template <typename R>
static int bb_fwd_t(const vh_problem* p, const vh_fwd_io* io, cudaStream_t s) {
  BbLauncher<R, false> f;
  if (const char* err = build_call<R>(p, io, nullptr, f.a)) {
    set_error("vh_elbo_terms_fwd (dr_blackbox): %s", err);
    return VH_ERR_INVALID;
  }
  f.stream = s;
  return bb_dispatch_solver<BbRhs<R, VH_BB_SHAPE> >(p->solver, f);
}

template <typename R>
static int bb_bwd_t(const vh_problem* p, const vh_bwd_io* io, cudaStream_t s) {
  BbLauncher<R, true> f;
  if (const char* err = build_call<R>(p, &io->fwd, io, f.a)) {
    set_error("vh_elbo_terms_bwd (dr_blackbox): %s", err);
    return VH_ERR_INVALID;
  }
  f.stream = s;
  return bb_dispatch_solver<BbRhs<R, VH_BB_SHAPE> >(p->solver, f);
}

#if VH_BB_DIR != 1
int launch_bb_fwd(const vh_problem* p, const vh_fwd_io* io, cudaStream_t s) {
  if (!bb_shape_ok(p)) return VH_ERR_UNSUPPORTED;
  return p->dtype == VH_F64 ? bb_fwd_t<double>(p, io, s) : bb_fwd_t<float>(p, io, s);
}
#endif
#if VH_BB_DIR != 0
int launch_bb_bwd(const vh_problem* p, const vh_bwd_io* io, cudaStream_t s) {
  if (!bb_shape_ok(p)) return VH_ERR_UNSUPPORTED;
  return p->dtype == VH_F64 ? bb_bwd_t<double>(p, io, s) : bb_bwd_t<float>(p, io, s);
}
#endif

}  // namespace vh
