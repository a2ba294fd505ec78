// __global__ wrappers around the per-trajectory code (vh_traj.cuh) and their launchers.
// Mapping: ONE THREAD PER TRAJECTORY (individual x importance sample).  The species state, the RHS constants and (in
// the reverse sweep) the adjoint state live in registers for all T steps; every global access is coalesced across
// the 32 trajectories of a warp because traces are laid out [T][S][N].  See DESIGN.md for why this mapping (and not
// a warp per trajectory) is the one that fills the machine: the white-box RHS is ~25 scalars of state with a long
// dependent chain, so the parallel axis is trajectories, not species.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include "vh_dispatch.cuh"
#include "vh_lane.cuh"
#include "vh_pdl.cuh"

// 1: accumulator warp in elbo_bwd_ws_kernel for the constant-precision models (measured: 65 us instead of 62 us at the icml
// size -- the instructions taken off the consumer warp were filling its dependency stalls; kept for the record, off)
#ifndef VH_WS_SPLIT
#define VH_WS_SPLIT 0
#endif

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
    for (int i = threadIdx.x; i < a.nw; i += blockDim.x) w[i] = a.weights[i];
    __syncthreads();
  }
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const SlotScratch<R> sc{w + ((a.nw + 3) & ~3) + threadIdx.x, (int)blockDim.x};
  if (n < a.N) traj_forward<M, TB>(a, n, w, sc);
}

constexpr int WS_PF = 4;  // checkpoint prefetch distance of the producer warp (time steps)
__device__ __forceinline__ void cp_async_elem(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_elem(double* dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NN>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(NN) : "memory");
}


// 3 resident CTAs of 128 threads per SM (<= 168 registers) for the fp32 8-species models: measured faster than 2 CTAs
// at 190 registers and than 4 CTAs with spills; the wider models keep the full register file.
#ifndef VH_BWD_PF
#define VH_BWD_PF 3
#endif
template <class M>
struct BwdBounds {
  static constexpr int min_blocks = (sizeof(typename M::real) == 4 && !M::DYN && !M::RELAY) ? 3 : 1;
};
// Latency-bound launches (fewer warps than schedulers): FWD_TEAM warps share one group of 32 trajectories for the
// sample / clip / log-prob prologue -- warp r takes theta columns r, r + FWD_TEAM, ... and leaves values (by slot) and
// its partial log q / log p in shared memory -- then warp 0 integrates alone.  Column k of slot s is written only by
// the warp that owns column k, slots without a column only by the warp that owns slot s (build_call guarantees one
// column per slot), so the prologue needs a single barrier.  fp32 / fp64; the dynamic-precision models keep their
// NeuralPrecisions weights in shared memory next to it.
#ifndef VH_FWD_TEAM
#define VH_FWD_TEAM 4
#endif
#ifndef VH_FWD_LANE_DEFAULT
#define VH_FWD_LANE_DEFAULT 0
#endif
// named barriers (two warps unless a count is given) and 16-byte ring slots: shared by the team forward kernel and the
// warp-specialised reverse kernels below
__device__ __forceinline__ void named_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
template <int NT>
__device__ __forceinline__ void named_bar_sync_n(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NT) : "memory"); }
template <int NT>
__device__ __forceinline__ void named_bar_arrive_n(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(NT) : "memory"); }

// Ring slots are arrays of 16-byte vectors, [vector][lane]: item i of a lane sits in vector i / V, element i % V
// (V = 4 fp32 / 2 fp64).  One 128-bit shared-memory access moves V items (conflict-free: a quarter warp per wavefront);
// as 32-bit accesses the hand-off was 36 + 36 instructions per step and the warps queued on the load/store unit.
template <typename R>
struct RingVec {
  static constexpr int V = 16 / sizeof(R);
  struct alignas(16) Vec {
    R v[V];
  };
  static constexpr int vectors(int items) { return (items + V - 1) / V; }
  // elements a slot of `items` items occupies
  static constexpr int slot_elems(int items) { return vectors(items) * V * 32; }
  template <int N>
  __device__ static void store(R* slot, int lane, const R (&buf)[N]) {
    Vec* s = reinterpret_cast<Vec*>(slot) + lane;
#pragma unroll
    for (int g = 0; g < vectors(N); ++g) {
      Vec q;
#pragma unroll
      for (int e = 0; e < V; ++e) q.v[e] = g * V + e < N ? buf[g * V + e] : R(0);
      s[g * 32] = q;
    }
  }
  // items [first, first + N) of the slot; `first` must be a multiple of V
  template <int N>
  __device__ static void load(const R* slot, int lane, int first, R (&buf)[N]) {
    const Vec* s = reinterpret_cast<const Vec*>(slot) + lane + (first / V) * 32;
#pragma unroll
    for (int g = 0; g < vectors(N); ++g) {
      const Vec q = s[g * 32];
#pragma unroll
      for (int e = 0; e < V; ++e)
        if (g * V + e < N) buf[g * V + e] = q.v[e];
    }
  }
};

constexpr int FWD_TEAM = VH_FWD_TEAM;
// SCRIBE: after the prologue warp 1 stays as the "scribe" of the time loop.  Warp 0 only integrates (one RK step per time
// point) and hands every state x_k over through a two-slot shared-memory ring; the scribe stores the trace, applies the
// observation map, accumulates the Gaussian log-likelihood and fetches the observations.  That takes ~55 of the ~210
// instructions per step off the warp whose in-order instruction stream IS the launch time at this size.
template <class M, class TB, bool SCRIBE>
__global__ void __launch_bounds__(FWD_TEAM * 32) elbo_fwd_team_kernel(const Call<typename M::real> a) {
  typedef typename M::real R;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* sm = reinterpret_cast<R*>(smem_raw);
  // launched under the tail of the encoder forward when that is the preceding kernel of the stream (vh_pdl.cuh); everything
  // below reads its q tables
  pdl_wait();
  pdl_trigger();  // the reverse launch may become resident now; it blocks in pdl_wait() until this grid has completed
  const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 32 + lane;
  const bool active = n0 < a.N;
  const int n = active ? n0 : a.N - 1;
  const int b = n / a.IW;
  constexpr int S = M::S, NS = M::NS;
  constexpr int XSLOT = RingVec<R>::slot_elems(S);
  R* xring = sm;  // [2][XSLOT] (scribe form), 16-byte aligned
  R* slots = sm + (SCRIBE ? 2 * XSLOT : 0);
  const SlotScratch<R> loc{slots + lane, 32};
  R* part = slots + M::NSLOT * 32;    // [2][FWD_TEAM][32]
  R* wsm = part + 2 * FWD_TEAM * 32;  // NeuralPrecisions weights (dynamic-precision models)
  if (M::DYN) {
    for (int i = threadIdx.x; i < a.nw; i += blockDim.x) wsm[i] = a.weights[i];
  }
  for (int s = role; s < M::NSLOT; s += FWD_TEAM) {
    const int src = a.slot_src[s];
    if (src < 0) loc[s] = src != VH_SLOT_UNUSED ? a.extra[(size_t)(-1 - src) * a.N + n] : R(0);
  }
  R lq = R(0), lp = R(0);
#pragma unroll 3
  for (int k = role; k < a.P; k += FWD_TEAM) {
    const R v = sample_column(a, n, b, k, lq, lp, active);
    const int s = a.col_slot[k];
    if (s >= 0) loc[s] = v;
  }
  part[role * 32 + lane] = lq;
  part[(FWD_TEAM + role) * 32 + lane] = lp;
  __syncthreads();
  if constexpr (!SCRIBE) {
    if (role != 0 || !active) return;
    lq = R(0);
    lp = R(0);
#pragma unroll
    for (int r = 0; r < FWD_TEAM; ++r) {
      lq += part[r * 32 + lane];
      lp += part[(FWD_TEAM + r) * 32 + lane];
    }
    R th[M::NSLOT];
#pragma unroll
    for (int s = 0; s < M::NSLOT; ++s) th[s] = M::uses(s) ? loc[s] : R(0);
    traj_forward_from<M, TB>(a, n, M::DYN ? wsm : nullptr, th, lq, lp);
  } else {
    if (role > 1) return;
    // whole warps from here on (the named barriers count threads); lanes past the batch shadow its last trajectory
    enum { FULL0 = 1, EMPTY0 = 3 };
    const size_t N = a.N;
    const int T = a.T;
    if (role == 0) {
      // ---------------- integrator ----------------
      lq = R(0);
      lp = R(0);
#pragma unroll
      for (int r = 0; r < FWD_TEAM; ++r) {
        lq += part[r * 32 + lane];
        lp += part[(FWD_TEAM + r) * 32 + lane];
      }
      Rhs<M> f;
      f.w = M::DYN ? wsm : nullptr;
      f.nh = a.n_hidden;
      R x[S];
      {
        R th[M::NSLOT];
        R tc[3];
#pragma unroll
        for (int s = 0; s < M::NSLOT; ++s) th[s] = M::uses(s) ? loc[s] : R(0);
        M::treatments(a.treatments + (size_t)b * a.C, tc);
        M::setup(th, tc, f.c);
        M::init_state(th, tc, x);
      }
      const R h0 = a.times[1] - a.times[0];
      R t0 = a.times[0], t1 = a.times[1];
      for (int k = 0; k < T; ++k) {
        const int slot = k & 1;
        const R t2 = ld_early(a.times + (k + 2 < T ? k + 2 : T - 1));
        if (k >= 2) named_bar_sync(EMPTY0 + slot);  // the scribe has taken x_{k-2} out of this slot
        RingVec<R>::store(xring + slot * XSLOT, lane, x);
        __threadfence_block();
        named_bar_arrive(FULL0 + slot);
        if (k + 1 < T) rk_step<Rhs<M>, TB>(f, t0, t1, TB::const_h ? h0 : (t1 - t0), x);
        t0 = t1;
        t1 = t2;
      }
      if (active) {
        if (a.logp_theta) a.logp_theta[n] = lp;
        if (a.logq_theta) a.logq_theta[n] = lq;
      }
    } else {
      // ---------------- scribe: trace, observation map, log-likelihood (vihds/ode.py:84-93, training.py:24-44) ----------------
      R prec[4], lprec[4], ll[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        prec[o] = M::DYN ? R(1) : loc[S_prec_x + o];
        lprec[o] = M::DYN ? R(0) : vlog(prec[o]);
        ll[o] = R(0);
      }
      const R* obs = a.obs ? a.obs + (size_t)b * 4 * T : nullptr;
      R* xs = (a.x_states && active) ? a.x_states + n : nullptr;
      R* xpr = (a.x_predict && active) ? a.x_predict + n : nullptr;
      R ob[4] = {R(0), R(0), R(0), R(0)};
      if (obs) {
#pragma unroll
        for (int o = 0; o < 4; ++o) ob[o] = obs[o * T];
      }
      for (int k = 0; k < T; ++k) {
        const int slot = k & 1;
        R obn[4] = {R(0), R(0), R(0), R(0)};
        const int kn = k + 1 < T ? k + 1 : k;
        if (obs) {
#pragma unroll
          for (int o = 0; o < 4; ++o) obn[o] = ld_early(obs + o * T + kn);
        }
        R x[S];
        named_bar_sync(FULL0 + slot);
        RingVec<R>::load(xring + slot * XSLOT, lane, 0, x);
        if (k + 2 < T) {
          __threadfence_block();
          named_bar_arrive(EMPTY0 + slot);
        }
        if (xs) {
#pragma unroll
          for (int q = 0; q < S; ++q) xs[(size_t)q * N] = x[q];
          xs += (size_t)S * N;
        }
        R xp[4];
        M::observe(x, xp);
        if (xpr) {
#pragma unroll
          for (int o = 0; o < 4; ++o) xpr[(size_t)o * N] = xp[o];
          xpr += (size_t)4 * N;
        }
        if (obs) {
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const R pr = M::DYN ? x[NS + o] : prec[o];
            const R lpr = M::DYN ? vlog(pr) : lprec[o];
            const R d = xp[o] - ob[o];
            ll[o] = loglik_add(ll[o], pr, lpr, d);
          }
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) ob[o] = obn[o];
      }
      if (a.logp_species && active) {
#pragma unroll
        for (int o = 0; o < 4; ++o) a.logp_species[(size_t)n * 4 + o] = ll[o];
      }
    }
  }
}

template <class M, class TB>
__global__ void __launch_bounds__(128, BwdBounds<M>::min_blocks) elbo_bwd_kernel(const Call<typename M::real> a) {
  typedef typename M::real R;
  const int NW = a.nw;  // run-time: the NeuralPrecisions net may have a hidden layer of any width <= HidPrecNet::MAXH
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
  const SlotScratch<R> sc{w + ((NW * ((int)blockDim.x + 1) + 3) & ~3) + threadIdx.x, (int)blockDim.x};
  if (M::DYN) {
    StridedGW<R> h{gw + threadIdx.x, (int)blockDim.x};
    DirectCk<R, M::S> ck;
    traj_backward<M, TB>(a, nn, active, w, h, red, sc, ck);
    __syncthreads();
    for (int k = threadIdx.x; k < NW; k += blockDim.x) {
      R s = R(0);
      for (int t = 0; t < (int)blockDim.x; ++t) s += gw[k * blockDim.x + ((t + threadIdx.x) % blockDim.x)];
      atomicAdd(a.d_weights + k, s);
    }
  } else {
    NoGW<R> nogw;
    // (a cp.async staging ring as in the warp-specialised kernel was measured here too: 1.340 vs 1.306 ms at
    // N = 131,072 -- with 12 resident warps per SM the one-step register prefetch already hides the latency)
    DirectCk<R, M::S, VH_BWD_PF> ck;
    traj_backward<M, TB>(a, nn, active, w, nogw, red, sc, ck);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-specialised reverse sweep for latency-bound launches (N small enough that every scheduler holds at most one
// warp).  A CTA is a PAIR of warps over the same 32 trajectories:
//   warp 0 (producer)  loads checkpoint x(k) and re-evaluates the stages of step k -> k+1 (rk_stages_forward): this
//                      needs the checkpoint only, so it runs ahead of the adjoint;
//   warp 1 (consumer)  takes x(k), the stage derivatives and the kept RHS intermediates from a two-slot shared-memory
//                      ring and does the part that is serial in lambda: rk_step_adjoint + the emission adjoint.
// The two halves of a time step (~200 and ~260 instructions) run on two schedulers instead of one after the other on
// one.  Hand-off: named barriers (full / empty per slot, bar.arrive on one side, bar.sync on the other), data laid out
// [item][lane] (conflict-free).  fp32 / fp64.  Dynamic-precision models: the NeuralPrecisions weights sit in shared
// memory too, and a third warp accumulates their gradient (WgradRing).
// ---------------------------------------------------------------------------------------------------------------

template <class M, class TB>
struct WsRing {
  typedef typename M::real R;
  typedef StageData<Rhs<M>, TB> SD;
  static constexpr int KN = sizeof(typename Rhs<M>::Kept) / sizeof(R);  // species intermediates (+ NeuralPrecisions activations)
  static constexpr int NITEM = M::S + SD::nk * M::S + TB::s * KN;
  static constexpr int SLOT = RingVec<R>::slot_elems(NITEM);  // elements per ring slot
  // flat item order: x | stage derivatives | kept intermediates of every stage
  __device__ static void pack(const R* x, const SD& sd, R (&buf)[NITEM]) {
    int it = 0;
#pragma unroll
    for (int q = 0; q < M::S; ++q) buf[it++] = x[q];
#pragma unroll
    for (int i = 0; i < SD::nk; ++i)
#pragma unroll
      for (int q = 0; q < M::S; ++q) buf[it++] = sd.k[i][q];
#pragma unroll
    for (int i = 0; i < TB::s; ++i) {
      const R* m = reinterpret_cast<const R*>(&sd.kept[i]);
#pragma unroll
      for (int j = 0; j < KN; ++j) buf[it++] = m[j];
    }
  }
  __device__ static void unpack(const R (&buf)[NITEM], R* x, SD& sd) {
    int it = 0;
#pragma unroll
    for (int q = 0; q < M::S; ++q) x[q] = buf[it++];
#pragma unroll
    for (int i = 0; i < SD::nk; ++i)
#pragma unroll
      for (int q = 0; q < M::S; ++q) sd.k[i][q] = buf[it++];
#pragma unroll
    for (int i = 0; i < TB::s; ++i) {
      R* m = reinterpret_cast<R*>(&sd.kept[i]);
#pragma unroll
      for (int j = 0; j < KN; ++j) m[j] = buf[it++];
    }
  }
  __device__ static void put(R* slot, int lane, const R* x, const SD& sd) {
    R buf[NITEM];
    pack(x, sd, buf);
    RingVec<R>::store(slot, lane, buf);
  }
  __device__ static void get(const R* slot, int lane, R* x, SD& sd) {
    R buf[NITEM];
    RingVec<R>::load(slot, lane, 0, buf);
    unpack(buf, x, sd);
  }
};

// models whose reverse step is split over a consumer and an accumulator warp (see elbo_bwd_ws_kernel)
template <class M>
struct WsSplit {
  static constexpr bool value = !M::DYN && VH_WS_SPLIT;
};

// WS_WARPS warps per CTA share one group of 32 trajectories: warp 0 = producer, warp 1 = consumer, the others sleep on
// a named barrier until the lambda recurrence is finished and then take their share of the theta columns of the
// chain-rule epilogue (35 columns x ~260 instructions were 42 % of the consumer warp's time when it did them alone).
#ifndef VH_WS_WARPS
#define VH_WS_WARPS 4
#endif
constexpr int WS_WARPS = VH_WS_WARPS;
__device__ __forceinline__ void named_bar_sync_all(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(WS_WARPS * 32) : "memory");
}

// Dynamic-precision models: the consumer warp does not accumulate the NeuralPrecisions weight gradient itself (112
// shared-memory read-modify-writes per VJP at 13 inputs -- 40 % of its instructions); it hands the two factors of the
// outer product (a[NIN], gzp[4], gzd[4]) to a third warp through a two-slot ring, and that warp -- otherwise asleep
// until the epilogue -- keeps all NW accumulators in registers.
enum { WG_FULL0 = 7, WG_EMPTY0 = 9 };
template <typename R, int NIN>
struct WgradRing {
  static constexpr int NITEM = NIN + 8;
  R* buf;  // [2][NITEM][32] + lane
  int it;
  __device__ void add(int, R) const {}
  template <int N2>
  __device__ void outer(const R* a, const R* gzp, const R* gzd) {
    const int slot = it & 1;
    if (it >= 2) named_bar_sync(WG_EMPTY0 + slot);
    R* s = buf + slot * NITEM * 32;
#pragma unroll
    for (int j = 0; j < NIN; ++j) s[j * 32] = a[j];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      s[(NIN + o) * 32] = gzp[o];
      s[(NIN + 4 + o) * 32] = gzd[o];
    }
    __threadfence_block();
    named_bar_arrive(WG_FULL0 + slot);
    ++it;
  }
};

// IWAE reduction in the prologue of the reverse launch (vihds/training.py:134-148 and the unit upstream gradient of
// elbo.backward(), :334): every CTA reduces the IW log-weights of the (one or two) individuals its 32 trajectories belong
// to -- 1,200 floats from L2, all warps of the team -- instead of a separate one-block-per-individual launch.  Returns
// d cost / d log_w of this lane's trajectory; the CTA that holds an individual's first sample adds its term to the cost.
// Same arithmetic as iwae_fwd_kernel (vh_api.cu): max-shifted logsumexp, NaN log-weights surface in the cost.
template <typename R>
__device__ R iwae_upstream_in_kernel(const Call<R>& a, int n, bool active, R* red /* >= 2 * nwarps + 2 elements */) {
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int n_first = blockIdx.x * 32, n_last = min(a.N - 1, n_first + 31);
  const int b_first = n_first / a.IW, b_last = n_last / a.IW, b_mine = n / a.IW;
  const R inv_b = R(1) / R(a.iw_b_total);
  auto logw = [&](size_t m) {
    const R* l = a.logp_species + m * 4;
    return ((l[0] + l[1]) + (l[2] + l[3])) + a.logp_theta[m] - a.logq_theta[m];
  };
  R lse_mine = R(0);
  for (int bb = b_first; bb <= b_last; ++bb) {
    const size_t base = (size_t)bb * a.IW;
    R mx = -INFINITY;
    bool has_nan = false;
    for (int i = tid; i < a.IW; i += blockDim.x) {
      const R v = logw(base + i);
      has_nan |= (v != v);
      mx = v > mx ? v : mx;
    }
    R nanf = has_nan ? R(1) : R(0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const R other = __shfl_xor_sync(full, mx, o);
      mx = other > mx ? other : mx;
      nanf += __shfl_xor_sync(full, nanf, o);
    }
    __syncthreads();
    if (lane == 0) {
      red[warp] = mx;
      red[nwarp + warp] = nanf;
    }
    __syncthreads();
    mx = red[0];
    nanf = red[nwarp];
    for (int w = 1; w < nwarp; ++w) {
      mx = red[w] > mx ? red[w] : mx;
      nanf += red[nwarp + w];
    }
    const R shift = (mx == -INFINITY || mx == INFINITY) ? R(0) : mx;
    R se = R(0);
    for (int i = tid; i < a.IW; i += blockDim.x) se += vexp(logw(base + i) - shift);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(full, se, o);
    __syncthreads();
    if (lane == 0) red[warp] = se;
    __syncthreads();
    se = red[0];
    for (int w = 1; w < nwarp; ++w) se += red[w];
    R lse = vlog(se) + shift;
    if (nanf > R(0)) lse = NAN;
    if (bb == b_mine) lse_mine = lse;
    if (tid == 0 && base >= (size_t)n_first) atomicAdd(a.iw_cost, -(lse - vlog(R(a.IW))) * inv_b);
  }
  __syncthreads();
  return active ? -vexp(logw((size_t)n) - lse_mine) * inv_b : R(0);
}

template <class M, class TB>
__global__ void __launch_bounds__(WS_WARPS * 32) elbo_bwd_ws_kernel(const Call<typename M::real> a) {
  typedef typename M::real R;
  typedef WsRing<M, TB> Ring;
  constexpr int S = M::S;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* ring = reinterpret_cast<R*>(smem_raw);
  pdl_wait();     // launched under the forward kernel's tail (vh_bwd_io.outputs_cleared): everything below reads its outputs
  pdl_trigger();  // the next launch on the stream (encoder backward) may stage its weights while this grid runs
  const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 32 + lane;
  const bool active = n0 < a.N;
  const int n = active ? n0 : a.N - 1;  // every lane runs (the named barriers count whole warps)
  const int b = n / a.IW;
  const size_t N = a.N;
  const int T = a.T;
  enum { FULL0 = 1, EMPTY0 = 3, EPILOGUE = 5, PROLOGUE = 6 };
  const SlotScratch<R> thv{ring + 2 * Ring::SLOT + lane, 64};        // theta by slot, filled once by the whole team
  const SlotScratch<R> gloc{ring + 2 * Ring::SLOT + 32 + lane, 64};  // its cotangent, written by the consumer
  // dynamic-precision models: NeuralPrecisions weights and the hand-off ring of the weight-gradient warp
  constexpr int NW = NetInfo<M>::NW;
  typedef WgradRing<R, M::NIN> WG;
  R* wsm = ring + 2 * Ring::SLOT + M::NSLOT * 64 + (WS_PF + 1) * S * 32;
  R* wgbuf = wsm + ((NW + 3) & ~3);
  // Constant-precision models: the parameter cotangents (23 accumulators for dr_constant, ~50 of the consumer's ~360
  // instructions per step and VJP) only ACCUMULATE -- they are not part of the recurrence in lambda.  A third warp (the
  // "accumulator", otherwise asleep until the epilogue) reads the same ring slots as the consumer, takes the cotangents
  // of the stage derivatives from the consumer through a second two-slot ring, and keeps those accumulators; the
  // consumer's step shrinks to the state cotangents (rk_step_adjoint_x / rk_step_adjoint_c, vh_traj.cuh).
  constexpr bool SPLIT = WsSplit<M>::value;
  constexpr int RD = SPLIT ? 96 : 64;                 // threads on the full / empty barriers of the main ring
  constexpr int GN = TB::s * S;                       // items per slot of the consumer -> accumulator ring
  R* gring = wgbuf + 2 * (M::NIN + 8) * 32;           // [2][GN][32]
  R* gcsm = gring + 2 * GN * 32;                      // [NC][32]: the accumulator's result, read by the consumer's epilogue
  enum { G_FULL0 = 7, G_EMPTY0 = 9, GC_READY = 11 };  // ids 7..10 are the weight-gradient ring's in the DYN kernels
  if (M::DYN) {
    for (int i = threadIdx.x; i < NW; i += blockDim.x) wsm[i] = a.weights[i];
  }
  // upstream gradients: handed in, or (fused IWAE) derived here from the forward call's per-sample terms
  const bool iwae = a.iw_b_total > 0;
  const R gup = iwae ? iwae_upstream_in_kernel(a, n, active, ring) : R(0);
  const R glq = iwae ? -gup : ((a.g_logq_theta && active) ? a.g_logq_theta[n] : R(0));
  const R glp = iwae ? gup : ((a.g_logp_theta && active) ? a.g_logp_theta[n] : R(0));
  // theta of this trajectory: warp r fetches columns r, r + WS_WARPS, ... (read back from the forward's theta planes,
  // or re-sampled) and the slots without a column that it owns; the values stay in shared memory for the epilogue
  for (int s = role; s < M::NSLOT; s += WS_WARPS) {
    const int src = a.slot_src[s];
    if (src < 0) thv[s] = src != VH_SLOT_UNUSED ? a.extra[(size_t)(-1 - src) * N + n] : R(0);
  }
#pragma unroll 3
  for (int k = role; k < a.P; k += WS_WARPS) {
    R lq = R(0), lp = R(0);
    const R v = a.theta_in ? a.theta_in[(size_t)k * N + n] : sample_column(a, n, b, k, lq, lp, false);
    const int s = a.col_slot[k];
    if (s >= 0) thv[s] = v;
  }
  named_bar_sync_all(PROLOGUE);
  if (role < 2 || (SPLIT && role == 2)) {
    // every role of the recurrence needs the RHS constants
    Rhs<M> f;
    f.w = M::DYN ? wsm : nullptr;
    f.nh = 0;  // the warp-specialised form keeps the no-hidden-layer net only (its weight-gradient warp is sized for it)
    R prec[4], iprec[4];
    {
      R th[M::NSLOT];
      R tc[3];
#pragma unroll
      for (int s = 0; s < M::NSLOT; ++s) th[s] = M::uses(s) ? thv[s] : R(0);
      M::treatments(a.treatments + (size_t)b * a.C, tc);
      M::setup(th, tc, f.c);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        prec[o] = M::DYN ? R(1) : th[S_prec_x + o];
        iprec[o] = R(1) / prec[o];
      }
    }
    const R h0 = a.times[1] - a.times[0];
    const size_t slab = (size_t)S * N;
    if (role == 0) {
      // ---------------- producer: checkpoints + stage re-evaluation, one step ahead ----------------
      // Checkpoints x_k come through a cp.async staging ring WS_PF steps ahead: with a one-step register prefetch
      // the producer sat 58 % of its time on the long scoreboard (one warp per scheduler cannot hide an L2 / HBM
      // round trip behind ~190 instructions) and the consumer, in turn, 46 % of its loop on the FULL barrier.
      // Each thread reads back only what it copied itself, so cp.async.wait_group is all the synchronisation needed.
      R* ck = ring + 2 * Ring::SLOT + M::NSLOT * 64 + lane;  // [WS_PF + 1][S][32]
      const R* xs = a.x_states + (size_t)(T - 2) * slab + n;
      int kw = T - 2, sw = 0, sr = 0;  // step / ring slot of the next copy, ring slot of the next read
      auto issue = [&]() {
        if (kw >= 0) {
#pragma unroll
          for (int q = 0; q < S; ++q) cp_async_elem(ck + (sw * S + q) * 32, xs + (size_t)q * N);
          xs -= slab;
        }
        cp_async_commit();
        --kw;
        sw = sw == WS_PF ? 0 : sw + 1;
      };
#pragma unroll
      for (int d = 0; d < WS_PF; ++d) issue();
      R x[S];
      R t1 = a.times[T - 1], t0 = a.times[T - 2];
      for (int k = T - 2; k >= 0; --k) {
        const int it = T - 2 - k, slot = it & 1;
        const int kp = k > 0 ? k - 1 : 0;
        issue();  // refills the slot that was read one iteration ago
        cp_async_wait<WS_PF>();
#pragma unroll
        for (int q = 0; q < S; ++q) x[q] = ck[(sr * S + q) * 32];
        sr = sr == WS_PF ? 0 : sr + 1;
        const R tp = ld_early(a.times + kp);
        typename Ring::SD sd;
        rk_stages_forward<Rhs<M>, TB>(f, t0, t1, TB::const_h ? h0 : (t1 - t0), x, sd);
        if (it >= 2) named_bar_sync_n<RD>(EMPTY0 + slot);  // the reader(s) have released this slot
        Ring::put(ring + slot * Ring::SLOT, lane, x, sd);
        __threadfence_block();
        named_bar_arrive_n<RD>(FULL0 + slot);
        t1 = t0;
        t0 = tp;
      }
    } else if (role == 1) {
      // ---------------- consumer: everything that is serial in lambda ----------------
      R gl[4], gprec[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        gprec[o] = R(0);
        gl[o] = iwae ? gup : ((a.g_logp_species && active) ? a.g_logp_species[(size_t)n * 4 + o] : R(0));
      }
      const R* obs = a.obs ? a.obs + (size_t)b * 4 * T : nullptr;
      const R* gxs = a.g_x_states ? a.g_x_states + (size_t)(T - 1) * slab + n : nullptr;
      const R* gxpr = a.g_x_predict ? a.g_x_predict + (size_t)(T - 1) * 4 * N + n : nullptr;
      typename M::Consts gc;
#pragma unroll
      for (int i = 0; i < M::NC; ++i) gc.v[i] = R(0);
      NoGW<R> nogw;
      WG sgw{wgbuf + lane, 0};
      R lam[S], x[S];
      R ob[4] = {R(0), R(0), R(0), R(0)}, obp[4] = {R(0), R(0), R(0), R(0)};
#pragma unroll
      for (int q = 0; q < S; ++q) {
        lam[q] = R(0);
        x[q] = a.x_states[(size_t)(T - 1) * slab + (size_t)q * N + n];
      }
      if (obs) {
#pragma unroll
        for (int o = 0; o < 4; ++o) ob[o] = obs[o * T + T - 1];
      }
      R t1 = a.times[T - 1], t0 = t1;
      for (int k = T - 1; k >= 0; --k) {
        const int kp = k > 0 ? k - 1 : 0;
        if (obs) {
#pragma unroll
          for (int o = 0; o < 4; ++o) obp[o] = ld_early(obs + o * T + kp);
        }
        const R tp = ld_early(a.times + kp);
        if (k + 1 < T) {
          const int it = T - 2 - k, slot = it & 1;
          typename Ring::SD sd;
          named_bar_sync_n<RD>(FULL0 + slot);
          Ring::get(ring + slot * Ring::SLOT, lane, x, sd);
          if (k >= 2) {  // slot will be refilled with step k-2; the last two fills are never waited for
            __threadfence_block();
            named_bar_arrive_n<RD>(EMPTY0 + slot);
          }
          if constexpr (SPLIT) {
            // the cotangent of every stage derivative goes to the accumulator warp as soon as it is final
            R* gs = gring + slot * GN * 32 + lane;
            auto pub = [&](int i, const R* g) {
              if (i == TB::s - 1 && it >= 2) named_bar_sync(G_EMPTY0 + slot);  // the accumulator has read iteration it - 2
#pragma unroll
              for (int q = 0; q < S; ++q) gs[(i * S + q) * 32] = g[q];
              if (i == 0) {
                __threadfence_block();
                named_bar_arrive(G_FULL0 + slot);
              }
            };
            rk_step_adjoint_x<Rhs<M>, TB>(f, t0, t1, TB::const_h ? h0 : (t1 - t0), x, sd, lam, pub);
          } else if (M::DYN)
            rk_step_adjoint<Rhs<M>, TB>(f, t0, t1, TB::const_h ? h0 : (t1 - t0), x, sd, lam, gc, sgw);
          else
            rk_step_adjoint<Rhs<M>, TB>(f, t0, t1, TB::const_h ? h0 : (t1 - t0), x, sd, lam, gc, nogw);
        }
        // emission at time k
        R xp[4], gxp[4];
        M::observe(x, xp);
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          gxp[o] = (gxpr && active) ? gxpr[(size_t)o * N] : R(0);
          if (obs) {
            const R pr = M::DYN ? x[M::NS + o] : prec[o];
            const R ipr = M::DYN ? vdiv(R(1), pr) : iprec[o];
            const R d = xp[o] - ob[o];
            gxp[o] -= gl[o] * pr * d;
            const R gp = gl[o] * R(0.5) * (ipr - d * d);
            if (M::DYN)
              lam[M::NS + o] += gp;
            else
              gprec[o] += gp;
          }
        }
        M::observe_vjp(x, gxp, lam);
        if (gxs) {
          if (active) {
#pragma unroll
            for (int q = 0; q < S; ++q) lam[q] += gxs[(size_t)q * N];
          }
          gxs -= slab;
        }
        if (gxpr) gxpr -= (size_t)4 * N;
#pragma unroll
        for (int o = 0; o < 4; ++o) ob[o] = obp[o];
        t1 = t0;
        t0 = tp;
      }
      // chain rule back to theta + scatter (as traj_backward)
      R gth[M::NSLOT];
#pragma unroll
      for (int s = 0; s < M::NSLOT; ++s) gth[s] = R(0);
      {
        R th[M::NSLOT];
        R tc[3];
#pragma unroll
        for (int s = 0; s < M::NSLOT; ++s) th[s] = M::uses(s) ? thv[s] : R(0);
        M::treatments(a.treatments + (size_t)b * a.C, tc);
        M::init_state_vjp(lam, gth);
        if constexpr (SPLIT) {
          named_bar_sync(GC_READY);  // the accumulator warp has finished
#pragma unroll
          for (int i = 0; i < M::NC; ++i) gc.v[i] = gcsm[i * 32 + lane];
        }
        M::setup_vjp(th, tc, f.c, gc, gth);
        if (!M::DYN) {
#pragma unroll
          for (int o = 0; o < 4; ++o) gth[S_prec_x + o] += gprec[o];
        }
      }
#pragma unroll
      for (int s = 0; s < M::NSLOT; ++s) gloc[s] = M::uses(s) ? gth[s] : R(0);
    } else if constexpr (SPLIT) {
      // ---------------- accumulator: parameter cotangents of every stage VJP ----------------
      typename M::Consts gc;
#pragma unroll
      for (int i = 0; i < M::NC; ++i) gc.v[i] = R(0);
      R x[S];
      R t1 = a.times[T - 1], t0 = a.times[T > 1 ? T - 2 : 0];
      for (int k = T - 2; k >= 0; --k) {
        const int it = T - 2 - k, slot = it & 1;
        const R tp = ld_early(a.times + (k > 0 ? k - 1 : 0));
        typename Ring::SD sd;
        named_bar_sync_n<RD>(FULL0 + slot);
        Ring::get(ring + slot * Ring::SLOT, lane, x, sd);
        if (k >= 2) {
          __threadfence_block();
          named_bar_arrive_n<RD>(EMPTY0 + slot);
        }
        R gk[TB::s][S];
        named_bar_sync(G_FULL0 + slot);
        {
          const R* gs = gring + slot * GN * 32 + lane;
#pragma unroll
          for (int i = 0; i < TB::s; ++i)
#pragma unroll
            for (int q = 0; q < S; ++q) gk[i][q] = gs[(i * S + q) * 32];
        }
        if (k >= 2) {  // the consumer waits for this before it writes iteration it + 2
          __threadfence_block();
          named_bar_arrive(G_EMPTY0 + slot);
        }
        rk_step_adjoint_c<Rhs<M>, TB>(f, t0, t1, TB::const_h ? h0 : (t1 - t0), x, sd, gk, gc);
        t1 = t0;
        t0 = tp;
      }
#pragma unroll
      for (int i = 0; i < M::NC; ++i) gcsm[i * 32 + lane] = gc.v[i];
      __threadfence_block();
      named_bar_arrive(GC_READY);
    }  // consumer / accumulator
  }    // producer / consumer / accumulator
  if (M::DYN && role == 2) {
    // ---------------- weight-gradient warp: acc[k] += outer product of every NeuralPrecisions VJP ----------------
    constexpr int NIN = M::NIN, H = 4 * NIN + 4;
    R acc[NW > 0 ? NW : 1];
#pragma unroll
    for (int i = 0; i < NW; ++i) acc[i] = R(0);
    const int nev = TB::s * (T - 1);
    for (int e = 0; e < nev; ++e) {
      const int slot = e & 1;
      const R* sb = wgbuf + lane + slot * WG::NITEM * 32;
      R av[NIN], gp[4], gd[4];
      named_bar_sync(WG_FULL0 + slot);
#pragma unroll
      for (int j = 0; j < NIN; ++j) av[j] = sb[j * 32];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        gp[o] = sb[(NIN + o) * 32];
        gd[o] = sb[(NIN + 4 + o) * 32];
      }
      if (e + 2 < nev) {
        __threadfence_block();
        named_bar_arrive(WG_EMPTY0 + slot);
      }
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        acc[4 * NIN + o] += gp[o];
        acc[H + 4 * NIN + o] += gd[o];
#pragma unroll
        for (int j = 0; j < NIN; ++j) {
          acc[o * NIN + j] += gp[o] * av[j];
          acc[H + o * NIN + j] += gd[o] * av[j];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      R v = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) atomicAdd(a.d_weights + i, v);
    }
  }
  named_bar_sync_all(EPILOGUE);  // gloc is complete (bar.sync orders the shared-memory writes)
  WarpSegRed<R> red(a.d_q_mu, a.d_q_prec, a.P, b, active);
#pragma unroll 3
  for (int k = role; k < a.P; k += WS_WARPS) {
    const int s = a.col_slot[k];
    R dmu = R(0), dprec = R(0);
    if (active) column_vjp(a, n, b, k, s >= 0 ? gloc[s] : R(0), glq, glp, dmu, dprec);
    red(b, k, dmu, dprec, active);
  }
  if (a.d_extra && active) {
    for (int s = role; s < M::NSLOT; s += WS_WARPS) {
      const int src = a.slot_src[s];
      if (src < 0 && src != VH_SLOT_UNUSED) a.d_extra[(size_t)(-1 - src) * N + n] = gloc[s];
    }
  }
}

}  // namespace vh
#include "vh_bwd_mx.cuh"
namespace vh {

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
  // VIHDS_FWD_TEAM=0|1 overrides (tests / measurements; read per call so that one process can compare the forms)
  template <class M>
  static bool use_team(int block) {
    const char* m = getenv("VIHDS_FWD_TEAM");
    const int mode = (!m || !*m) ? -1 : atoi(m);
    return mode >= 0 ? mode != 0 : block == 32;
  }
  // VIHDS_FWD_LANE=0|1: the lane-split forward kernel (vh_lane.cuh: 8 lanes per trajectory) for dr_constant v1 / v2 in
  // fp32; default: on for latency-bound launches (see DESIGN.md section 4 for the measurement)
  template <class M>
  static bool use_lane(int block) {
    const char* m = getenv("VIHDS_FWD_LANE");
    const int mode = (!m || !*m) ? -1 : atoi(m);
    return mode >= 0 ? mode != 0 : VH_FWD_LANE_DEFAULT && block == 32;
  }
  template <class M, class TB>
  int run() {
    const int block = pick_block(a.N);
    const int grid = (a.N + block - 1) / block;
    const size_t smem = sizeof(R) * (((a.nw + 3) & ~3) + (size_t)M::NSLOT * block);  // weights | slot scratch
    if constexpr (LaneOk<M>::value) {
      if (use_lane<M>(block)) {
        constexpr int TPC = LANE_WARPS * LANE_TRAJ_PER_WARP;
        elbo_fwd_lane_kernel<M, TB><<<(a.N + TPC - 1) / TPC, LANE_WARPS * 32, sizeof(float) * M::NSLOT * TPC, stream>>>(a);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
          set_error("elbo_fwd_lane_kernel launch failed: %s", cudaGetErrorString(e));
          return VH_ERR_CUDA;
        }
        return VH_OK;
      }
    }
    if (use_team<M>(block)) {
      // slot values | partial log-probs | NeuralPrecisions weights
      // VIHDS_FWD_SCRIBE=0|1 (read per call): second warp of the time loop (see elbo_fwd_team_kernel); default on
      const char* sc = getenv("VIHDS_FWD_SCRIBE");
      const bool scribe = !(sc && *sc == '0') && a.T >= 2;
      const size_t tsm = sizeof(R) * ((scribe ? 2 * (size_t)RingVec<R>::slot_elems(M::S) : 0) + (size_t)M::NSLOT * 32 +
                                      2 * FWD_TEAM * 32 + a.nw);
      if (scribe) {
        if (tsm > 48 * 1024)
          cudaFuncSetAttribute(elbo_fwd_team_kernel<M, TB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm);
        launch_maybe_pdl(elbo_fwd_team_kernel<M, TB, true>, dim3((a.N + 31) / 32), dim3(FWD_TEAM * 32), tsm, stream, true, a);
      } else {
        if (tsm > 48 * 1024)
          cudaFuncSetAttribute(elbo_fwd_team_kernel<M, TB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm);
        launch_maybe_pdl(elbo_fwd_team_kernel<M, TB, false>, dim3((a.N + 31) / 32), dim3(FWD_TEAM * 32), tsm, stream, true, a);
      }
    } else {
      if (smem > 48 * 1024)
        cudaFuncSetAttribute(elbo_fwd_kernel<M, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      elbo_fwd_kernel<M, TB><<<grid, block, smem, stream>>>(a);
    }
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
  bool outputs_cleared = false;  // vh_bwd_io.outputs_cleared: no memset nodes; the latency-form kernel launches under the forward's tail
  // warp-specialised kernel: latency-bound launches (the 32-thread-CTA regime), every white-box model.
  // VIHDS_BWD_WS=0|1 overrides (tests / measurements; read per call).
  template <class M>
  static bool use_ws(int block) {
    const char* m = getenv("VIHDS_BWD_WS");
    const int mode = (!m || !*m) ? -1 : atoi(m);
    return mode >= 0 ? mode != 0 : block == 32;
  }
  // VIHDS_BWD_MX=0|1 (read per call): the matrix form of the latency-bound reverse kernel (vh_bwd_mx.cuh) where it exists
  static bool use_mx() {
    const char* m = getenv("VIHDS_BWD_MX");
    return !(m && *m == '0');
  }
  template <class M, class TB>
  void launch_bwd_variant(bool ws, int grid, int block, size_t smem) {
    if constexpr (MxOk<M, TB>::value) {
      // training-shaped calls only: observations present, no upstream gradients on the trajectories themselves
      if (ws && use_mx() && a.obs && !a.g_x_states && !a.g_x_predict && a.T >= 2) {
        const size_t sm = mx_smem_bytes<M, TB>();
        if (sm > 48 * 1024)
          cudaFuncSetAttribute(elbo_bwd_mx_kernel<M, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        launch_maybe_pdl(elbo_bwd_mx_kernel<M, TB>, dim3((a.N + 31) / 32), dim3(MX_WARPS * 32), sm, stream, outputs_cleared, a);
        return;
      }
    }
    if (ws) {
      // hand-off ring | slot scratch | checkpoint staging ring
      // | NeuralPrecisions weights + hand-off ring of the weight-gradient warp
      const size_t ring = sizeof(R) * (2 * WsRing<M, TB>::SLOT + (size_t)M::NSLOT * 64 + (size_t)(WS_PF + 1) * M::S * 32 +
                                       (size_t)((NetInfo<M>::NW + 3) & ~3) + 2 * (M::NIN + 8) * 32 +
                                       (WsSplit<M>::value ? 2 * TB::s * M::S * 32 + M::NC * 32 : 0));
      if (ring > 48 * 1024)
        cudaFuncSetAttribute(elbo_bwd_ws_kernel<M, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring);
      launch_maybe_pdl(elbo_bwd_ws_kernel<M, TB>, dim3((a.N + 31) / 32), dim3(WS_WARPS * 32), ring, stream, outputs_cleared, a);
    } else {
      elbo_bwd_kernel<M, TB><<<grid, block, smem, stream>>>(a);
    }
  }
  template <class M, class TB>
  int run() {
    const int NW = a.nw;
    int block = pick_block(a.N);
    if (NW > 0 && sizeof(R) * (NW * (block + 1) + M::NSLOT * block) > 200 * 1024) block = 64;
    const int grid = (a.N + block - 1) / block;
    const size_t smem = sizeof(R) * (((NW * (block + 1) + 3) & ~3) + (size_t)M::NSLOT * block);  // w | gw | slot scratch
    cudaError_t e;
    const bool ws = use_ws<M>(block) && !(M::DYN && a.n_hidden > 0);  // hidden-layer precision nets: throughput form only
    if (a.iw_b_total > 0 && !ws) {
      set_error("vh_elbo_terms_bwd_iwae: the fused IWAE reduction exists in the latency-form reverse kernel only "
                "(N <= %d, no hidden-layer precision net); use vh_iwae_fwd_bwd + vh_elbo_terms_bwd", 148 * 4 * 32);
      return VH_ERR_UNSUPPORTED;
    }
    if (!ws && smem > 48 * 1024) {
      e = cudaFuncSetAttribute(elbo_bwd_kernel<M, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(smem=%zu) failed: %s", smem, cudaGetErrorString(e));
        return VH_ERR_CUDA;
      }
    }
    bool cost_cleared = a.iw_b_total == 0 || outputs_cleared;
    if (a.d_q_mu && a.P > 0 && !outputs_cleared) {
      const size_t nq = (size_t)a.B * a.P;
      if (a.d_q_prec == a.d_q_mu + nq) {  // adjacent tables (and the cost right behind them): one memset node
        const bool with_cost = !cost_cleared && a.iw_cost == a.d_q_mu + 2 * nq;
        cudaMemsetAsync(a.d_q_mu, 0, sizeof(R) * (2 * nq + (with_cost ? 1 : 0)), stream);
        cost_cleared |= with_cost;
      } else {
        cudaMemsetAsync(a.d_q_mu, 0, sizeof(R) * nq, stream);
        cudaMemsetAsync(a.d_q_prec, 0, sizeof(R) * nq, stream);
      }
    }
    if (!cost_cleared) cudaMemsetAsync(a.iw_cost, 0, sizeof(R), stream);
    if (NW > 0 && !outputs_cleared) cudaMemsetAsync(a.d_weights, 0, sizeof(R) * NW, stream);
    launch_bwd_variant<M, TB>(ws, grid, block, smem);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("elbo_bwd_kernel launch failed: %s", cudaGetErrorString(e));
      return VH_ERR_CUDA;
    }
    return VH_OK;
  }
};

// NeuralPrecisions net of a dynamic-precision white-box model: hidden width and flat weight count of this call
template <class M>
const char* set_precision_net(const vh_problem* p, Call<typename M::real>& a) {
  typedef typename M::real R;
  a.n_hidden = M::DYN ? p->n_hidden : 0;
  if (a.n_hidden < 0 || a.n_hidden > HidPrecNet<R, M::NIN>::MAXH)
    return "NeuralPrecisions hidden width (n_hidden_decoder_precisions) must be in 0..32 for the white-box models";
  a.nw = !M::DYN ? 0 : (a.n_hidden == 0 ? NetInfo<M>::NW : HidPrecNet<R, M::NIN>::num_weights(a.n_hidden));
  return nullptr;
}

template <class M>
int launch_fwd_model(const vh_problem* p, const vh_fwd_io* io, cudaStream_t stream) {
  typedef typename M::real R;
  FwdLauncher<R> f;
  if (const char* err = build_call<R>(p, io, nullptr, f.a)) {
    set_error("vh_elbo_terms_fwd: %s", err);
    return VH_ERR_INVALID;
  }
  f.stream = stream;
  if (const char* err = set_precision_net<M>(p, f.a)) {
    set_error("%s", err);
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
  f.outputs_cleared = io->outputs_cleared != 0;
  if (const char* err = set_precision_net<M>(p, f.a)) {
    set_error("%s", err);
    return VH_ERR_UNSUPPORTED;
  }
  return dispatch_solver<M>(p->solver, f);
}

}  // namespace vh
