// Matrix form of the latency-bound reverse sweep (double-receiver family with constant precisions, midpoint rule):
// the launch `elbo_bwd_mx_kernel`.  Included by vh_launch.cuh (uses its rings, prologue and epilogue helpers).
//
// What bounds the reverse launch at the icml size (7,200 trajectories = 225 warps on 592 schedulers) is the time ONE warp
// needs per time step for the part that is serial in the adjoint state lambda: in elbo_bwd_ws_kernel the consumer warp
// issues ~300 instructions per step at IPC 0.37 (`wait` = fixed-latency dependency stalls: a single warp, in order) --
// 855 cycles per step, 37 of the launch's 62 us (profiles/r02_ws_bwd_icml_*).  Taking instructions that merely accumulate
// away from it changes nothing (measured: VH_WS_SPLIT), only a shorter dependent chain does.
//
// The step adjoint is LINEAR in lambda.  For the midpoint rule x1 = x0 + h f(x0 + h/2 f(x0)):
//     lambda0 = (I + h A + h^2/2 A B)^T lambda1 + e,      A = df/dx at the mid-point state, B = df/dx at x0,
// e = the emission (log-likelihood) cotangent at t0.  A, B and e depend on the checkpoint x0 only -- not on lambda -- so the
// whole matrix N = I + h A + h^2/2 A B is formed AHEAD of the recurrence by producer warps (independent per time step:
// several of them, round-robin), and the recurrence itself shrinks to one sparse 8 x 8 matrix-vector product: 20
// multiply-adds, dependent depth 5.  The Jacobian of models/dr_constant.py:77-112 has 20 non-zeros (diagonal; column 0
// through the growth rate gamma(x0); rows yfp / cfp x columns luxR / lasR through the promoter activities) and the
// pattern is closed under the product, so N has the same 20.
//
// The parameter cotangents (23 accumulators: they need the stage VJPs with the actual lambda) only ACCUMULATE; accumulator
// warps run the unchanged rk_step_adjoint on (x0, stages, kept intermediates) from the producers' ring and lambda1 from the
// consumer's ring, one step each, round-robin, and discard its lambda output.
//
//   warps 0 .. NP-1     producers     checkpoint (cp.async ring) -> stages, kept intermediates, A, B, N, e, d prec
//   warp  NP            consumer      lambda <- N^T lambda + e; publishes every lambda1
//   warps NP+1 ..       accumulators  d constants
// Rings: [slot][item][lane] in shared memory, one mbarrier pair (full / empty) per slot.
#pragma once
#include "vh_mx_math.cuh"

namespace vh {

// warps per role (8 warps at 128 registers keep two CTAs per SM): 4 + 3 and 5 + 2 measure the same (46.7 / 46.2 us at the
// icml size), 3 + 2 in six warps 48 us
#ifndef VH_MX_NP
#define VH_MX_NP 4
#endif
#ifndef VH_MX_NA
#define VH_MX_NA 3
#endif
constexpr int MX_NP = VH_MX_NP, MX_NA = VH_MX_NA;
constexpr int MX_WARPS = MX_NP + 1 + MX_NA;
// two CTAs per SM must stay resident (225 CTAs on 148 SMs at the icml size): register budget per thread, in units of 8;
// warps are allocated in pairs (measured: 7 warps x 144 registers left room for ONE CTA per SM)
constexpr int MX_MAXREG = (65536 / (2 * ((MX_WARPS + 1) / 2 * 2) * 32)) / 8 * 8;
constexpr int MX_D1 = MX_NP + 2;  // slots of the producers' ring (7.4 KB each: two CTAs must fit one SM)
constexpr int MX_DL = MX_D1;       // slots of the consumer's lambda ring (same count: the consumer's loop is unrolled over it)

template <class M, class TB>
struct MxOk {
  static constexpr bool value = false;
};
template <int VER>
struct MxOk<DrModel<float, VER, 0, false>, TabMidpoint<float>> {  // fp32: the fp64 build needs 247 registers x 192 threads
  static constexpr bool value = true;
};

// ---- mbarrier (shared memory, CTA scope) ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count) : "memory");
}
// VH_MX_ARRIVE_ALL 1: every lane arrives for itself (barrier counts are per lane; each lane's release covers its own
// shared-memory accesses -- the form compute-sanitizer's racecheck can follow); 0: __syncwarp() + one arrival by lane 0
// (ordered through the warp barrier and the cumulativity of the release).
#ifndef VH_MX_ARRIVE_ALL
#define VH_MX_ARRIVE_ALL 1
#endif
constexpr int MBAR_PER_WARP = VH_MX_ARRIVE_ALL ? 32 : 1;
__device__ __forceinline__ void mbar_arrive_warp(unsigned long long* b, int lane) {
#if !VH_MX_ARRIVE_ALL
  __syncwarp();
  if (lane == 0)
#endif
  {
    unsigned long long state;
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 %0, [%1];"
                 : "=l"(state)
                 : "r"((unsigned)__cvta_generic_to_shared(b))
                 : "memory");
    (void)state;
  }
}
// VH_MX_SPIN 1: poll with test_wait (returns at once); 0: try_wait (the hardware may suspend the warp before it looks again).
// Measured at the icml size: 56.5 us spinning, 53.0 us with try_wait.
#ifndef VH_MX_SPIN
#define VH_MX_SPIN 0
#endif
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(b);
  unsigned done;
  do {
#if VH_MX_SPIN
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.test_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
#else
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
#endif
  } while (!done);
}

template <class M, class TB>
struct MxRing {
  typedef typename M::real R;
  typedef WsRing<M, TB> Inner;                             // x0, stage derivatives, kept intermediates: the accumulators' share
  static constexpr int NITEM = MXN_PAD + Inner::NITEM;
  static constexpr int SLOT = RingVec<R>::slot_elems(NITEM);  // elements
};

template <class M, class TB>
__global__ void __maxnreg__(MX_MAXREG) elbo_bwd_mx_kernel(const Call<typename M::real> a) {
  typedef typename M::real R;
  typedef WsRing<M, TB> Ring;
  typedef MxRing<M, TB> MR;
  constexpr int S = M::S;
  static_assert(S == 8 && !M::DYN, "matrix-form reverse kernel: dr_constant family");
  static_assert(MX_NP <= MX_D1 && MX_NA <= MX_D1 && MX_DL == MX_D1, "ring bookkeeping below");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw);
  unsigned long long* full1 = bars;                    // [MX_D1] producers -> consumer + accumulator
  unsigned long long* empty1 = bars + MX_D1;           // [MX_D1]
  unsigned long long* fullL = bars + 2 * MX_D1;        // [MX_DL] consumer -> accumulator
  unsigned long long* emptyL = bars + 2 * MX_D1 + MX_DL;
  R* ring = reinterpret_cast<R*>(smem_raw + 8 * (2 * MX_D1 + 2 * MX_DL + 2));
  pdl_wait();     // see elbo_bwd_ws_kernel
  pdl_trigger();
  const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 32 + lane;
  const bool active = n0 < a.N;
  const int n = active ? n0 : a.N - 1;
  const int b = n / a.IW;
  const size_t N = a.N;
  const int T = a.T;
  R* lring = ring + MX_D1 * MR::SLOT;                                 // [MX_DL][S][32]
  R* slots = lring + MX_DL * S * 32;                                  // theta by slot | its cotangent: [NSLOT][64]
  const SlotScratch<R> thv{slots + lane, 64};
  const SlotScratch<R> gloc{slots + 32 + lane, 64};
  R* ckbase = slots + M::NSLOT * 64;                                  // [MX_NP][WS_PF + 1][S][32]
  R* gcsm = ckbase + MX_NP * (WS_PF + 1) * S * 32;                    // [MX_NA][NC][32]
  R* gpsm = gcsm + MX_NA * M::NC * 32;                                // [MX_NP][4][32]
  if (threadIdx.x == 0) {
    for (int i = 0; i < MX_D1; ++i) {
      mbar_init(full1 + i, MBAR_PER_WARP);
      mbar_init(empty1 + i, 2 * MBAR_PER_WARP);  // the consumer and the accumulator of that step
    }
    for (int i = 0; i < MX_DL; ++i) {
      mbar_init(fullL + i, MBAR_PER_WARP);
      mbar_init(emptyL + i, MBAR_PER_WARP);
    }
  }
  // upstream gradients (fused IWAE or handed in) and theta, as elbo_bwd_ws_kernel
  const bool iwae = a.iw_b_total > 0;
  const R gup = iwae ? iwae_upstream_in_kernel(a, n, active, ring) : R(0);
  const R glq = iwae ? -gup : ((a.g_logq_theta && active) ? a.g_logq_theta[n] : R(0));
  const R glp = iwae ? gup : ((a.g_logp_theta && active) ? a.g_logp_theta[n] : R(0));
  for (int s = role; s < M::NSLOT; s += MX_WARPS) {
    const int src = a.slot_src[s];
    if (src < 0) thv[s] = src != VH_SLOT_UNUSED ? a.extra[(size_t)(-1 - src) * N + n] : R(0);
  }
#pragma unroll 3
  for (int k = role; k < a.P; k += MX_WARPS) {
    R lq = R(0), lp = R(0);
    const R v = a.theta_in ? a.theta_in[(size_t)k * N + n] : sample_column(a, n, b, k, lq, lp, false);
    const int s = a.col_slot[k];
    if (s >= 0) thv[s] = v;
  }
  __syncthreads();  // theta complete, barriers initialised
  Rhs<M> f;
  f.w = nullptr;
  f.nh = 0;
  R prec[4], gl[4];
  {
    // the RHS constants (six powf in the Hill fractions) are formed by ONE warp and handed to the others through shared
    // memory: computed by all eight warps they were 8 % of the launch's instructions, all at the same moment
    static_assert(sizeof(typename M::Consts) / sizeof(R) <= MX_NA * M::NC, "constants are staged in the accumulators' area");
    constexpr int NCW = sizeof(typename M::Consts) / sizeof(R);
    if (role == 0) {
      R th[M::NSLOT];
      R tc[3];
#pragma unroll
      for (int s = 0; s < M::NSLOT; ++s) th[s] = M::uses(s) ? thv[s] : R(0);
      M::treatments(a.treatments + (size_t)b * a.C, tc);
      M::setup(th, tc, f.c);
      const R* cv = reinterpret_cast<const R*>(&f.c);
#pragma unroll
      for (int i = 0; i < NCW; ++i) gcsm[i * 32 + lane] = cv[i];
    }
    __syncthreads();
    if (role != 0) {
      R* cv = reinterpret_cast<R*>(&f.c);
#pragma unroll
      for (int i = 0; i < NCW; ++i) cv[i] = gcsm[i * 32 + lane];
    }
    __syncthreads();  // gcsm is the accumulators' again
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      prec[o] = thv[S_prec_x + o];
      gl[o] = iwae ? gup : ((a.g_logp_species && active) ? a.g_logp_species[(size_t)n * 4 + o] : R(0));
    }
  }
  const size_t slab = (size_t)S * N;
  const R* obs = a.obs + (size_t)b * 4 * T;
  const int nit = T - 1;  // time steps; iteration it handles step k = T - 2 - it
  // emission cotangent at state x against the observations of time index k: e[0..3] (cotangent of x0, x1, x2 = x4, x3 = x5)
  // and the precision cotangents; vihds/ode.py:84-93 + the Gaussian log-likelihood (training.py:150-160)
  R glp_o[4], hgl[4], iprec[4];  // gl * prec, gl / 2, 1 / prec: per-trajectory constants of the emission cotangent
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    glp_o[o] = gl[o] * prec[o];
    hgl[o] = gl[o] * R(0.5);
    iprec[o] = R(1) / prec[o];
  }
  auto emission = [&](const R* x, const R* ob, R* e, R* gprec) {
    R xp[4];
    M::observe(x, xp);
    R gxp[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const R d = xp[o] - ob[o];
      gxp[o] = -(glp_o[o] * d);
      gprec[o] += hgl[o] * (iprec[o] - d * d);
    }
    e[0] = gxp[0] + gxp[1] * x[1] + gxp[2] * (x[2] + x[4]) + gxp[3] * (x[3] + x[5]);
    e[1] = gxp[1] * x[0];
    e[2] = gxp[2] * x[0];
    e[3] = gxp[3] * x[0];
  };
  if (role < MX_NP) {
    // ---------------- producers ----------------
    R* ck = ckbase + role * (WS_PF + 1) * S * 32 + lane;
    R gprec[4] = {R(0), R(0), R(0), R(0)};
    int itw = role, sw = 0, sr = 0;  // iteration of the next checkpoint copy, staging slots of the next copy / read
    auto issue = [&]() {
      if (itw < nit) {
        const R* xs = a.x_states + (size_t)(T - 2 - itw) * slab + n;
#pragma unroll
        for (int q = 0; q < S; ++q) cp_async_elem(ck + (sw * S + q) * 32, xs + (size_t)q * N);
      }
      cp_async_commit();
      itw += MX_NP;
      sw = sw == WS_PF ? 0 : sw + 1;
    };
#pragma unroll
    for (int d = 0; d < WS_PF; ++d) issue();
    // grid times and observations of the step after this one are fetched one (own) iteration ahead: as plain loads at
    // their point of use they held ~45 % of the producers' stall samples (long scoreboard)
    R t0 = R(0), t1 = R(0), ob[4] = {R(0), R(0), R(0), R(0)};
    if (role < nit) {
      const int k = T - 2 - role;
      t0 = a.times[k];
      t1 = a.times[k + 1];
#pragma unroll
      for (int o = 0; o < 4; ++o) ob[o] = obs[o * T + k];
    }
    int slot = role % MX_D1, use = role / MX_D1;
    for (int it = role; it < nit; it += MX_NP) {
      const int k = T - 2 - it;
      issue();
      const int kn = k >= MX_NP ? k - MX_NP : 0;
      const R t0n = ld_early(a.times + kn), t1n = ld_early(a.times + kn + 1);
      R obn[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) obn[o] = ld_early(obs + o * T + kn);
      const R h = t1 - t0;
      cp_async_wait<WS_PF>();
      R x[S];
#pragma unroll
      for (int q = 0; q < S; ++q) x[q] = ck[(sr * S + q) * 32];
      sr = sr == WS_PF ? 0 : sr + 1;
      typename Ring::SD sd;
      rk_stages_forward<Rhs<M>, TB>(f, t0, t1, h, x, sd);
      R Xm[S];
#pragma unroll
      for (int q = 0; q < S; ++q) Xm[q] = x[q] + (h * TB::a(1, 0)) * sd.k[0][q];
      R Nv[MXN_ITEMS];
      mx_step_matrix<M>(x, Xm, f.c, sd.kept[0].m, sd.kept[1].m, h, TB::a(1, 0), Nv);
      emission(x, ob, Nv + MXN_E, gprec);
      if (use > 0) mbar_wait(empty1 + slot, (use - 1) & 1);  // both readers have released the slot
      {
        R buf[MR::NITEM];
#pragma unroll
        for (int i = 0; i < MXN_PAD; ++i) buf[i] = i < MXN_ITEMS ? Nv[i] : R(0);
        R(&inner)[Ring::NITEM] = *reinterpret_cast<R(*)[Ring::NITEM]>(buf + MXN_PAD);
        Ring::pack(x, sd, inner);
        RingVec<R>::store(ring + slot * MR::SLOT, lane, buf);
      }
      mbar_arrive_warp(full1 + slot, lane);
      t0 = t0n;
      t1 = t1n;
#pragma unroll
      for (int o = 0; o < 4; ++o) ob[o] = obn[o];
      slot += MX_NP;
      if (slot >= MX_D1) {
        slot -= MX_D1;
        ++use;
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) gpsm[(role * 4 + o) * 32 + lane] = gprec[o];
  } else if (role == MX_NP) {
    // ---------------- consumer: lambda <- N^T lambda + e ----------------
    R lam[S];
    R gprec[4] = {R(0), R(0), R(0), R(0)};
    {
      R x[S], e[4], ob[4];
#pragma unroll
      for (int q = 0; q < S; ++q) x[q] = a.x_states[(size_t)(T - 1) * slab + (size_t)q * N + n];
#pragma unroll
      for (int o = 0; o < 4; ++o) ob[o] = obs[o * T + T - 1];
      emission(x, ob, e, gprec);
      lam[0] = e[0]; lam[1] = e[1]; lam[2] = e[2]; lam[3] = e[3]; lam[4] = e[2]; lam[5] = e[3];
      lam[6] = R(0); lam[7] = R(0);
    }
    // unrolled over the ring: slot numbers and shared-memory addresses are immediates, one backward branch per MX_D1
    // steps (the branch and the index arithmetic were ~30 % of this warp's samples)
    for (int it0 = 0, use = 0; it0 < nit; it0 += MX_D1, ++use)
#pragma unroll
    for (int slot = 0; slot < MX_D1; ++slot) {
      const int it = it0 + slot;
      if (it >= nit) break;
      const int ls = slot, luse = use;
      // lambda1 of this step for its accumulator
      if (luse > 0) mbar_wait(emptyL + ls, (luse - 1) & 1);
      RingVec<R>::store(lring + ls * S * 32, lane, lam);
      mbar_arrive_warp(fullL + ls, lane);
      mbar_wait(full1 + slot, use & 1);
      R Nv[MXN_ITEMS];
      RingVec<R>::load(ring + slot * MR::SLOT, lane, 0, Nv);
      mbar_arrive_warp(empty1 + slot, lane);
      mx_apply(Nv, lam);
    }
    // chain rule back to theta (as elbo_bwd_ws_kernel); the constants' cotangents come from the accumulators
    named_bar_sync_n<MX_WARPS * 32>(1);  // producers' d prec and accumulators' d constants are in shared memory
    typename M::Consts gc;
#pragma unroll
    for (int i = 0; i < M::NC; ++i) {
      R v = R(0);
#pragma unroll
      for (int q = 0; q < MX_NA; ++q) v += gcsm[(q * M::NC + i) * 32 + lane];
      gc.v[i] = v;
    }
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int p = 0; p < MX_NP; ++p) gprec[o] += gpsm[(p * 4 + o) * 32 + lane];
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
      M::setup_vjp(th, tc, f.c, gc, gth, 4 | 1);  // the LasR Hill fraction's share: first accumulator warp, below
#pragma unroll
      for (int o = 0; o < 4; ++o) gth[S_prec_x + o] += gprec[o];
    }
#pragma unroll
    for (int s = 0; s < M::NSLOT; ++s) gloc[s] = M::uses(s) ? gth[s] : R(0);
  } else {
    // ---------------- accumulators: parameter cotangents of the stage VJPs ----------------
    const int q = role - MX_NP - 1;
    typename M::Consts gc;
#pragma unroll
    for (int i = 0; i < M::NC; ++i) gc.v[i] = R(0);
    NoGW<R> nogw;
    R t0 = R(0), t1 = R(0);
    if (q < nit) {
      t0 = a.times[T - 2 - q];
      t1 = a.times[T - 1 - q];
    }
    int slot = q % MX_D1, use = q / MX_D1;
    for (int it = q; it < nit; it += MX_NA) {
      const int k = T - 2 - it;
      const int ls = slot, luse = use;
      const int kn = k >= MX_NA ? k - MX_NA : 0;
      const R t0n = ld_early(a.times + kn), t1n = ld_early(a.times + kn + 1);
      R x[S], lam[S];
      typename Ring::SD sd;
      mbar_wait(full1 + slot, use & 1);
      {
        R buf[Ring::NITEM];
        RingVec<R>::load(ring + slot * MR::SLOT, lane, MXN_PAD, buf);
        Ring::unpack(buf, x, sd);
      }
      mbar_arrive_warp(empty1 + slot, lane);
      mbar_wait(fullL + ls, luse & 1);
      RingVec<R>::load(lring + ls * S * 32, lane, 0, lam);
      mbar_arrive_warp(emptyL + ls, lane);
      rk_step_adjoint<Rhs<M>, TB>(f, t0, t1, t1 - t0, x, sd, lam, gc, nogw);  // its lambda output is not used
      t0 = t0n;
      t1 = t1n;
      slot += MX_NA;
      if (slot >= MX_D1) {
        slot -= MX_D1;
        ++use;
      }
    }
#pragma unroll
    for (int i = 0; i < M::NC; ++i) gcsm[(q * M::NC + i) * 32 + lane] = gc.v[i];
  }
  // this warp's columns of the chain-rule epilogue: their inputs are fetched NOW, while the consumer warp is still busy
  // with the chain rule of the constants (nine dependent-latency loads per column otherwise sit behind the barrier)
  constexpr int CPW = 6;  // columns per warp held in registers (P <= CPW * (MX_WARPS - 1) = 42; more: plain path)
  const bool pre = a.P <= CPW * (MX_WARPS - 1);
  ColumnIn<R> cin[CPW];
  int cslot[CPW];
  const int cw = role < MX_NP ? role : role - 1;  // the consumer takes no columns in this form
  if (pre && role != MX_NP) {
#pragma unroll
    for (int j = 0; j < CPW; ++j) {
      const int k = cw + j * (MX_WARPS - 1);
      if (k < a.P) {
        column_load(a, n, b, k, cin[j]);
        cslot[j] = a.col_slot[k];
      }
    }
  }
  if (role != MX_NP) named_bar_sync_n<MX_WARPS * 32>(1);
  // second half of the constants' chain rule (the LasR Hill fraction: six powf, three logf) on the first accumulator warp,
  // next to the consumer's half; its slot cotangents go to gloc2 and are added where the columns read them
  R* gloc2 = ring + lane;  // [NSLOT][32]: the rings are idle from here on
  if (role == MX_NP + 1) {
    typename M::Consts gc;
#pragma unroll
    for (int i = 0; i < M::NC; ++i) {
      R v = R(0);
#pragma unroll
      for (int q = 0; q < MX_NA; ++q) v += gcsm[(q * M::NC + i) * 32 + lane];
      gc.v[i] = v;
    }
    R gth[M::NSLOT], th[M::NSLOT], tc[3];
#pragma unroll
    for (int s = 0; s < M::NSLOT; ++s) {
      gth[s] = R(0);
      th[s] = M::uses(s) ? thv[s] : R(0);
    }
    M::treatments(a.treatments + (size_t)b * a.C, tc);
    M::setup_vjp(th, tc, f.c, gc, gth, 2);
#pragma unroll
    for (int s = 0; s < M::NSLOT; ++s) gloc2[s * 32] = gth[s];
  }
  __syncthreads();  // gloc, gloc2 complete
  WarpSegRed<R> red(a.d_q_mu, a.d_q_prec, a.P, b, active);
  if (pre) {
#pragma unroll
    for (int j = 0; j < CPW; ++j) {
      const int k = cw + j * (MX_WARPS - 1);
      if (role != MX_NP && k < a.P) {
        R dmu = R(0), dprec = R(0);
        if (active) column_vjp_from(cin[j], cslot[j] >= 0 ? gloc[cslot[j]] + gloc2[cslot[j] * 32] : R(0), glq, glp, dmu, dprec);
        red(b, k, dmu, dprec, active);
      }
    }
  } else {
#pragma unroll 3
    for (int k = role; k < a.P; k += MX_WARPS) {
      const int s = a.col_slot[k];
      R dmu = R(0), dprec = R(0);
      if (active) column_vjp(a, n, b, k, s >= 0 ? gloc[s] + gloc2[s * 32] : R(0), glq, glp, dmu, dprec);
      red(b, k, dmu, dprec, active);
    }
  }
  if (a.d_extra && active) {
    for (int s = role; s < M::NSLOT; s += MX_WARPS) {
      const int src = a.slot_src[s];
      if (src < 0 && src != VH_SLOT_UNUSED) a.d_extra[(size_t)(-1 - src) * N + n] = gloc[s] + gloc2[s * 32];
    }
  }
}

template <class M, class TB>
inline size_t mx_smem_bytes() {
  typedef typename M::real R;
  return 8 * (2 * MX_D1 + 2 * MX_DL + 2) +
         sizeof(R) * ((size_t)MX_D1 * MxRing<M, TB>::SLOT + (size_t)MX_DL * M::S * 32 + (size_t)M::NSLOT * 64 +
                      (size_t)MX_NP * (WS_PF + 1) * M::S * 32 + (size_t)MX_NA * M::NC * 32 + (size_t)MX_NP * 4 * 32);
}

}  // namespace vh
