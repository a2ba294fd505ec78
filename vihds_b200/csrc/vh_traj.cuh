// One trajectory (individual b, importance sample i) of the hot path, forward and reverse.
//
//   forward : sample+clip theta (vihds/distributions.py:119-142, :76-85), log q / log p (:64-74, :338-345, :373-375),
//             RHS constants, fixed-step solve (vihds/solvers.py:9-41; torchdiffeq 0.1 fixed-grid schemes called at
//             vihds/ode.py:80-81), observe (vihds/ode.py:84-93), Gaussian log-likelihood summed over time
//             (vihds/training.py:24-44).
//   reverse : discrete adjoint of exactly that computation (what autograd's replay yields, vihds/training.py:334),
//             re-reading the state trace the forward pass wrote.
//
// VH_HD throughout: the device kernels (vh_kernels.cu) call these with n = blockIdx.x*blockDim.x+threadIdx.x; the
// host-side math check (tests/hostcheck) loops over n on the CPU.
#pragma once
#include "../../include/vihds_b200.h"
#include "vh_math.cuh"
#include "vh_models.cuh"

namespace vh {

// ---------------------------------------------------------------------------------------------------------------
// Butcher tableaux of the supported fixed-step schemes
// ---------------------------------------------------------------------------------------------------------------
template <typename R>
struct TabEuler {
  static constexpr int s = 1;
  static constexpr bool const_h = false, end_is_t1 = false;
  VH_HD static constexpr R a(int, int) { return R(0); }
  VH_HD static constexpr R b(int) { return R(1); }
  VH_HD static constexpr R c(int) { return R(0); }
};
template <typename R>
struct TabMidpoint {
  static constexpr int s = 2;
  static constexpr bool const_h = false, end_is_t1 = false;
  VH_HD static constexpr R a(int i, int j) { return (i == 1 && j == 0) ? R(0.5) : R(0); }
  VH_HD static constexpr R b(int i) { return i == 1 ? R(1) : R(0); }
  VH_HD static constexpr R c(int i) { return i == 1 ? R(0.5) : R(0); }
};
template <typename R>
struct TabRK4_38 {  // torchdiffeq 0.1 "rk4" (rk4_alt_step_func): the 3/8 rule
  static constexpr int s = 4;
  static constexpr bool const_h = false, end_is_t1 = false;
  VH_HD static constexpr R a(int i, int j) {
    return i == 1 ? (j == 0 ? R(1) / R(3) : R(0))
                  : i == 2 ? (j == 0 ? R(-1) / R(3) : (j == 1 ? R(1) : R(0)))
                           : i == 3 ? (j == 0 ? R(1) : (j == 1 ? R(-1) : (j == 2 ? R(1) : R(0)))) : R(0);
  }
  VH_HD static constexpr R b(int i) { return (i == 0 || i == 3) ? R(0.125) : R(0.375); }
  VH_HD static constexpr R c(int i) { return i == 1 ? R(1) / R(3) : (i == 2 ? R(2) / R(3) : (i == 3 ? R(1) : R(0))); }
};
template <typename R, bool CONST_H>
struct TabHeun {  // vihds/solvers.py: modeuler (CONST_H: h = times[1]-times[0] for every step) / modeulerwhile
  static constexpr int s = 2;
  static constexpr bool const_h = CONST_H, end_is_t1 = true;
  VH_HD static constexpr R a(int i, int j) { return (i == 1 && j == 0) ? R(1) : R(0); }
  VH_HD static constexpr R b(int) { return R(0.5); }
  VH_HD static constexpr R c(int i) { return i == 1 ? R(1) : R(0); }
};

// ---------------------------------------------------------------------------------------------------------------
// typed view of one call (device pointers; built by the launcher from vh_problem + vh_fwd_io / vh_bwd_io)
// ---------------------------------------------------------------------------------------------------------------
template <typename R>
struct Call {
  int B, IW, N, T, P, C, D, E;
  int slot_src[VH_MAX_SLOTS];
  int bb_nlat, bb_ny, bb_noff;   // dr_blackbox: latent-parameter count (n_z + n_x + n_y), its conditioned tail n_y, offset rows
  R bb_init_latent, bb_init_prec;
  int n_hidden, nw;              // NeuralPrecisions of the white-box models: hidden width (0 = none), flat weight count
  int n_free;                    // theta columns no model slot reads (they still carry log-prob terms)
  int free_cols[VH_MAX_SLOTS];
  int col_slot[VH_MAX_SLOTS];    // inverse of slot_src: the model slot column k feeds, or -1
  const R *times, *u, *q_mu, *q_prec, *p_mu, *p_prec, *clip_lo, *clip_hi, *extra, *treatments, *dev_1hot, *obs, *weights;
  const int* kind;
  R *theta, *x_states, *x_predict, *logp_species, *logp_theta, *logq_theta;
  const R* theta_in;  // reverse only: the forward call's theta planes [P][N] (NULL: re-sample from u)
  // reverse only
  const R *g_logp_species, *g_logp_theta, *g_logq_theta, *g_theta, *g_x_states, *g_x_predict;
  R *d_q_mu, *d_q_prec, *d_extra, *d_weights;
  // reverse with the IWAE reduction fused in (vh_elbo_terms_bwd_iwae): the upstream gradients are derived in the kernel
  // from the forward call's per-sample terms (logp_species / logp_theta / logq_theta above, read as inputs)
  int iw_b_total;  // denominator of the batch mean; 0: upstream gradients come from g_logp_* as usual
  R* iw_cost;      // [1], accumulated: -(logsumexp_i log_w - log IW) / b_total per individual
};

// weight-gradient accumulator handles -------------------------------------------------------------------------
// outer<NIN>(a, gzp, gzd): one NeuralPrecisions evaluation's contribution to the flat weight gradient
//   Wp[o][j] += gzp[o] a[j], bp[o] += gzp[o], Wd[o][j] += gzd[o] a[j], bd[o] += gzd[o]      (layout: LinPrecNet)
template <int NIN, typename GW, typename R>
VH_HD void prec_outer_adds(GW& gw, const R* a, const R* gzp, const R* gzd) {
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    gw.add(4 * NIN + o, gzp[o]);
    gw.add((4 * NIN + 4) + 4 * NIN + o, gzd[o]);
#pragma unroll
    for (int j = 0; j < NIN; ++j) {
      gw.add(o * NIN + j, gzp[o] * a[j]);
      gw.add((4 * NIN + 4) + o * NIN + j, gzd[o] * a[j]);
    }
  }
}
template <typename R>
struct NoGW {
  VH_HD void add(int, R) const {}
  template <int NIN>
  VH_HD void outer(const R*, const R*, const R*) const {}
};
template <typename R>
struct StridedGW {  // element k of this thread's accumulators lives at base[k * stride]
  R* base;
  int stride;
  VH_HD void add(int k, R v) const { base[k * stride] += v; }
  template <int NIN>
  VH_HD void outer(const R* a, const R* gzp, const R* gzd) const {
    prec_outer_adds<NIN>(*this, a, gzp, gzd);
  }
};

// full right-hand side over the ODE state (species + dynamic-precision states) ---------------------------------
// what a stage evaluation hands to its VJP: the species intermediates and, for dynamic precisions, the activations of the
// NeuralPrecisions net (no hidden layer).  Plain R members only: the warp-specialised kernel ships it as an R array.
template <class M, bool DYN>
struct RhsKept {
  typename M::Mid m;
  typename LinPrecNet<typename M::real, M::NIN>::Kept p;
};
template <class M>
struct RhsKept<M, false> {
  typename M::Mid m;
};

template <class M>
struct Rhs {
  typedef typename M::real R;
  typedef R real;
  typedef typename M::Mid Mid;
  typedef RhsKept<M, M::DYN> Kept;
  typedef typename M::Consts Grad;   // cotangent accumulator of the per-trajectory constants
  static constexpr int S = M::S;
  typename M::Consts c;
  const R* w;  // NeuralPrecisions weights (shared memory on the device), unused for constant precisions
  int nh;      // hidden width of the NeuralPrecisions net (0: no hidden layer)

  VH_HD void eval(R t, const R* x, R* dx) const {
    Mid m;
    M::mid(t, x, c, m);
    M::rhs_from(x, c, m, dx);
    net_rhs(t, x, dx);
  }
  VH_HD void net_rhs(R t, const R* x, R* dx) const {
    if (M::DYN) {
      if (nh == 0)
        LinPrecNet<R, M::NIN>::rhs(t, x, x + M::NS, w, dx + M::NS);
      else
        HidPrecNet<R, M::NIN>::rhs(t, x, x + M::NS, w, nh, dx + M::NS);
    }
  }
  // same, handing back the intermediates so that the reverse sweep need not recompute them
  VH_HD void eval_keep(R t, const R* x, R* dx, Kept& k) const {
    M::mid(t, x, c, k.m);
    M::rhs_from(x, c, k.m, dx);
    keep_net(t, x, dx + M::NS, k);
  }
  // intermediates only (last stage of a step: its derivative is not needed to rebuild any stage state)
  VH_HD void keep_only(R t, const R* x, Kept& k) const {
    M::mid(t, x, c, k.m);
    keep_net(t, x, (R*)nullptr, k);
  }
  VH_HD void keep_net(R t, const R* x, R* dv, RhsKept<M, true>& k) const {
    if (nh == 0)
      LinPrecNet<R, M::NIN>::rhs_keep(t, x, x + M::NS, w, dv, k.p);
    else if (dv)
      HidPrecNet<R, M::NIN>::rhs(t, x, x + M::NS, w, nh, dv);  // hidden-layer nets recompute in their VJP
  }
  VH_HD void keep_net(R, const R*, R*, RhsKept<M, false>&) const {}
  template <typename GW>
  VH_HD void vjp_kept(R t, const R* x, const Kept& k, const R* g, R* gx, typename M::Consts& gc, GW& gw) const {
    M::rhs_vjp_from(x, c, k.m, g, gx, gc);
    net_vjp(t, x, k, g, gx, gw);
  }
  template <typename GW>
  VH_HD void net_vjp(R t, const R* x, const RhsKept<M, true>& k, const R* g, R* gx, GW& gw) const {
    if (nh == 0)
      LinPrecNet<R, M::NIN>::rhs_vjp_kept(k.p, x + M::NS, w, g + M::NS, gx, gx + M::NS, gw);
    else
      HidPrecNet<R, M::NIN>::rhs_vjp(t, x, x + M::NS, w, nh, g + M::NS, gx, gx + M::NS, gw);
  }
  template <typename GW>
  VH_HD void net_vjp(R, const R*, const RhsKept<M, false>&, const R*, R*, GW&) const {}
  template <typename GW>
  VH_HD void vjp(R t, const R* x, const R* g, R* gx, typename M::Consts& gc, GW& gw) const {
    Kept k;
    keep_only(t, x, k);
    vjp_kept(t, x, k, g, gx, gc, gw);
  }
};

template <class TB, typename R>
VH_HD R stage_time(int i, R t0, R t1) {
  if (TB::c(i) == R(0)) return t0;
  if (TB::end_is_t1 && TB::c(i) == R(1)) return t1;
  return t0 + (t1 - t0) * TB::c(i);
}

// x <- one step of the scheme from t0 to t1 with step size h
template <class F, class TB>
VH_HD void rk_step(const F& f, typename F::real t0, typename F::real t1, typename F::real h, typename F::real* x) {
  typedef typename F::real R;
  constexpr int S = F::S;
  R k[TB::s][S];
#pragma unroll
  for (int i = 0; i < TB::s; ++i) {
    R X[S];
#pragma unroll
    for (int q = 0; q < S; ++q) X[q] = x[q];
#pragma unroll
    for (int j = 0; j < i; ++j)
      if (TB::a(i, j) != R(0)) {
        const R ha = h * TB::a(i, j);
#pragma unroll
        for (int q = 0; q < S; ++q) X[q] += ha * k[j][q];
      }
    f.eval(stage_time<TB, R>(i, t0, t1), X, k[i]);
  }
#pragma unroll
  for (int i = 0; i < TB::s; ++i)
    if (TB::b(i) != R(0)) {
      const R hb = h * TB::b(i);
#pragma unroll
      for (int q = 0; q < S; ++q) x[q] += hb * k[i][q];
    }
}

// The reverse of one step has two phases.  Phase 1 (rk_stages_forward) re-evaluates the stages from the checkpoint
// x(t0): it depends on the checkpoint only, NOT on the adjoint state.  Phase 2 (rk_step_adjoint) is linear in lam and
// uses only the kept intermediates.  rk_step_vjp runs them back to back in one thread; the warp-specialised reverse
// kernel (vh_launch.cuh, elbo_bwd_ws_kernel) gives phase 1 to a producer warp that runs one step ahead of the
// consumer warp doing phase 2.
template <class F, class TB>
struct StageData {
  typedef typename F::real R;
  static constexpr int s = TB::s;
  static constexpr int nk = s > 1 ? s - 1 : 1;
  R k[nk][F::S];                 // stage derivatives 0 .. s-2 (the last one is never needed to rebuild a stage state)
  typename F::Kept kept[TB::s];  // RHS intermediates of every stage, reused by its vjp
};

template <class F, class TB>
VH_HD void rk_stages_forward(const F& f, typename F::real t0, typename F::real t1, typename F::real h,
                             const typename F::real* x, StageData<F, TB>& sd) {
  typedef typename F::real R;
  constexpr int S = F::S;
  constexpr int s = TB::s;
#pragma unroll
  for (int i = 0; i < s; ++i) {
    R X[S];
#pragma unroll
    for (int q = 0; q < S; ++q) X[q] = x[q];
#pragma unroll
    for (int j = 0; j < i; ++j)
      if (TB::a(i, j) != R(0)) {
        const R ha = h * TB::a(i, j);
#pragma unroll
        for (int q = 0; q < S; ++q) X[q] += ha * sd.k[j][q];
      }
    if (i + 1 < s)
      f.eval_keep(stage_time<TB, R>(i, t0, t1), X, sd.k[i], sd.kept[i]);
    else
      f.keep_only(stage_time<TB, R>(i, t0, t1), X, sd.kept[i]);
  }
}

// lam: in = dL/dx(t1), out = dL/dx(t0);  x = state at t0;  gc/gw accumulate parameter cotangents
template <class F, class TB, typename GW>
VH_HD void rk_step_adjoint(const F& f, typename F::real t0, typename F::real t1, typename F::real h,
                           const typename F::real* x, const StageData<F, TB>& sd, typename F::real* lam,
                           typename F::Grad& gc, GW& gw) {
  typedef typename F::real R;
  constexpr int S = F::S;
  constexpr int s = TB::s;
  R gk[s][S];
#pragma unroll
  for (int i = 0; i < s; ++i)
#pragma unroll
    for (int q = 0; q < S; ++q) gk[i][q] = (h * TB::b(i)) * lam[q];
#pragma unroll
  for (int i = s - 1; i >= 0; --i) {
    R X[S], gX[S];
#pragma unroll
    for (int q = 0; q < S; ++q) {
      X[q] = x[q];
      gX[q] = R(0);
    }
#pragma unroll
    for (int j = 0; j < i; ++j)
      if (TB::a(i, j) != R(0)) {
        const R ha = h * TB::a(i, j);
#pragma unroll
        for (int q = 0; q < S; ++q) X[q] += ha * sd.k[j][q];
      }
    f.vjp_kept(stage_time<TB, R>(i, t0, t1), X, sd.kept[i], gk[i], gX, gc, gw);
#pragma unroll
    for (int q = 0; q < S; ++q) lam[q] += gX[q];
#pragma unroll
    for (int j = 0; j < i; ++j)
      if (TB::a(i, j) != R(0)) {
        const R ha = h * TB::a(i, j);
#pragma unroll
        for (int q = 0; q < S; ++q) gk[j][q] += ha * gX[q];
      }
  }
}

// The adjoint of one step SPLIT over two executors (warp-specialised reverse kernel, constant-precision models): the part
// that is serial in lam (state cotangents only) and the parameter cotangents, which only ACCUMULATE.  Both call the same
// vjp_kept; each discards one half of its results, and with everything inlined the compiler drops the instructions that
// feed only the discarded half (no per-model code).  rk_step_adjoint_x hands the cotangent of every stage derivative,
// gk[i] -- final at the moment stage i's VJP runs -- to `pub(i, gk[i])`; rk_step_adjoint_c takes them back.
template <class F, class TB, typename Pub>
VH_HD void rk_step_adjoint_x(const F& f, typename F::real t0, typename F::real t1, typename F::real h,
                             const typename F::real* x, const StageData<F, TB>& sd, typename F::real* lam, Pub& pub) {
  typedef typename F::real R;
  constexpr int S = F::S;
  constexpr int s = TB::s;
  NoGW<R> nosink;
  typename F::Grad unused;
#pragma unroll
  for (int i = 0; i < (int)(sizeof(unused.v) / sizeof(R)); ++i) unused.v[i] = R(0);
  R gk[s][S];
#pragma unroll
  for (int i = 0; i < s; ++i)
#pragma unroll
    for (int q = 0; q < S; ++q) gk[i][q] = (h * TB::b(i)) * lam[q];
#pragma unroll
  for (int i = s - 1; i >= 0; --i) {
    R X[S], gX[S];
#pragma unroll
    for (int q = 0; q < S; ++q) {
      X[q] = x[q];
      gX[q] = R(0);
    }
#pragma unroll
    for (int j = 0; j < i; ++j)
      if (TB::a(i, j) != R(0)) {
        const R ha = h * TB::a(i, j);
#pragma unroll
        for (int q = 0; q < S; ++q) X[q] += ha * sd.k[j][q];
      }
    pub(i, gk[i]);
    f.vjp_kept(stage_time<TB, R>(i, t0, t1), X, sd.kept[i], gk[i], gX, unused, nosink);
#pragma unroll
    for (int q = 0; q < S; ++q) lam[q] += gX[q];
#pragma unroll
    for (int j = 0; j < i; ++j)
      if (TB::a(i, j) != R(0)) {
        const R ha = h * TB::a(i, j);
#pragma unroll
        for (int q = 0; q < S; ++q) gk[j][q] += ha * gX[q];
      }
  }
}

template <class F, class TB>
VH_HD void rk_step_adjoint_c(const F& f, typename F::real t0, typename F::real t1, typename F::real h,
                             const typename F::real* x, const StageData<F, TB>& sd,
                             const typename F::real (*gk)[F::S], typename F::Grad& gc) {
  typedef typename F::real R;
  constexpr int S = F::S;
  constexpr int s = TB::s;
  NoGW<R> nosink;
#pragma unroll
  for (int i = s - 1; i >= 0; --i) {
    R X[S], unused[S];
#pragma unroll
    for (int q = 0; q < S; ++q) {
      X[q] = x[q];
      unused[q] = R(0);
    }
#pragma unroll
    for (int j = 0; j < i; ++j)
      if (TB::a(i, j) != R(0)) {
        const R ha = h * TB::a(i, j);
#pragma unroll
        for (int q = 0; q < S; ++q) X[q] += ha * sd.k[j][q];
      }
    f.vjp_kept(stage_time<TB, R>(i, t0, t1), X, sd.kept[i], gk[i], unused, gc, nosink);
  }
}

// lam: in = dL/dx(t1), out = dL/dx(t0);  x = state at t0;  gc/gw accumulate parameter cotangents
template <class F, class TB, typename GW>
VH_HD void rk_step_vjp(const F& f, typename F::real t0, typename F::real t1, typename F::real h,
                       const typename F::real* x, typename F::real* lam, typename F::Grad& gc, GW& gw) {
  typedef typename F::real R;
  constexpr int S = F::S;
  constexpr int s = TB::s;
  constexpr int nk = s > 1 ? s - 1 : 1;
  R k[nk][S];               // the last stage derivative is never needed to rebuild a stage state
  typename F::Kept kept[nk];  // intermediates of the stages that had to be re-evaluated: reused by their vjp
#pragma unroll
  for (int i = 0; i + 1 < s; ++i) {
    R X[S];
#pragma unroll
    for (int q = 0; q < S; ++q) X[q] = x[q];
#pragma unroll
    for (int j = 0; j < i; ++j)
      if (TB::a(i, j) != R(0)) {
        const R ha = h * TB::a(i, j);
#pragma unroll
        for (int q = 0; q < S; ++q) X[q] += ha * k[j][q];
      }
    f.eval_keep(stage_time<TB, R>(i, t0, t1), X, k[i], kept[i]);
  }
  R gk[s][S];
#pragma unroll
  for (int i = 0; i < s; ++i)
#pragma unroll
    for (int q = 0; q < S; ++q) gk[i][q] = (h * TB::b(i)) * lam[q];
#pragma unroll
  for (int i = s - 1; i >= 0; --i) {
    R X[S], gX[S];
#pragma unroll
    for (int q = 0; q < S; ++q) {
      X[q] = x[q];
      gX[q] = R(0);
    }
#pragma unroll
    for (int j = 0; j < i; ++j)
      if (TB::a(i, j) != R(0)) {
        const R ha = h * TB::a(i, j);
#pragma unroll
        for (int q = 0; q < S; ++q) X[q] += ha * k[j][q];
      }
    if (i + 1 < s)
      f.vjp_kept(stage_time<TB, R>(i, t0, t1), X, kept[i], gk[i], gX, gc, gw);
    else
      f.vjp(stage_time<TB, R>(i, t0, t1), X, gk[i], gX, gc, gw);
#pragma unroll
    for (int q = 0; q < S; ++q) lam[q] += gX[q];
#pragma unroll
    for (int j = 0; j < i; ++j)
      if (TB::a(i, j) != R(0)) {
        const R ha = h * TB::a(i, j);
#pragma unroll
        for (int q = 0; q < S; ++q) gk[j][q] += ha * gX[q];
      }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// theta: sample, clip, log-probabilities
// ---------------------------------------------------------------------------------------------------------------
template <typename R>
VH_HD R sample_column(const Call<R>& a, int n, int b, int k, R& lq, R& lp, bool store = true) {
  const int kind = a.kind[k];
  const R mu = a.q_mu[b * a.P + k];
  R th;
  if (kind == VH_KIND_CONSTANT) {
    th = mu;  // TfConstant.sample: 0*u + value; log_prob 0 (vihds/distributions.py:242-246)
  } else {
    const R prec = a.q_prec[b * a.P + k];
    const R sigma = R(1) / vsqrt(prec);
    const R sv = mu + sigma * a.u[(size_t)n * a.P + k];
    const R raw = kind == VH_KIND_LOGNORMAL ? vexp(sv) : sv;
    th = clampv(raw, a.clip_lo[k], a.clip_hi[k]);
    const R x = kind == VH_KIND_LOGNORMAL ? vlog(th + R(1e-12)) : th;
    const R jac = kind == VH_KIND_LOGNORMAL ? x : R(0);
    const R pm = a.p_mu[k], pp = a.p_prec[k];
    lq += -Lim<R>::log2pi + R(0.5) * vlog(prec + R(1e-12)) - R(0.5) * prec * (mu - x) * (mu - x) - jac;
    lp += -Lim<R>::log2pi + R(0.5) * vlog(pp + R(1e-12)) - R(0.5) * pp * (pm - x) * (pm - x) - jac;
  }
  if (store && a.theta) a.theta[(size_t)k * a.N + n] = th;
  return th;
}

// ROLLED over the P sampled columns (one copy of sample_column in the instruction stream instead of one per slot):
// values land in a small per-thread array indexed by slot (local memory, L1-resident), from which the model's slot
// registers are filled with constant indices.
// Per-thread array indexed by a RUN-TIME slot number.  It must live in addressable memory (shared memory on the
// device, element s of thread t at p[s * stride]): as a register array ptxas turned every dynamic access into a chain of
// ~47 compare + predicated-move pairs, ~100 instructions per theta column (measured: the rolled prologue / epilogue
// was 58 % of the reverse kernel's samples at the icml size).
template <typename R>
struct SlotScratch {
  R* p;
  int stride;
  VH_HD R& operator[](int s) const { return p[s * stride]; }
};

template <class M, bool UNROLL = false>
VH_HD void load_theta(const Call<typename M::real>& a, int n, int b, typename M::real* th, typename M::real& lq,
                      typename M::real& lp, const SlotScratch<typename M::real>& loc, bool store = true) {
  typedef typename M::real R;
  for (int s = 0; s < M::NSLOT; ++s) {
    const int src = a.slot_src[s];
    loc[s] = (src < 0 && src != VH_SLOT_UNUSED) ? a.extra[(size_t)(-1 - src) * a.N + n] : R(0);
  }
  if (a.theta_in) {  // reverse sweep with the forward's theta at hand: P coalesced loads instead of re-sampling
#pragma unroll 5
    for (int k = 0; k < a.P; ++k) {
      const int s = a.col_slot[k];
      const R v = a.theta_in[(size_t)k * a.N + n];
      if (s >= 0) loc[s] = v;
    }
#pragma unroll
    for (int s = 0; s < M::NSLOT; ++s) th[s] = M::uses(s) ? loc[s] : R(0);
    return;
  }
  // The columns are independent: partial unrolling lets the ~8 dependent-latency loads of several columns overlap.
  // Used by the reverse kernels, where at the icml size this prologue (run twice) and the matching epilogue had grown
  // to 58 % of the samples; the forward kernels keep the rolled loop (unrolled it costs 8 % at N = 131,072).
  if (UNROLL) {
#pragma unroll 5
    for (int k = 0; k < a.P; ++k) {
      const R v = sample_column(a, n, b, k, lq, lp, store);
      const int s = a.col_slot[k];
      if (s >= 0) loc[s] = v;
    }
  } else {
#pragma unroll 1
    for (int k = 0; k < a.P; ++k) {
      const R v = sample_column(a, n, b, k, lq, lp, store);
      const int s = a.col_slot[k];
      if (s >= 0) loc[s] = v;
    }
  }
#pragma unroll
  for (int s = 0; s < M::NSLOT; ++s) th[s] = M::uses(s) ? loc[s] : R(0);
}

// column_vjp in two halves: everything it reads from global memory (nine loads per column, the one of u strided by P), and
// the arithmetic.  The latency-bound reverse kernels issue the loads of all their columns while the warps are still waiting
// for the lambda recurrence to finish -- the cotangent gth is the only late input.
template <typename R>
struct ColumnIn {
  int kind;
  R mu, prec, uu, lo, hi, th_in, pm, pp, gth_up;
  bool have_th;
};
template <typename R>
VH_HD void column_load(const Call<R>& a, int n, int b, int k, ColumnIn<R>& c) {
  c.kind = a.kind[k];
  c.gth_up = a.g_theta ? a.g_theta[(size_t)k * a.N + n] : R(0);
  c.mu = a.q_mu[b * a.P + k];
  c.prec = a.q_prec[b * a.P + k];
  c.uu = a.u[(size_t)n * a.P + k];
  c.lo = a.clip_lo[k];
  c.hi = a.clip_hi[k];
  c.have_th = a.theta_in != nullptr;
  c.th_in = c.have_th ? a.theta_in[(size_t)k * a.N + n] : R(0);
  c.pm = a.p_mu[k];
  c.pp = a.p_prec[k];
}
template <typename R>
VH_HD void column_vjp_from(const ColumnIn<R>& c, R gth, R glq, R glp, R& dmu, R& dprec) {
  dmu = R(0);
  dprec = R(0);
  const int kind = c.kind;
  gth += c.gth_up;
  const R mu = c.mu;
  if (kind == VH_KIND_CONSTANT) {
    dmu = gth;
    return;
  }
  const R prec = c.prec;
  const R sigma = R(1) / vsqrt(prec);
  const R uu = c.uu;
  const R lo = c.lo, hi = c.hi;
  R raw, th;
  bool have = false;
  if (c.have_th) {
    th = c.th_in;
    raw = th;
    have = th > lo && th < hi;
  }
  if (!have) {
    const R sv = mu + sigma * uu;
    raw = kind == VH_KIND_LOGNORMAL ? vexp(sv) : sv;
    th = clampv(raw, lo, hi);
  }
  const R pm = c.pm, pp = c.pp;
  R x, gx_to_th;
  if (kind == VH_KIND_LOGNORMAL) {
    x = vlog(th + R(1e-12));
    gx_to_th = R(1) / (th + R(1e-12));
  } else {
    x = th;
    gx_to_th = R(1);
  }
  const R jac = kind == VH_KIND_LOGNORMAL ? R(1) : R(0);
  const R dq = mu - x, dp = pm - x;
  const R gx = glq * (prec * dq - jac) + glp * (pp * dp - jac);
  const R gtot = gth + gx * gx_to_th;
  const R graw = gtot * clampmask(raw, lo, hi);
  const R gs = kind == VH_KIND_LOGNORMAL ? graw * raw : graw;
  dmu = gs - glq * prec * dq;
  dprec = -R(0.5) * gs * uu * sigma / prec + glq * (R(0.5) / (prec + R(1e-12)) - R(0.5) * dq * dq);
}

// cotangent of one sampled column -> (d mu, d prec) of q for this trajectory
template <typename R>
VH_HD void column_vjp(const Call<R>& a, int n, int b, int k, R gth, R glq, R glp, R& dmu, R& dprec) {
  dmu = R(0);
  dprec = R(0);
  const int kind = a.kind[k];
  if (a.g_theta) gth += a.g_theta[(size_t)k * a.N + n];
  const R mu = a.q_mu[b * a.P + k];
  if (kind == VH_KIND_CONSTANT) {
    dmu = gth;  // value of a constant: no trainable parameter behind it, reported for completeness
    return;
  }
  const R prec = a.q_prec[b * a.P + k];
  const R sigma = R(1) / vsqrt(prec);
  const R uu = a.u[(size_t)n * a.P + k];
  const R lo = a.clip_lo[k], hi = a.clip_hi[k];
  R raw, th;
  bool have = false;
  if (a.theta_in) {  // strictly inside the clip interval the clamp was the identity: raw == theta, no exp needed
    th = a.theta_in[(size_t)k * a.N + n];
    raw = th;
    have = th > lo && th < hi;
  }
  if (!have) {
    const R sv = mu + sigma * uu;
    raw = kind == VH_KIND_LOGNORMAL ? vexp(sv) : sv;
    th = clampv(raw, lo, hi);
  }
  const R pm = a.p_mu[k], pp = a.p_prec[k];
  R x, gx_to_th;
  if (kind == VH_KIND_LOGNORMAL) {
    x = vlog(th + R(1e-12));
    gx_to_th = R(1) / (th + R(1e-12));
  } else {
    x = th;
    gx_to_th = R(1);
  }
  const R jac = kind == VH_KIND_LOGNORMAL ? R(1) : R(0);
  const R dq = mu - x, dp = pm - x;
  // d logq / dx and d logp / dx (x = log theta for LogNormal, including the -x Jacobian term)
  const R gx = glq * (prec * dq - jac) + glp * (pp * dp - jac);
  const R gtot = gth + gx * gx_to_th;
  const R graw = gtot * clampmask(raw, lo, hi);
  const R gs = kind == VH_KIND_LOGNORMAL ? graw * raw : graw;
  dmu = gs - glq * prec * dq;
  dprec = -R(0.5) * gs * uu * sigma / prec + glq * (R(0.5) / (prec + R(1e-12)) - R(0.5) * dq * dq);
}

// ---------------------------------------------------------------------------------------------------------------
// forward trajectory.  Trace pointers are bumped by one time slab (S*N resp. 4*N elements) per step and the next
// step's observations / grid time are fetched one iteration ahead, so the loop carries no 64-bit index arithmetic
// and no exposed load latency.
// ---------------------------------------------------------------------------------------------------------------
// th: the model's slot values (sampled, clipped); lq / lp: log q(theta), log p(theta) of this trajectory
template <class M, class TB>
VH_HD void traj_forward_from(const Call<typename M::real>& a, int n, const typename M::real* w,
                             const typename M::real* th, typename M::real lq, typename M::real lp) {
  typedef typename M::real R;
  constexpr int S = M::S, NS = M::NS;
  const int b = n / a.IW;
  const size_t N = a.N;
  const int T = a.T;
  Rhs<M> f;
  f.w = w;
  f.nh = a.n_hidden;
  R x[S];
  R prec[4], lprec[4], ll[4];
  {
    R tc[3];
    M::treatments(a.treatments + (size_t)b * a.C, tc);
    M::setup(th, tc, f.c);
    M::init_state(th, tc, x);
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      prec[o] = M::DYN ? R(1) : th[S_prec_x + o];
      lprec[o] = M::DYN ? R(0) : vlog(prec[o]);
      ll[o] = R(0);
    }
  }
  const R h0 = a.times[1] - a.times[0];
  const R* obs = a.obs ? a.obs + (size_t)b * 4 * T : nullptr;
  R* xs = a.x_states ? a.x_states + n : nullptr;
  R* xpr = a.x_predict ? a.x_predict + n : nullptr;
  R ob[4] = {R(0), R(0), R(0), R(0)};
  if (obs) {
#pragma unroll
    for (int o = 0; o < 4; ++o) ob[o] = obs[o * T];
  }
  R t0 = a.times[0], t1 = a.times[1];
  for (int k = 0; k < T; ++k) {
    // fetch what the NEXT iteration consumes
    R obn[4] = {R(0), R(0), R(0), R(0)};
    const int kn = k + 1 < T ? k + 1 : k;
    if (obs) {
#pragma unroll
      for (int o = 0; o < 4; ++o) obn[o] = ld_early(obs + o * T + kn);
    }
    const R t2 = ld_early(a.times + (k + 2 < T ? k + 2 : T - 1));
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
    if (k + 1 < T) rk_step<Rhs<M>, TB>(f, t0, t1, TB::const_h ? h0 : (t1 - t0), x);
    t0 = t1;
    t1 = t2;
#pragma unroll
    for (int o = 0; o < 4; ++o) ob[o] = obn[o];
  }
  if (a.logp_species) {
#pragma unroll
    for (int o = 0; o < 4; ++o) a.logp_species[(size_t)n * 4 + o] = ll[o];
  }
  if (a.logp_theta) a.logp_theta[n] = lp;
  if (a.logq_theta) a.logq_theta[n] = lq;
}

template <class M, class TB>
VH_HD void traj_forward(const Call<typename M::real>& a, int n, const typename M::real* w,
                        const SlotScratch<typename M::real>& sc) {
  typedef typename M::real R;
  R th[M::NSLOT];
  R lq = R(0), lp = R(0);
  load_theta<M>(a, n, n / a.IW, th, lq, lp, sc);
  traj_forward_from<M, TB>(a, n, w, th, lq, lp);
}

// ---------------------------------------------------------------------------------------------------------------
// reverse trajectory.  RED: functor that folds (d mu, d prec) of column k of individual b into d_q_mu/d_q_prec
// (warp-aggregated atomics on the device, plain adds on the host).
// Register budget: theta (up to NSLOT values) is only needed before the time loop (RHS constants) and after it
// (chain rule through setup / clip / sample), so it is re-derived from u after the loop instead of being kept
// alive across it; the loop itself carries x, the prefetched previous checkpoint, lambda, the constants and their
// cotangents.
// ---------------------------------------------------------------------------------------------------------------
// Source of the checkpoints x_k for the reverse sweep, newest first: plain loads, register-prefetched one step ahead.
// (The warp-specialised kernel has its own source: a cp.async ring in shared memory, vh_launch.cuh.)
// PF > 0 (device, throughput kernels): additionally pull the checkpoint of PF steps further back into L2 (no
// destination registers) so that the one-step-ahead register load is an L2 hit, not an HBM round trip.
template <typename R, int S, int PF = 0>
struct DirectCk {
  const R* xs;
  size_t N;
  int k;
  R nx[S];
  VH_HD void start(const R* x_states, int n, size_t N_, int T) {
    N = N_;
    k = T - 1;
    xs = x_states + (size_t)k * S * N + n;
#pragma unroll
    for (int q = 0; q < S; ++q) nx[q] = xs[(size_t)q * N];
  }
  VH_HD void next(R* x) {
#pragma unroll
    for (int q = 0; q < S; ++q) x[q] = nx[q];
    if (k > 0) {
      xs -= (size_t)S * N;
#pragma unroll
      for (int q = 0; q < S; ++q) nx[q] = xs[(size_t)q * N];
#if defined(__CUDA_ARCH__)
      if (PF > 0 && k > PF) {
        const R* pf = xs - (size_t)PF * S * N;
#pragma unroll
        for (int q = 0; q < S; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + (size_t)q * N));
      }
#endif
    }
    --k;
  }
};

template <class M, class TB, typename GW, typename RED, typename CKS>
VH_HD void traj_backward(const Call<typename M::real>& a, int n, bool active, const typename M::real* w, GW& gw, RED& red,
                         const SlotScratch<typename M::real>& sc, CKS& ck) {
  typedef typename M::real R;
  constexpr int S = M::S, NS = M::NS;
  const int b = n / a.IW;
  const size_t N = a.N;
  const int T = a.T;
  R gth[M::NSLOT];
#pragma unroll
  for (int s = 0; s < M::NSLOT; ++s) gth[s] = R(0);
  R glq = R(0), glp = R(0);
  if (active) {
    Rhs<M> f;
    f.w = w;
    f.nh = a.n_hidden;
    R prec[4], iprec[4];
    {
      R th[M::NSLOT];
      R lq = R(0), lp = R(0), tc[3];
      load_theta<M, true>(a, n, b, th, lq, lp, sc, false);  // never rewrite theta in the reverse pass
      M::treatments(a.treatments + (size_t)b * a.C, tc);
      M::setup(th, tc, f.c);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        prec[o] = M::DYN ? R(1) : th[S_prec_x + o];
        iprec[o] = R(1) / prec[o];
      }
    }
    typename M::Consts gc;
#pragma unroll
    for (int i = 0; i < M::NC; ++i) gc.v[i] = R(0);
    R gprec[4], gl[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      gprec[o] = R(0);
      gl[o] = a.g_logp_species ? a.g_logp_species[(size_t)n * 4 + o] : R(0);
    }
    glq = a.g_logq_theta ? a.g_logq_theta[n] : R(0);
    glp = a.g_logp_theta ? a.g_logp_theta[n] : R(0);
    const R h0 = a.times[1] - a.times[0];
    const R* obs = a.obs ? a.obs + (size_t)b * 4 * T : nullptr;
    const size_t slab = (size_t)S * N;
    const R* gxs = a.g_x_states ? a.g_x_states + (size_t)(T - 1) * slab + n : nullptr;
    const R* gxpr = a.g_x_predict ? a.g_x_predict + (size_t)(T - 1) * 4 * N + n : nullptr;
    R lam[S], x[S];
    R ob[4] = {R(0), R(0), R(0), R(0)}, obp[4] = {R(0), R(0), R(0), R(0)};
#pragma unroll
    for (int q = 0; q < S; ++q) lam[q] = R(0);
    ck.start(a.x_states, n, N, T);
    if (obs) {
#pragma unroll
      for (int o = 0; o < 4; ++o) ob[o] = obs[o * T + T - 1];
    }
    R t0 = a.times[T - 1], t1 = t0;  // interval [t0, t1] = [times[k], times[k+1]]; unused at k = T-1
    for (int k = T - 1; k >= 0; --k) {
      // checkpoint x_k (and the fetch of what later iterations consume while this step's adjoint is computed)
      const int kp = k > 0 ? k - 1 : 0;
      ck.next(x);
      if (obs) {
#pragma unroll
        for (int o = 0; o < 4; ++o) obp[o] = ld_early(obs + o * T + kp);
      }
      const R tp = ld_early(a.times + kp);
      if (k + 1 < T) rk_step_vjp<Rhs<M>, TB>(f, t0, t1, TB::const_h ? h0 : (t1 - t0), x, lam, gc, gw);
      // emission at time k
      R xp[4], gxp[4];
      M::observe(x, xp);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        gxp[o] = gxpr ? gxpr[(size_t)o * N] : R(0);
        if (obs) {
          const R pr = M::DYN ? x[NS + o] : prec[o];
          const R ipr = M::DYN ? vdiv(R(1), pr) : iprec[o];
          const R d = xp[o] - ob[o];
          gxp[o] -= gl[o] * pr * d;
          const R gp = gl[o] * R(0.5) * (ipr - d * d);
          if (M::DYN)
            lam[NS + o] += gp;
          else
            gprec[o] += gp;
        }
      }
      M::observe_vjp(x, gxp, lam);
      if (gxs) {
#pragma unroll
        for (int q = 0; q < S; ++q) lam[q] += gxs[(size_t)q * N];
        gxs -= slab;
      }
      if (gxpr) gxpr -= (size_t)4 * N;
#pragma unroll
      for (int o = 0; o < 4; ++o) ob[o] = obp[o];
      t1 = t0;
      t0 = tp;
    }
    // chain rule back to theta: re-derive theta (cheap) rather than keep it live across the loop
    R th[M::NSLOT];
    R lq = R(0), lp = R(0), tc[3];
    load_theta<M, true>(a, n, b, th, lq, lp, sc, false);
    M::treatments(a.treatments + (size_t)b * a.C, tc);
    M::init_state_vjp(lam, gth);
    M::setup_vjp(th, tc, f.c, gc, gth);
    if (!M::DYN) {
#pragma unroll
      for (int o = 0; o < 4; ++o) gth[S_prec_x + o] += gprec[o];
    }
  }
  // scatter slot cotangents: sampled columns -> (d q_mu, d q_prec); extras -> d_extra.  Rolled over the columns
  // (one copy of column_vjp + the segmented warp reduction), reading the slot cotangents from a per-thread array.
  const SlotScratch<R>& gloc = sc;
#pragma unroll
  for (int s = 0; s < M::NSLOT; ++s) gloc[s] = M::uses(s) ? gth[s] : R(0);
#pragma unroll 5
  for (int k = 0; k < a.P; ++k) {
    const int s = a.col_slot[k];
    R dmu = R(0), dprec = R(0);
    if (active) column_vjp(a, n, b, k, s >= 0 ? gloc[s] : R(0), glq, glp, dmu, dprec);
    red(b, k, dmu, dprec, active);
  }
  if (a.d_extra && active) {
    for (int s = 0; s < M::NSLOT; ++s) {
      const int src = a.slot_src[s];
      if (src < 0 && src != VH_SLOT_UNUSED) a.d_extra[(size_t)(-1 - src) * N + n] = gloc[s];
    }
  }
}

}  // namespace vh
