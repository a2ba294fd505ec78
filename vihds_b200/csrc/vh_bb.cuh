// dr_blackbox: neural right-hand side (models/dr_blackbox.py:15-58, vihds/ode.py:119-138 NeuralStates,
// vihds/precisions.py:44-94 NeuralPrecisions with a hidden layer and ReLU), one trajectory per thread.
//
//   aug   = [x (NST = 4 + n_latent_species), c]        c = [z.., x.., y.. (+ device offset), treatments, dev_1hot]
//   hid   = relu(W1 aug + b1)                          H units
//   dx    = sigmoid(Wp hid + bp) - sigmoid(Wd hid + bd) * x
//   hp    = relu(Q1 [t, x, c] + qb1)                   HP units
//   dv    = sigmoid(Qp hp + qbp) - sigmoid(Qd hp + qbd) * v          v = the 4 precision states
//
// c is constant along a trajectory, so the c-columns of W1 / Q1 (and the biases) are folded ONCE per trajectory into
// hc = W1[:, NST:] c + b1 and hpc = Q1[:, 1+NST:] c + qb1: an RHS evaluation then costs NST*H + 2*NST*H + (1+NST)*HP
// + 8*HP multiply-adds (750 for the icml spec) instead of 1,695, and the reverse sweep only accumulates the cotangents
// of hc / hpc per step; the outer products with c are formed once after the time loop.
//
// Flat weight layout (vh_num_weights): W1[H][nin] b1[H] Wp[NST][H] bp[NST] Wd[NST][H] bd[NST]
//                                      Q1[HP][nin+1] qb1[HP] Qp[4][HP] qbp[4] Qd[4][HP] qbd[4],   nin = NST + nc.
// Weight gradients are sums over trajectories AND RHS evaluations of outer products (pre-activation cotangent) x
// (activation): the sink `GW` receives both vectors per evaluation.  On the device it is a warp-level GEMM through
// shared memory (BbWarpWgrad, vh_blackbox.cu); in the host-side math check it accumulates directly.
#pragma once
#include "vh_traj.cuh"

namespace vh {

// compile-time layout of the flat weight vector (so that every weight access is base + immediate)
template <int NST, int H, int HP, int NC>
struct BbLayoutT {
  static constexpr int nc = NC, nin = NST + NC;
  static constexpr int W1 = 0;
  static constexpr int b1 = W1 + H * nin;
  static constexpr int Wp = b1 + H;
  static constexpr int bp = Wp + NST * H;
  static constexpr int Wd = bp + NST;
  static constexpr int bd = Wd + NST * H;
  static constexpr int Q1 = bd + NST;
  static constexpr int qb1 = Q1 + HP * (nin + 1);
  static constexpr int Qp = qb1 + HP;
  static constexpr int qbp = Qp + 4 * HP;
  static constexpr int Qd = qbp + 4;
  static constexpr int qbd = Qd + 4 * HP;
  static constexpr int total = qbd + 4;
};

// relu'(0) = 0, as torch.relu's backward
template <typename R>
VH_HD R relu_mask(R pre, R g) {
  return pre > R(0) ? g : R(0);
}

// Per-thread scratch row (shared memory on the device, odd row stride => conflict-free; a plain array on the host):
//   [0, RS)              staging area, doubles as the hidden-layer scratch of the current evaluation:
//                          states net : x[NST] | hid[H] | gpre[H] | gzp[NST] | gzd[NST]
//                          precisions : t | x[NST] | hp[HP] | gpre[HP] | gzp[4] | gzd[4]
//                          constants  : c[NC]                       (read by fold() and by the final outer products)
//   [RS, RS+H+HP)        hc | hpc   folded constant parts of the two hidden pre-activations
//   [RS+H+HP, ROW)       cotangents of hc | hpc, accumulated over all evaluations of the reverse sweep
template <int NST, int H, int HP, int NC>
struct BbRow {
  // staging regions are padded to tensor-core tile multiples (8 for the n side, 32 for the m side); entry H of the
  // hidden vector is a constant 1 so that the bias gradients fall out of the same GEMM, the rest of the padding is 0
  static constexpr int XP = 8, HM = 32, GZ = 16, GZP = 8;
  static_assert(NST <= XP && 1 + NST <= XP && H + 1 <= HM && HP + 1 <= HM && 2 * NST <= GZ, "tile padding");
  // states staging
  static constexpr int sX = 0, sHID = XP, sGPRE = XP + HM, sGZP = XP + 2 * HM, sGZD = sGZP + NST;
  static constexpr int RS1 = XP + 2 * HM + GZ;
  // precisions staging ([t, x] is one operand; gzp | gzd are contiguous: one 8-wide operand)
  static constexpr int pT = 0, pX = 1, pHP = XP, pGPRE = XP + HM, pGZP = XP + 2 * HM, pGZD = pGZP + 4;
  static constexpr int RS2 = XP + 2 * HM + GZP;
  static constexpr int RS = (RS1 > RS2 ? RS1 : RS2) > NC ? (RS1 > RS2 ? RS1 : RS2) : NC;
  static constexpr int HC = RS, HPC = RS + H, GHC = RS + H + HP, GHPC = RS + 2 * H + HP;
  static constexpr int ROW = (RS + 2 * (H + HP)) | 1;
};

// The hidden layers are ROLLED loops over scratch in shared memory (small code, few registers); only the short
// inner loops over the NST states / 4 precisions are unrolled with their accumulators in registers.
template <typename R, int NLS_, int H_, int HP_, int NC_>
struct BbRhs {
  typedef R real;
  static constexpr int NLS = NLS_, H = H_, HP = HP_, NC = NC_;
  static constexpr int NST = 4 + NLS_;  // neural states
  static constexpr int S = NST + 4;     // + precision states
  typedef BbLayoutT<NST, H_, HP_, NC_> L;
  typedef BbRow<NST, H_, HP_, NC_> ROWL;
  struct Kept {};  // nothing kept across stages: the vjp recomputes its hidden layer into the scratch row
  struct Grad {};  // cotangents of hc / hpc live in the scratch row
  const R* w;      // flat weights (shared memory on the device)
  R* row;          // this thread's scratch row

  // fold the constant inputs (already written to row[0..NC)) into the hidden pre-activations; zero the cotangents
  VH_HD void fold() {
    for (int h = 0; h < H; ++h) {
      R a = w[L::b1 + h];
      const R* wr = w + L::W1 + h * L::nin + NST;
      for (int j = 0; j < NC; ++j) a += wr[j] * row[j];
      row[ROWL::HC + h] = a;
      row[ROWL::GHC + h] = R(0);
    }
    for (int h = 0; h < HP; ++h) {
      R a = w[L::qb1 + h];
      const R* wr = w + L::Q1 + h * (L::nin + 1) + 1 + NST;
      for (int j = 0; j < NC; ++j) a += wr[j] * row[j];
      row[ROWL::HPC + h] = a;
      row[ROWL::GHPC + h] = R(0);
    }
  }

  // states net forward: leaves hid (and, if KEEP_PRE, the relu mask in gpre's slot) in the row; zp/zd in registers
  template <bool KEEP_PRE>
  VH_HD void states_fwd(const R* x, R* zp, R* zd) const {
#pragma unroll
    for (int o = 0; o < NST; ++o) {
      zp[o] = w[L::bp + o];
      zd[o] = w[L::bd + o];
    }
    for (int h = 0; h < H; ++h) {
      const R* wr = w + L::W1 + h * L::nin;
      R a = row[ROWL::HC + h];
#pragma unroll
      for (int s = 0; s < NST; ++s) a += wr[s] * x[s];
      const R hv = a > R(0) ? a : R(0);
      if (KEEP_PRE) {
        row[ROWL::sHID + h] = hv;
        row[ROWL::sGPRE + h] = a;  // pre-activation for now; overwritten by its cotangent in the vjp
      }
#pragma unroll
      for (int o = 0; o < NST; ++o) {
        zp[o] += w[L::Wp + o * H + h] * hv;
        zd[o] += w[L::Wd + o * H + h] * hv;
      }
    }
  }
  template <bool KEEP_PRE>
  VH_HD void prec_fwd(R t, const R* x, R* zp, R* zd) const {
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      zp[o] = w[L::qbp + o];
      zd[o] = w[L::qbd + o];
    }
    for (int h = 0; h < HP; ++h) {
      const R* q = w + L::Q1 + h * (L::nin + 1);
      R a = row[ROWL::HPC + h] + q[0] * t;
#pragma unroll
      for (int s = 0; s < NST; ++s) a += q[1 + s] * x[s];
      const R hv = a > R(0) ? a : R(0);
      if (KEEP_PRE) {
        row[ROWL::pHP + h] = hv;
        row[ROWL::pGPRE + h] = a;
      }
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        zp[o] += w[L::Qp + o * HP + h] * hv;
        zd[o] += w[L::Qd + o * HP + h] * hv;
      }
    }
  }

  VH_HD void eval(R t, const R* x, R* dx) const {
    {
      R zp[NST], zd[NST];
      states_fwd<false>(x, zp, zd);
#pragma unroll
      for (int o = 0; o < NST; ++o) dx[o] = sigmoid(zp[o]) - sigmoid(zd[o]) * x[o];
    }
    {
      R zp[4], zd[4];
      prec_fwd<false>(t, x, zp, zd);
#pragma unroll
      for (int o = 0; o < 4; ++o) dx[NST + o] = sigmoid(zp[o]) - sigmoid(zd[o]) * x[NST + o];
    }
  }
  VH_HD void eval_keep(R t, const R* x, R* dx, Kept&) const { eval(t, x, dx); }  // eval() never writes the row
  VH_HD void keep_only(R, const R*, Kept&) const {}

  // g: cotangent of dx  ->  gx (accumulated); the staged rows go to the weight-gradient sink gw
  template <typename GW>
  VH_HD void vjp(R t, const R* x, const R* g, R* gx, Grad&, GW& gw) const {
    {  // states net
      gw.begin();  // the row is about to be overwritten: the previous outer-product pass must have consumed it
      R zp[NST], zd[NST], gzp[NST], gzd[NST];
      states_fwd<true>(x, zp, zd);
#pragma unroll
      for (int o = 0; o < NST; ++o) {
        const R sp = sigmoid(zp[o]), sd = sigmoid(zd[o]);
        gx[o] -= g[o] * sd;
        gzp[o] = g[o] * sp * (R(1) - sp);
        gzd[o] = -g[o] * x[o] * sd * (R(1) - sd);
        row[ROWL::sGZP + o] = gzp[o];
        row[ROWL::sGZD + o] = gzd[o];
        row[ROWL::sX + o] = x[o];
      }
      for (int h = 0; h < H; ++h) {
        R gh = R(0);
#pragma unroll
        for (int o = 0; o < NST; ++o) gh += w[L::Wp + o * H + h] * gzp[o] + w[L::Wd + o * H + h] * gzd[o];
        gh = relu_mask(row[ROWL::sGPRE + h], gh);
        row[ROWL::sGPRE + h] = gh;
        row[ROWL::GHC + h] += gh;
        const R* wr = w + L::W1 + h * L::nin;
#pragma unroll
        for (int s = 0; s < NST; ++s) gx[s] += wr[s] * gh;
      }
      // tile padding (the precision phase reuses this space): hidden unit H is the constant 1 of the bias
      // gradients, everything else zero
#pragma unroll
      for (int i = NST; i < ROWL::XP; ++i) row[ROWL::sX + i] = R(0);
#pragma unroll
      for (int i = H; i < ROWL::HM; ++i) {
        row[ROWL::sHID + i] = i == H ? R(1) : R(0);
        row[ROWL::sGPRE + i] = R(0);
      }
#pragma unroll
      for (int i = 2 * NST; i < ROWL::GZ; ++i) row[ROWL::sGZP + i] = R(0);
      gw.states(row);
    }
    {  // precision net
      gw.begin();
      R zp[4], zd[4], gzp[4], gzd[4];
      prec_fwd<true>(t, x, zp, zd);
      row[ROWL::pT] = t;
#pragma unroll
      for (int s = 0; s < NST; ++s) row[ROWL::pX + s] = x[s];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const R sp = sigmoid(zp[o]), sd = sigmoid(zd[o]);
        gx[NST + o] -= g[NST + o] * sd;
        gzp[o] = g[NST + o] * sp * (R(1) - sp);
        gzd[o] = -g[NST + o] * x[NST + o] * sd * (R(1) - sd);
        row[ROWL::pGZP + o] = gzp[o];
        row[ROWL::pGZD + o] = gzd[o];
      }
      for (int h = 0; h < HP; ++h) {
        R gh = R(0);
#pragma unroll
        for (int o = 0; o < 4; ++o) gh += w[L::Qp + o * HP + h] * gzp[o] + w[L::Qd + o * HP + h] * gzd[o];
        gh = relu_mask(row[ROWL::pGPRE + h], gh);
        row[ROWL::pGPRE + h] = gh;
        row[ROWL::GHPC + h] += gh;
        const R* q = w + L::Q1 + h * (L::nin + 1);
#pragma unroll
        for (int s = 0; s < NST; ++s) gx[s] += q[1 + s] * gh;
      }
#pragma unroll
      for (int i = 1 + NST; i < ROWL::XP; ++i) row[ROWL::pT + i] = R(0);
#pragma unroll
      for (int i = HP; i < ROWL::HM; ++i) {
        row[ROWL::pHP + i] = i == HP ? R(1) : R(0);
        row[ROWL::pGPRE + i] = R(0);
      }
      gw.precisions(row);
    }
  }
  template <typename GW>
  VH_HD void vjp_kept(R t, const R* x, const Kept&, const R* g, R* gx, Grad& gc, GW& gw) const {
    vjp(t, x, g, gx, gc, gw);
  }

  // observe, models/dr_blackbox.py:112-121
  VH_HD static void observe(const R* x, R* xp) {
    xp[0] = x[0];
    xp[1] = x[0] * x[1];
    xp[2] = x[0] * x[2];
    xp[3] = x[0] * x[3];
  }
  VH_HD static void observe_vjp(const R* x, const R* gxp, R* gx) {
    gx[0] += gxp[0] + gxp[1] * x[1] + gxp[2] * x[2] + gxp[3] * x[3];
    gx[1] += gxp[1] * x[0];
    gx[2] += gxp[2] * x[0];
    gx[3] += gxp[3] * x[0];
  }
};

// direct accumulation of the staged rows into a flat gradient vector (host-side math check; one trajectory at a time)
template <class F>
struct BbDirectWgrad {
  typedef typename F::real R;
  typedef typename F::L L;
  typedef typename F::ROWL RW;
  R* d;
  VH_HD void begin() const {}
  VH_HD void states(const R* r) const {
    for (int h = 0; h < F::H; ++h)
      for (int s = 0; s < F::NST; ++s) d[L::W1 + h * L::nin + s] += r[RW::sGPRE + h] * r[RW::sX + s];
    for (int o = 0; o < F::NST; ++o) {
      d[L::bp + o] += r[RW::sGZP + o];
      d[L::bd + o] += r[RW::sGZD + o];
      for (int h = 0; h < F::H; ++h) {
        d[L::Wp + o * F::H + h] += r[RW::sGZP + o] * r[RW::sHID + h];
        d[L::Wd + o * F::H + h] += r[RW::sGZD + o] * r[RW::sHID + h];
      }
    }
  }
  VH_HD void precisions(const R* r) const {
    for (int h = 0; h < F::HP; ++h)
      for (int i = 0; i < 1 + F::NST; ++i) d[L::Q1 + h * (L::nin + 1) + i] += r[RW::pGPRE + h] * r[i];
    for (int o = 0; o < 4; ++o) {
      d[L::qbp + o] += r[RW::pGZP + o];
      d[L::qbd + o] += r[RW::pGZD + o];
      for (int h = 0; h < F::HP; ++h) {
        d[L::Qp + o * F::HP + h] += r[RW::pGZP + o] * r[RW::pHP + h];
        d[L::Qd + o * F::HP + h] += r[RW::pGZD + o] * r[RW::pHP + h];
      }
    }
  }
  // after the time loop (row[0..NC) = c again): outer products of the folded-constant cotangents with c, hidden biases
  VH_HD void consts(const R* r) const {
    for (int h = 0; h < F::H; ++h) {
      d[L::b1 + h] += r[RW::GHC + h];
      for (int j = 0; j < F::NC; ++j) d[L::W1 + h * L::nin + F::NST + j] += r[RW::GHC + h] * r[j];
    }
    for (int h = 0; h < F::HP; ++h) {
      d[L::qb1 + h] += r[RW::GHPC + h];
      for (int j = 0; j < F::NC; ++j) d[L::Q1 + h * (L::nin + 1) + 1 + F::NST + j] += r[RW::GHPC + h] * r[j];
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// trajectories.  Slots: 0..3 = init_x, init_rfp, init_yfp, init_cfp; 4.. = latent parameters in the order
// z1..z_nz, x1..x_nx, y1..y_ny (models/dr_blackbox.py:35-51).  `extra` rows (P > 0, E == n_y): the device offsets
// offset_layer(dev_1hot)[:, k] that condition_theta adds to y_k (models/dr_blackbox.py:86-96), one [N] plane each;
// the sampled y_k keeps its own log-prob terms (they are evaluated on the un-conditioned sample in the reference).
// ---------------------------------------------------------------------------------------------------------------
// c = [latents (+ offsets on the y block), treatments, dev_1hot] (written to the scratch row);  th4 = the four
// initial-state parameters
template <typename R>
VH_HD void bb_load(const Call<R>& a, int n, int b, R* th4, R* c, R& lq, R& lp, bool store) {
  const int n_lat = a.bb_nlat;
  for (int s = 0; s < 4 + n_lat; ++s) {
    const int src = a.slot_src[s];
    R v = R(0);
    if (src >= 0)
      v = sample_column(a, n, b, src, lq, lp, store);
    else if (src != VH_SLOT_UNUSED)
      v = a.extra[(size_t)(-1 - src) * a.N + n];
    if (s < 4)
      th4[s] = v;
    else
      c[s - 4] = v;
  }
  for (int k = 0; k < a.bb_noff; ++k) c[n_lat - a.bb_ny + k] += a.extra[(size_t)k * a.N + n];
  for (int j = 0; j < a.n_free; ++j) sample_column(a, n, b, a.free_cols[j], lq, lp, store);
  for (int j = 0; j < a.C; ++j) c[n_lat + j] = a.treatments[(size_t)b * a.C + j];
  for (int j = 0; j < a.D; ++j) c[n_lat + a.C + j] = a.dev_1hot[(size_t)b * a.D + j];
}

template <class F>
VH_HD void bb_init_state(const Call<typename F::real>& a, const typename F::real* th4, typename F::real* x) {
  typedef typename F::real R;
#pragma unroll
  for (int s = 0; s < 4; ++s) x[s] = th4[s];
#pragma unroll
  for (int s = 4; s < F::NST; ++s) x[s] = a.bb_init_latent;  // models/dr_blackbox.py:101-104
#pragma unroll
  for (int o = 0; o < 4; ++o) x[F::NST + o] = a.bb_init_prec;
}

template <class F, class TB>
VH_HD void bb_traj_forward(const Call<typename F::real>& a, int n, const typename F::real* w, typename F::real* row) {
  typedef typename F::real R;
  constexpr int S = F::S, NST = F::NST;
  const int b = n / a.IW;
  const size_t N = a.N;
  const int T = a.T;
  F f;
  f.w = w;
  f.row = row;
  R x[S], ll[4];
  R lq = R(0), lp = R(0);
  {
    R th4[4];
    bb_load(a, n, b, th4, row, lq, lp, true);
    f.fold();
    bb_init_state<F>(a, th4, x);
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) ll[o] = R(0);
  const R h0 = a.times[1] - a.times[0];
  const R* obs = a.obs ? a.obs + (size_t)b * 4 * T : nullptr;
  R* xs = a.x_states ? a.x_states + n : nullptr;
  R* xpr = a.x_predict ? a.x_predict + n : nullptr;
  R t0 = a.times[0], t1 = a.times[1];
  for (int k = 0; k < T; ++k) {
    const R t2 = a.times[k + 2 < T ? k + 2 : T - 1];
    if (xs) {
#pragma unroll
      for (int q = 0; q < S; ++q) xs[(size_t)q * N] = x[q];
      xs += (size_t)S * N;
    }
    R xp[4];
    F::observe(x, xp);
    if (xpr) {
#pragma unroll
      for (int o = 0; o < 4; ++o) xpr[(size_t)o * N] = xp[o];
      xpr += (size_t)4 * N;
    }
    if (obs) {
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const R pr = x[NST + o];
        const R d = xp[o] - obs[o * T + k];
        ll[o] += R(-0.5) * (Lim<R>::log2pi - vlog(pr) + pr * d * d);
      }
    }
    if (k + 1 < T) rk_step<F, TB>(f, t0, t1, TB::const_h ? h0 : (t1 - t0), x);
    t0 = t1;
    t1 = t2;
  }
  if (a.logp_species) {
#pragma unroll
    for (int o = 0; o < 4; ++o) a.logp_species[(size_t)n * 4 + o] = ll[o];
  }
  if (a.logp_theta) a.logp_theta[n] = lp;
  if (a.logq_theta) a.logq_theta[n] = lq;
}

// Reverse sweep.  EVERY lane of a warp must run this (the weight-gradient sink is warp-cooperative): lanes past the end
// of the batch (`active` false) re-run the last trajectory with zero upstream gradients, so all their contributions
// are exact zeros.
template <class F, class TB, typename GW, typename RED>
VH_HD void bb_traj_backward(const Call<typename F::real>& a, int n, bool active, const typename F::real* w,
                            typename F::real* row, GW& gw, RED& red) {
  typedef typename F::real R;
  constexpr int S = F::S, NST = F::NST;
  const int b = n / a.IW;
  const size_t N = a.N;
  const int T = a.T;
  const int n_lat = a.bb_nlat;
  const R on = active ? R(1) : R(0);
  typedef typename F::L L;
  typedef typename F::ROWL RW;
  F f;
  f.w = w;
  f.row = row;
  {
    R th4[4], lq = R(0), lp = R(0);
    bb_load(a, n, b, th4, row, lq, lp, false);
    f.fold();
  }
  typename F::Grad gc;
  R gl[4];
#pragma unroll
  for (int o = 0; o < 4; ++o) gl[o] = (a.g_logp_species ? a.g_logp_species[(size_t)n * 4 + o] : R(0)) * on;
  const R glq = (a.g_logq_theta ? a.g_logq_theta[n] : R(0)) * on;
  const R glp = (a.g_logp_theta ? a.g_logp_theta[n] : R(0)) * on;
  const R h0 = a.times[1] - a.times[0];
  const R* obs = a.obs ? a.obs + (size_t)b * 4 * T : nullptr;
  const size_t slab = (size_t)S * N;
  const R* xs = a.x_states + (size_t)(T - 1) * slab + n;
  const R* gxs = a.g_x_states ? a.g_x_states + (size_t)(T - 1) * slab + n : nullptr;
  const R* gxpr = a.g_x_predict ? a.g_x_predict + (size_t)(T - 1) * 4 * N + n : nullptr;
  R lam[S], x[S], xprev[S];
#pragma unroll
  for (int q = 0; q < S; ++q) {
    lam[q] = R(0);
    x[q] = xs[(size_t)q * N];
    xprev[q] = x[q];
  }
  R t0 = a.times[T - 1], t1 = t0;
  for (int k = T - 1; k >= 0; --k) {
    const int kp = k > 0 ? k - 1 : 0;
    if (k > 0) {
      xs -= slab;
#pragma unroll
      for (int q = 0; q < S; ++q) xprev[q] = xs[(size_t)q * N];
    }
    const R tp = a.times[kp];
    if (k + 1 < T) rk_step_vjp<F, TB>(f, t0, t1, TB::const_h ? h0 : (t1 - t0), x, lam, gc, gw);
    R xp[4], gxp[4];
    F::observe(x, xp);
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      gxp[o] = gxpr ? gxpr[(size_t)o * N] * on : R(0);
      if (obs) {
        const R pr = x[NST + o];
        const R d = xp[o] - obs[o * T + k];
        gxp[o] -= gl[o] * pr * d;
        lam[NST + o] += gl[o] * R(0.5) * (vdiv(R(1), pr) - d * d);
      }
    }
    F::observe_vjp(x, gxp, lam);
    if (gxs) {
#pragma unroll
      for (int q = 0; q < S; ++q) lam[q] += gxs[(size_t)q * N] * on;
      gxs -= slab;
    }
    if (gxpr) gxpr -= (size_t)4 * N;
#pragma unroll
    for (int q = 0; q < S; ++q) x[q] = xprev[q];
    t1 = t0;
    t0 = tp;
  }
  // weight gradients of the folded constant columns + biases of the two hidden layers (needs c in the row again)
  gw.begin();
  {
    R th4[4], lq = R(0), lp = R(0);
    bb_load(a, n, b, th4, row, lq, lp, false);
  }
  gw.consts(row);
  // cotangents of the slots: initial state (slots 0..3) and latent parameters (through the folded columns)
  for (int s = 0; s < 4 + n_lat; ++s) {
    R gth;
    if (s < 4) {
      gth = lam[s];
    } else {
      const int j = s - 4;
      gth = R(0);
      for (int h = 0; h < F::H; ++h) gth += w[L::W1 + h * L::nin + NST + j] * row[RW::GHC + h];
      for (int h = 0; h < F::HP; ++h) gth += w[L::Q1 + h * (L::nin + 1) + 1 + NST + j] * row[RW::GHPC + h];
      const int ky = j - (n_lat - a.bb_ny);
      if (ky >= 0 && ky < a.bb_noff && a.d_extra && active) a.d_extra[(size_t)ky * N + n] = gth;
    }
    const int src = a.slot_src[s];
    if (src >= 0) {
      R dmu = R(0), dprec = R(0);
      if (active) column_vjp(a, n, b, src, gth, glq, glp, dmu, dprec);
      red(b, src, dmu, dprec, active);
    } else if (src != VH_SLOT_UNUSED && a.d_extra && active) {
      a.d_extra[(size_t)(-1 - src) * N + n] = gth;  // simulate seam: theta handed in as extra rows
    }
  }
  for (int j = 0; j < a.n_free; ++j) {
    R dmu = R(0), dprec = R(0);
    if (active) column_vjp(a, n, b, a.free_cols[j], R(0), glq, glp, dmu, dprec);
    red(b, a.free_cols[j], dmu, dprec, active);
  }
}

}  // namespace vh
