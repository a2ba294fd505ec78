// Lane-split forward kernel for the double-receiver family (dr_constant v1 / v2, constant precisions): EIGHT LANES PER
// TRAJECTORY, one species per lane -- the mapping BASELINE.json's north star names ("one warp per trajectory", here a
// quarter warp: the model has 8 species), built to MEASURE it against the one-thread-per-trajectory kernels
// (captures under profiles/r02_lane_*).
//
//   species i = lane & 7:   dx_i = prod_i - (sgn_i * gamma + d_i) * x_i,     gamma = r sigmoid(4 (t - tlag)) (1 - x_0 / K)
//   (models/dr_constant.py:77-112: sgn_0 = -1, prod_0 = d_0 = 0 gives dx_0 = gamma x_0; prod_2 = cY * P81 and
//   prod_3 = cC * P76 depend on luxR / lasR, every other production term is a per-trajectory constant.)
// Per evaluation a lane needs x_0 (growth / dilution), luxR and lasR (promoter activities): three width-8 shuffles.
// The promoter activity is ONE formula evaluated with per-lane constants (lane 2 holds the P81 constants, lane 3 the
// P76 ones), the sigmoid / gamma part is computed redundantly by the 8 lanes (SIMT: it costs the same as once).
// Sampling / clipping / log-probabilities of the P theta columns are split over the 8 lanes of a trajectory; the
// RHS constants (clamps, Hill fractions: 6 powf) are computed redundantly per lane from the slot values in shared memory.
#pragma once
#include "vh_traj.cuh"

namespace vh {

constexpr int LANE_TRAJ_PER_WARP = 4;  // 8 lanes each
constexpr int LANE_WARPS = 4;          // warps per CTA

template <class M>
struct LaneRhs {
  typedef float real;
  static constexpr int S = 1;  // what rk_step integrates per lane
  float r, tlag, iK, fR, fS;   // shared by the 8 lanes of a trajectory
  float prod, d, sgn;          // this lane's species: constant production, decay, sign of the dilution term
  float pe, pkr, pks, pc;      // promoter activity constants of this lane (lanes 2, 3; zero elsewhere)
  int base;                    // first lane of this trajectory's group
  __device__ void eval(float t, const float* x, float* dx) const {
    const unsigned full = 0xffffffffu;
    const float x0 = __shfl_sync(full, x[0], base);
    const float xr = __shfl_sync(full, x[0], base + 6);
    const float xs = __shfl_sync(full, x[0], base + 7);
    const float gam = r * sigmoid(4.f * (t - tlag)) * (1.f - x0 * iK);
    const float a = pkr * (xr * xr * fR), b = pks * (xs * xs * fS);
    const float P = (pe + a + b) * vdiv(1.f, 1.f + a + b);  // models/dr_constant.py:90-95
    dx[0] = (prod + pc * P) - (sgn * gam + d) * x[0];
  }
};

template <class M, class TB>
__global__ void __launch_bounds__(LANE_WARPS * 32) elbo_fwd_lane_kernel(const Call<float> a) {
  static_assert(!M::DYN && !M::RELAY && !M::DEGR && M::NS == 8, "lane-split kernel: dr_constant v1 / v2");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sm = reinterpret_cast<float*>(smem_raw);  // slot values [NSLOT][16 trajectories of this CTA]
  constexpr int TPC = LANE_WARPS * LANE_TRAJ_PER_WARP;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sp = lane & 7, tl = warp * LANE_TRAJ_PER_WARP + (lane >> 3);  // species, trajectory within the CTA
  const int n0 = blockIdx.x * TPC + tl;
  const bool active = n0 < a.N;
  const int n = active ? n0 : a.N - 1;
  const int b = n / a.IW;
  const size_t N = a.N;
  const int T = a.T;
  const SlotScratch<float> loc{sm + tl, TPC};
  // theta: lane j of the group samples columns j, j + 8, ...; slots without a column are filled by the lane that owns them
  for (int s = sp; s < M::NSLOT; s += 8) {
    const int src = a.slot_src[s];
    if (src < 0) loc[s] = src != VH_SLOT_UNUSED ? a.extra[(size_t)(-1 - src) * N + n] : 0.f;
  }
  float lq = 0.f, lp = 0.f;
  for (int k = sp; k < a.P; k += 8) {
    const float v = sample_column(a, n, b, k, lq, lp, active);
    const int s = a.col_slot[k];
    if (s >= 0) loc[s] = v;
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    lq += __shfl_xor_sync(full, lq, o);
    lp += __shfl_xor_sync(full, lp, o);
  }
  __syncwarp();
  LaneRhs<M> f;
  float x, prec;
  {
    float th[M::NSLOT];
#pragma unroll
    for (int s = 0; s < M::NSLOT; ++s) th[s] = M::uses(s) ? loc[s] : 0.f;
    float tc[3];
    M::treatments(a.treatments + (size_t)b * a.C, tc);
    typename M::Consts c;
    M::setup(th, tc, c);
    float x0[M::S];
    M::init_state(th, tc, x0);
    const float* v = c.v;
    f.r = v[C_r];
    f.tlag = v[C_tlag];
    f.iK = c.iK;
    f.fR = v[C_fR];
    f.fS = v[C_fS];
    f.base = lane & ~7;
    // per-lane species constants (selected with predicates: no dynamic register indexing)
    f.prod = sp == 1 ? v[C_rc] : sp == 4 ? v[C_p4] : sp == 5 ? v[C_p5] : sp == 6 ? v[C_p6] : sp == 7 ? v[C_p7] : 0.f;
    f.d = sp == 1 ? v[C_drfp] : sp == 2 ? v[C_dyfp] : sp == 3 ? v[C_dcfp] : sp == 6 ? v[C_dR] : sp == 7 ? v[C_dS] : 0.f;
    f.sgn = sp == 0 ? -1.f : 1.f;
    f.pc = sp == 2 ? v[C_cY] : sp == 3 ? v[C_cC] : 0.f;
    f.pe = sp == 2 ? v[C_e81] : v[C_e76];
    f.pkr = sp == 2 ? v[C_KGR81] : v[C_KGR76];
    f.pks = sp == 2 ? v[C_KGS81] : v[C_KGS76];
    x = sp == 0 ? x0[0] : sp == 1 ? x0[1] : sp == 2 ? x0[2] : sp == 3 ? x0[3] : sp == 4 ? x0[4] : sp == 5 ? x0[5] : sp == 6 ? x0[6] : x0[7];
    prec = sp == 0 ? th[S_prec_x] : sp == 1 ? th[S_prec_rfp] : sp == 2 ? th[S_prec_yfp] : th[S_prec_cfp];  // lanes 0..3
  }
  const float lprec = vlog(prec);
  float ll = 0.f;
  const float h0 = a.times[1] - a.times[0];
  const bool obsl = a.obs != nullptr && sp < 4;
  const float* obs = obsl ? a.obs + ((size_t)b * 4 + sp) * T : nullptr;
  float* xs = (a.x_states && active) ? a.x_states + (size_t)sp * N + n : nullptr;
  float* xpr = (a.x_predict && active && sp < 4) ? a.x_predict + (size_t)sp * N + n : nullptr;
  float ob = obsl ? obs[0] : 0.f;
  float t0 = a.times[0], t1 = a.times[1];
  for (int k = 0; k < T; ++k) {
    const float obn = obsl ? ld_early(obs + (k + 1 < T ? k + 1 : k)) : 0.f;
    const float t2 = ld_early(a.times + (k + 2 < T ? k + 2 : T - 1));
    if (xs) {
      *xs = x;
      xs += (size_t)M::S * N;
    }
    // observe (vihds/ode.py:84-93): [OD, OD RFP, OD (YFP + F530), OD (CFP + F480)] on lanes 0..3
    const float xod = __shfl_sync(full, x, f.base);
    const float xaf = __shfl_down_sync(full, x, 2, 8);  // lane 2 <- F530 (lane 4), lane 3 <- F480 (lane 5)
    const float xp = sp == 0 ? x : (sp == 1 ? xod * x : xod * (x + xaf));
    if (xpr) {
      *xpr = xp;
      xpr += (size_t)4 * N;
    }
    if (obsl) {
      const float dd = xp - ob;
      ll += -0.5f * (Lim<float>::log2pi - lprec + prec * dd * dd);
    }
    if (k + 1 < T) rk_step<LaneRhs<M>, TB>(f, t0, t1, TB::const_h ? h0 : (t1 - t0), &x);
    t0 = t1;
    t1 = t2;
    ob = obn;
  }
  if (active) {
    if (a.logp_species && sp < 4) a.logp_species[(size_t)n * 4 + sp] = ll;
    if (sp == 0) {
      if (a.logp_theta) a.logp_theta[n] = lp;
      if (a.logq_theta) a.logq_theta[n] = lq;
    }
  }
}

template <class M>
struct LaneOk {
  static constexpr bool value = sizeof(typename M::real) == 4 && !M::DYN && !M::RELAY && !M::BLACKBOX && M::NS == 8;
};

}  // namespace vh
