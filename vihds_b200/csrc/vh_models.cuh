// Model right-hand sides and their hand-written reverse-mode (VJP) counterparts, one trajectory per call.
//
// White-box "double receiver" family (dr_constant v1/v2, relay_constant, degrader_constant, +- NeuralPrecisions states):
//   forward maths   <- models/dr_constant.py:14-112, models/relay_constant.py:13-134, models/degrader_constant.py:13-146
//                      (reference, read-only); growth-only family: auto / prpr / inducer_constant (models/inducer_constant.py)
//   initial state   <- models/dr_constant.py:133-150, :178-199; models/relay_constant.py:150-178, :220-250
//   observe         <- vihds/ode.py:84-93
//   NeuralPrecisions<- vihds/precisions.py:44-94
// Everything is VH_HD so that the identical code is compiled for the device kernels and for the host-side math check
// (tests/hostcheck), which compares it with the CPU oracle without needing a GPU.
#pragma once
#include "vh_math.cuh"

namespace vh {

// ---------------------------------------------------------------------------------------------------------------
// theta slots of the double-receiver family (names as the RHS reads them off `theta`)
// ---------------------------------------------------------------------------------------------------------------
enum DrSlot : int {
  S_r = 0, S_K, S_tlag, S_rc, S_a530, S_a480, S_drfp, S_dyfp, S_dcfp, S_dR, S_dS,
  S_e76, S_e81, S_aCFP, S_aYFP, S_KGR_76, S_KGS_76, S_KGR_81, S_KGS_81, S_aR, S_aS, S_nR, S_nS,
  S_KR6, S_KR12, S_KS6, S_KS12,  // version 1
  S_eS6, S_eR12,                 // version 2
  S_init_x, S_init_rfp, S_init_yfp, S_init_cfp, S_init_luxR, S_init_lasR,
  S_prec_x, S_prec_rfp, S_prec_yfp, S_prec_cfp,  // constant precisions: prec_*; dynamic: init_prec_*
  S_dlasI, S_dluxI, S_KC6, S_KC12, S_Klux, S_Klas, S_init_luxI, S_init_lasI,  // relay
  S_aYFP_PR, S_aCFP_PR,                                                           // prpr_constant
  S_nA, S_eA, S_KAra,                                                             // arabinose promoter PBAD (inducer, degrader)
  S_aYFP_Inducer,                                                                 // inducer_constant
  S_aI, S_dA6, S_dA12, S_daiiA, S_init_aiiA,                                      // degrader_constant
  DR_NSLOT
};

static const char* const kDrSlotNames[DR_NSLOT] = {
    "r", "K", "tlag", "rc", "a530", "a480", "drfp", "dyfp", "dcfp", "dR", "dS",
    "e76", "e81", "aCFP", "aYFP", "KGR_76", "KGS_76", "KGR_81", "KGS_81", "aR", "aS", "nR", "nS",
    "KR6", "KR12", "KS6", "KS12", "eS6", "eR12",
    "init_x", "init_rfp", "init_yfp", "init_cfp", "init_luxR", "init_lasR",
    "prec_x", "prec_rfp", "prec_yfp", "prec_cfp",
    "dlasI", "dluxI", "KC6", "KC12", "Klux", "Klas", "init_luxI", "init_lasI",
    "aYFP_PR", "aCFP_PR",
    "nA", "eA", "KAra", "aYFP_Inducer", "aI", "dA6", "dA12", "daiiA", "init_aiiA"};
static const char* const kDrDynPrecNames[4] = {"init_prec_x", "init_prec_rfp", "init_prec_yfp", "init_prec_cfp"};

// per-trajectory constants of the RHS (clamped parameters, Hill fractions, pre-multiplied production rates)
enum DrConst : int {
  C_r = 0, C_K, C_tlag, C_rc, C_drfp, C_dyfp, C_dcfp, C_dR, C_dS,
  C_e76, C_e81, C_KGR76, C_KGS76, C_KGR81, C_KGS81, C_fR, C_fS,
  C_cY,  // rc * aYFP
  C_cC,  // rc * aCFP
  C_p4,  // rc * a530
  C_p5,  // rc * a480
  C_p6,  // rc * aR
  C_p7,  // rc * aS
  C_dluxI, C_dlasI, C_k6, /* KC6*rc */ C_k12, /* KC12*rc */ C_Klux, C_Klas,
  DR_NCONST,
  // degrader_constant reuses the extension block: production of AiiA rc*aI*PBAD, its constant loss, dA6*c6, dA12*c12
  C_cI = C_dluxI, C_daiiA, C_rC6, C_rC12, DG_NCONST
};

// treatments as the models read them: c = clamp(exp(x) - 1, 1e-12, 1e6) (models/dr_constant.py:26); tc[0..2] = C6, C12, Ara
// for the receiver models, tc[0] = Ara for inducer_constant
template <typename R>
VH_HD R treat_conc(R x) {
  return clampv(vexp(x) - R(1), R(1e-12), R(1e6));
}

// PBAD = (A^n + eA K^n) / (A^n + K^n), A = arabinose, n = clamp(nA, 0.5, 3) (models/inducer_constant.py:52-55,
// models/degrader_constant.py:83-86); vjp w.r.t. (nA, eA, KAra) following torch's pow backward
template <typename R>
VH_HD R pbad(R A, R nA_raw, R eA, R K) {
  const R n = clampv(nA_raw, R(0.5), R(3));
  const R An = vpow(A, n), Kn = vpow(K, n);
  return (An + eA * Kn) / (An + Kn);
}
template <typename R>
VH_HD void pbad_vjp(R A, R nA_raw, R eA, R K, R gP, R& gn_raw, R& geA, R& gK) {
  const R n = clampv(nA_raw, R(0.5), R(3));
  const R An = vpow(A, n), Kn = vpow(K, n);
  const R D = An + Kn, P = (An + eA * Kn) / D;
  const R gN = gP / D, gD = -gN * P;
  geA = gN * Kn;
  const R gKn = gN * eA + gD, gAn = gN + gD;
  gK = gKn * n * vpow(K, n - R(1));
  gn_raw = (gAn * An * vlog(A) + gKn * Kn * vlog(K)) * clampmask(nA_raw, R(0.5), R(3));
}

// EXT: 0 = dr_constant, 1 = relay_constant (+ LuxI, LasI, C6, C12), 2 = degrader_constant (+ AiiA, C6, C12)
template <typename R, int VERSION_, int EXT_, bool DYN_>
struct DrModel {
  typedef R real;
  static constexpr int VERSION = VERSION_;
  static constexpr bool RELAY = EXT_ == 1;
  static constexpr bool DEGR = EXT_ == 2;
  static constexpr bool DYN = DYN_;
  static constexpr bool BLACKBOX = false;
  static constexpr int NSLOT = DR_NSLOT;
  static constexpr int NS = RELAY ? 12 : (DEGR ? 11 : 8);  // species (OdeModel.n_species)
  static constexpr int S = NS + (DYN ? 4 : 0);             // ODE state width
  static constexpr int NC = RELAY ? DR_NCONST : (DEGR ? DG_NCONST : C_dluxI);
  static constexpr int NIN = NS + 1;                       // NeuralPrecisions inputs: [t, species]

  struct Consts {
    R v[NC];
    R iK, iKlux, iKlas;  // reciprocals of K, Klux, Klas: not differentiated themselves (their cotangent goes to v[C_K]...)
  };

  VH_HD static constexpr bool uses(int s) {
    return (s <= S_nS) || (VERSION == 1 && s >= S_KR6 && s <= S_KS12) || (VERSION == 2 && (s == S_eS6 || s == S_eR12)) ||
           (s >= S_init_x && s <= S_prec_cfp) || (RELAY && s >= S_dlasI && s <= S_init_lasI) ||
           (DEGR && (s == S_nA || s == S_eA || s == S_KAra || (s >= S_aI && s <= S_init_aiiA)));
  }
  static constexpr int NTREAT = DEGR ? 3 : 2;

  // treatments -> inducer concentrations, models/dr_constant.py:26 (degrader: + arabinose, degrader_constant.py:28-32)
  VH_HD static void treatments(const R* tr, R* tc) {
    tc[0] = treat_conc(tr[0]);
    tc[1] = treat_conc(tr[1]);
    tc[2] = DEGR ? treat_conc(tr[2]) : R(0);
  }

  VH_HD static void hill(R Ka, R Kb, R n, R c6, R c12, R& f) {  // version 1 fraction
    const R A = Ka * c6, Bq = Kb * c12;
    const R D = R(1) + A + Bq;
    f = (vpow(A, n) + vpow(Bq, n)) / vpow(D, n);
  }

  VH_HD static void setup(const R* th, const R* tc, Consts& c) {
    const R c6 = tc[0], c12 = tc[1];
    R* v = c.v;
    v[C_r] = clampv(th[S_r], R(0), R(4));
    v[C_K] = clampv(th[S_K], R(0), R(4));
    v[C_tlag] = th[S_tlag];
    v[C_rc] = th[S_rc];
    v[C_drfp] = clampv(th[S_drfp], R(1e-12), R(2));
    v[C_dyfp] = clampv(th[S_dyfp], R(1e-12), R(2));
    v[C_dcfp] = clampv(th[S_dcfp], R(1e-12), R(2));
    v[C_dR] = clampv(th[S_dR], R(1e-12), R(5));
    v[C_dS] = clampv(th[S_dS], R(1e-12), R(5));
    v[C_e76] = th[S_e76];
    v[C_e81] = th[S_e81];
    v[C_KGR76] = th[S_KGR_76];
    v[C_KGS76] = th[S_KGS_76];
    v[C_KGR81] = th[S_KGR_81];
    v[C_KGS81] = th[S_KGS_81];
    const R nR = clampv(th[S_nR], R(0.5), R(3)), nS = clampv(th[S_nS], R(0.5), R(3));
    if (VERSION == 1) {
      const R lb = R(1e-12), ub = R(1);
      hill(clampv(th[S_KR6], lb, ub), clampv(th[S_KR12], lb, ub), nR, c6, c12, v[C_fR]);
      hill(clampv(th[S_KS6], lb, ub), clampv(th[S_KS12], lb, ub), nS, c6, c12, v[C_fS]);
    } else {
      const R eS6 = clampv(th[S_eS6], R(1e-12), R(1)), eR12 = clampv(th[S_eR12], R(1e-12), R(1));
      v[C_fR] = vpow(c6, nR) + vpow(eR12 * c12, nR);
      v[C_fS] = vpow(eS6 * c6, nS) + vpow(c12, nS);
    }
    v[C_cY] = th[S_rc] * th[S_aYFP];
    v[C_cC] = th[S_rc] * th[S_aCFP];
    v[C_p4] = th[S_rc] * th[S_a530];
    v[C_p5] = th[S_rc] * th[S_a480];
    v[C_p6] = th[S_rc] * th[S_aR];
    v[C_p7] = th[S_rc] * th[S_aS];
    if (RELAY) {
      v[C_dluxI] = clampv(th[S_dluxI], R(1e-12), R(5));
      v[C_dlasI] = clampv(th[S_dlasI], R(1e-12), R(5));
      v[C_k6] = th[S_KC6] * th[S_rc];
      v[C_k12] = th[S_KC12] * th[S_rc];
      v[C_Klux] = th[S_Klux];
      v[C_Klas] = th[S_Klas];
      c.iKlux = R(1) / v[C_Klux];
      c.iKlas = R(1) / v[C_Klas];
    } else {
      c.iKlux = c.iKlas = R(0);
    }
    if (DEGR) {  // models/degrader_constant.py:76-88
      v[C_cI] = th[S_rc] * th[S_aI] * pbad(tc[2], th[S_nA], th[S_eA], th[S_KAra]);
      v[C_daiiA] = th[S_daiiA];
      v[C_rC6] = th[S_dA6] * c6;
      v[C_rC12] = th[S_dA12] * c12;
    }
    c.iK = R(1) / v[C_K];
  }

  // d f / d (Ka, Kb, n) for the version-1 Hill fraction, following torch's pow backward
  VH_HD static void hill_vjp(R Ka, R Kb, R n, R c6, R c12, R f, R gf, R& gKa, R& gKb, R& gn) {
    const R A = Ka * c6, Bq = Kb * c12;
    const R D = R(1) + A + Bq;
    const R An = vpow(A, n), Bn = vpow(Bq, n), Dn = vpow(D, n);
    const R gnum = gf / Dn;
    const R gDn = -gnum * f;
    const R gD = gDn * n * vpow(D, n - R(1));
    const R gA = gnum * n * vpow(A, n - R(1)) + gD;
    const R gB = gnum * n * vpow(Bq, n - R(1)) + gD;
    gn = gnum * (An * vlog(A) + Bn * vlog(Bq)) + gDn * Dn * vlog(D);
    gKa = gA * c6;
    gKb = gB * c12;
  }

  // gc: cotangent of the constants  ->  gth: cotangent of the theta slots (accumulated).
  // parts (compile-time constant at every call site): the two Hill fractions carry most of the cost (six powf + three logf
  // each) and are independent of each other and of the rest, so a kernel may give them to different warps:
  // 1 = the LuxR fraction (fR), 2 = the LasR fraction (fS), 4 = everything else; 7 = all.
  VH_HD static void setup_vjp(const R* th, const R* tc, const Consts& c, const Consts& gc, R* gth, int parts = 7) {
    const R c6 = tc[0], c12 = tc[1];
    const R* g = gc.v;
    if (parts & 4) {
    gth[S_r] += g[C_r] * clampmask(th[S_r], R(0), R(4));
    gth[S_K] += g[C_K] * clampmask(th[S_K], R(0), R(4));
    gth[S_tlag] += g[C_tlag];
    gth[S_drfp] += g[C_drfp] * clampmask(th[S_drfp], R(1e-12), R(2));
    gth[S_dyfp] += g[C_dyfp] * clampmask(th[S_dyfp], R(1e-12), R(2));
    gth[S_dcfp] += g[C_dcfp] * clampmask(th[S_dcfp], R(1e-12), R(2));
    gth[S_dR] += g[C_dR] * clampmask(th[S_dR], R(1e-12), R(5));
    gth[S_dS] += g[C_dS] * clampmask(th[S_dS], R(1e-12), R(5));
    gth[S_e76] += g[C_e76];
    gth[S_e81] += g[C_e81];
    gth[S_KGR_76] += g[C_KGR76];
    gth[S_KGS_76] += g[C_KGS76];
    gth[S_KGR_81] += g[C_KGR81];
    gth[S_KGS_81] += g[C_KGS81];
    const R rc = th[S_rc];
    R grc = g[C_rc] + g[C_cY] * th[S_aYFP] + g[C_cC] * th[S_aCFP] + g[C_p4] * th[S_a530] + g[C_p5] * th[S_a480] +
            g[C_p6] * th[S_aR] + g[C_p7] * th[S_aS];
    gth[S_aYFP] += g[C_cY] * rc;
    gth[S_aCFP] += g[C_cC] * rc;
    gth[S_a530] += g[C_p4] * rc;
    gth[S_a480] += g[C_p5] * rc;
    gth[S_aR] += g[C_p6] * rc;
    gth[S_aS] += g[C_p7] * rc;
    gth[S_rc] += grc;
    }  // parts & 4
    const R nR = clampv(th[S_nR], R(0.5), R(3)), nS = clampv(th[S_nS], R(0.5), R(3));
    R gnR = R(0), gnS = R(0);
    if (VERSION == 1) {
      const R lb = R(1e-12), ub = R(1);
      R ga, gb;
      if (parts & 1) {
        hill_vjp(clampv(th[S_KR6], lb, ub), clampv(th[S_KR12], lb, ub), nR, c6, c12, c.v[C_fR], g[C_fR], ga, gb, gnR);
        gth[S_KR6] += ga * clampmask(th[S_KR6], lb, ub);
        gth[S_KR12] += gb * clampmask(th[S_KR12], lb, ub);
      }
      if (parts & 2) {
        hill_vjp(clampv(th[S_KS6], lb, ub), clampv(th[S_KS12], lb, ub), nS, c6, c12, c.v[C_fS], g[C_fS], ga, gb, gnS);
        gth[S_KS6] += ga * clampmask(th[S_KS6], lb, ub);
        gth[S_KS12] += gb * clampmask(th[S_KS12], lb, ub);
      }
    } else {
      const R eS6 = clampv(th[S_eS6], R(1e-12), R(1)), eR12 = clampv(th[S_eR12], R(1e-12), R(1));
      const R E = eR12 * c12, F = eS6 * c6;
      if (parts & 1) {
        gnR = g[C_fR] * (vpow(c6, nR) * vlog(c6) + vpow(E, nR) * vlog(E));
        gth[S_eR12] += g[C_fR] * nR * vpow(E, nR - R(1)) * c12 * clampmask(th[S_eR12], R(1e-12), R(1));
      }
      if (parts & 2) {
        gnS = g[C_fS] * (vpow(F, nS) * vlog(F) + vpow(c12, nS) * vlog(c12));
        gth[S_eS6] += g[C_fS] * nS * vpow(F, nS - R(1)) * c6 * clampmask(th[S_eS6], R(1e-12), R(1));
      }
    }
    if (parts & 1) gth[S_nR] += gnR * clampmask(th[S_nR], R(0.5), R(3));
    if (parts & 2) gth[S_nS] += gnS * clampmask(th[S_nS], R(0.5), R(3));
    if (!(parts & 4)) return;
    R grc = R(0);
    const R rc = th[S_rc];
    if (RELAY) {
      gth[S_dluxI] += g[C_dluxI] * clampmask(th[S_dluxI], R(1e-12), R(5));
      gth[S_dlasI] += g[C_dlasI] * clampmask(th[S_dlasI], R(1e-12), R(5));
      gth[S_KC6] += g[C_k6] * rc;
      gth[S_KC12] += g[C_k12] * rc;
      grc += g[C_k6] * th[S_KC6] + g[C_k12] * th[S_KC12];
      gth[S_Klux] += g[C_Klux];
      gth[S_Klas] += g[C_Klas];
    }
    if (DEGR) {
      const R P = pbad(tc[2], th[S_nA], th[S_eA], th[S_KAra]);
      grc += g[C_cI] * th[S_aI] * P;
      gth[S_aI] += g[C_cI] * rc * P;
      R gn, ge, gk;
      pbad_vjp(tc[2], th[S_nA], th[S_eA], th[S_KAra], g[C_cI] * rc * th[S_aI], gn, ge, gk);
      gth[S_nA] += gn;
      gth[S_eA] += ge;
      gth[S_KAra] += gk;
      gth[S_daiiA] += g[C_daiiA];
      gth[S_dA6] += g[C_rC6] * c6;
      gth[S_dA12] += g[C_rC12] * c12;
    }
    gth[S_rc] += grc;
  }

  VH_HD static void init_state(const R* th, const R* tc, R* x) {
    const R c6 = tc[0], c12 = tc[1];
    x[0] = th[S_init_x];
    x[1] = th[S_init_rfp];
    x[2] = th[S_init_yfp];
    x[3] = th[S_init_cfp];
    x[4] = R(0);
    x[5] = R(0);
    x[6] = th[S_init_luxR];
    x[7] = th[S_init_lasR];
    if (RELAY) {
      x[8] = th[S_init_luxI];
      x[9] = th[S_init_lasI];
      x[10] = c6;
      x[11] = c12;
    }
    if (DEGR) {  // models/degrader_constant.py:180-196
      x[8] = th[S_init_aiiA];
      x[9] = c6;
      x[10] = c12;
    }
    if (DYN) {
#pragma unroll
      for (int o = 0; o < 4; ++o) x[NS + o] = th[S_prec_x + o];
    }
  }

  VH_HD static void init_state_vjp(const R* gx, R* gth) {
    if (DEGR) gth[S_init_aiiA] += gx[8];
    gth[S_init_x] += gx[0];
    gth[S_init_rfp] += gx[1];
    gth[S_init_yfp] += gx[2];
    gth[S_init_cfp] += gx[3];
    gth[S_init_luxR] += gx[6];
    gth[S_init_lasR] += gx[7];
    if (RELAY) {
      gth[S_init_luxI] += gx[8];
      gth[S_init_lasI] += gx[9];
    }
    if (DYN) {
#pragma unroll
      for (int o = 0; o < 4; ++o) gth[S_prec_x + o] += gx[NS + o];
    }
  }

  // intermediates shared by rhs and rhs_vjp (the reverse sweep keeps them between the stage evaluation and its vjp)
  struct Mid {
    R sg, gr, g, gam, bR, bS, i76, i81, P76, P81;
  };

  VH_HD static void mid(R t, const R* x, const Consts& c, Mid& m) {
    const R* v = c.v;
    m.sg = sigmoid(R(4) * (t - v[C_tlag]));
    m.gr = v[C_r] * m.sg;
    m.g = R(1) - x[0] * c.iK;
    m.gam = m.gr * m.g;
    m.bR = x[6] * x[6] * v[C_fR];
    m.bS = x[7] * x[7] * v[C_fS];
    const R a76 = v[C_KGR76] * m.bR, b76 = v[C_KGS76] * m.bS;
    const R a81 = v[C_KGR81] * m.bR, b81 = v[C_KGS81] * m.bS;
    m.i76 = vdiv(R(1), R(1) + a76 + b76);
    m.i81 = vdiv(R(1), R(1) + a81 + b81);
    m.P76 = (v[C_e76] + a76 + b76) * m.i76;
    m.P81 = (v[C_e81] + a81 + b81) * m.i81;
  }

  // species part of the right-hand side (dx[0..NS)) from precomputed intermediates
  VH_HD static void rhs_from(const R* x, const Consts& c, const Mid& m, R* dx) {
    const R* v = c.v;
    dx[0] = m.gam * x[0];
    dx[1] = v[C_rc] - (m.gam + v[C_drfp]) * x[1];
    dx[2] = v[C_cY] * m.P81 - (m.gam + v[C_dyfp]) * x[2];
    dx[3] = v[C_cC] * m.P76 - (m.gam + v[C_dcfp]) * x[3];
    dx[4] = v[C_p4] - m.gam * x[4];
    dx[5] = v[C_p5] - m.gam * x[5];
    dx[6] = v[C_p6] - (m.gam + v[C_dR]) * x[6];
    dx[7] = v[C_p7] - (m.gam + v[C_dS]) * x[7];
    if (RELAY) {
      dx[8] = v[C_rc] * m.P81 - (m.gam + v[C_dluxI]) * x[8];
      dx[9] = v[C_rc] * m.P76 - (m.gam + v[C_dlasI]) * x[9];
      dx[10] = vdiv(v[C_k6] * x[0] * x[8], R(1) + x[8] * c.iKlux);
      dx[11] = vdiv(v[C_k12] * x[0] * x[9], R(1) + x[9] * c.iKlas);
    }
    if (DEGR) {  // models/degrader_constant.py:128-131 (as written there: the constant loss is not multiplied by AiiA)
      dx[8] = v[C_cI] - (v[C_daiiA] + m.gam * x[8]);
      dx[9] = x[0] * v[C_rC6] * x[8];
      dx[10] = x[0] * v[C_rC12] * x[8];
    }
  }

  VH_HD static void rhs(R t, const R* x, const Consts& c, R* dx) {
    Mid m;
    mid(t, x, c, m);
    rhs_from(x, c, m, dx);
  }

  // g: cotangent of dx[0..NS)  ->  gx (accumulated), gc (accumulated); m = mid(t, x, c)
  VH_HD static void rhs_vjp_from(const R* x, const Consts& c, const Mid& m, const R* g, R* gx, Consts& gcs) {
    const R* v = c.v;
    R* gc = gcs.v;
    R ggam = g[0] * x[0] - g[1] * x[1] - g[2] * x[2] - g[3] * x[3] - g[4] * x[4] - g[5] * x[5] - g[6] * x[6] - g[7] * x[7];
    gx[0] += g[0] * m.gam;
    gx[1] -= g[1] * (m.gam + v[C_drfp]);
    gx[2] -= g[2] * (m.gam + v[C_dyfp]);
    gx[3] -= g[3] * (m.gam + v[C_dcfp]);
    gx[4] -= g[4] * m.gam;
    gx[5] -= g[5] * m.gam;
    gx[6] -= g[6] * (m.gam + v[C_dR]);
    gx[7] -= g[7] * (m.gam + v[C_dS]);
    gc[C_drfp] -= g[1] * x[1];
    gc[C_dyfp] -= g[2] * x[2];
    gc[C_dcfp] -= g[3] * x[3];
    gc[C_dR] -= g[6] * x[6];
    gc[C_dS] -= g[7] * x[7];
    gc[C_rc] += g[1];
    gc[C_cY] += g[2] * m.P81;
    gc[C_cC] += g[3] * m.P76;
    gc[C_p4] += g[4];
    gc[C_p5] += g[5];
    gc[C_p6] += g[6];
    gc[C_p7] += g[7];
    R gP81 = g[2] * v[C_cY];
    R gP76 = g[3] * v[C_cC];
    if (RELAY) {
      ggam -= g[8] * x[8] + g[9] * x[9];
      gx[8] -= g[8] * (m.gam + v[C_dluxI]);
      gx[9] -= g[9] * (m.gam + v[C_dlasI]);
      gc[C_dluxI] -= g[8] * x[8];
      gc[C_dlasI] -= g[9] * x[9];
      gc[C_rc] += g[8] * m.P81 + g[9] * m.P76;
      gP81 += g[8] * v[C_rc];
      gP76 += g[9] * v[C_rc];
      {
        const R iden = vdiv(R(1), R(1) + x[8] * c.iKlux);
        const R gnum = g[10] * iden;
        const R val = (v[C_k6] * x[0] * x[8]) * iden;
        const R gden = -gnum * val;
        gc[C_k6] += gnum * x[0] * x[8];
        gx[0] += gnum * v[C_k6] * x[8];
        gx[8] += gnum * v[C_k6] * x[0] + gden * c.iKlux;
        gc[C_Klux] -= gden * x[8] * (c.iKlux * c.iKlux);
      }
      {
        const R iden = vdiv(R(1), R(1) + x[9] * c.iKlas);
        const R gnum = g[11] * iden;
        const R val = (v[C_k12] * x[0] * x[9]) * iden;
        const R gden = -gnum * val;
        gc[C_k12] += gnum * x[0] * x[9];
        gx[0] += gnum * v[C_k12] * x[9];
        gx[9] += gnum * v[C_k12] * x[0] + gden * c.iKlas;
        gc[C_Klas] -= gden * x[9] * (c.iKlas * c.iKlas);
      }
    }
    if (DEGR) {
      ggam -= g[8] * x[8];
      gx[8] -= g[8] * m.gam;
      gc[C_cI] += g[8];
      gc[C_daiiA] -= g[8];
      const R q = g[9] * v[C_rC6] + g[10] * v[C_rC12];
      gx[0] += q * x[8];
      gx[8] += q * x[0];
      gc[C_rC6] += g[9] * x[0] * x[8];
      gc[C_rC12] += g[10] * x[0] * x[8];
    }
    // gamma = gr * g,  gr = r * sg,  g = 1 - x0 / K
    const R ggr = ggam * m.g, gg = ggam * m.gr;
    gc[C_r] += ggr * m.sg;
    gc[C_tlag] -= R(4) * ggr * v[C_r] * m.sg * (R(1) - m.sg);
    gx[0] -= gg * c.iK;
    gc[C_K] += gg * x[0] * (c.iK * c.iK);
    // promoter activities P = (e + a + b) / (1 + a + b)
    const R gn76 = gP76 * m.i76, gn81 = gP81 * m.i81;
    gc[C_e76] += gn76;
    gc[C_e81] += gn81;
    const R ga76 = gn76 * (R(1) - m.P76), ga81 = gn81 * (R(1) - m.P81);
    gc[C_KGR76] += ga76 * m.bR;
    gc[C_KGS76] += ga76 * m.bS;
    gc[C_KGR81] += ga81 * m.bR;
    gc[C_KGS81] += ga81 * m.bS;
    const R gbR = ga76 * v[C_KGR76] + ga81 * v[C_KGR81];
    const R gbS = ga76 * v[C_KGS76] + ga81 * v[C_KGS81];
    gx[6] += gbR * R(2) * x[6] * v[C_fR];
    gx[7] += gbS * R(2) * x[7] * v[C_fS];
    gc[C_fR] += gbR * x[6] * x[6];
    gc[C_fS] += gbS * x[7] * x[7];
  }

  VH_HD static void rhs_vjp(R t, const R* x, const Consts& c, const R* g, R* gx, Consts& gcs) {
    Mid m;
    mid(t, x, c, m);
    rhs_vjp_from(x, c, m, g, gx, gcs);
  }

  // observe, vihds/ode.py:84-93
  VH_HD static void observe(const R* x, R* xp) {
    xp[0] = x[0];
    xp[1] = vmul_rn(x[0], x[1]);
    xp[2] = vmul_rn(x[0], x[2] + x[4]);
    xp[3] = vmul_rn(x[0], x[3] + x[5]);
  }
  VH_HD static void observe_vjp(const R* x, const R* gxp, R* gx) {
    gx[0] += gxp[0] + gxp[1] * x[1] + gxp[2] * (x[2] + x[4]) + gxp[3] * (x[3] + x[5]);
    gx[1] += gxp[1] * x[0];
    gx[2] += gxp[2] * x[0];
    gx[4] += gxp[2] * x[0];
    gx[3] += gxp[3] * x[0];
    gx[5] += gxp[3] * x[0];
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Growth-only family: auto_constant (NSP = 4: OD, RFP, F530, F480; models/auto_constant.py:11-60) and prpr_constant
// (NSP = 6: + YFP, CFP expressed constitutively; models/prpr_constant.py:11-58).  The double-receiver right-hand side
// with the receiver / promoter terms removed; same slot names, same interface as DrModel.
// ---------------------------------------------------------------------------------------------------------------
// inducer_constant (NSP = 5: OD, RFP, YFP, F530, F480; models/inducer_constant.py:12-84): YFP expressed from PBAD, a
// per-trajectory constant of the arabinose treatment: G_cY = rc * aYFP_Inducer * PBAD.
enum GrConst : int { G_r = 0, G_K, G_tlag, G_rc, G_drfp, G_dyfp, G_dcfp, G_cY, G_cC, G_p530, G_p480, GR_NCONST };

template <typename R, int NSP_, bool DYN_>
struct GrowthModel {
  typedef R real;
  static constexpr bool DYN = DYN_, RELAY = false, BLACKBOX = false;
  static constexpr bool PRPR = NSP_ == 6;
  static constexpr bool IND = NSP_ == 5;
  static constexpr int NSLOT = DR_NSLOT;
  static constexpr int NS = NSP_;
  static constexpr int S = NS + (DYN ? 4 : 0);
  static constexpr int NC = GR_NCONST;
  static constexpr int NIN = NS + 1;
  static constexpr int NTREAT = IND ? 1 : 0;
  // state index of the autofluorescence species
  static constexpr int I530 = PRPR ? 4 : (IND ? 3 : 2), I480 = PRPR ? 5 : (IND ? 4 : 3);
  static_assert(NSP_ == 4 || NSP_ == 6 || NSP_ == 5, "auto_constant has 4 species, inducer_constant 5, prpr_constant 6");

  struct Consts {
    R v[NC];
    R iK;
  };
  struct Mid {
    R sg, gr, g, gam;
  };

  VH_HD static constexpr bool uses(int s) {
    return s == S_r || s == S_K || s == S_tlag || s == S_rc || s == S_a530 || s == S_a480 || s == S_drfp || s == S_init_x ||
           s == S_init_rfp || (s >= S_prec_x && s <= S_prec_cfp) ||
           (PRPR && (s == S_dyfp || s == S_dcfp || s == S_aYFP_PR || s == S_aCFP_PR || s == S_init_yfp || s == S_init_cfp)) ||
           (IND && (s == S_dyfp || s == S_aYFP_Inducer || s == S_nA || s == S_eA || s == S_KAra || s == S_init_yfp));
  }
  // treatments do not enter auto / prpr; inducer: arabinose (models/inducer_constant.py:28)
  VH_HD static void treatments(const R* tr, R* tc) {
    tc[0] = IND ? treat_conc(tr[0]) : R(0);
    tc[1] = tc[2] = R(0);
  }

  VH_HD static void setup(const R* th, const R* tc, Consts& c) {
    R* v = c.v;
    v[G_r] = clampv(th[S_r], R(0), R(4));
    v[G_K] = clampv(th[S_K], R(0), R(4));
    v[G_tlag] = th[S_tlag];
    v[G_rc] = th[S_rc];
    v[G_drfp] = clampv(th[S_drfp], R(1e-12), R(2));
    v[G_dyfp] = (PRPR || IND) ? clampv(th[S_dyfp], R(1e-12), R(2)) : R(0);
    v[G_dcfp] = PRPR ? clampv(th[S_dcfp], R(1e-12), R(2)) : R(0);
    v[G_cY] = PRPR ? th[S_rc] * th[S_aYFP_PR] : (IND ? th[S_rc] * th[S_aYFP_Inducer] * pbad(tc[0], th[S_nA], th[S_eA], th[S_KAra]) : R(0));
    v[G_cC] = PRPR ? th[S_rc] * th[S_aCFP_PR] : R(0);
    v[G_p530] = th[S_rc] * th[S_a530];
    v[G_p480] = th[S_rc] * th[S_a480];
    c.iK = R(1) / v[G_K];
  }
  VH_HD static void setup_vjp(const R* th, const R* tc, const Consts&, const Consts& gc, R* gth) {
    const R* g = gc.v;
    gth[S_r] += g[G_r] * clampmask(th[S_r], R(0), R(4));
    gth[S_K] += g[G_K] * clampmask(th[S_K], R(0), R(4));
    gth[S_tlag] += g[G_tlag];
    gth[S_drfp] += g[G_drfp] * clampmask(th[S_drfp], R(1e-12), R(2));
    R grc = g[G_rc] + g[G_p530] * th[S_a530] + g[G_p480] * th[S_a480];
    gth[S_a530] += g[G_p530] * th[S_rc];
    gth[S_a480] += g[G_p480] * th[S_rc];
    if (PRPR) {
      gth[S_dyfp] += g[G_dyfp] * clampmask(th[S_dyfp], R(1e-12), R(2));
      gth[S_dcfp] += g[G_dcfp] * clampmask(th[S_dcfp], R(1e-12), R(2));
      grc += g[G_cY] * th[S_aYFP_PR] + g[G_cC] * th[S_aCFP_PR];
      gth[S_aYFP_PR] += g[G_cY] * th[S_rc];
      gth[S_aCFP_PR] += g[G_cC] * th[S_rc];
    }
    if (IND) {
      gth[S_dyfp] += g[G_dyfp] * clampmask(th[S_dyfp], R(1e-12), R(2));
      const R P = pbad(tc[0], th[S_nA], th[S_eA], th[S_KAra]);
      grc += g[G_cY] * th[S_aYFP_Inducer] * P;
      gth[S_aYFP_Inducer] += g[G_cY] * th[S_rc] * P;
      R gn, ge, gk;
      pbad_vjp(tc[0], th[S_nA], th[S_eA], th[S_KAra], g[G_cY] * th[S_rc] * th[S_aYFP_Inducer], gn, ge, gk);
      gth[S_nA] += gn;
      gth[S_eA] += ge;
      gth[S_KAra] += gk;
    }
    gth[S_rc] += grc;
  }

  VH_HD static void init_state(const R* th, const R*, R* x) {
    x[0] = th[S_init_x];
    x[1] = th[S_init_rfp];
    if (PRPR) {
      x[2] = th[S_init_yfp];
      x[3] = th[S_init_cfp];
    }
    if (IND) x[2] = th[S_init_yfp];
    x[I530] = R(0);
    x[I480] = R(0);
    if (DYN) {
#pragma unroll
      for (int o = 0; o < 4; ++o) x[NS + o] = th[S_prec_x + o];
    }
  }
  VH_HD static void init_state_vjp(const R* gx, R* gth) {
    gth[S_init_x] += gx[0];
    gth[S_init_rfp] += gx[1];
    if (PRPR) {
      gth[S_init_yfp] += gx[2];
      gth[S_init_cfp] += gx[3];
    }
    if (IND) gth[S_init_yfp] += gx[2];
    if (DYN) {
#pragma unroll
      for (int o = 0; o < 4; ++o) gth[S_prec_x + o] += gx[NS + o];
    }
  }

  VH_HD static void mid(R t, const R* x, const Consts& c, Mid& m) {
    m.sg = sigmoid(R(4) * (t - c.v[G_tlag]));
    m.gr = c.v[G_r] * m.sg;
    m.g = R(1) - x[0] * c.iK;
    m.gam = m.gr * m.g;
  }
  VH_HD static void rhs_from(const R* x, const Consts& c, const Mid& m, R* dx) {
    const R* v = c.v;
    dx[0] = m.gam * x[0];
    dx[1] = v[G_rc] - (m.gam + v[G_drfp]) * x[1];
    if (PRPR) {
      dx[2] = v[G_cY] - (m.gam + v[G_dyfp]) * x[2];
      dx[3] = v[G_cC] - (m.gam + v[G_dcfp]) * x[3];
    }
    if (IND) dx[2] = v[G_cY] - (m.gam + v[G_dyfp]) * x[2];
    dx[I530] = v[G_p530] - m.gam * x[I530];
    dx[I480] = v[G_p480] - m.gam * x[I480];
  }
  VH_HD static void rhs(R t, const R* x, const Consts& c, R* dx) {
    Mid m;
    mid(t, x, c, m);
    rhs_from(x, c, m, dx);
  }
  VH_HD static void rhs_vjp_from(const R* x, const Consts& c, const Mid& m, const R* g, R* gx, Consts& gcs) {
    const R* v = c.v;
    R* gc = gcs.v;
    R ggam = g[0] * x[0] - g[1] * x[1] - g[I530] * x[I530] - g[I480] * x[I480];
    gx[0] += g[0] * m.gam;
    gx[1] -= g[1] * (m.gam + v[G_drfp]);
    gx[I530] -= g[I530] * m.gam;
    gx[I480] -= g[I480] * m.gam;
    gc[G_drfp] -= g[1] * x[1];
    gc[G_rc] += g[1];
    gc[G_p530] += g[I530];
    gc[G_p480] += g[I480];
    if (PRPR) {
      ggam -= g[2] * x[2] + g[3] * x[3];
      gx[2] -= g[2] * (m.gam + v[G_dyfp]);
      gx[3] -= g[3] * (m.gam + v[G_dcfp]);
      gc[G_dyfp] -= g[2] * x[2];
      gc[G_dcfp] -= g[3] * x[3];
      gc[G_cY] += g[2];
      gc[G_cC] += g[3];
    }
    if (IND) {
      ggam -= g[2] * x[2];
      gx[2] -= g[2] * (m.gam + v[G_dyfp]);
      gc[G_dyfp] -= g[2] * x[2];
      gc[G_cY] += g[2];
    }
    const R ggr = ggam * m.g, gg = ggam * m.gr;
    gc[G_r] += ggr * m.sg;
    gc[G_tlag] -= R(4) * ggr * v[G_r] * m.sg * (R(1) - m.sg);
    gx[0] -= gg * c.iK;
    gc[G_K] += gg * x[0] * (c.iK * c.iK);
  }
  VH_HD static void rhs_vjp(R t, const R* x, const Consts& c, const R* g, R* gx, Consts& gcs) {
    Mid m;
    mid(t, x, c, m);
    rhs_vjp_from(x, c, m, g, gx, gcs);
  }

  // prpr: vihds/ode.py:84-93 (default observe); auto: models/auto_constant.py:81-89; inducer: inducer_constant.py:102-110
  VH_HD static void observe(const R* x, R* xp) {
    xp[0] = x[0];
    xp[1] = vmul_rn(x[0], x[1]);
    if (PRPR) {
      xp[2] = vmul_rn(x[0], x[2] + x[4]);
      xp[3] = vmul_rn(x[0], x[3] + x[5]);
    } else if (IND) {
      xp[2] = vmul_rn(x[0], x[2] + x[3]);
      xp[3] = vmul_rn(x[0], x[4]);
    } else {
      xp[2] = vmul_rn(x[0], x[2]);
      xp[3] = vmul_rn(x[0], x[3]);
    }
  }
  VH_HD static void observe_vjp(const R* x, const R* gxp, R* gx) {
    if (PRPR) {
      gx[0] += gxp[0] + gxp[1] * x[1] + gxp[2] * (x[2] + x[4]) + gxp[3] * (x[3] + x[5]);
      gx[1] += gxp[1] * x[0];
      gx[2] += gxp[2] * x[0];
      gx[4] += gxp[2] * x[0];
      gx[3] += gxp[3] * x[0];
      gx[5] += gxp[3] * x[0];
    } else if (IND) {
      gx[0] += gxp[0] + gxp[1] * x[1] + gxp[2] * (x[2] + x[3]) + gxp[3] * x[4];
      gx[1] += gxp[1] * x[0];
      gx[2] += gxp[2] * x[0];
      gx[3] += gxp[2] * x[0];
      gx[4] += gxp[3] * x[0];
    } else {
      gx[0] += gxp[0] + gxp[1] * x[1] + gxp[2] * x[2] + gxp[3] * x[3];
      gx[1] += gxp[1] * x[0];
      gx[2] += gxp[2] * x[0];
      gx[3] += gxp[3] * x[0];
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// NeuralPrecisions with no hidden layer (n_hidden_decoder_precisions: 0; vihds/precisions.py:55-61, :76-87):
//   a = tanh([t, species]);  prod = sigmoid(Wp a + bp);  degr = sigmoid(Wd a + bd);  dv = prod - degr * v
// flat weight layout: Wp[4][NIN], bp[4], Wd[4][NIN], bd[4]
// ---------------------------------------------------------------------------------------------------------------
template <typename R, int NIN>
struct LinPrecNet {
  static constexpr int NW = 2 * (4 * NIN + 4);
  // what an evaluation leaves for its VJP: the tanh'd inputs and the eight sigmoids (recomputing them was 13 tanh + 8
  // sigmoids per VJP at relay's 13 inputs -- the bulk of the reverse sweep of the *_precisions models)
  struct Kept {
    R a[NIN], sp[4], sd[4];
  };
  // species = xin[1..NIN), v = current precision states
  VH_HD static void rhs_keep(R t, const R* species, const R* v, const R* w, R* dv, Kept& k) {
    k.a[0] = vtanh(t);
#pragma unroll
    for (int j = 1; j < NIN; ++j) k.a[j] = vtanh(species[j - 1]);
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      R zp = w[4 * NIN + o], zd = w[(4 * NIN + 4) + 4 * NIN + o];
#pragma unroll
      for (int j = 0; j < NIN; ++j) {
        zp += w[o * NIN + j] * k.a[j];
        zd += w[(4 * NIN + 4) + o * NIN + j] * k.a[j];
      }
      k.sp[o] = sigmoid(zp);
      k.sd[o] = sigmoid(zd);
      if (dv) dv[o] = k.sp[o] - k.sd[o] * v[o];
    }
  }
  VH_HD static void rhs(R t, const R* species, const R* v, const R* w, R* dv) {
    Kept k;
    rhs_keep(t, species, v, w, dv, k);
  }
  // gw: this thread's weight-gradient accumulators, element k at gw[k * gstride]
  template <typename GW>
  VH_HD static void rhs_vjp_kept(const Kept& k, const R* v, const R* w, const R* g, R* gspecies, R* gv, GW& gw) {
    R ga[NIN];
#pragma unroll
    for (int j = 0; j < NIN; ++j) ga[j] = R(0);
    R gzp[4], gzd[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const R sp = k.sp[o], sd = k.sd[o];
      gv[o] -= g[o] * sd;
      gzp[o] = g[o] * sp * (R(1) - sp);
      gzd[o] = -g[o] * v[o] * sd * (R(1) - sd);
#pragma unroll
      for (int j = 0; j < NIN; ++j) ga[j] += gzp[o] * w[o * NIN + j] + gzd[o] * w[(4 * NIN + 4) + o * NIN + j];
    }
    gw.template outer<NIN>(k.a, gzp, gzd);  // weight gradients (in place, or handed to another warp)
#pragma unroll
    for (int j = 1; j < NIN; ++j) gspecies[j - 1] += ga[j] * (R(1) - k.a[j] * k.a[j]);
  }
  template <typename GW>
  VH_HD static void rhs_vjp(R t, const R* species, const R* v, const R* w, const R* g, R* gspecies, R* gv, GW& gw) {
    Kept k;
    rhs_keep(t, species, v, w, (R*)nullptr, k);
    rhs_vjp_kept(k, v, w, g, gspecies, gv, gw);
  }
};

// ---------------------------------------------------------------------------------------------------------------
// NeuralPrecisions WITH a hidden layer (n_hidden_decoder_precisions = H >= 1; vihds/precisions.py:63-74, :76-87;
// reachable through --precision_hidden_layers, run_xval.py:38, config.py:157-158):
//   h = tanh(W1 [t, species] + b1);  prod = sigmoid(Wp h + bp);  degr = sigmoid(Wd h + bd);  dv = prod - degr * v
// (no activation on the inputs in this branch: nn.Sequential(prec_hidden, act, prec_production, Sigmoid)).
// flat weight layout: W1[H][NIN], b1[H], Wp[4][H], bp[4], Wd[4][H], bd[4].  H is a run-time value (<= MAXH).
// ---------------------------------------------------------------------------------------------------------------
template <typename R, int NIN>
struct HidPrecNet {
  static constexpr int MAXH = 32;
  VH_HD static int num_weights(int H) { return H * (NIN + 1) + 2 * (4 * H + 4); }
  VH_HD static void hidden(R t, const R* species, const R* w, int H, R* h) {
    for (int k = 0; k < H; ++k) {
      const R* w1 = w + k * NIN;
      R a = w[H * NIN + k] + w1[0] * t;
#pragma unroll
      for (int j = 1; j < NIN; ++j) a += w1[j] * species[j - 1];
      h[k] = vtanh(a);
    }
  }
  VH_HD static void rhs(R t, const R* species, const R* v, const R* w, int H, R* dv) {
    R h[MAXH];
    hidden(t, species, w, H, h);
    const R* wp = w + H * (NIN + 1);
    const R* wd = wp + 4 * H + 4;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      R zp = wp[4 * H + o], zd = wd[4 * H + o];
      for (int k = 0; k < H; ++k) {
        zp += wp[o * H + k] * h[k];
        zd += wd[o * H + k] * h[k];
      }
      dv[o] = sigmoid(zp) - sigmoid(zd) * v[o];
    }
  }
  // gw.add(k, value): element k of the flat weight gradient
  template <typename GW>
  VH_HD static void rhs_vjp(R t, const R* species, const R* v, const R* w, int H, const R* g, R* gspecies, R* gv, GW& gw) {
    R h[MAXH], gh[MAXH];
    hidden(t, species, w, H, h);
    for (int k = 0; k < H; ++k) gh[k] = R(0);
    const int oP = H * (NIN + 1), oD = oP + 4 * H + 4;
    const R* wp = w + oP;
    const R* wd = w + oD;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      R zp = wp[4 * H + o], zd = wd[4 * H + o];
      for (int k = 0; k < H; ++k) {
        zp += wp[o * H + k] * h[k];
        zd += wd[o * H + k] * h[k];
      }
      const R sp = sigmoid(zp), sd = sigmoid(zd);
      gv[o] -= g[o] * sd;
      const R gzp = g[o] * sp * (R(1) - sp), gzd = -g[o] * v[o] * sd * (R(1) - sd);
      gw.add(oP + 4 * H + o, gzp);
      gw.add(oD + 4 * H + o, gzd);
      for (int k = 0; k < H; ++k) {
        gw.add(oP + o * H + k, gzp * h[k]);
        gw.add(oD + o * H + k, gzd * h[k]);
        gh[k] += gzp * wp[o * H + k] + gzd * wd[o * H + k];
      }
    }
    for (int k = 0; k < H; ++k) {
      const R gpre = gh[k] * (R(1) - h[k] * h[k]);
      const R* w1 = w + k * NIN;
      gw.add(H * NIN + k, gpre);
      gw.add(k * NIN, gpre * t);
#pragma unroll
      for (int j = 1; j < NIN; ++j) {
        gw.add(k * NIN + j, gpre * species[j - 1]);
        gspecies[j - 1] += gpre * w1[j];
      }
    }
  }
};

}  // namespace vh
