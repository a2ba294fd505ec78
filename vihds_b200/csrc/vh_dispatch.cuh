// Host-side helpers shared by the CUDA launchers and the host math check: argument validation, building the typed
// Call<R> view from the C-ABI structs, and the (model x solver) -> template dispatch.
#pragma once
#include <string.h>

#include "vh_traj.cuh"

namespace vh {

// white-box families (hand-written RHS + VJP templates): double receiver / relay (DrModel) and growth-only (GrowthModel)
inline bool model_is_dr_family(int model) {
  return (model >= VH_MODEL_DR_CONSTANT && model <= VH_MODEL_RELAY_CONSTANT_PRECISIONS) ||
         (model >= VH_MODEL_AUTO_CONSTANT && model <= VH_MODEL_DEGRADER_CONSTANT_PRECISIONS);
}
inline bool model_is_dyn(int model) {
  return model == VH_MODEL_DR_CONSTANT_PRECISIONS || model == VH_MODEL_DR_CONSTANT_PRECISIONS_V2 ||
         model == VH_MODEL_RELAY_CONSTANT_PRECISIONS || model == VH_MODEL_DR_BLACKBOX ||
         model == VH_MODEL_AUTO_CONSTANT_PRECISIONS || model == VH_MODEL_PRPR_CONSTANT_PRECISIONS ||
         model == VH_MODEL_INDUCER_CONSTANT_PRECISIONS || model == VH_MODEL_DEGRADER_CONSTANT_PRECISIONS;
}
// treatments the right-hand side reads (columns of the batch's `inputs`)
inline int model_min_treatments(int model) {
  if (model >= VH_MODEL_DR_CONSTANT && model <= VH_MODEL_RELAY_CONSTANT_PRECISIONS) return 2;
  if (model == VH_MODEL_DEGRADER_CONSTANT || model == VH_MODEL_DEGRADER_CONSTANT_PRECISIONS) return 3;
  if (model == VH_MODEL_INDUCER_CONSTANT || model == VH_MODEL_INDUCER_CONSTANT_PRECISIONS) return 1;
  return 0;
}
inline int model_species(int model) {
  switch (model) {
    case VH_MODEL_RELAY_CONSTANT:
    case VH_MODEL_RELAY_CONSTANT_PRECISIONS:
      return 12;
    case VH_MODEL_DR_BLACKBOX:
    case VH_MODEL_AUTO_CONSTANT:
    case VH_MODEL_AUTO_CONSTANT_PRECISIONS:
      return 4;
    case VH_MODEL_PRPR_CONSTANT:
    case VH_MODEL_PRPR_CONSTANT_PRECISIONS:
      return 6;
    case VH_MODEL_INDUCER_CONSTANT:
    case VH_MODEL_INDUCER_CONSTANT_PRECISIONS:
      return 5;
    case VH_MODEL_DEGRADER_CONSTANT:
    case VH_MODEL_DEGRADER_CONSTANT_PRECISIONS:
      return 11;
    default:
      return 8;
  }
}

template <typename R>
inline const char* build_call(const vh_problem* p, const vh_fwd_io* io, const vh_bwd_io* bio, Call<R>& a) {
  memset(&a, 0, sizeof(a));
  if (!p || !io) return "null problem / io";
  if (p->B <= 0 || p->IW <= 0 || p->T < 2) return "B, IW must be positive and T >= 2";
  if (p->P < 0 || p->P > VH_MAX_SLOTS) return "P out of range (0..VH_MAX_SLOTS)";
  if ((long long)p->B * p->IW > 0x7fffffffLL) return "B*IW exceeds int32";
  a.B = p->B; a.IW = p->IW; a.N = p->B * p->IW; a.T = p->T; a.P = p->P; a.C = p->C; a.D = p->D; a.E = p->E;
  if (p->C < model_min_treatments(p->model))
    return "C (treatments): the double-receiver / relay models read 2 (C6, C12), degrader 3 (+ Ara), inducer 1 (Ara)";
  a.bb_nlat = p->n_z + p->n_x + p->n_y;
  a.bb_ny = p->n_y;
  a.bb_noff = (p->model == VH_MODEL_DR_BLACKBOX && p->P > 0) ? p->E : 0;
  a.bb_init_latent = (R)p->init_latent_species;
  a.bb_init_prec = (R)p->init_prec;
  bool used[VH_MAX_SLOTS];
  for (int k = 0; k < VH_MAX_SLOTS; ++k) used[k] = false;
  for (int s = 0; s < VH_MAX_SLOTS; ++s) {
    const int src = p->slot_src[s];
    a.slot_src[s] = src;
    if (src >= 0) {
      if (src >= p->P) return "slot_src column >= P";
      used[src] = true;
    } else if (src != VH_SLOT_UNUSED && (-1 - src) >= p->E) {
      return "slot_src extra row >= E";
    }
  }
  a.n_free = 0;
  for (int k = 0; k < VH_MAX_SLOTS; ++k) a.col_slot[k] = -1;
  for (int k = 0; k < p->P; ++k)
    if (!used[k]) a.free_cols[a.n_free++] = k;
  for (int s = 0; s < VH_MAX_SLOTS; ++s)
    if (p->slot_src[s] >= 0) {
      if (a.col_slot[p->slot_src[s]] >= 0) return "a theta column feeds two model slots";
      a.col_slot[p->slot_src[s]] = s;
    }
  a.times = (const R*)io->times; a.u = (const R*)io->u; a.q_mu = (const R*)io->q_mu; a.q_prec = (const R*)io->q_prec;
  a.p_mu = (const R*)io->p_mu; a.p_prec = (const R*)io->p_prec; a.clip_lo = (const R*)io->clip_lo;
  a.clip_hi = (const R*)io->clip_hi; a.kind = io->kind; a.extra = (const R*)io->extra;
  a.treatments = (const R*)io->treatments; a.dev_1hot = (const R*)io->dev_1hot; a.obs = (const R*)io->observations;
  a.weights = (const R*)io->weights;
  a.theta = (R*)io->theta; a.x_states = (R*)io->x_states; a.x_predict = (R*)io->x_predict;
  a.logp_species = (R*)io->logp_by_species; a.logp_theta = (R*)io->logp_theta; a.logq_theta = (R*)io->logq_theta;
  if (!a.times || !a.treatments) return "times / treatments must not be NULL";
  if (p->P > 0 && (!a.u || !a.q_mu || !a.q_prec || !a.p_mu || !a.p_prec || !a.clip_lo || !a.clip_hi || !a.kind))
    return "P > 0 needs u, q_mu, q_prec, p_mu, p_prec, clip_lo, clip_hi, kind";
  if (p->E > 0 && !a.extra) return "E > 0 needs extra";
  if (model_is_dyn(p->model) && !a.weights) return "dynamic-precision / black-box models need weights";
  if (bio) {
    a.g_logp_species = (const R*)bio->g_logp_by_species; a.g_logp_theta = (const R*)bio->g_logp_theta;
    a.g_logq_theta = (const R*)bio->g_logq_theta; a.g_theta = (const R*)bio->g_theta;
    a.g_x_states = (const R*)bio->g_x_states; a.g_x_predict = (const R*)bio->g_x_predict;
    a.d_q_mu = (R*)bio->d_q_mu; a.d_q_prec = (R*)bio->d_q_prec; a.d_extra = (R*)bio->d_extra; a.d_weights = (R*)bio->d_weights;
    if (!a.x_states) return "backward needs the forward x_states trace";
    a.theta_in = a.theta;  // the forward call's theta output, if handed back: reused instead of re-sampling
    a.theta = nullptr;     // never written by the reverse sweep
    if (p->P > 0 && (!a.d_q_mu || !a.d_q_prec)) return "backward with P > 0 needs d_q_mu and d_q_prec";
    if (model_is_dyn(p->model) && !a.d_weights) return "backward of a dynamic-precision model needs d_weights";
    if (bio->iwae_b_total > 0) {
      if (!bio->iwae_cost || !a.logp_species || !a.logp_theta || !a.logq_theta)
        return "fused IWAE: iwae_cost and the forward call's logp_by_species / logp_theta / logq_theta are required";
      if (a.g_logp_species || a.g_logp_theta || a.g_logq_theta) return "fused IWAE: g_logp_* must be NULL";
      if (p->model == VH_MODEL_DR_BLACKBOX) return "fused IWAE: not available for dr_blackbox";
      a.iw_b_total = bio->iwae_b_total;
      a.iw_cost = (R*)bio->iwae_cost;
    }
  }
  if (p->model == VH_MODEL_DR_BLACKBOX) {
    if (a.bb_nlat + p->C + p->D > 48) return "black-box: n_z + n_x + n_y + C + D exceeds 48";
    if (4 + a.bb_nlat > VH_MAX_SLOTS) return "black-box: too many latent parameters";
    if (p->P > 0 && p->E != 0 && p->E != p->n_y) return "black-box: E must be 0 or n_y (device offsets of the y parameters)";
    if (p->D > 0 && !a.dev_1hot) return "black-box needs dev_1hot";
  }
  return nullptr;
}

// F must provide:  template <class M, class TB> int run();
template <class M, class F>
inline int dispatch_solver(int solver, F& f) {
  typedef typename M::real R;
  switch (solver) {
    case VH_SOLVER_EULER: return f.template run<M, TabEuler<R> >();
    case VH_SOLVER_MIDPOINT: return f.template run<M, TabMidpoint<R> >();
    case VH_SOLVER_RK4: return f.template run<M, TabRK4_38<R> >();
    case VH_SOLVER_MODEULER: return f.template run<M, TabHeun<R, true> >();
    case VH_SOLVER_MODEULERWHILE: return f.template run<M, TabHeun<R, false> >();
    default: return VH_ERR_UNSUPPORTED;
  }
}

template <typename R, class F>
inline int dispatch_dr(int model, int solver, F& f) {
  switch (model) {
    case VH_MODEL_DR_CONSTANT: return dispatch_solver<DrModel<R, 1, 0, false> >(solver, f);
    case VH_MODEL_DR_CONSTANT_V2: return dispatch_solver<DrModel<R, 2, 0, false> >(solver, f);
    case VH_MODEL_DR_CONSTANT_PRECISIONS: return dispatch_solver<DrModel<R, 1, 0, true> >(solver, f);
    case VH_MODEL_DR_CONSTANT_PRECISIONS_V2: return dispatch_solver<DrModel<R, 2, 0, true> >(solver, f);
    case VH_MODEL_RELAY_CONSTANT: return dispatch_solver<DrModel<R, 1, 1, false> >(solver, f);
    case VH_MODEL_RELAY_CONSTANT_PRECISIONS: return dispatch_solver<DrModel<R, 1, 1, true> >(solver, f);
    case VH_MODEL_DEGRADER_CONSTANT: return dispatch_solver<DrModel<R, 1, 2, false> >(solver, f);
    case VH_MODEL_DEGRADER_CONSTANT_PRECISIONS: return dispatch_solver<DrModel<R, 1, 2, true> >(solver, f);
    case VH_MODEL_INDUCER_CONSTANT: return dispatch_solver<GrowthModel<R, 5, false> >(solver, f);
    case VH_MODEL_INDUCER_CONSTANT_PRECISIONS: return dispatch_solver<GrowthModel<R, 5, true> >(solver, f);
    case VH_MODEL_AUTO_CONSTANT: return dispatch_solver<GrowthModel<R, 4, false> >(solver, f);
    case VH_MODEL_AUTO_CONSTANT_PRECISIONS: return dispatch_solver<GrowthModel<R, 4, true> >(solver, f);
    case VH_MODEL_PRPR_CONSTANT: return dispatch_solver<GrowthModel<R, 6, false> >(solver, f);
    case VH_MODEL_PRPR_CONSTANT_PRECISIONS: return dispatch_solver<GrowthModel<R, 6, true> >(solver, f);
    default: return VH_ERR_UNSUPPORTED;
  }
}

}  // namespace vh
