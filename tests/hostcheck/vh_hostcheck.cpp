// TEST INFRASTRUCTURE ONLY -- never linked into libvihds_b200.so and never used by the product path.
// Compiles the VH_HD trajectory maths of vihds_b200/csrc (the very same headers the CUDA kernels instantiate) for the
// host, so that the hand-written RHS / stepper / adjoint code can be checked against the CPU oracle in the build
// container, which has no GPU.  Pointers in the io structs are HOST pointers here.
#include <vector>

#include "../../vihds_b200/csrc/vh_dispatch.cuh"
#include "../../vihds_b200/csrc/vh_bb.cuh"
#include "../../vihds_b200/csrc/vh_mx_math.cuh"

namespace {
using namespace vh;

template <typename R>
struct HostRed {
  const Call<R>* a;
  void operator()(int b, int k, R dmu, R dprec, bool active) const {
    if (!active) return;
    a->d_q_mu[b * a->P + k] += dmu;
    a->d_q_prec[b * a->P + k] += dprec;
  }
};

template <typename R>
struct FwdRunner {
  const Call<R>* a;
  template <class M, class TB>
  int run() {
    std::vector<R> scratch(M::NSLOT);
    SlotScratch<R> sc{scratch.data(), 1};
    for (int n = 0; n < a->N; ++n) traj_forward<M, TB>(*a, n, a->weights, sc);
    return 0;
  }
};

template <typename R>
struct BwdRunner {
  const Call<R>* a;
  size_t nw;
  template <class M, class TB>
  int run() {
    std::vector<R> gw(nw ? nw : 1);
    for (size_t i = 0; i < (size_t)a->B * a->P; ++i) a->d_q_mu[i] = a->d_q_prec[i] = R(0);
    for (size_t i = 0; i < nw; ++i) a->d_weights[i] = R(0);
    HostRed<R> red{a};
    for (int n = 0; n < a->N; ++n) {
      for (size_t i = 0; i < nw; ++i) gw[i] = R(0);
      StridedGW<R> h{gw.data(), 1};
      std::vector<R> scratch(M::NSLOT);
      SlotScratch<R> sc{scratch.data(), 1};
      DirectCk<R, M::S> ck;
      traj_backward<M, TB>(*a, n, true, a->weights, h, red, sc, ck);
      for (size_t i = 0; i < nw; ++i) a->d_weights[i] += gw[i];
    }
    return 0;
  }
};

// dr_blackbox on the host: same BbRhs / bb_traj_* code, weight gradients accumulated directly
template <typename R>
struct BbFwdRunner {
  const Call<R>* a;
  template <class F, class TB>
  int run() {
    std::vector<R> row(F::ROWL::ROW);
    for (int n = 0; n < a->N; ++n) bb_traj_forward<F, TB>(*a, n, a->weights, row.data());
    return 0;
  }
};
template <typename R>
struct BbBwdRunner {
  const Call<R>* a;
  template <class F, class TB>
  int run() {
    if (a->bb_nlat + a->C + a->D != F::NC) return VH_ERR_UNSUPPORTED;
    for (size_t i = 0; i < (size_t)a->B * a->P; ++i) a->d_q_mu[i] = a->d_q_prec[i] = R(0);
    for (int i = 0; i < F::L::total; ++i) a->d_weights[i] = R(0);
    HostRed<R> red{a};
    BbDirectWgrad<F> sink{a->d_weights};
    std::vector<R> row(F::ROWL::ROW);
    for (int n = 0; n < a->N; ++n) bb_traj_backward<F, TB>(*a, n, true, a->weights, row.data(), sink, red);
    return 0;
  }
};
template <class F, class L>
int bb_solver(int solver, L& f) {
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

template <typename R>
int fwd_t(const vh_problem* p, const vh_fwd_io* io) {
  Call<R> a;
  if (build_call<R>(p, io, nullptr, a)) return VH_ERR_INVALID;
  if (p->model == VH_MODEL_DR_BLACKBOX) {
    BbFwdRunner<R> f{&a};
    return bb_solver<BbRhs<R, 2, 25, 20, 21> >(p->solver, f);
  }
  a.n_hidden = (model_is_dyn(p->model) && p->model != VH_MODEL_DR_BLACKBOX) ? p->n_hidden : 0;
  FwdRunner<R> f{&a};
  return dispatch_dr<R>(p->model, p->solver, f);
}
template <typename R>
int bwd_t(const vh_problem* p, const vh_bwd_io* io) {
  Call<R> a;
  if (build_call<R>(p, &io->fwd, io, a)) return VH_ERR_INVALID;
  if (p->model == VH_MODEL_DR_BLACKBOX) {
    BbBwdRunner<R> f{&a};
    return bb_solver<BbRhs<R, 2, 25, 20, 21> >(p->solver, f);
  }
  const int nin = model_species(p->model) + 1, H = model_is_dyn(p->model) ? p->n_hidden : 0;
  a.n_hidden = H;
  // LinPrecNet::NW, or HidPrecNet::num_weights(H) with a hidden layer
  size_t nw = !model_is_dyn(p->model) ? 0 : (H == 0 ? (size_t)2 * (4 * nin + 4) : (size_t)H * (nin + 1) + 2 * (4 * H + 4));
  a.nw = (int)nw;
  BwdRunner<R> f{&a, nw};
  return dispatch_dr<R>(p->model, p->solver, f);
}
}  // namespace

extern "C" int hc_fwd(const vh_problem* p, const vh_fwd_io* io) {
  return p->dtype == VH_F64 ? fwd_t<double>(p, io) : fwd_t<float>(p, io);
}
extern "C" int hc_bwd(const vh_problem* p, const vh_bwd_io* io) {
  return p->dtype == VH_F64 ? bwd_t<double>(p, io) : bwd_t<float>(p, io);
}
extern "C" int hc_num_slots() { return vh::DR_NSLOT; }
extern "C" const char* hc_build_error(const vh_problem* p, const vh_fwd_io* io) {
  Call<float> a;
  const char* e = build_call<float>(p, io, nullptr, a);
  return e ? e : "";
}
extern "C" const char* hc_slot_name(int model, int s) {
  if (vh::model_is_dyn(model) && s >= vh::S_prec_x && s <= vh::S_prec_cfp) return vh::kDrDynPrecNames[s - vh::S_prec_x];
  return vh::kDrSlotNames[s];
}

// One midpoint step of the adjoint two ways (fp64): rk_step_adjoint on the kept stage data -- the code every reverse kernel
// runs -- and the matrix form of vh_bwd_mx.cuh, lambda0 = (I + h A + h^2/2 A B)^T lambda1.  th: DR_NSLOT slot values,
// tc: {C6, C12, Ara}, x: state at t0, lam1: cotangent at t1.  Outputs: lam_ref[8], lam_mx[8].
template <int VER>
static void mx_step_t(const double* th, const double* tc, const double* x, const double* lam1, double t0, double t1,
                      double* lam_ref, double* lam_mx) {
  typedef DrModel<double, VER, 0, false> M;
  typedef TabMidpoint<double> TB;
  Rhs<M> f;
  f.w = nullptr;
  f.nh = 0;
  M::setup(th, tc, f.c);
  const double h = t1 - t0;
  StageData<Rhs<M>, TB> sd;
  rk_stages_forward<Rhs<M>, TB>(f, t0, t1, h, x, sd);
  typename M::Consts gc;
  for (int i = 0; i < M::NC; ++i) gc.v[i] = 0.0;
  NoGW<double> nogw;
  for (int q = 0; q < 8; ++q) lam_ref[q] = lam1[q];
  rk_step_adjoint<Rhs<M>, TB>(f, t0, t1, h, x, sd, lam_ref, gc, nogw);
  double Xm[8], Nv[MXN_ITEMS];
  for (int q = 0; q < 8; ++q) Xm[q] = x[q] + (h * TB::a(1, 0)) * sd.k[0][q];
  mx_step_matrix<M>(x, Xm, f.c, sd.kept[0].m, sd.kept[1].m, h, TB::a(1, 0), Nv);
  for (int i = 0; i < 4; ++i) Nv[MXN_E + i] = 0.0;
  for (int q = 0; q < 8; ++q) lam_mx[q] = lam1[q];
  mx_apply(Nv, lam_mx);
}
extern "C" void hc_mx_step(int version, const double* th, const double* tc, const double* x, const double* lam1, double t0,
                           double t1, double* lam_ref, double* lam_mx) {
  if (version == 2)
    mx_step_t<2>(th, tc, x, lam1, t0, t1, lam_ref, lam_mx);
  else
    mx_step_t<1>(th, tc, x, lam1, t0, t1, lam_ref, lam_mx);
}
