// The arithmetic of the matrix-form reverse sweep (vh_bwd_mx.cuh), host-compilable: the sparse Jacobian of the
// double-receiver right-hand side, the step matrix N = I + h A + a10 h^2 A B of the midpoint rule, and the recurrence
// lambda0 = N^T lambda1 + e.  tests/hostcheck builds these with g++ and compares one step against rk_step_adjoint.
#pragma once
#include "vh_models.cuh"

namespace vh {

// ---- Jacobian of the species right-hand side (DrModel, no extension) at state X with intermediates m -------------------
template <typename R>
struct DrJac {
  R j00;          // d f0 / d x0
  R c0[8];        // d f_i / d x0, i = 1..7
  R dg[8];        // d f_i / d x_i, i = 1..7
  R j26, j27, j36, j37;
};
template <class M>
VH_HD void dr_jacobian(const typename M::real* X, const typename M::Consts& c, const typename M::Mid& m,
                                            DrJac<typename M::real>& J) {
  typedef typename M::real R;
  const R* v = c.v;
  const R dgam = -(m.gr * c.iK);  // d gamma / d x0
  J.j00 = m.gam + X[0] * dgam;
#pragma unroll
  for (int i = 1; i < 8; ++i) J.c0[i] = -(X[i] * dgam);
  J.dg[1] = -(m.gam + v[C_drfp]);
  J.dg[2] = -(m.gam + v[C_dyfp]);
  J.dg[3] = -(m.gam + v[C_dcfp]);
  J.dg[4] = -m.gam;
  J.dg[5] = -m.gam;
  J.dg[6] = -(m.gam + v[C_dR]);
  J.dg[7] = -(m.gam + v[C_dS]);
  // promoter activities P = (e + a + b) / (1 + a + b), a = KGR x6^2 fR, b = KGS x7^2 fS:  dP/da = (1 - P) / (1 + a + b)
  const R q81 = v[C_cY] * ((R(1) - m.P81) * m.i81), q76 = v[C_cC] * ((R(1) - m.P76) * m.i76);
  const R s6 = R(2) * X[6] * v[C_fR], s7 = R(2) * X[7] * v[C_fS];
  J.j26 = q81 * v[C_KGR81] * s6;
  J.j27 = q81 * v[C_KGS81] * s7;
  J.j36 = q76 * v[C_KGR76] * s6;
  J.j37 = q76 * v[C_KGS76] * s7;
}

// the 20 + 4 numbers a step hands to the recurrence; order = ring item order
enum {
  MXN_00 = 0,   // N_00
  MXN_C0 = 0,   // N_i0 at MXN_C0 + i, i = 1..7
  MXN_DG = 7,   // N_ii at MXN_DG + i, i = 1..7
  MXN_26 = 15, MXN_27 = 16, MXN_36 = 17, MXN_37 = 18,
  MXN_E = 19,   // emission cotangent: e0, e1, e2 (= e4), e3 (= e5)
  MXN_ITEMS = 23,
  MXN_PAD = 24  // the accumulators' share of a slot starts on a vector boundary
};


// N = I + h A + (a10 h^2) A B on the common sparsity pattern (A at the mid-point state Xm, B at x0); Nv[MXN_E ..] untouched
template <class M>
VH_HD void mx_step_matrix(const typename M::real* x, const typename M::real* Xm, const typename M::Consts& c,
                          const typename M::Mid& m0, const typename M::Mid& m1, typename M::real h, typename M::real a10,
                          typename M::real* Nv) {
  typedef typename M::real R;
  DrJac<R> A, B;
  dr_jacobian<M>(Xm, c, m1, A);
  dr_jacobian<M>(x, c, m0, B);
  const R hh = h * h * a10;
  Nv[MXN_00] = R(1) + h * A.j00 + hh * (A.j00 * B.j00);
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    R ab = A.c0[i] * B.j00 + A.dg[i] * B.c0[i];
    if (i == 2) ab += A.j26 * B.c0[6] + A.j27 * B.c0[7];
    if (i == 3) ab += A.j36 * B.c0[6] + A.j37 * B.c0[7];
    Nv[MXN_C0 + i] = h * A.c0[i] + hh * ab;
    Nv[MXN_DG + i] = R(1) + h * A.dg[i] + hh * (A.dg[i] * B.dg[i]);
  }
  Nv[MXN_26] = h * A.j26 + hh * (A.dg[2] * B.j26 + A.j26 * B.dg[6]);
  Nv[MXN_27] = h * A.j27 + hh * (A.dg[2] * B.j27 + A.j27 * B.dg[7]);
  Nv[MXN_36] = h * A.j36 + hh * (A.dg[3] * B.j36 + A.j36 * B.dg[6]);
  Nv[MXN_37] = h * A.j37 + hh * (A.dg[3] * B.j37 + A.j37 * B.dg[7]);
}

// lambda <- N^T lambda + e (e0, e1, e2 = e4, e3 = e5 at Nv[MXN_E ..]).  Column 0 gathers all eight components
// (pairwise: dependent depth 4), the others one or three.
template <typename R>
VH_HD void mx_apply(const R* Nv, R* lam) {
  const R s01 = Nv[MXN_00] * lam[0] + Nv[MXN_C0 + 1] * lam[1];
  const R s23 = Nv[MXN_C0 + 2] * lam[2] + Nv[MXN_C0 + 3] * lam[3];
  const R s45 = Nv[MXN_C0 + 4] * lam[4] + Nv[MXN_C0 + 5] * lam[5];
  const R s67 = Nv[MXN_C0 + 6] * lam[6] + Nv[MXN_C0 + 7] * lam[7];
  const R l6 = Nv[MXN_DG + 6] * lam[6] + (Nv[MXN_26] * lam[2] + Nv[MXN_36] * lam[3]);
  const R l7 = Nv[MXN_DG + 7] * lam[7] + (Nv[MXN_27] * lam[2] + Nv[MXN_37] * lam[3]);
  lam[0] = ((s01 + s23) + (s45 + s67)) + Nv[MXN_E + 0];
  lam[1] = Nv[MXN_DG + 1] * lam[1] + Nv[MXN_E + 1];
  lam[2] = Nv[MXN_DG + 2] * lam[2] + Nv[MXN_E + 2];
  lam[3] = Nv[MXN_DG + 3] * lam[3] + Nv[MXN_E + 3];
  lam[4] = Nv[MXN_DG + 4] * lam[4] + Nv[MXN_E + 2];
  lam[5] = Nv[MXN_DG + 5] * lam[5] + Nv[MXN_E + 3];
  lam[6] = l6;
  lam[7] = l7;
}

}  // namespace vh
