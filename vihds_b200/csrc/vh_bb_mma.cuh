// dr_blackbox on the tensor cores: the MLP right-hand side (models/dr_blackbox.py:15-58, vihds/ode.py:119-138,
// vihds/precisions.py:63-87) as a warp-level GEMM, forward and reverse, fp32 via mma.sync.m16n8k8 3xTF32.
//
// Mapping.  A warp owns 16 trajectories = the M dimension of one mma tile; lane = 4 g + tg.  Trajectory row g (and
// g + 8) is shared by the four lanes of a quad, each of which owns two of the eight "input slots"
//      slot j:  0..5 = neural states x_j      6 = t      7 = unused
//      lane tg holds slots tg and tg + 4   (A-fragment columns)   and the precision state v_tg,
// so a trajectory's ODE state is spread over a quad and NOTHING is replicated.  The fragment layouts are chosen so
// that every product chains into the next without a shuffle:
//   * a C tile (row g / g+8, columns 2tg, 2tg+1) is the A fragment of the next GEMM's k-step when that GEMM's weight
//     rows are permuted by  k-slot kappa <-> column col_of(kappa)  (A regs = c0, c2, c1, c3);
//   * output column c of the last layer carries the derivative of slot slot_of(c), so dx lands on the lane that owns x.
// Both networks take the same inputs, so their hidden layers are ONE GEMM [16 x 8] x [8 x (32 + 24)] whose accumulator
// is initialised with the per-trajectory folded constants hc / hpc (vh_bb.cuh); the output biases ride on a constant-1
// hidden unit (states: unit 25, precisions: unit 20), which also yields the bias gradients from the weight-gradient GEMM.
// Weights are split once per CTA into TF32 hi / lo parts and stored in shared memory in fragment order (one
// conflict-free LDS.128 per B tile); activations are split on the fly (hi = value & 0xffffe000, lo = value - hi).
// Products: lo*hi + hi*lo + hi*hi, fp32 accumulate (the lo*lo term is 2^-22 relative).
//
// Reverse sweep.  A CTA is a PAIR of warps over the same 16 trajectories: warp 0 runs the discrete adjoint (forward
// re-evaluation + VJP GEMMs through the transposed weights); after each VJP it leaves the operand panels of that
// evaluation's weight-gradient outer products (inputs, hidden activations, pre-activation cotangents, output
// cotangents) in a two-slot shared-memory ring, and warp 1 -- which keeps all 40 weight-gradient accumulator registers
// -- runs the K = 16-trajectory GEMMs  dW += cotangent^T x activation  off the critical path.
#pragma once
#include "vh_bb.cuh"

namespace vh {
namespace bbm {

constexpr int ROWS = 16;  // trajectories per warp
constexpr int NT_FWD = 18, NT_ALL = 36;
// operand panel of one evaluation (per trajectory row), columns:
constexpr int PX = 0, PH = 8, PG = 40, PZP = 72, PZD = 80, PHP = 88, PGQ = 112, PQ = 136, PCOLS = 144;
constexpr int RS = 168;  // row stride: == 8 (mod 32) -> the float2 stores and the fragment loads are conflict-free

VH_HD constexpr int col_of(int kappa) { return kappa < 4 ? 2 * kappa : 2 * (kappa - 4) + 1; }
VH_HD constexpr int slot_of(int c) { return (c & 1) ? (c >> 1) + 4 : (c >> 1); }

// B[k][n] of weight tile `tile` (see the tile table in WarpMlp), from the flat weight vector
template <class F>
VH_HD float btile(const float* w, int tile, int k, int n) {
  typedef typename F::L L;
  constexpr int NST = F::NST, H = F::H, HP = F::HP;
  if (tile < 4) {  // layer 1, states net: k = input slot, column n = hidden unit 8 tile + n
    const int h = 8 * tile + n;
    return (k < NST && h < H) ? w[L::W1 + h * L::nin + k] : 0.f;
  }
  if (tile < 7) {  // layer 1, precision net (input slot NST = t)
    const int h = 8 * (tile - 4) + n;
    if (h >= HP || k > NST) return 0.f;
    return k < NST ? w[L::Q1 + h * (L::nin + 1) + 1 + k] : w[L::Q1 + h * (L::nin + 1)];
  }
  if (tile < 15) {  // layer 2, states net: tile = 7 + 2 j + pd;  k <-> hidden 8 j + col_of(k);  n <-> state slot_of(n)
    const int j = (tile - 7) >> 1, pd = (tile - 7) & 1;
    const int h = 8 * j + col_of(k), o = slot_of(n);
    if (o >= NST) return 0.f;
    if (h < H) return w[(pd ? L::Wd : L::Wp) + o * H + h];
    return h == H ? w[(pd ? L::bd : L::bp) + o] : 0.f;  // the constant-1 hidden unit carries the bias
  }
  if (tile < 18) {  // layer 2, precision net: column n = 2 o + pd
    const int h = 8 * (tile - 15) + col_of(k), o = n >> 1, pd = n & 1;
    if (h < HP) return w[(pd ? L::Qd : L::Qp) + o * HP + h];
    return h == HP ? w[(pd ? L::qbd : L::qbp) + o] : 0.f;
  }
  if (tile < 26) {  // VJP through layer 2, states: tile = 18 + 4 jd + nt;  k = state (gzp: jd 0, gzd: jd 1)
    const int jd = (tile - 18) >> 2, h = 8 * ((tile - 18) & 3) + n;
    return (k < NST && h < H) ? w[(jd ? L::Wd : L::Wp) + k * H + h] : 0.f;
  }
  if (tile < 29) {  // VJP through layer 2, precisions: k < 4: production o = k, else degradation o = k - 4
    const int h = 8 * (tile - 26) + n;
    if (h >= HP) return 0.f;
    return k < 4 ? w[L::Qp + k * HP + h] : w[L::Qd + (k - 4) * HP + h];
  }
  if (tile < 33) {  // VJP through layer 1, states: k <-> hidden 8 j + col_of(k);  n <-> input slot slot_of(n)
    const int h = 8 * (tile - 29) + col_of(k), s = slot_of(n);
    return (h < H && s < NST) ? w[L::W1 + h * L::nin + s] : 0.f;
  }
  const int h = 8 * (tile - 33) + col_of(k), s = slot_of(n);
  return (h < HP && s < NST) ? w[L::Q1 + h * (L::nin + 1) + 1 + s] : 0.f;
}

// folding of the constants c (once per trajectory): hidden pre-activation += W[:, c-columns] c.  B[k][n] of (q, nt):
// k <-> constant 8 q + k, n <-> hidden 8 nt + n
template <class F>
VH_HD float fold_b(const float* w, bool prec, int q, int nt, int k, int n) {
  typedef typename F::L L;
  const int j = 8 * q + k, h = 8 * nt + n;
  if (j >= F::NC) return 0.f;
  if (!prec) return h < F::H ? w[L::W1 + h * L::nin + F::NST + j] : 0.f;
  return h < F::HP ? w[L::Q1 + h * (L::nin + 1) + 1 + F::NST + j] : 0.f;
}
// its transpose (cotangent of c from the accumulated cotangents of hc / hpc): k <-> hidden 8 jt + col_of(k),
// n <-> constant 8 q + slot_of(n)
template <class F>
VH_HD float gc_b(const float* w, bool prec, int jt, int q, int k, int n) {
  typedef typename F::L L;
  const int h = 8 * jt + col_of(k), j = 8 * q + slot_of(n);
  if (j >= F::NC) return 0.f;
  if (!prec) return h < F::H ? w[L::W1 + h * L::nin + F::NST + j] : 0.f;
  return h < F::HP ? w[L::Q1 + h * (L::nin + 1) + 1 + F::NST + j] : 0.f;
}

#if defined(__CUDACC__)
__device__ __forceinline__ void split(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  lo = v - hi;
}
// d = c + A B, one m16n8k8 TF32 tile (operands are fp32 bit patterns; the tensor core reads their upper 19 bits).
// d and c are separate register quadruples so that a persistent accumulator initialiser (the folded constants) or a
// literal zero can be the C operand without a copy.
__device__ __forceinline__ void mma8(float* d, const float* a, float b0, float b1, const float* c) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
        "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}
__device__ __forceinline__ void mma8(float* c, const float* a, float b0, float b1) { mma8(c, a, b0, b1, c); }
// 3xTF32 with a pre-split B tile b = (b0 hi, b1 hi, b0 lo, b1 lo) -- the register pairs the instruction wants.
// One accumulator chain, small terms first:  d = ((c + lo*hi) + hi*lo) + hi*hi
__device__ __forceinline__ void mma3c(float* d, const float* ah, const float* al, const float4 b, const float* c) {
  mma8(d, al, b.x, b.y, c);
  mma8(d, ah, b.z, b.w);
  mma8(d, ah, b.x, b.y);
}
// three independent accumulators (hi*hi, lo*hi, hi*lo): dependent chains a third as long, summed by the caller
__device__ __forceinline__ void mma3x(float (*acc)[4], const float* ah, const float* al, const float4 b) {
  mma8(acc[0], ah, b.x, b.y);
  mma8(acc[1], al, b.x, b.y);
  mma8(acc[2], ah, b.z, b.w);
}
// sigmoid on the SFU: ex2.approx + rcp.approx, 4 instructions instead of the ~14 of 1 / (1 + expf(-z)).  The argument
// product carries |z| * 2^-24 of error, i.e. an absolute error in sigma of at most sigma (1 - sigma) |z| 1e-7 < 2e-7:
// fp32 round-off class, three orders of magnitude inside the 1e-4 parity bar.  NaN propagates; +-inf saturate.
__device__ __forceinline__ float sigmoid_fast(float z) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}
// a C tile as the A fragment of the next GEMM: (c0, c2, c1, c3), split
__device__ __forceinline__ void frag_of(const float* c, float* ah, float* al) {
  split(c[0], ah[0], al[0]);
  split(c[2], ah[1], al[1]);
  split(c[1], ah[2], al[2]);
  split(c[3], ah[3], al[3]);
}
// one-off GEMMs whose B tile is computed on the fly (fold / gc): c += A B with 3xTF32
__device__ __forceinline__ void mma3_fly(float* c, const float* ah, const float* al, float b0, float b1) {
  float b0h, b0l, b1h, b1l;
  split(b0, b0h, b0l);
  split(b1, b1h, b1l);
  mma8(c, al, b0h, b1h);
  mma8(c, ah, b0l, b1l);
  mma8(c, ah, b0h, b1h);
}

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
enum { BAR_FULL0 = 1, BAR_EMPTY0 = 3, BAR_TAIL1 = 5, BAR_TAIL2 = 6 };

// fragment-ordered, pre-split weight tiles: tiles[tile * 32 + lane] = (b0 hi, b1 hi, b0 lo, b1 lo)
template <class F>
__device__ void build_tiles(const float* w_smem, float4* tiles, int ntile) {
  for (int i = threadIdx.x; i < ntile * 32; i += blockDim.x) {
    const int tile = i >> 5, ln = i & 31, g = ln >> 2, tg = ln & 3;
    float4 v;
    split(btile<F>(w_smem, tile, tg, g), v.x, v.z);
    split(btile<F>(w_smem, tile, tg + 4, g), v.y, v.w);
    tiles[i] = v;
  }
}

// hand-off of one evaluation's weight-gradient operands from the adjoint warp to the weight-gradient warp
struct PanelSink {
  float* ring;  // [2][ROWS][RS]
  int it, g, tg, bar0;  // bar0: first named-barrier id of this warp pair
  __device__ void put(const float* in, const float (*hid)[4], const float (*gpre)[4], const float* gzP, const float* gzD,
                      const float* gq) {
    const int slot = it & 1;
    if (it >= 2) bar_sync(bar0 + BAR_EMPTY0 + slot, 64);
    float* p0 = ring + (slot * ROWS + g) * RS;
    float* p1 = p0 + 8 * RS;
    p0[PX + tg] = in[0];
    p1[PX + tg] = in[1];
    p0[PX + tg + 4] = in[2];
    p1[PX + tg + 4] = in[3];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      *reinterpret_cast<float2*>(p0 + PH + 8 * nt + 2 * tg) = make_float2(hid[nt][0], hid[nt][1]);
      *reinterpret_cast<float2*>(p1 + PH + 8 * nt + 2 * tg) = make_float2(hid[nt][2], hid[nt][3]);
      *reinterpret_cast<float2*>(p0 + PG + 8 * nt + 2 * tg) = make_float2(gpre[nt][0], gpre[nt][1]);
      *reinterpret_cast<float2*>(p1 + PG + 8 * nt + 2 * tg) = make_float2(gpre[nt][2], gpre[nt][3]);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      *reinterpret_cast<float2*>(p0 + PHP + 8 * j + 2 * tg) = make_float2(hid[4 + j][0], hid[4 + j][1]);
      *reinterpret_cast<float2*>(p1 + PHP + 8 * j + 2 * tg) = make_float2(hid[4 + j][2], hid[4 + j][3]);
      *reinterpret_cast<float2*>(p0 + PGQ + 8 * j + 2 * tg) = make_float2(gpre[4 + j][0], gpre[4 + j][1]);
      *reinterpret_cast<float2*>(p1 + PGQ + 8 * j + 2 * tg) = make_float2(gpre[4 + j][2], gpre[4 + j][3]);
    }
    *reinterpret_cast<float2*>(p0 + PZP + 2 * tg) = make_float2(gzP[0], gzP[1]);
    *reinterpret_cast<float2*>(p1 + PZP + 2 * tg) = make_float2(gzP[2], gzP[3]);
    *reinterpret_cast<float2*>(p0 + PZD + 2 * tg) = make_float2(gzD[0], gzD[1]);
    *reinterpret_cast<float2*>(p1 + PZD + 2 * tg) = make_float2(gzD[2], gzD[3]);
    *reinterpret_cast<float2*>(p0 + PQ + 2 * tg) = make_float2(gq[0], gq[1]);
    *reinterpret_cast<float2*>(p1 + PQ + 2 * tg) = make_float2(gq[2], gq[3]);
    __threadfence_block();
    bar_arrive(bar0 + BAR_FULL0 + slot, 64);
    ++it;
  }
};
struct NoSink {
  __device__ void put(const float*, const float (*)[4], const float (*)[4], const float*, const float*, const float*) const {}
};

// ---------------------------------------------------------------------------------------------------------------
// The right-hand side of 16 trajectories, evaluated by one warp.  Per-lane state vector Y (what rk_step / rk_step_vjp
// of vh_traj.cuh integrate): [xa, xb, v] of row g, then of row g + 8; xa = x_tg, xb = x_{tg+4} (lanes tg < NST - 4;
// a dead component otherwise: derivative 0), v = precision state tg.
//
// Weight tiles (fragment order, shared memory):
//    0.. 3  layer 1 states   (n-tile = hidden 8 nt ..)      4.. 6  layer 1 precisions
//    7..14  layer 2 states   (7 + 2 j + {P, D})             15..17  layer 2 precisions (columns 2 o + {p, d})
//   18..25  VJP layer 2 states (18 + 4 {gzp, gzd} + nt)     26..28  VJP layer 2 precisions
//   29..32  VJP layer 1 states (k-step = hidden tile j)     33..35  VJP layer 1 precisions
// ---------------------------------------------------------------------------------------------------------------
template <class F>
struct WarpMlp {
  typedef float real;
  static constexpr int NST = F::NST, H = F::H, HP = F::HP, NC = F::NC;
  static constexpr int S = 6;
  static_assert(NST == 6 && H + 1 <= 32 && HP + 1 <= 24 && NC + 1 <= 24, "tile shape of the compiled network");
  struct Kept {
    float hid[7][4];  // relu'd hidden activations (C layout), states tiles 0..3, precision tiles 4..6
    float sP[4], sD[4], sQ[4];
  };
  struct Grad {
    float ghc[7][4];  // accumulated cotangent of the folded constants hc | hpc (C layout)
  };
  const float4* bt;  // weight tiles + lane
  float hc[7][4];    // folded constants = accumulator init of layer 1
  int tg;
  bool hasB;

  __device__ float xb_in(float t, float y) const { return hasB ? y : (tg == NST - 4 ? t : 0.f); }
  __device__ void inputs(float t, const float* Y, float* in) const {
    in[0] = Y[0];
    in[1] = Y[3];
    in[2] = xb_in(t, Y[1]);
    in[3] = xb_in(t, Y[4]);
  }

  __device__ void forward(float t, const float* Y, Kept& k) const {
    float in[4], ah[4], al[4];
    inputs(t, Y, in);
#pragma unroll
    for (int i = 0; i < 4; ++i) split(in[i], ah[i], al[i]);
#pragma unroll
    for (int nt = 0; nt < 7; ++nt) {
      float c[4];
      mma3c(c, ah, al, bt[nt * 32], hc[nt]);
#pragma unroll
      for (int i = 0; i < 4; ++i) k.hid[nt][i] = fmaxf(c[i], 0.f);
    }
    const float zero[4] = {0.f, 0.f, 0.f, 0.f};
    float aP[3][4], aD[3][4], aQ[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float hh[4], hl[4];
      frag_of(k.hid[j], hh, hl);
      const float4 bp = bt[(7 + 2 * j) * 32], bd = bt[(8 + 2 * j) * 32];
      if (j == 0) {
        mma8(aP[0], hh, bp.x, bp.y, zero);
        mma8(aP[1], hl, bp.x, bp.y, zero);
        mma8(aP[2], hh, bp.z, bp.w, zero);
        mma8(aD[0], hh, bd.x, bd.y, zero);
        mma8(aD[1], hl, bd.x, bd.y, zero);
        mma8(aD[2], hh, bd.z, bd.w, zero);
      } else {
        mma3x(aP, hh, hl, bp);
        mma3x(aD, hh, hl, bd);
      }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float hh[4], hl[4];
      frag_of(k.hid[4 + j], hh, hl);
      const float4 bq = bt[(15 + j) * 32];
      if (j == 0) {
        mma8(aQ[0], hh, bq.x, bq.y, zero);
        mma8(aQ[1], hl, bq.x, bq.y, zero);
        mma8(aQ[2], hh, bq.z, bq.w, zero);
      } else {
        mma3x(aQ, hh, hl, bq);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      k.sP[i] = sigmoid_fast((aP[1][i] + aP[2][i]) + aP[0][i]);
      k.sD[i] = sigmoid_fast((aD[1][i] + aD[2][i]) + aD[0][i]);
      k.sQ[i] = sigmoid_fast((aQ[1][i] + aQ[2][i]) + aQ[0][i]);
    }
  }
  __device__ void derivative(const float* Y, const Kept& k, float* dY) const {
    dY[0] = k.sP[0] - k.sD[0] * Y[0];
    dY[1] = hasB ? k.sP[1] - k.sD[1] * Y[1] : 0.f;
    dY[2] = k.sQ[0] - k.sQ[1] * Y[2];
    dY[3] = k.sP[2] - k.sD[2] * Y[3];
    dY[4] = hasB ? k.sP[3] - k.sD[3] * Y[4] : 0.f;
    dY[5] = k.sQ[2] - k.sQ[3] * Y[5];
  }
  __device__ void eval(float t, const float* Y, float* dY) const {
    Kept k;
    forward(t, Y, k);
    derivative(Y, k, dY);
  }
  __device__ void eval_keep(float t, const float* Y, float* dY, Kept& k) const {
    forward(t, Y, k);
    derivative(Y, k, dY);
  }
  __device__ void keep_only(float t, const float* Y, Kept& k) const { forward(t, Y, k); }

  template <typename GW>
  __device__ void vjp_kept(float t, const float* Y, const Kept& k, const float* g, float* gX, Grad& gc, GW& gw) const {
    // output layer (lane-local): row g = entries 0, 1; row g + 8 = entries 2, 3
    float gzP[4], gzD[4], gq[4];
    const float gB0 = hasB ? g[1] : 0.f, gB1 = hasB ? g[4] : 0.f;
    const float gg[4] = {g[0], gB0, g[3], gB1};
    const float yy[4] = {Y[0], Y[1], Y[3], Y[4]};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      gzP[i] = gg[i] * k.sP[i] * (1.f - k.sP[i]);
      gzD[i] = -gg[i] * yy[i] * k.sD[i] * (1.f - k.sD[i]);
    }
    gX[0] -= gg[0] * k.sD[0];
    gX[1] -= gg[1] * k.sD[1];
    gX[3] -= gg[2] * k.sD[2];
    gX[4] -= gg[3] * k.sD[3];
    gq[0] = g[2] * k.sQ[0] * (1.f - k.sQ[0]);
    gq[1] = -g[2] * Y[2] * k.sQ[1] * (1.f - k.sQ[1]);
    gq[2] = g[5] * k.sQ[2] * (1.f - k.sQ[2]);
    gq[3] = -g[5] * Y[5] * k.sQ[3] * (1.f - k.sQ[3]);
    gX[2] -= g[2] * k.sQ[1];
    gX[5] -= g[5] * k.sQ[3];
    // hidden cotangents: gpre = relu'(pre) * (Wp^T gzp + Wd^T gzd)
    float gpre[7][4];
    {
      float ph[4], pl[4], dh[4], dl[4], qh[4], ql[4];
      frag_of(gzP, ph, pl);
      frag_of(gzD, dh, dl);
      frag_of(gq, qh, ql);
      const float zero[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float m[4];
        mma3c(m, ph, pl, bt[(18 + nt) * 32], zero);
        mma3c(m, dh, dl, bt[(22 + nt) * 32], m);
#pragma unroll
        for (int i = 0; i < 4; ++i) gpre[nt][i] = k.hid[nt][i] > 0.f ? m[i] : 0.f;
      }
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        float m[4];
        mma3c(m, qh, ql, bt[(26 + nt) * 32], zero);
#pragma unroll
        for (int i = 0; i < 4; ++i) gpre[4 + nt][i] = k.hid[4 + nt][i] > 0.f ? m[i] : 0.f;
      }
    }
#pragma unroll
    for (int nt = 0; nt < 7; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) gc.ghc[nt][i] += gpre[nt][i];
    // input cotangents: gx = W1x^T gpre + Q1x^T gqpre  (six independent accumulator chains)
    float aS[3][4], aQ[3][4];
    {
      const float zero[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float ah[4], al[4];
        frag_of(gpre[j], ah, al);
        const float4 b = bt[(29 + j) * 32];
        if (j == 0) {
          mma8(aS[0], ah, b.x, b.y, zero);
          mma8(aS[1], al, b.x, b.y, zero);
          mma8(aS[2], ah, b.z, b.w, zero);
        } else {
          mma3x(aS, ah, al, b);
        }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float ah[4], al[4];
        frag_of(gpre[4 + j], ah, al);
        const float4 b = bt[(33 + j) * 32];
        if (j == 0) {
          mma8(aQ[0], ah, b.x, b.y, zero);
          mma8(aQ[1], al, b.x, b.y, zero);
          mma8(aQ[2], ah, b.z, b.w, zero);
        } else {
          mma3x(aQ, ah, al, b);
        }
      }
    }
    float r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = ((aS[1][i] + aS[2][i]) + (aQ[1][i] + aQ[2][i])) + (aS[0][i] + aQ[0][i]);
    gX[0] += r[0];
    if (hasB) gX[1] += r[1];
    gX[3] += r[2];
    if (hasB) gX[4] += r[3];
    float in[4];
    inputs(t, Y, in);
    gw.put(in, k.hid, gpre, gzP, gzD, gq);
  }
  template <typename GW>
  __device__ void vjp(float t, const float* Y, const float* g, float* gX, Grad& gc, GW& gw) const {
    Kept k;
    forward(t, Y, k);
    vjp_kept(t, Y, k, g, gX, gc, gw);
  }
};

// Reverse of one step (the scheme of rk_step_vjp, vh_traj.cuh) with the kept activations of the re-evaluated stages
// parked in shared memory ([stage][item][lane], conflict-free) instead of registers: 40 values per stage would not fit
// next to the adjoint's own working set.
template <class F>
struct KeptStore {
  static constexpr int ITEMS = 40;
  __device__ static void put(float* sm, const typename WarpMlp<F>::Kept& k) {
#pragma unroll
    for (int nt = 0; nt < 7; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) sm[(nt * 4 + i) * 32] = k.hid[nt][i];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sm[(28 + i) * 32] = k.sP[i];
      sm[(32 + i) * 32] = k.sD[i];
      sm[(36 + i) * 32] = k.sQ[i];
    }
  }
  __device__ static void get(const float* sm, typename WarpMlp<F>::Kept& k) {
#pragma unroll
    for (int nt = 0; nt < 7; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) k.hid[nt][i] = sm[(nt * 4 + i) * 32];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      k.sP[i] = sm[(28 + i) * 32];
      k.sD[i] = sm[(32 + i) * 32];
      k.sQ[i] = sm[(36 + i) * 32];
    }
  }
};

template <class F, class TB, typename GW>
__device__ void step_vjp(const WarpMlp<F>& f, float* kept_sm, float t0, float t1, float h, const float* x, float* lam,
                         typename WarpMlp<F>::Grad& gc, GW& gw) {
  constexpr int S = WarpMlp<F>::S, s = TB::s, nk = s > 1 ? s - 1 : 1;
  float k[nk][S];
#pragma unroll
  for (int i = 0; i + 1 < s; ++i) {
    float X[S];
#pragma unroll
    for (int q = 0; q < S; ++q) X[q] = x[q];
#pragma unroll
    for (int j = 0; j < i; ++j)
      if (TB::a(i, j) != 0.f) {
        const float ha = h * TB::a(i, j);
#pragma unroll
        for (int q = 0; q < S; ++q) X[q] += ha * k[j][q];
      }
    typename WarpMlp<F>::Kept kp;
    f.eval_keep(stage_time<TB, float>(i, t0, t1), X, k[i], kp);
    KeptStore<F>::put(kept_sm + i * KeptStore<F>::ITEMS * 32, kp);
  }
  float gk[s][S];
#pragma unroll
  for (int i = 0; i < s; ++i)
#pragma unroll
    for (int q = 0; q < S; ++q) gk[i][q] = (h * TB::b(i)) * lam[q];
#pragma unroll
  for (int i = s - 1; i >= 0; --i) {
    float X[S], gX[S];
#pragma unroll
    for (int q = 0; q < S; ++q) {
      X[q] = x[q];
      gX[q] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < i; ++j)
      if (TB::a(i, j) != 0.f) {
        const float ha = h * TB::a(i, j);
#pragma unroll
        for (int q = 0; q < S; ++q) X[q] += ha * k[j][q];
      }
    typename WarpMlp<F>::Kept kp;
    if (i + 1 < s)
      KeptStore<F>::get(kept_sm + i * KeptStore<F>::ITEMS * 32, kp);
    else
      f.forward(stage_time<TB, float>(i, t0, t1), X, kp);
    f.vjp_kept(stage_time<TB, float>(i, t0, t1), X, kp, gk[i], gX, gc, gw);
#pragma unroll
    for (int q = 0; q < S; ++q) lam[q] += gX[q];
#pragma unroll
    for (int j = 0; j < i; ++j)
      if (TB::a(i, j) != 0.f) {
        const float ha = h * TB::a(i, j);
#pragma unroll
        for (int q = 0; q < S; ++q) gk[j][q] += ha * gX[q];
      }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// per-trajectory constants c = [latent parameters (+ device offsets on the y block), treatments, dev_1hot], lane
// distributed as the A fragment of the folding GEMM: lane tg holds c[8 q + tg] and c[8 q + tg + 4], q = 0..2, of its
// two rows; the four initial-state parameters: lane tg samples slot tg.  Each theta column is sampled by exactly one
// lane of the quad; log q / log p partial sums are combined across the quad by the caller.
// ---------------------------------------------------------------------------------------------------------------
struct RowInfo {
  int n[2], b[2];
  bool active[2];
};

template <class F>
__device__ void load_consts(const Call<float>& a, const RowInfo& ri, int tg, bool store, float (*cv)[3][2], float* x0,
                            float* lq, float* lp) {
  const int n_lat = a.bb_nlat;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int n = ri.n[hf], b = ri.b[hf];
    const bool st = store && ri.active[hf];
    lq[hf] = 0.f;
    lp[hf] = 0.f;
    {
      const int src = a.slot_src[tg];
      float v = 0.f;
      if (src >= 0)
        v = sample_column(a, n, b, src, lq[hf], lp[hf], st);
      else if (src != VH_SLOT_UNUSED)
        v = a.extra[(size_t)(-1 - src) * a.N + n];
      x0[hf] = v;
    }
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 8 * q + tg + 4 * e;
        float v = 0.f;
        if (j < n_lat) {
          const int src = a.slot_src[4 + j];
          if (src >= 0)
            v = sample_column(a, n, b, src, lq[hf], lp[hf], st);
          else if (src != VH_SLOT_UNUSED)
            v = a.extra[(size_t)(-1 - src) * a.N + n];
          const int ky = j - (n_lat - a.bb_ny);
          if (ky >= 0 && ky < a.bb_noff) v += a.extra[(size_t)ky * a.N + n];
        } else if (j < n_lat + a.C) {
          v = a.treatments[(size_t)b * a.C + (j - n_lat)];
        } else if (j < n_lat + a.C + a.D) {
          v = a.dev_1hot[(size_t)b * a.D + (j - n_lat - a.C)];
        }
        cv[hf][q][e] = v;
      }
    for (int i = tg; i < a.n_free; i += 4) sample_column(a, n, b, a.free_cols[i], lq[hf], lp[hf], st);
  }
}

// hc | hpc = bias + W[:, c-columns] c  (A = c, B from the raw weights in shared memory), plus the constant-1 units
template <class F>
__device__ void fold_consts(const float* w, const float (*cv)[3][2], int g, int tg, float (*hc)[4]) {
  typedef typename F::L L;
#pragma unroll
  for (int nt = 0; nt < 7; ++nt) {
    const bool prec = nt >= 4;
    const int nn = prec ? nt - 4 : nt;
    const int h0 = 8 * nn + 2 * tg;
    const int HH = prec ? F::HP : F::H;
    const float bias0 = h0 < HH ? w[(prec ? L::qb1 : L::b1) + h0] : (h0 == HH ? 1.f : 0.f);
    const float bias1 = h0 + 1 < HH ? w[(prec ? L::qb1 : L::b1) + h0 + 1] : (h0 + 1 == HH ? 1.f : 0.f);
    float c[4] = {bias0, bias1, bias0, bias1};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      float ah[4], al[4];
      split(cv[0][q][0], ah[0], al[0]);
      split(cv[1][q][0], ah[1], al[1]);
      split(cv[0][q][1], ah[2], al[2]);
      split(cv[1][q][1], ah[3], al[3]);
      mma3_fly(c, ah, al, fold_b<F>(w, prec, q, nn, tg, g), fold_b<F>(w, prec, q, nn, tg + 4, g));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) hc[nt][i] = c[i];
  }
}

// segmented (by individual) reduction over the 8 rows held by the lanes of equal tg, one atomic per segment
struct QuadColRed {
  float *d_mu, *d_prec;
  int P;
  unsigned same[2];
  bool head[2];
  int b[2];
  __device__ void init(const Call<float>& a, const RowInfo& ri, int lane) {
    d_mu = a.d_q_mu;
    d_prec = a.d_q_prec;
    P = a.P;
    const unsigned full = 0xffffffffu;
    const int g = lane >> 2;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int bl = ri.active[hf] ? ri.b[hf] : -1;
      b[hf] = ri.b[hf];
      same[hf] = 0;
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        const int other = __shfl_down_sync(full, bl, 4 << o);
        if (g + (1 << o) < 8 && other == bl) same[hf] |= 1u << o;
      }
      const int prev = __shfl_up_sync(full, bl, 4);
      head[hf] = ri.active[hf] && (g == 0 || prev != bl);
    }
  }
  // every lane of the warp calls this together; k: column (may differ between lanes of different tg; < 0: none)
  __device__ void add(int k, const float* dmu, const float* dprec) const {
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float m = dmu[hf], p = dprec[hf];
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        const float mo = __shfl_down_sync(full, m, 4 << o);
        const float po = __shfl_down_sync(full, p, 4 << o);
        if (same[hf] & (1u << o)) {
          m += mo;
          p += po;
        }
      }
      if (head[hf] && k >= 0) {
        atomicAdd(d_mu + (size_t)b[hf] * P + k, m);
        atomicAdd(d_prec + (size_t)b[hf] * P + k, p);
      }
    }
  }
};

__device__ __forceinline__ RowInfo row_info(const Call<float>& a, int n_base, int g) {
  RowInfo ri;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int n = n_base + g + 8 * hf;
    ri.active[hf] = n < a.N;
    ri.n[hf] = ri.active[hf] ? n : a.N - 1;
    ri.b[hf] = ri.n[hf] / a.IW;
  }
  return ri;
}

// ---------------------------------------------------------------------------------------------------------------
// forward: sample / clip / log-probs, fold, integrate, observe + log-likelihood; one warp per 16 trajectories
// ---------------------------------------------------------------------------------------------------------------
template <class F, class TB>
__device__ void warp_forward(const Call<float>& a, int n_base, const float* w, const float4* tiles) {
  constexpr int NST = F::NST;
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const size_t N = a.N;
  const int T = a.T;
  const RowInfo ri = row_info(a, n_base, g);
  WarpMlp<F> f;
  f.bt = tiles + lane;
  f.tg = tg;
  f.hasB = tg + 4 < NST;
  float Y[6];
  {
    float cv[2][3][2], x0[2], lq[2], lp[2];
    load_consts<F>(a, ri, tg, true, cv, x0, lq, lp);
    fold_consts<F>(w, cv, g, tg, f.hc);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      lq[hf] += __shfl_xor_sync(0xffffffffu, lq[hf], 1);
      lq[hf] += __shfl_xor_sync(0xffffffffu, lq[hf], 2);
      lp[hf] += __shfl_xor_sync(0xffffffffu, lp[hf], 1);
      lp[hf] += __shfl_xor_sync(0xffffffffu, lp[hf], 2);
      if (tg == 0 && ri.active[hf]) {
        if (a.logp_theta) a.logp_theta[ri.n[hf]] = lp[hf];
        if (a.logq_theta) a.logq_theta[ri.n[hf]] = lq[hf];
      }
      Y[3 * hf + 0] = x0[hf];
      Y[3 * hf + 1] = f.hasB ? a.bb_init_latent : 0.f;  // models/dr_blackbox.py:101-104
      Y[3 * hf + 2] = a.bb_init_prec;
    }
  }
  float ll[2] = {0.f, 0.f};
  const float h0 = a.times[1] - a.times[0];
  const float* obs[2];
  float* xs[2];
  float* xpr[2];
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    obs[hf] = a.obs ? a.obs + ((size_t)ri.b[hf] * 4 + tg) * T : nullptr;
    xs[hf] = (a.x_states && ri.active[hf]) ? a.x_states + ri.n[hf] : nullptr;
    xpr[hf] = (a.x_predict && ri.active[hf]) ? a.x_predict + (size_t)tg * N + ri.n[hf] : nullptr;
  }
  float ob[2] = {0.f, 0.f};
  if (a.obs) {
    ob[0] = obs[0][0];
    ob[1] = obs[1][0];
  }
  float t0 = a.times[0], t1 = a.times[1];
  for (int k = 0; k < T; ++k) {
    const int kn = k + 1 < T ? k + 1 : k;
    float obn[2] = {0.f, 0.f};
    if (a.obs) {
      obn[0] = ld_early(obs[0] + kn);
      obn[1] = ld_early(obs[1] + kn);
    }
    const float t2 = ld_early(a.times + (k + 2 < T ? k + 2 : T - 1));
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const float xa = Y[3 * hf], xb = Y[3 * hf + 1], v = Y[3 * hf + 2];
      if (xs[hf]) {
        xs[hf][(size_t)tg * N] = xa;
        if (f.hasB) xs[hf][(size_t)(tg + 4) * N] = xb;
        xs[hf][(size_t)(NST + tg) * N] = v;
        xs[hf] += (size_t)F::S * N;
      }
      const float x0 = __shfl_sync(0xffffffffu, xa, lane & ~3);
      const float xp = tg == 0 ? xa : x0 * xa;  // models/dr_blackbox.py:112-121
      if (xpr[hf]) {
        *xpr[hf] = xp;
        xpr[hf] += (size_t)4 * N;
      }
      if (a.obs) {
        const float d = xp - ob[hf];
        ll[hf] += -0.5f * (Lim<float>::log2pi - vlog(v) + v * d * d);
      }
    }
    if (k + 1 < T) rk_step<WarpMlp<F>, TB>(f, t0, t1, TB::const_h ? h0 : (t1 - t0), Y);
    t0 = t1;
    t1 = t2;
    ob[0] = obn[0];
    ob[1] = obn[1];
  }
  if (a.logp_species) {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf)
      if (ri.active[hf]) a.logp_species[(size_t)ri.n[hf] * 4 + tg] = ll[hf];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// reverse, adjoint warp
// ---------------------------------------------------------------------------------------------------------------
template <class F, class TB>
__device__ void warp_backward(const Call<float>& a, int n_base, const float* w, const float4* tiles, float* ring,
                              float* kept_sm, int bar0) {
  typedef typename F::L L;
  constexpr int NST = F::NST;
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const size_t N = a.N;
  const int T = a.T;
  const RowInfo ri = row_info(a, n_base, g);
  WarpMlp<F> f;
  f.bt = tiles + lane;
  f.tg = tg;
  f.hasB = tg + 4 < NST;
  {
    float cv[2][3][2], x0[2], lq[2], lp[2];
    load_consts<F>(a, ri, tg, false, cv, x0, lq, lp);
    fold_consts<F>(w, cv, g, tg, f.hc);
  }
  typename WarpMlp<F>::Grad gc;
#pragma unroll
  for (int nt = 0; nt < 7; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) gc.ghc[nt][i] = 0.f;
  PanelSink sink{ring, 0, g, tg, bar0};
  float gl[2], glq[2], glp[2];
  const float* obs[2];
  const float* xs[2];
  const float* gxs[2];
  const float* gxpr[2];
  const size_t slab = (size_t)F::S * N;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int n = ri.n[hf];
    const bool act = ri.active[hf];
    gl[hf] = (a.g_logp_species && act) ? a.g_logp_species[(size_t)n * 4 + tg] : 0.f;
    glq[hf] = (a.g_logq_theta && act) ? a.g_logq_theta[n] : 0.f;
    glp[hf] = (a.g_logp_theta && act) ? a.g_logp_theta[n] : 0.f;
    obs[hf] = a.obs ? a.obs + ((size_t)ri.b[hf] * 4 + tg) * T : nullptr;
    xs[hf] = a.x_states + (size_t)(T - 1) * slab + n;
    gxs[hf] = (a.g_x_states && act) ? a.g_x_states + (size_t)(T - 1) * slab + n : nullptr;
    gxpr[hf] = (a.g_x_predict && act) ? a.g_x_predict + ((size_t)(T - 1) * 4 + tg) * N + n : nullptr;
  }
  auto load_state = [&](float* dst) {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      dst[3 * hf + 0] = xs[hf][(size_t)tg * N];
      dst[3 * hf + 1] = f.hasB ? xs[hf][(size_t)(tg + 4) * N] : 0.f;
      dst[3 * hf + 2] = xs[hf][(size_t)(NST + tg) * N];
    }
  };
  float lam[6], Y[6], Yp[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) lam[q] = 0.f;
  load_state(Y);
#pragma unroll
  for (int q = 0; q < 6; ++q) Yp[q] = Y[q];
  float ob[2] = {0.f, 0.f};
  if (a.obs) {
    ob[0] = obs[0][T - 1];
    ob[1] = obs[1][T - 1];
  }
  const float h0 = a.times[1] - a.times[0];
  float t0 = a.times[T - 1], t1 = t0;
  for (int k = T - 1; k >= 0; --k) {
    const int kp = k > 0 ? k - 1 : 0;
    if (k > 0) {
      xs[0] -= slab;
      xs[1] -= slab;
      load_state(Yp);
    }
    float obp[2] = {0.f, 0.f};
    if (a.obs) {
      obp[0] = ld_early(obs[0] + kp);
      obp[1] = ld_early(obs[1] + kp);
    }
    const float tp = ld_early(a.times + kp);
    if (k + 1 < T) step_vjp<F, TB>(f, kept_sm + lane, t0, t1, TB::const_h ? h0 : (t1 - t0), Y, lam, gc, sink);
    // emission at time k: lane tg owns observed signal tg
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const float xa = Y[3 * hf], v = Y[3 * hf + 2];
      const float x0 = __shfl_sync(0xffffffffu, xa, lane & ~3);
      const float xp = tg == 0 ? xa : x0 * xa;
      float gxp = gxpr[hf] ? *gxpr[hf] : 0.f;
      if (a.obs) {
        const float d = xp - ob[hf];
        gxp -= gl[hf] * v * d;
        lam[3 * hf + 2] += gl[hf] * 0.5f * (vdiv(1.f, v) - d * d);
      }
      // observe_vjp: gx0 += gxp0 + sum_o gxp_o x_o;  gx_o += gxp_o x0
      float to0 = tg == 0 ? 0.f : gxp * xa;
      to0 += __shfl_xor_sync(0xffffffffu, to0, 1);
      to0 += __shfl_xor_sync(0xffffffffu, to0, 2);
      lam[3 * hf] += tg == 0 ? gxp + to0 : gxp * x0;
      if (gxs[hf]) {
        lam[3 * hf + 0] += gxs[hf][(size_t)tg * N];
        if (f.hasB) lam[3 * hf + 1] += gxs[hf][(size_t)(tg + 4) * N];
        lam[3 * hf + 2] += gxs[hf][(size_t)(NST + tg) * N];
        gxs[hf] -= slab;
      }
      if (gxpr[hf]) gxpr[hf] -= (size_t)4 * N;
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) Y[q] = Yp[q];
    ob[0] = obp[0];
    ob[1] = obp[1];
    t1 = t0;
    t0 = tp;
  }
  // ---- tail: constants' weight gradients (through the weight-gradient warp) and the chain rule back to theta ----
  float cv[2][3][2], x0[2], lq[2], lp[2];
  load_consts<F>(a, ri, tg, false, cv, x0, lq, lp);
  bar_sync(bar0 + BAR_TAIL1, 64);  // the weight-gradient warp has consumed every evaluation panel
  {
    float* p0 = ring + g * RS;
    float* p1 = p0 + 8 * RS;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      *reinterpret_cast<float2*>(p0 + PG + 8 * nt + 2 * tg) = make_float2(gc.ghc[nt][0], gc.ghc[nt][1]);
      *reinterpret_cast<float2*>(p1 + PG + 8 * nt + 2 * tg) = make_float2(gc.ghc[nt][2], gc.ghc[nt][3]);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      *reinterpret_cast<float2*>(p0 + PGQ + 8 * j + 2 * tg) = make_float2(gc.ghc[4 + j][0], gc.ghc[4 + j][1]);
      *reinterpret_cast<float2*>(p1 + PGQ + 8 * j + 2 * tg) = make_float2(gc.ghc[4 + j][2], gc.ghc[4 + j][3]);
    }
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 8 * q + tg + 4 * e;
        const float one = j == F::NC ? 1.f : 0.f;  // column NC of the constants panel = 1: hidden-bias gradients
        p0[PH + j] = j < F::NC ? (ri.active[0] ? cv[0][q][e] : 0.f) : (ri.active[0] ? one : 0.f);
        p1[PH + j] = j < F::NC ? (ri.active[1] ? cv[1][q][e] : 0.f) : (ri.active[1] ? one : 0.f);
      }
  }
  __threadfence_block();
  bar_sync(bar0 + BAR_TAIL2, 64);
  // cotangent of c: gc = W1c^T ghc + Q1c^T ghpc, landing on the lane that sampled c
  float gcv[2][3][2];
  {
    float ah[7][4], al[7][4];
#pragma unroll
    for (int jt = 0; jt < 7; ++jt) frag_of(gc.ghc[jt], ah[jt], al[jt]);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int jt = 0; jt < 7; ++jt) {
        const bool prec = jt >= 4;
        const int jj = prec ? jt - 4 : jt;
        mma3_fly(c, ah[jt], al[jt], gc_b<F>(w, prec, jj, q, tg, g), gc_b<F>(w, prec, jj, q, tg + 4, g));
      }
      gcv[0][q][0] = c[0];
      gcv[0][q][1] = c[1];
      gcv[1][q][0] = c[2];
      gcv[1][q][1] = c[3];
    }
  }
  QuadColRed red;
  red.init(a, ri, lane);
  const int n_lat = a.bb_nlat;
  {  // initial-state slots 0..3: cotangent = lambda(t0) of x_tg
    const int src = a.slot_src[tg];
    float dmu[2] = {0.f, 0.f}, dprec[2] = {0.f, 0.f};
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const float gth = lam[3 * hf];
      if (!ri.active[hf]) continue;
      if (src >= 0)
        column_vjp(a, ri.n[hf], ri.b[hf], src, gth, glq[hf], glp[hf], dmu[hf], dprec[hf]);
      else if (src != VH_SLOT_UNUSED && a.d_extra)
        a.d_extra[(size_t)(-1 - src) * N + ri.n[hf]] = gth;
    }
    red.add(src >= 0 ? src : -1, dmu, dprec);
  }
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = 8 * q + tg + 4 * e;
      const int src = j < n_lat ? a.slot_src[4 + j] : VH_SLOT_UNUSED;
      float dmu[2] = {0.f, 0.f}, dprec[2] = {0.f, 0.f};
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        if (!ri.active[hf] || j >= n_lat) continue;
        const float gth = gcv[hf][q][e];
        const int ky = j - (n_lat - a.bb_ny);
        if (ky >= 0 && ky < a.bb_noff && a.d_extra) a.d_extra[(size_t)ky * N + ri.n[hf]] = gth;
        if (src >= 0)
          column_vjp(a, ri.n[hf], ri.b[hf], src, gth, glq[hf], glp[hf], dmu[hf], dprec[hf]);
        else if (src != VH_SLOT_UNUSED && a.d_extra)
          a.d_extra[(size_t)(-1 - src) * N + ri.n[hf]] = gth;
      }
      red.add(src >= 0 ? src : -1, dmu, dprec);
    }
  for (int i0 = 0; i0 < a.n_free; i0 += 4) {
    const int i = i0 + tg;
    const int k = i < a.n_free ? a.free_cols[i] : -1;
    float dmu[2] = {0.f, 0.f}, dprec[2] = {0.f, 0.f};
#pragma unroll
    for (int hf = 0; hf < 2; ++hf)
      if (ri.active[hf] && k >= 0) column_vjp(a, ri.n[hf], ri.b[hf], k, 0.f, glq[hf], glp[hf], dmu[hf], dprec[hf]);
    red.add(k, dmu, dprec);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// reverse, weight-gradient warp
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_a(const float* panel, int ks, int off, bool upper, float* ah, float* al) {
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const float* r0 = panel + (ks * 8 + tg) * RS + off + g;
  const float* r1 = r0 + 4 * RS;
  split(r0[0], ah[0], al[0]);
  split(r1[0], ah[2], al[2]);
  if (upper) {
    split(r0[8], ah[1], al[1]);
    split(r1[8], ah[3], al[3]);
  } else {
    ah[1] = al[1] = ah[3] = al[3] = 0.f;
  }
}
__device__ __forceinline__ void load_b(const float* panel, int ks, int off, float* bh, float* bl) {
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const float* r0 = panel + (ks * 8 + tg) * RS + off + g;
  split(r0[0], bh[0], bl[0]);
  split(r0[4 * RS], bh[1], bl[1]);
}
__device__ __forceinline__ void mma3_ab(float* c, const float* ah, const float* al, const float* bh, const float* bl) {
  mma8(c, al, bh[0], bh[1]);
  mma8(c, ah, bl[0], bl[1]);
  mma8(c, ah, bh[0], bh[1]);
}

template <class F>
__device__ void warp_wgrad(const Call<float>& a, float* ring, int nev, int bar0) {
  typedef typename F::L L;
  constexpr int NST = F::NST, H = F::H, HP = F::HP, NC = F::NC;
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  float cW1[2][4], cWpd[2][2][4], cQ1[2][4], cQpd[2][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int i = 0; i < 4; ++i) cW1[m][i] = cWpd[m][0][i] = cWpd[m][1][i] = cQ1[m][i] = cQpd[m][i] = 0.f;
  for (int e = 0; e < nev; ++e) {
    const int slot = e & 1;
    const float* panel = ring + slot * ROWS * RS;
    bar_sync(bar0 + BAR_FULL0 + slot, 64);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      float bxh[2], bxl[2], bph[2], bpl[2], bdh[2], bdl[2], bqh[2], bql[2];
      load_b(panel, ks, PX, bxh, bxl);
      load_b(panel, ks, PZP, bph, bpl);
      load_b(panel, ks, PZD, bdh, bdl);
      load_b(panel, ks, PQ, bqh, bql);
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        float ah[4], al[4];
        load_a(panel, ks, PG + 16 * m, true, ah, al);
        mma3_ab(cW1[m], ah, al, bxh, bxl);
        load_a(panel, ks, PH + 16 * m, true, ah, al);
        mma3_ab(cWpd[m][0], ah, al, bph, bpl);
        mma3_ab(cWpd[m][1], ah, al, bdh, bdl);
        load_a(panel, ks, PGQ + 16 * m, m == 0, ah, al);
        mma3_ab(cQ1[m], ah, al, bxh, bxl);
        load_a(panel, ks, PHP + 16 * m, m == 0, ah, al);
        mma3_ab(cQpd[m], ah, al, bqh, bql);
      }
    }
    if (e + 2 < nev) {
      __threadfence_block();
      bar_arrive(bar0 + BAR_EMPTY0 + slot, 64);
    }
  }
  float* d = a.d_weights;
  // C element i of m-tile m: (hidden unit, column) = (16 m + g + 8 (i >> 1), 2 tg + (i & 1))
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int h = 16 * m + g + 8 * (i >> 1), n = 2 * tg + (i & 1);
      if (h < H && n < NST) atomicAdd(d + L::W1 + h * L::nin + n, cW1[m][i]);
      if (h < HP && n <= NST) atomicAdd(d + L::Q1 + h * (L::nin + 1) + (n < NST ? 1 + n : 0), cQ1[m][i]);
      if (h <= HP) {
        const int o = n >> 1, pd = n & 1;
        atomicAdd(h < HP ? d + (pd ? L::Qd : L::Qp) + o * HP + h : d + (pd ? L::qbd : L::qbp) + o, cQpd[m][i]);
      }
      const int o = slot_of(n);
      if (h <= H && o < NST) {
        atomicAdd(h < H ? d + L::Wp + o * H + h : d + L::bp + o, cWpd[m][0][i]);
        atomicAdd(h < H ? d + L::Wd + o * H + h : d + L::bd + o, cWpd[m][1][i]);
      }
    }
  bar_sync(bar0 + BAR_TAIL1, 64);
  bar_sync(bar0 + BAR_TAIL2, 64);  // the adjoint warp has left ghc | ghpc and the constants panel in ring slot 0
  // constants' columns of the hidden layers and their biases: dW[h][NST + j] += sum_r ghc[r][h] c[r][j]; column NC = 1
  const float* panel = ring;
#pragma unroll
  for (int net = 0; net < 2; ++net) {
    float acc[2][3][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[m][q][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      float bh[3][2], bl[3][2];
#pragma unroll
      for (int q = 0; q < 3; ++q) load_b(panel, ks, PH + 8 * q, bh[q], bl[q]);
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        float ah[4], al[4];
        load_a(panel, ks, (net ? PGQ : PG) + 16 * m, net == 0 || m == 0, ah, al);
#pragma unroll
        for (int q = 0; q < 3; ++q) mma3_ab(acc[m][q], ah, al, bh[q], bl[q]);
      }
    }
    const int HH = net ? HP : H;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int h = 16 * m + g + 8 * (i >> 1), j = 8 * q + 2 * tg + (i & 1);
          if (h >= HH || j > NC) continue;
          float* dst;
          if (net == 0)
            dst = j < NC ? d + L::W1 + h * L::nin + NST + j : d + L::b1 + h;
          else
            dst = j < NC ? d + L::Q1 + h * (L::nin + 1) + 1 + NST + j : d + L::qb1 + h;
          atomicAdd(dst, acc[m][q][i]);
        }
  }
}

// A reverse CTA holds BWD_PAIRS (adjoint, weight-gradient) warp pairs = 4 warps with the two HMMA-heavy adjoint warps at
// warp 0 and 2: measured (tools/micro/warp_placement.cu), 225 such CTAs put every adjoint warp on a sub-partition of
// its own, whereas 450 two-warp CTAs leave two adjoint warps on one tensor pipe in a third of the sub-partitions.
constexpr int BWD_PAIRS = 2;
template <class F>
struct Smem {
  static constexpr int WRAW = (F::L::total + 3) & ~3;
  static constexpr size_t fwd_bytes = sizeof(float) * WRAW + sizeof(float4) * NT_FWD * 32;
  // per pair: the two-slot panel ring + the parked activations of the re-evaluated stages (KeptStore): [stages - 1][40][32]
  static constexpr int pair_floats(int stages) { return 2 * ROWS * RS + (stages > 1 ? stages - 1 : 0) * 40 * 32; }
  static constexpr size_t bwd_bytes(int stages) {
    return sizeof(float) * WRAW + sizeof(float4) * NT_ALL * 32 + sizeof(float) * BWD_PAIRS * pair_floats(stages);
  }
};

template <class F, class TB>
__global__ void __launch_bounds__(128) bbm_fwd_kernel(const Call<float> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* w = reinterpret_cast<float*>(smem_raw);
  float4* tiles = reinterpret_cast<float4*>(w + Smem<F>::WRAW);
  for (int i = threadIdx.x; i < F::L::total; i += blockDim.x) w[i] = a.weights[i];
  __syncthreads();
  build_tiles<F>(w, tiles, NT_FWD);
  __syncthreads();
  const int n_base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ROWS;
  if (n_base < a.N) warp_forward<F, TB>(a, n_base, w, tiles);
}

template <class F, class TB>
__global__ void __launch_bounds__(BWD_PAIRS * 64) bbm_bwd_kernel(const Call<float> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* w = reinterpret_cast<float*>(smem_raw);
  float4* tiles = reinterpret_cast<float4*>(w + Smem<F>::WRAW);
  for (int i = threadIdx.x; i < F::L::total; i += blockDim.x) w[i] = a.weights[i];
  __syncthreads();
  build_tiles<F>(w, tiles, NT_ALL);
  __syncthreads();
  const int warp = threadIdx.x >> 5, pair = warp >> 1;
  const int n_base = (blockIdx.x * BWD_PAIRS + pair) * ROWS;
  if (n_base >= a.N) return;  // a whole pair leaves together: its named barriers are its own
  float* ring = reinterpret_cast<float*>(tiles + NT_ALL * 32) + pair * Smem<F>::pair_floats(TB::s);
  const int bar0 = pair * 6;
  if ((warp & 1) == 0)
    warp_backward<F, TB>(a, n_base, w, tiles, ring, ring + 2 * ROWS * RS, bar0);
  else
    warp_wgrad<F>(a, ring, TB::s * (a.T - 1), bar0);
}
#endif  // __CUDACC__

}  // namespace bbm
}  // namespace vh
