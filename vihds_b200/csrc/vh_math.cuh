// Scalar helpers shared by device kernels and the host-side math check (tests/hostcheck).
// IEEE-accurate expf/logf/powf/division only: the parity target (<= 1e-4 relative on states and the ELBO against the
// reference's fp32 CPU path) is met in fp32 if and only if no fast-math approximations are used.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define VH_HD __host__ __device__ __forceinline__
#else
#define VH_HD inline
#endif

namespace vh {

VH_HD float vexp(float x) { return expf(x); }
VH_HD double vexp(double x) { return exp(x); }
VH_HD float vlog(float x) { return logf(x); }
VH_HD double vlog(double x) { return log(x); }
// powf / pow are ~150 SASS instructions inlined; they are only used in the per-trajectory set-up (Hill fractions) and
// its chain rule, ~30 call sites: one out-of-line copy keeps the kernels' straight-line prologue / epilogue (which a
// latency-bound launch executes with a cold instruction cache) small.
#if defined(__CUDACC__)
#define VH_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define VH_HD_NOINLINE inline
#endif
VH_HD_NOINLINE float vpow(float x, float y) { return powf(x, y); }
VH_HD_NOINLINE double vpow(double x, double y) { return pow(x, y); }
VH_HD float vsqrt(float x) { return sqrtf(x); }
VH_HD double vsqrt(double x) { return sqrt(x); }
// fp32 on the device: tanh on the SFU, 1 - 2 / (exp(2x) + 1) with ex2.approx + rcp.approx (6 instructions instead of the
// ~25 of tanhf; absolute error <= 3e-7, the class of fp32 round-off).  The NeuralPrecisions nets evaluate 9..14 of them
// per right-hand-side evaluation.  Host build and fp64 keep the library function.
VH_HD float vtanh(float x) {
#if defined(__CUDA_ARCH__)
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return fmaf(-2.f, r, 1.f);
#else
  return tanhf(x);
#endif
}
VH_HD double vtanh(double x) { return tanh(x); }

// Division inside the time loop.  IEEE `a / b` in fp32 costs ~12 SASS instructions, a convergence barrier and a
// slow-path subroutine that denormal operands (tiny importance weights in the reverse sweep) actually take; here:
// MUFU.RCP (<= 1 ulp) + one Newton correction of the quotient = 4 instructions, no branch, error < 1 ulp -- two
// orders of magnitude inside the 1e-4 parity tolerance.  b = 0 gives NaN instead of +-inf (the ELBO is non-finite
// either way, vihds/training.py:331).  fp64 and the host-side math check keep the plain quotient.
VH_HD float vdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  const float q = a * r;
  return fmaf(fmaf(-b, q, a), r, q);
#else
  return a / b;
#endif
}
VH_HD double vdiv(double a, double b) { return a / b; }

// A global load the compiler must issue where it is written.  Used for the one-step-ahead fetch of the small shared
// inputs (observations, grid time): as plain loads ptxas sank them below the ~400-instruction adjoint body, right in
// front of their first use, which exposed a full memory latency per time step (41 % of the reverse loop's stall
// samples at the icml size sat on that one move).
VH_HD float ld_early(const float* p) {
#if defined(__CUDA_ARCH__)
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
#else
  return *p;
#endif
}
VH_HD double ld_early(const double* p) {
#if defined(__CUDA_ARCH__)
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
#else
  return *p;
#endif
}

// fp32 on the device: sigmoid on the SFU (ex2.approx + rcp.approx, 4 instructions instead of ~14, and the head of the
// dependent chain of every right-hand-side evaluation).  The argument product carries |z| 2^-24 of error, i.e. an absolute
// error in sigma of at most sigma (1 - sigma) |z| 1e-7 < 4e-7; +-inf saturate, NaN propagates.  Host build and fp64:
// 1 / (1 + exp(-z)).
VH_HD float sigmoid_impl(float z) {
#if defined(__CUDA_ARCH__)
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
#else
  return vdiv(1.f, 1.f + vexp(-z));
#endif
}
VH_HD double sigmoid_impl(double z) { return vdiv(1.0, 1.0 + vexp(-z)); }
template <typename R>
VH_HD R sigmoid(R z) {
  return sigmoid_impl(z);
}

// Arithmetic whose rounding must not depend on the surrounding code: the emission (observation map + Gaussian
// log-likelihood) is compiled into several kernels -- the throughput form, the one-warp latency form and the scribe warp of
// the team kernel -- and a batch must give bit-identical per-sample terms whichever of them runs it
// (tests/test_gpu_properties.py: batch-composition invariance).  vmul_rn is a product the compiler may not contract into a
// following add / subtract; vfma is an explicit fused multiply-add.
VH_HD float vmul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
VH_HD double vmul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}
VH_HD float vfma(float a, float b, float c) { return fmaf(a, b, c); }
VH_HD double vfma(double a, double b, double c) { return fma(a, b, c); }

// torch.clamp semantics: NaN propagates (both comparisons false); gradient passes on the CLOSED interval.
template <typename R>
VH_HD R clampv(R x, R lo, R hi) {
  return x < lo ? lo : (x > hi ? hi : x);
}
template <typename R>
VH_HD R clampmask(R x, R lo, R hi) {
  return (x >= lo && x <= hi) ? R(1) : R(0);
}

template <typename R>
struct Lim;
template <typename R>
VH_HD R loglik_add(R ll, R pr, R lpr, R d);
template <>
struct Lim<float> {
  static constexpr float log2pi = 1.8378770664093453f;
};
template <>
struct Lim<double> {
  static constexpr double log2pi = 1.8378770664093453;
};

// ll + log N(d; 0, 1 / pr) with a pinned operation order (see vmul_rn above)
template <typename R>
VH_HD R loglik_add(R ll, R pr, R lpr, R d) {
  const R q = vmul_rn(pr, d);
  const R e = vfma(q, d, Lim<R>::log2pi - lpr);
  return vfma(R(-0.5), e, ll);
}

}  // namespace vh
