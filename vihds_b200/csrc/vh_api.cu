// C ABI of libvihds_b200.so (include/vihds_b200.h) + the small reduction / optimiser kernels around the ODE kernels.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "vh_dispatch.cuh"

namespace vh {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// white-box instantiations, one translation unit each (vh_inst.cu): inst_<dir>_<dtype>_m<vh_model>
#define VH_DECL_MODEL(m)                                                      \
  int inst_fwd_f32_m##m(const vh_problem*, const vh_fwd_io*, cudaStream_t); \
  int inst_fwd_f64_m##m(const vh_problem*, const vh_fwd_io*, cudaStream_t); \
  int inst_bwd_f32_m##m(const vh_problem*, const vh_bwd_io*, cudaStream_t); \
  int inst_bwd_f64_m##m(const vh_problem*, const vh_bwd_io*, cudaStream_t);
VH_DECL_MODEL(0) VH_DECL_MODEL(1) VH_DECL_MODEL(2) VH_DECL_MODEL(3) VH_DECL_MODEL(4) VH_DECL_MODEL(5)
VH_DECL_MODEL(7) VH_DECL_MODEL(8) VH_DECL_MODEL(9) VH_DECL_MODEL(10) VH_DECL_MODEL(11) VH_DECL_MODEL(12) VH_DECL_MODEL(13)
VH_DECL_MODEL(14)
#undef VH_DECL_MODEL
typedef int (*fwd_fn)(const vh_problem*, const vh_fwd_io*, cudaStream_t);
typedef int (*bwd_fn)(const vh_problem*, const vh_bwd_io*, cudaStream_t);
#define VH_ROW(d, t) {inst_##d##_##t##_m0, inst_##d##_##t##_m1, inst_##d##_##t##_m2, inst_##d##_##t##_m3, inst_##d##_##t##_m4, \
                     inst_##d##_##t##_m5, nullptr /* dr_blackbox: vh_blackbox.cu */, inst_##d##_##t##_m7, inst_##d##_##t##_m8,  \
                     inst_##d##_##t##_m9, inst_##d##_##t##_m10, inst_##d##_##t##_m11, inst_##d##_##t##_m12,                     \
                     inst_##d##_##t##_m13, inst_##d##_##t##_m14}
static const fwd_fn kFwd[2][VH_MODEL_COUNT] = {VH_ROW(fwd, f32), VH_ROW(fwd, f64)};
static const bwd_fn kBwd[2][VH_MODEL_COUNT] = {VH_ROW(bwd, f32), VH_ROW(bwd, f64)};
#undef VH_ROW
int launch_bb_fwd(const vh_problem*, const vh_fwd_io*, cudaStream_t);
int launch_bb_bwd(const vh_problem*, const vh_bwd_io*, cudaStream_t);

static int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return VH_ERR_CUDA;
  }
  return VH_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// IWAE reduction: one block per individual (vihds/training.py:134-148)
// ---------------------------------------------------------------------------------------------------------------
template <typename R>
__device__ R block_reduce(R v, bool is_max, R* sh) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const R other = __shfl_xor_sync(full, v, o);
    v = is_max ? (other > v ? other : v) : v + other;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  R r = sh[0];
  for (int i = 1; i < nwarp; ++i) r = is_max ? (sh[i] > r ? sh[i] : r) : r + sh[i];
  return r;
}

template <typename R>
__global__ void iwae_fwd_kernel(int IW, R inv_b_total, const R* __restrict__ lpx, const R* __restrict__ lp,
                                const R* __restrict__ lq, R* cost, R* log_w, R* w, R* g_lpx, R* g_lp, R* g_lq) {
  __shared__ R sh[32];
  const int b = blockIdx.x;
  const R NEG = -INFINITY;
  R mx = NEG;
  bool has_nan = false;
  for (int i = threadIdx.x; i < IW; i += blockDim.x) {
    const size_t n = (size_t)b * IW + i;
    const R v = ((lpx[n * 4 + 0] + lpx[n * 4 + 1]) + (lpx[n * 4 + 2] + lpx[n * 4 + 3])) + lp[n] - lq[n];
    if (log_w) log_w[n] = v;
    has_nan |= (v != v);
    mx = v > mx ? v : mx;
  }
  mx = block_reduce(mx, true, sh);
  const R nanflag = block_reduce(has_nan ? R(1) : R(0), false, sh);
  const R shift = (mx == NEG || mx == INFINITY) ? R(0) : mx;  // torch.logsumexp: guard the all -inf / +inf rows
  R se = R(0);
  for (int i = threadIdx.x; i < IW; i += blockDim.x) {
    const size_t n = (size_t)b * IW + i;
    const R v = ((lpx[n * 4 + 0] + lpx[n * 4 + 1]) + (lpx[n * 4 + 2] + lpx[n * 4 + 3])) + lp[n] - lq[n];
    se += vexp(v - shift);
  }
  se = block_reduce(se, false, sh);
  R lse = vlog(se) + shift;
  if (nanflag > R(0)) lse = NAN;  // NaN contract: an ELBO term that is NaN must surface in the cost (training.py:331)
  if (w || g_lpx || g_lp || g_lq) {
    // normalised importance weights and (vh_iwae_fwd_bwd) the gradient of the cost for a unit upstream gradient
    for (int i = threadIdx.x; i < IW; i += blockDim.x) {
      const size_t n = (size_t)b * IW + i;
      const R v = ((lpx[n * 4 + 0] + lpx[n * 4 + 1]) + (lpx[n * 4 + 2] + lpx[n * 4 + 3])) + lp[n] - lq[n];
      const R wn = vexp(v - lse);
      if (w) w[n] = wn;
      const R g = -wn * inv_b_total;
      if (g_lpx) {
#pragma unroll
        for (int o = 0; o < 4; ++o) g_lpx[n * 4 + o] = g;
      }
      if (g_lp) g_lp[n] = g;
      if (g_lq) g_lq[n] = -g;
    }
  }
  if (threadIdx.x == 0) atomicAdd(cost, -(lse - vlog(R(IW))) * inv_b_total);
}

template <typename R>
__global__ void iwae_bwd_kernel(size_t N, R inv_b_total, const R* __restrict__ w, const R* __restrict__ g, R* g_lpx,
                                R* g_lp, R* g_lq) {
  const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const R gg = g ? g[0] : R(1);
  const R v = -w[n] * inv_b_total * gg;
  if (g_lpx) {
#pragma unroll
    for (int o = 0; o < 4; ++o) g_lpx[n * 4 + o] = v;
  }
  if (g_lp) g_lp[n] = v;
  if (g_lq) g_lq[n] = -v;
}

// ---------------------------------------------------------------------------------------------------------------
// Importance-weighted trace moments (vihds/utils.py:79-99): one warp per (time, row, individual), lanes over IW
// ---------------------------------------------------------------------------------------------------------------
template <typename R>
__global__ void iw_moments_kernel(int B, int IW, int T, int S, int NS, const R* __restrict__ w, const R* __restrict__ xs,
                                  const R* __restrict__ xp, const R* __restrict__ prec_const, R* mu, R* sd, R* st, R* var) {
  const int warps_per_block = blockDim.x >> 5;
  const long long gw = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int rows = S + 4;
  const long long total = (long long)T * rows * B;
  if (gw >= total) return;
  const int b = (int)(gw % B);
  const int row = (int)((gw / B) % rows);
  const int t = (int)(gw / ((long long)B * rows));
  const size_t N = (size_t)B * IW;
  const unsigned full = 0xffffffffu;
  if (row < S) {  // iw_states (dynamic-precision rows are not part of x_states in the reference, skipped by caller)
    R acc = R(0);
    for (int i = lane; i < IW; i += 32) acc += w[(size_t)b * IW + i] * xs[((size_t)t * S + row) * N + (size_t)b * IW + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(full, acc, o);
    if (lane == 0 && row < NS) st[((size_t)b * NS + row) * T + t] = acc;
  } else {
    const int o4 = row - S;
    R m1 = R(0), m2 = R(0), v = R(0);
    for (int i = lane; i < IW; i += 32) {
      const size_t n = (size_t)b * IW + i;
      const R wi = w[n];
      const R x = xp[((size_t)t * 4 + o4) * N + n];
      const R pr = prec_const ? prec_const[(size_t)o4 * N + n] : xs[((size_t)t * S + NS + o4) * N + n];
      m1 += wi * x;
      m2 += wi * (x * x + R(1) / pr);
      v += wi / pr;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m1 += __shfl_xor_sync(full, m1, o);
      m2 += __shfl_xor_sync(full, m2, o);
      v += __shfl_xor_sync(full, v, o);
    }
    if (lane == 0) {
      const size_t idx = ((size_t)b * 4 + o4) * T + t;
      mu[idx] = m1;
      sd[idx] = vsqrt(m2 - m1 * m1);
      var[idx] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fused Adam (torch.optim.Adam, default flags) over one flat vector
// ---------------------------------------------------------------------------------------------------------------
template <typename R>
__global__ void adam_kernel(size_t n, R* __restrict__ p, const R* __restrict__ g, R* __restrict__ m, R* __restrict__ v,
                            R lr, R b1, R b2, R eps, R bc1, R bc2_sqrt) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const R gi = g[i];
  const R mi = m[i] + (gi - m[i]) * (R(1) - b1);  // lerp form used by torch
  const R vi = b2 * v[i] + (R(1) - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const R denom = vsqrt(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

// graph-capturable variant: hyper-parameters and step counter are read from device memory.  step[0] = updates done so
// far, step[1] = ticket counter: every CTA reads step[0] before it takes a ticket, the last one bumps the counter
// (no separate increment launch).  zero_grad: the gradient is cleared once consumed (no separate fill launch).
// guard: the step's cost on the device (or NULL).  A NaN cost means NaN gradients (training.py:331-333 stops BEFORE
// optimizer.step()): the update of parameters and moments is skipped, the gradient is still cleared, step[0] stays and
// step[2] counts the skipped call -- the host reads it when it next looks at the cost.  The refusal is sticky (every
// later guarded call is skipped too until the host clears step[2]): the reference stops training at the first NaN.
template <typename R>
__global__ void adam_dev_kernel(size_t n, R* __restrict__ p, R* __restrict__ g, R* __restrict__ m, R* __restrict__ v,
                                const double* __restrict__ hyper, long long* step, int zero_grad, const R* __restrict__ guard) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ R s_bc[2];
  __shared__ int s_skip;
  if (threadIdx.x == 0) {  // the bias corrections: two double-precision pow per CTA, not per thread
    const double t = (double)(*(volatile long long*)step + 1);
    s_bc[0] = (R)(1.0 - pow(hyper[1], t));
    s_bc[1] = (R)sqrt(1.0 - pow(hyper[2], t));
    const R c = guard ? *guard : R(0);
    s_skip = (c != c) || (guard && *(volatile long long*)(step + 2) != 0);  // sticky: training stops at the first NaN
  }
  __syncthreads();
  const bool skip = s_skip != 0;
  if (i < n) {
    if (!skip) {
      const double lr = hyper[0], b1d = hyper[1], b2d = hyper[2];
      const R b1 = (R)b1d, b2 = (R)b2d, eps = (R)hyper[3];
      const R bc1 = s_bc[0], bc2_sqrt = s_bc[1];
      const R gi = g[i];
      const R mi = m[i] + (gi - m[i]) * (R(1) - b1);
      const R vi = b2 * v[i] + (R(1) - b2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      const R denom = vsqrt(vi) / bc2_sqrt + eps;
      p[i] -= ((R)lr / bc1) * (mi / denom);
    }
    if (zero_grad) g[i] = R(0);
  }
  __syncthreads();  // this CTA has read step[0] (thread 0, above) and is done
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd((unsigned long long*)(step + 1), 1ULL);
    if (ticket == (unsigned long long)gridDim.x - 1) {
      step[1] = 0;
      step[skip ? 2 : 0] += 1;
    }
  }
}
__global__ void step_inc_kernel(long long* step) { step[0] += 1; }

static int dr_num_weights(const vh_problem* p) {
  if (!model_is_dyn(p->model) || p->model == VH_MODEL_DR_BLACKBOX) return 0;
  const int nin = model_species(p->model) + 1;
  if (p->n_hidden == 0) return 2 * (4 * nin + 4);
  return p->n_hidden * (nin + 1) + 2 * (4 * p->n_hidden + 4);
}

}  // namespace vh

using namespace vh;

extern "C" {

int vh_abi_version(void) { return VH_ABI_VERSION; }
const char* vh_last_error(void) { return g_err; }

int vh_model_id(const char* key) {
  static const char* const names[VH_MODEL_COUNT] = {"dr_constant", "dr_constant_v2", "dr_constant_precisions",
                                                    "dr_constant_precisions_v2", "relay_constant",
                                                    "relay_constant_precisions", "dr_blackbox", "auto_constant",
                                                    "auto_constant_precisions", "prpr_constant", "prpr_constant_precisions",
                                                    "inducer_constant", "inducer_constant_precisions", "degrader_constant",
                                                    "degrader_constant_precisions"};
  for (int i = 0; i < VH_MODEL_COUNT; ++i)
    if (key && !strcmp(key, names[i])) return i;
  set_error("no kernel for model '%s'", key ? key : "(null)");
  return VH_ERR_UNSUPPORTED;
}

int vh_solver_id(const char* name) {
  static const char* const names[VH_SOLVER_COUNT] = {"euler", "midpoint", "rk4", "modeuler", "modeulerwhile"};
  for (int i = 0; i < VH_SOLVER_COUNT; ++i)
    if (name && !strcmp(name, names[i])) return i;
  set_error("solver '%s' has no fixed-step kernel (adaptive solvers are unsupported; there is no CPU fallback)",
            name ? name : "(null)");
  return VH_ERR_UNSUPPORTED;
}

int vh_num_slots(int model) {
  if (model_is_dr_family(model)) return DR_NSLOT;
  if (model == VH_MODEL_DR_BLACKBOX) return 4 + 16;  // init_x..init_cfp + up to 16 latent parameters (z, x, y order)
  return 0;
}

const char* vh_slot_name(int model, int s) {
  if (model_is_dr_family(model)) {
    if (s < 0 || s >= DR_NSLOT) return "";
    if (model_is_dyn(model) && s >= S_prec_x && s <= S_prec_cfp) return kDrDynPrecNames[s - S_prec_x];
    return kDrSlotNames[s];
  }
  if (model == VH_MODEL_DR_BLACKBOX) {
    static const char* const init[4] = {"init_x", "init_rfp", "init_yfp", "init_cfp"};
    static const char* const lat[16] = {"latent0", "latent1", "latent2", "latent3", "latent4", "latent5", "latent6", "latent7",
                                        "latent8", "latent9", "latent10", "latent11", "latent12", "latent13", "latent14", "latent15"};
    if (s >= 0 && s < 4) return init[s];
    if (s >= 4 && s < 20) return lat[s - 4];  // host maps z1..z_nz, x1..x_nx, y1..y_ny onto latent0.. in that order
  }
  return "";
}

int vh_num_species(int model) { return model_species(model); }

int vh_state_width(const vh_problem* p) {
  if (!p) return VH_ERR_INVALID;
  if (p->model == VH_MODEL_DR_BLACKBOX) return 4 + p->n_latent + 4;
  return model_species(p->model) + (model_is_dyn(p->model) ? 4 : 0);
}

size_t vh_num_weights(const vh_problem* p) {
  if (!p) return 0;
  if (p->model == VH_MODEL_DR_BLACKBOX) {
    const int nlat = p->n_z + p->n_x + p->n_y;
    const int nst = 4 + p->n_latent;
    const int nin = nst + nlat + p->C + p->D;
    // NeuralStates: W1[H][nin], b1[H], Wp[nst][H], bp[nst], Wd[nst][H], bd[nst];  precisions: W1[Hp][nin+1], b1[Hp],
    // Wp[4][Hp], bp[4], Wd[4][Hp], bd[4]
    const int H = p->n_hidden_states, Hp = p->n_hidden;
    return (size_t)H * (nin + 1) + 2 * (size_t)nst * (H + 1) + (size_t)Hp * (nin + 2) + 2 * 4 * (size_t)(Hp + 1);
  }
  return (size_t)dr_num_weights(p);
}

static int run_fwd(const vh_problem* p, const vh_fwd_io* io, void* stream) {
  if (!p || !io) {
    set_error("null argument");
    return VH_ERR_INVALID;
  }
  if (p->solver < 0 || p->solver >= VH_SOLVER_COUNT) {
    set_error("unknown solver id %d", p->solver);
    return VH_ERR_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (p->dtype != VH_F32 && p->dtype != VH_F64) {
    set_error("unknown dtype %d", p->dtype);
    return VH_ERR_INVALID;
  }
  if (p->model == VH_MODEL_DR_BLACKBOX) return launch_bb_fwd(p, io, s);
  if (!model_is_dr_family(p->model)) {
    set_error("model %d has no kernel", p->model);
    return VH_ERR_UNSUPPORTED;
  }
  return kFwd[p->dtype][p->model](p, io, s);
}

static int run_bwd(const vh_problem* p, const vh_bwd_io* io, void* stream) {
  if (!p || !io) {
    set_error("null argument");
    return VH_ERR_INVALID;
  }
  if (p->solver < 0 || p->solver >= VH_SOLVER_COUNT) {
    set_error("unknown solver id %d", p->solver);
    return VH_ERR_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (p->dtype != VH_F32 && p->dtype != VH_F64) {
    set_error("unknown dtype %d", p->dtype);
    return VH_ERR_INVALID;
  }
  if (p->model == VH_MODEL_DR_BLACKBOX) return launch_bb_bwd(p, io, s);
  if (!model_is_dr_family(p->model)) {
    set_error("model %d has no kernel", p->model);
    return VH_ERR_UNSUPPORTED;
  }
  return kBwd[p->dtype][p->model](p, io, s);
}

int vh_elbo_terms_fwd(const vh_problem* p, const vh_fwd_io* io, void* stream) { return run_fwd(p, io, stream); }
int vh_elbo_terms_bwd(const vh_problem* p, const vh_bwd_io* io, void* stream) { return run_bwd(p, io, stream); }

int vh_simulate(const vh_problem* p, const vh_fwd_io* io, void* stream) {
  if (p && p->P != 0) {
    set_error("vh_simulate: theta is passed as `extra` rows, P must be 0");
    return VH_ERR_INVALID;
  }
  return run_fwd(p, io, stream);
}
int vh_simulate_bwd(const vh_problem* p, const vh_bwd_io* io, void* stream) {
  if (p && p->P != 0) {
    set_error("vh_simulate_bwd: theta is passed as `extra` rows, P must be 0");
    return VH_ERR_INVALID;
  }
  return run_bwd(p, io, stream);
}

static int iwae_fwd_launch(const char* who, int dtype, int B, int IW, int b_total, const void* lpx, const void* lp,
                           const void* lq, void* cost, void* log_w, void* w, void* g_lpx, void* g_lp, void* g_lq, void* stream) {
  if (B <= 0 || IW <= 0 || b_total <= 0 || !lpx || !lp || !lq || !cost) {
    set_error("%s: bad arguments", who);
    return VH_ERR_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int block = IW >= 256 ? 256 : (IW >= 128 ? 128 : (IW >= 64 ? 64 : 32));
  if (dtype == VH_F32) {
    cudaMemsetAsync(cost, 0, sizeof(float), s);
    iwae_fwd_kernel<float><<<B, block, 0, s>>>(IW, 1.0f / (float)b_total, (const float*)lpx, (const float*)lp,
                                                (const float*)lq, (float*)cost, (float*)log_w, (float*)w, (float*)g_lpx,
                                                (float*)g_lp, (float*)g_lq);
  } else if (dtype == VH_F64) {
    cudaMemsetAsync(cost, 0, sizeof(double), s);
    iwae_fwd_kernel<double><<<B, block, 0, s>>>(IW, 1.0 / (double)b_total, (const double*)lpx, (const double*)lp,
                                                 (const double*)lq, (double*)cost, (double*)log_w, (double*)w, (double*)g_lpx,
                                                 (double*)g_lp, (double*)g_lq);
  } else {
    set_error("unknown dtype %d", dtype);
    return VH_ERR_INVALID;
  }
  return check_launch("iwae_fwd_kernel");
}

int vh_iwae_fwd(int dtype, int B, int IW, int b_total, const void* lpx, const void* lp, const void* lq, void* cost,
                void* log_w, void* w, void* stream) {
  return iwae_fwd_launch("vh_iwae_fwd", dtype, B, IW, b_total, lpx, lp, lq, cost, log_w, w, nullptr, nullptr, nullptr, stream);
}

int vh_iwae_fwd_bwd(int dtype, int B, int IW, int b_total, const void* lpx, const void* lp, const void* lq, void* cost,
                    void* log_w, void* w, void* g_lpx, void* g_lp, void* g_lq, void* stream) {
  return iwae_fwd_launch("vh_iwae_fwd_bwd", dtype, B, IW, b_total, lpx, lp, lq, cost, log_w, w, g_lpx, g_lp, g_lq, stream);
}

int vh_iwae_bwd(int dtype, int B, int IW, int b_total, const void* w, const void* g, void* g_lpx, void* g_lp, void* g_lq,
                void* stream) {
  if (B <= 0 || IW <= 0 || b_total <= 0 || !w) {
    set_error("vh_iwae_bwd: bad arguments");
    return VH_ERR_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t N = (size_t)B * IW;
  const int block = 128;
  const int grid = (int)((N + block - 1) / block);
  if (dtype == VH_F32)
    iwae_bwd_kernel<float><<<grid, block, 0, s>>>(N, 1.0f / (float)b_total, (const float*)w, (const float*)g, (float*)g_lpx,
                                                   (float*)g_lp, (float*)g_lq);
  else if (dtype == VH_F64)
    iwae_bwd_kernel<double><<<grid, block, 0, s>>>(N, 1.0 / (double)b_total, (const double*)w, (const double*)g,
                                                    (double*)g_lpx, (double*)g_lp, (double*)g_lq);
  else {
    set_error("unknown dtype %d", dtype);
    return VH_ERR_INVALID;
  }
  return check_launch("iwae_bwd_kernel");
}

int vh_iw_moments(const vh_problem* p, const void* w, const void* x_states, const void* x_predict, const void* prec_const,
                  void* mu, void* sd, void* st, void* var, void* stream) {
  if (!p || !w || !x_states || !x_predict || !mu || !sd || !st || !var) {
    set_error("vh_iw_moments: bad arguments");
    return VH_ERR_INVALID;
  }
  const int S = vh_state_width(p);
  const int NS = p->model == VH_MODEL_DR_BLACKBOX ? 4 + p->n_latent : model_species(p->model);
  if ((S == NS) != (prec_const != nullptr)) {
    set_error("vh_iw_moments: prec_const must be given exactly for constant-precision models");
    return VH_ERR_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const long long total = (long long)p->T * (S + 4) * p->B;
  const int block = 128;
  const long long grid = (total + 3) / 4;
  if (p->dtype == VH_F32)
    iw_moments_kernel<float><<<(unsigned)grid, block, 0, s>>>(p->B, p->IW, p->T, S, NS, (const float*)w, (const float*)x_states,
                                                               (const float*)x_predict, (const float*)prec_const, (float*)mu,
                                                               (float*)sd, (float*)st, (float*)var);
  else
    iw_moments_kernel<double><<<(unsigned)grid, block, 0, s>>>(p->B, p->IW, p->T, S, NS, (const double*)w,
                                                                (const double*)x_states, (const double*)x_predict,
                                                                (const double*)prec_const, (double*)mu, (double*)sd,
                                                                (double*)st, (double*)var);
  return check_launch("iw_moments_kernel");
}

int vh_adam_step(int dtype, size_t n, void* param, const void* grad, void* exp_avg, void* exp_avg_sq, double lr, double beta1,
                 double beta2, double eps, int step, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || step < 1) {
    set_error("vh_adam_step: bad arguments");
    return VH_ERR_INVALID;
  }
  if (n == 0) return VH_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2s = sqrt(1.0 - pow(beta2, (double)step));
  const int block = 256;
  const unsigned grid = (unsigned)((n + block - 1) / block);
  if (dtype == VH_F32)
    adam_kernel<float><<<grid, block, 0, s>>>(n, (float*)param, (const float*)grad, (float*)exp_avg, (float*)exp_avg_sq,
                                               (float)lr, (float)beta1, (float)beta2, (float)eps, (float)bc1, (float)bc2s);
  else if (dtype == VH_F64)
    adam_kernel<double><<<grid, block, 0, s>>>(n, (double*)param, (const double*)grad, (double*)exp_avg, (double*)exp_avg_sq,
                                                lr, beta1, beta2, eps, bc1, bc2s);
  else {
    set_error("unknown dtype %d", dtype);
    return VH_ERR_INVALID;
  }
  return check_launch("adam_kernel");
}

int vh_adam_step_dev(int dtype, size_t n, void* param, void* grad, void* exp_avg, void* exp_avg_sq, const void* hyper,
                     void* step, int zero_grad, const void* guard, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !hyper || !step) {
    set_error("vh_adam_step_dev: bad arguments");
    return VH_ERR_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) {
    step_inc_kernel<<<1, 1, 0, s>>>((long long*)step);
    return check_launch("step_inc_kernel");
  }
  const int block = 256;
  const unsigned grid = (unsigned)((n + block - 1) / block);
  if (dtype == VH_F32)
    adam_dev_kernel<float><<<grid, block, 0, s>>>(n, (float*)param, (float*)grad, (float*)exp_avg, (float*)exp_avg_sq,
                                                   (const double*)hyper, (long long*)step, zero_grad, (const float*)guard);
  else if (dtype == VH_F64)
    adam_dev_kernel<double><<<grid, block, 0, s>>>(n, (double*)param, (double*)grad, (double*)exp_avg, (double*)exp_avg_sq,
                                                    (const double*)hyper, (long long*)step, zero_grad, (const double*)guard);
  else {
    set_error("unknown dtype %d", dtype);
    return VH_ERR_INVALID;
  }
  return check_launch("adam_dev_kernel");
}

int vh_zero_async(void* dst, size_t bytes, void* stream) {
  if (!dst) {
    set_error("vh_zero_async: null pointer");
    return VH_ERR_INVALID;
  }
  cudaError_t e = cudaMemsetAsync(dst, 0, bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_error("vh_zero_async(%zu bytes): %s", bytes, cudaGetErrorString(e));
    return VH_ERR_CUDA;
  }
  return VH_OK;
}

int vh_copy_async(void* dst, const void* src, size_t bytes, void* stream) {
  if (!dst || !src) {
    set_error("vh_copy_async: null pointer");
    return VH_ERR_INVALID;
  }
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_error("vh_copy_async(%zu bytes): %s", bytes, cudaGetErrorString(e));
    return VH_ERR_CUDA;
  }
  return VH_OK;
}

}  // extern "C"
