/*
 * vihds_b200.h -- C ABI of libvihds_b200.so: the B200-native batched ODE-integration + ELBO engine that replaces the
 * hot path of microsoft/vi-hds.
 *
 * The reference has no FFI: its boundary is the Python plugin surface.  Each entry point below
 * names the reference interface (file:line under /root/reference) whose work it takes over; INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *  - every pointer in the *_io structs is a DEVICE pointer owned by the caller (PyTorch's allocator in practice);
 *    the library allocates nothing persistent and keeps no global mutable state; calls are asynchronous on `stream`
 *    (a cudaStream_t passed as void*; NULL = legacy default stream) and re-entrant;
 *  - host-side descriptors (vh_problem, slot maps, the io structs themselves) are read during the call only;
 *  - return value: 0 on success, a negative vh_status otherwise; vh_last_error() gives a thread-local message;
 *  - dtype: VH_F32 or VH_F64 (data.dtype, vihds/config.py:164-178); all floating-point buffers of one call share it;
 *  - trajectory index n = b * IW + i  (b individual, i importance sample), N = B * IW.
 *
 * Device layouts (chosen for coalescing across trajectories; the reference itself returns x_states as a permuted,
 * time-major view, vihds/ode.py:82):
 *    u          [N][P]        as produced by BaseVAE.sample_u (vihds/vae.py:22-24), row-major
 *    q_mu,q_prec[B][P]        encoder output per individual (global parameters repeated over B)
 *    theta      [P][N]        clipped samples, one contiguous [B,IW] plane per parameter
 *    extra      [E][N]        per-trajectory inputs that are not sampled (conditioned aR/aS, or all of theta in
 *                             vh_simulate)
 *    x_states   [T][S][N]     S = species + dynamic-precision states;   x_predict [T][4][N]
 *    observations [B][4][T]   as in the reference batch (vihds/training.py:47-68)
 *    logp_by_species [N][4];  logp_theta, logq_theta [N]
 */
#ifndef VIHDS_B200_H
#define VIHDS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VH_ABI_VERSION 4

enum vh_status {
  VH_OK = 0,
  VH_ERR_INVALID = -1,     /* bad argument / unsupported combination */
  VH_ERR_UNSUPPORTED = -2, /* e.g. adaptive solvers (dopri5/8): no fixed-step kernel, and no CPU fallback */
  VH_ERR_CUDA = -3         /* a CUDA runtime call failed */
};

enum vh_dtype { VH_F32 = 0, VH_F64 = 1 };

/* models.LOOKUP keys (models/__init__.py:19-35) that have a kernel */
enum vh_model {
  VH_MODEL_DR_CONSTANT = 0,               /* models/dr_constant.py:114-160, version 1 */
  VH_MODEL_DR_CONSTANT_V2 = 1,            /* models/dr_constant.py:163-166 */
  VH_MODEL_DR_CONSTANT_PRECISIONS = 2,    /* models/dr_constant.py:169-209 (NeuralPrecisions states) */
  VH_MODEL_DR_CONSTANT_PRECISIONS_V2 = 3, /* models/dr_constant.py:212-215 */
  VH_MODEL_RELAY_CONSTANT = 4,            /* models/relay_constant.py:137-196 */
  VH_MODEL_RELAY_CONSTANT_PRECISIONS = 5, /* models/relay_constant.py:199-263 */
  VH_MODEL_DR_BLACKBOX = 6,               /* models/dr_blackbox.py:61-125 */
  VH_MODEL_AUTO_CONSTANT = 7,             /* models/auto_constant.py:63-90 */
  VH_MODEL_AUTO_CONSTANT_PRECISIONS = 8,  /* models/auto_constant.py:93-132 */
  VH_MODEL_PRPR_CONSTANT = 9,             /* models/prpr_constant.py:61-81 */
  VH_MODEL_PRPR_CONSTANT_PRECISIONS = 10, /* models/prpr_constant.py:84-130 */
  /* the next four do not run in the reference as shipped (OdeModel.init_with_params / OdeFunc.__init__ arity,
   * SURVEY.md section 8c); their parity target is the reference under the two-line monkeypatch of oracle/ref_harness.py */
  VH_MODEL_INDUCER_CONSTANT = 11,             /* models/inducer_constant.py:82-111 */
  VH_MODEL_INDUCER_CONSTANT_PRECISIONS = 12,  /* models/inducer_constant.py:114-151 */
  VH_MODEL_DEGRADER_CONSTANT = 13,            /* models/degrader_constant.py:146-207 */
  VH_MODEL_DEGRADER_CONSTANT_PRECISIONS = 14, /* models/degrader_constant.py:210-269 */
  VH_MODEL_COUNT = 15
};

/* params.solver values (vihds/ode.py:75-81) */
enum vh_solver {
  VH_SOLVER_EULER = 0,         /* torchdiffeq fixed-grid euler */
  VH_SOLVER_MIDPOINT = 1,      /* torchdiffeq fixed-grid midpoint (spec default, vihds/config.py:59) */
  VH_SOLVER_RK4 = 2,           /* torchdiffeq 0.1 "rk4" = 3/8 rule */
  VH_SOLVER_MODEULER = 3,      /* vihds/solvers.py:9-17: Heun with CONSTANT h = times[1]-times[0] */
  VH_SOLVER_MODEULERWHILE = 4, /* vihds/solvers.py:20-41: Heun with per-step h */
  VH_SOLVER_COUNT = 5
};

enum vh_kind { VH_KIND_CONSTANT = 0, VH_KIND_NORMAL = 1, VH_KIND_LOGNORMAL = 2 };

#define VH_MAX_SLOTS 64
#define VH_SLOT_UNUSED (-1000000)

typedef struct vh_problem {
  int model;  /* vh_model */
  int solver; /* vh_solver */
  int dtype;  /* vh_dtype */
  int B, IW, T;
  int P; /* sampled parameters = columns of u (0 in vh_simulate) */
  int C; /* treatments per individual (2: C6, C12; degrader: 3: + Ara; inducer: 1: Ara) */
  int D; /* device one-hot width (used by the black-box model only) */
  int E; /* rows of `extra` (black-box: E == n_y device offsets ADDED to the sampled y_k, see vh_bb.cuh) */
  int n_hidden; /* NeuralPrecisions hidden width (0 = single linear layer), black-box: precision-net hidden width */
  int n_hidden_states; /* black-box NeuralStates hidden width */
  int n_latent; /* black-box latent species */
  int n_z, n_x, n_y; /* black-box latent-parameter counts */
  /* slot_src[s]: where model slot s (see vh_slot_name) takes its value from:
   *   >= 0  column of u / theta;   -1-e  row e of `extra`;   VH_SLOT_UNUSED  not provided (value 0). */
  int slot_src[VH_MAX_SLOTS];
  /* black-box only: initial value of the latent species and of the precision states (params.init_latent_species,
   * params.init_prec; vihds/config.py:74-75, models/dr_blackbox.py:101-104) */
  double init_latent_species, init_prec;
} vh_problem;

/* Forward: replaces, in ONE launch, q.sample + p.clip (vihds/distributions.py:119-142, :76-85), Decoder.forward
 * (vihds/decoders.py:28-45 -> OdeModel.simulate vihds/ode.py:66-82, expand_precisions, observe), and the per-sample
 * terms of Training.cost (vihds/training.py:130-136: log_prob_observations :24-44, q.log_prob, p.log_prob). */
typedef struct vh_fwd_io {
  const void* times;        /* [T] */
  const void* u;            /* [N][P]  (NULL iff P == 0) */
  const void* q_mu;         /* [B][P] */
  const void* q_prec;       /* [B][P] */
  const void* p_mu;         /* [P] prior mean */
  const void* p_prec;       /* [P] prior precision */
  const void* clip_lo;      /* [P] lower clip bound (prior mu - 4 sigma, exp'd for LogNormal); -inf = unclipped */
  const void* clip_hi;      /* [P] */
  const int* kind;          /* [P] vh_kind */
  const void* extra;        /* [E][N] or NULL */
  const void* treatments;   /* [B][C] log1p-transformed inputs (vihds/datasets.py:87) */
  const void* dev_1hot;     /* [B][D] or NULL (black-box only) */
  const void* observations; /* [B][4][T] or NULL (then no log-likelihood is accumulated) */
  const void* weights;      /* flat decoder weights (layout: vh_weight_layout) or NULL */
  void* theta;              /* out [P][N] or NULL */
  void* x_states;           /* out [T][S][N] or NULL */
  void* x_predict;          /* out [T][4][N] or NULL */
  void* logp_by_species;    /* out [N][4] or NULL */
  void* logp_theta;         /* out [N] or NULL */
  void* logq_theta;         /* out [N] or NULL */
} vh_fwd_io;

/* Backward (discrete adjoint of the exact stepper; what autograd's replay of the unrolled solve computes in the
 * reference, vihds/training.py:334).  Re-reads the x_states trace written by the forward call as its checkpoint. */
typedef struct vh_bwd_io {
  vh_fwd_io fwd;                 /* the forward call's inputs and its x_states output (x_predict/log* unused);
                                  * fwd.theta, if non-NULL, must be the forward call's theta OUTPUT: the reverse sweep
                                  * then reads it back (P coalesced loads) instead of re-sampling from u */
  const void* g_logp_by_species; /* [N][4] upstream gradients, any may be NULL (= zero) */
  const void* g_logp_theta;      /* [N] */
  const void* g_logq_theta;      /* [N] */
  const void* g_theta;           /* [P][N] */
  const void* g_x_states;        /* [T][S][N] */
  const void* g_x_predict;       /* [T][4][N] */
  void* d_q_mu;                  /* out [B][P], OVERWRITTEN (zeroed by the call, then accumulated) or NULL */
  void* d_q_prec;                /* out [B][P] */
  void* d_extra;                 /* out [E][N] or NULL */
  void* d_weights;               /* out flat, same layout as weights (zeroed by the call) or NULL */
  /* IWAE reduction fused into the reverse launch (the training step's vh_iwae_fwd_bwd + vh_elbo_terms_bwd as ONE launch;
   * vihds/training.py:134-148 and the unit upstream gradient of elbo.backward(), :334).  iwae_b_total > 0: the upstream
   * gradients are derived inside the kernel from the forward call's outputs fwd.logp_by_species / logp_theta / logq_theta
   * (all three required; g_logp_* must be NULL) and iwae_cost[0] receives -mean_b(logsumexp_i log_w - log IW) with the
   * mean's denominator iwae_b_total (zeroed by the call).  Available where the latency-form reverse kernel runs
   * (B*IW <= 18,944, white-box models without a hidden-layer precision net); VH_ERR_UNSUPPORTED otherwise. */
  void* iwae_cost;               /* out [1] or NULL */
  int iwae_b_total;              /* 0: not fused */
  /* non-zero: the caller zeroed d_q_mu, d_q_prec, d_weights and iwae_cost on this stream BEFORE the forward launch
   * (vh_elbo_terms_fwd) of the same step.  The call then adds no memset between the two kernels, and the latency-form
   * reverse kernel is launched with programmatic stream serialization: it becomes resident under the forward kernel's
   * tail and starts the moment that grid has completed (VIHDS_PDL=0 turns the launch attribute off).  Honoured by the
   * white-box models; the dr_blackbox launcher clears its outputs regardless. */
  int outputs_cleared;
} vh_bwd_io;

int vh_abi_version(void);
const char* vh_last_error(void);

/* registry helpers: the Python side builds slot maps from names, never from hard-coded indices */
int vh_model_id(const char* lookup_key);  /* models.LOOKUP key -> vh_model, or VH_ERR_UNSUPPORTED */
int vh_solver_id(const char* name);       /* params.solver -> vh_solver, VH_ERR_UNSUPPORTED for adaptive solvers */
int vh_num_slots(int model);
const char* vh_slot_name(int model, int slot); /* theta attribute name the RHS reads (e.g. "KGR_76") */
int vh_num_species(int model);                 /* OdeModel.n_species */
int vh_state_width(const vh_problem* p);       /* S = species + dynamic-precision states */
size_t vh_num_weights(const vh_problem* p);    /* length of the flat weight vector (0 for constant-precision models) */

int vh_elbo_terms_fwd(const vh_problem* p, const vh_fwd_io* io, void* stream);
int vh_elbo_terms_bwd(const vh_problem* p, const vh_bwd_io* io, void* stream);

/* Narrow seam: OdeModel.simulate (vihds/ode.py:66-82) alone -- theta supplied by the caller as `extra` rows
 * (P == 0); same kernels, no sampling / log-prob work.  x_states [T][S][N]. */
int vh_simulate(const vh_problem* p, const vh_fwd_io* io, void* stream);
int vh_simulate_bwd(const vh_problem* p, const vh_bwd_io* io, void* stream);

/* IWAE reduction (vihds/training.py:134-148): log_w = sum_species logp + logp_theta - logq_theta;
 * cost = -mean_b(logsumexp_i log_w - log IW).  One block per individual.  Outputs: cost[1] (zeroed by the call),
 * log_w[N], normalized importance weights w[N] (vihds/training.py:152-153).  b_total: denominator of the mean
 * (= B on one GPU; the global batch when individuals are sharded across ranks). */
int vh_iwae_fwd(int dtype, int B, int IW, int b_total, const void* logp_by_species, const void* logp_theta,
                const void* logq_theta, void* cost, void* log_w, void* w, void* stream);
/* gradient of cost*g w.r.t. the three term arrays: d/dlog_w = -w/b_total * g[0]; g is a device scalar (or NULL = 1) */
int vh_iwae_bwd(int dtype, int B, int IW, int b_total, const void* w, const void* g, void* g_logp_by_species,
                void* g_logp_theta, void* g_logq_theta, void* stream);

/* The training step's pair in ONE launch: vh_iwae_fwd followed by vh_iwae_bwd with g = 1 (elbo.backward() of
 * vihds/training.py:334 starts from a unit upstream gradient).  Same outputs; log_w, w and any of the three gradient
 * arrays may be NULL. */
int vh_iwae_fwd_bwd(int dtype, int B, int IW, int b_total, const void* logp_by_species, const void* logp_theta,
                    const void* logq_theta, void* cost, void* log_w, void* w, void* g_logp_by_species,
                    void* g_logp_theta, void* g_logq_theta, void* stream);

/* Evaluation path (vihds/utils.py:79-99, Results.init): importance-weighted moments of the traces, reduced over IW
 * on the device so the [B,IW,.,T] traces never leave HBM.
 *   iw_predict_mu [B][4][T], iw_predict_std [B][4][T], iw_states [B][S][T], iw_variance [B][4][T]
 * prec_const: [4][N] planes of theta for constant-precision models, NULL for dynamic-precision models. */
int vh_iw_moments(const vh_problem* p, const void* w, const void* x_states, const void* x_predict,
                  const void* prec_const, void* iw_predict_mu, void* iw_predict_std, void* iw_states,
                  void* iw_variance, void* stream);

/* Fused Adam over a flat parameter vector (torch.optim.Adam defaults; vihds/training.py:82, :336). step is 1-based. */
int vh_adam_step(int dtype, size_t n, void* param, const void* grad, void* exp_avg, void* exp_avg_sq, double lr,
                 double beta1, double beta2, double eps, int step, void* stream);
/* Same update with the hyper-parameters and the step counter ON THE DEVICE, so that the launch can be captured in a
 * CUDA graph and replayed: hyper = double[4] {lr, beta1, beta2, eps}; step = int64[4]: step[0] = the number of updates
 * done so far (the call uses step[0]+1 for the bias correction and increments it when its last thread block retires),
 * step[1] = scratch ticket counter, zero on entry and on exit, step[2] = number of calls skipped by the guard.
 * zero_grad != 0: grad is cleared once it has been consumed (optimizer.zero_grad() of the next step, training.py:333,
 * without a launch of its own).  guard: device pointer to the step's cost (or NULL): if it is NaN the parameters and
 * moments are left untouched -- the reference tests torch.isnan(elbo) BEFORE optimizer.step() (training.py:331-336) --
 * the gradient is still cleared and step[2] is incremented instead of step[0].  The refusal is sticky: while step[2] != 0
 * every later guarded call is refused as well (the reference stops training at the first NaN); the host clears step[2]
 * to resume. */
int vh_adam_step_dev(int dtype, size_t n, void* param, void* grad, void* exp_avg, void* exp_avg_sq, const void* hyper,
                     void* step, int zero_grad, const void* guard, void* stream);

/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream): the step's host <-> device copies (batch, u, conditioner
 * weights: `.to(device)` in vihds/training.py:326-329, vae.py:22-24) without a tensor library in between.  Host memory
 * should be pinned. */
int vh_copy_async(void* dst, const void* src, size_t bytes, void* stream);
/* cudaMemsetAsync(dst, 0, bytes): the `optimizer.zero_grad()` of training.py:333 for the reverse launch's accumulators
 * (vh_bwd_io.outputs_cleared) as a memset node of the captured step instead of a fill kernel */
int vh_zero_async(void* dst, size_t bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Multi-GPU exchange step fused with the optimiser (no counterpart in the single-device reference; it takes the place
 * of ncclAllReduce(flat gradient) + vh_adam_step_dev in the data-parallel step): every rank pushes its gradient vector
 * into all ranks' inboxes over NVLink peer memory, waits for the others' flags, sums the inbox slots in rank order
 * (bit-identical on every rank) and applies Adam -- ONE launch.  One process per GPU, all on one node.
 *
 *   vh_peer_buffer_bytes    size of a rank's exchange buffer for n parameters and `world` ranks
 *   vh_peer_buffer_create   cudaMalloc + zero it + CUDA-IPC handle (64 bytes) to send to the peers   [the only
 *   vh_peer_buffer_open     map a peer's buffer from its handle (peer access enabled lazily)           allocation the
 *   vh_peer_buffer_close / _destroy                                                                    library makes]
 *   vh_adam_allreduce_step  param / grad / exp_avg / exp_avg_sq / hyper / step / guard as for vh_adam_step_dev (the
 *                           gradient is cleared); state = int64[4] {epoch, ticket scratch, timed_out, skipped steps},
 *                           zero-initialised, owned by the caller and NEVER rewound (every rank must make the same
 *                           sequence of calls); peers = DEVICE array of `world` pointers, entry r = this process's
 *                           mapping of rank r's exchange buffer (entry `rank` = its own).  A NaN guard on ANY rank makes
 *                           EVERY rank skip the update of that call (state[3], step[2] count them).  timed_out != 0
 *                           (sticky): a peer's flag did not arrive within timeout_s seconds (<= 0: 10 s, measured with
 *                           %globaltimer); the thread blocks that saw the timeout skip their part of the update and
 *                           every later call returns without touching anything -- the host must check state[2] whenever
 *                           it synchronises and stop. */
#define VH_PEER_MAX_WORLD 16
size_t vh_peer_buffer_bytes(int dtype, size_t n, int world);
/* diagnostics: %globaltimer stamps (ns) of thread block 0 of the last exchange launch on the current device: start, after
 * griddepcontrol.wait, pushed, published, peers' flags seen, vote passed, update applied, (unused) */
int vh_peer_debug_times(unsigned long long* out8);
int vh_peer_buffer_create(size_t bytes, void** dev_ptr, void* handle64);
int vh_peer_buffer_open(const void* handle64, void** dev_ptr);
int vh_peer_buffer_close(void* dev_ptr);
int vh_peer_buffer_destroy(void* dev_ptr);
int vh_adam_allreduce_step(int dtype, size_t n, void* param, void* grad, void* exp_avg, void* exp_avg_sq,
                           const void* hyper, void* step, void* state, int rank, int world, const void* peers,
                           const void* guard, double timeout_s, void* stream);
/* Same, with the encoder's hidden-layer weight gradient formed on the way to the peers (the data-parallel counterpart of
 * vh_encoder_bwd_adam): entries [offset, offset + H*NLIN) of the flat gradient receive sum_b d_pre[b][o] * pooled[b][c]
 * (d_pre [B][H], pooled [B][NLIN]: what vh_encoder_bwd leaves / vh_encoder_fwd saved) before they are pushed.  The
 * caller runs the encoder backward WITHOUT that weight gradient (vh_encoder_grads.skip_lin_wgrad).  B <= 128. */
typedef struct vh_lin_wgrad {
  const void* d_pre;
  const void* pooled;
  int B, H, NLIN;
  long long offset; /* of the hidden-layer weight view inside the flat vector, in elements */
} vh_lin_wgrad;
int vh_adam_allreduce_step_wgrad(int dtype, size_t n, void* param, void* grad, void* exp_avg, void* exp_avg_sq,
                                 const void* hyper, void* step, void* state, int rank, int world, const void* peers,
                                 const void* guard, double timeout_s, const vh_lin_wgrad* wg, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused amortised encoder q(theta | x, d) (vihds/encoders.py:16-55 ConditionalEncoder, :126-253 Q_Local / Q_Global_Cond
 * / Q_Global / Q_Constant, :383-404 evaluate_q): delta-observations -> Conv1d -> AvgPool1d(stride 1) -> Linear -> tanh
 * -> packed (mu, log_prec) heads -> q_mu / q_prec [B][P] in the column order local, global-conditioned, global, constant.
 * Weight layouts are PyTorch's: conv_w [F][n_signals][K], lin_w [H][F*NP], local_w [2*n_local][H (+C) (+D)] with rows
 * (mu, log_prec) interleaved per parameter, gcond_w [2*n_gcond][(C) (+D)] (no bias), global_free [2*n_global].
 * vh_encoder_bwd ACCUMULATES into the g_* buffers (views of the flat gradient, zeroed by the optimizer). */
typedef struct vh_encoder_desc {
  int dtype;
  int B, T, n_signals, n_filters, filter_size, pool_size, n_hidden, C, D;
  int n_local, n_gcond, n_global, n_const;
  int local_cond_treatments, local_cond_devices, gcond_cond_treatments, gcond_cond_devices;
} vh_encoder_desc;

typedef struct vh_encoder_io {
  const void *observations /* [B][n_signals][T] */, *inputs /* [B][C] */, *dev_1hot /* [B][D] */;
  const void *conv_w, *conv_b, *lin_w, *lin_b, *local_w, *local_b, *gcond_w, *global_free, *const_values;
  void *q_mu, *q_prec; /* out [B][P] */
  void *pooled;        /* out, saved for the backward: [B][F*NP] */
  void *enc;           /* out, saved for the backward: [B][H] (tanh features) */
} vh_encoder_io;

typedef struct vh_encoder_grads {
  const void *d_q_mu, *d_q_prec; /* [B][P] */
  void *g_conv_w, *g_conv_b, *g_lin_w, *g_lin_b, *g_local_w, *g_local_b, *g_gcond_w, *g_global_free;
  void* d_pre; /* workspace [B][H]: cotangent of the hidden layer's pre-activations (kept: input of the weight gradient) */
  int skip_lin_wgrad; /* != 0: leave g_lin_w alone (vh_adam_allreduce_step_wgrad forms it inside the exchange launch) */
  void* dpool; /* optional workspace [B][F*NP]: with it, batches of >= 256 individuals run the hidden layer's backward as
                * GEMMs (as the forward call does on its own); NULL: one monolithic launch per individual group */
} vh_encoder_grads;

int vh_encoder_fwd(const vh_encoder_desc* e, const vh_encoder_io* io, void* stream);
int vh_encoder_bwd(const vh_encoder_desc* e, const vh_encoder_io* io, const vh_encoder_grads* g, void* stream);
/* vh_encoder_bwd followed by vh_adam_step_dev(zero_grad = 1) over the flat parameter vector the g_* buffers are views of,
 * with the hidden-layer weight gradient formed inside the optimiser launch (two launches instead of three: the end of
 * Training._run_batch, vihds/training.py:334-337, on one GPU).  param / grad / exp_avg / exp_avg_sq / hyper / step /
 * guard as for vh_adam_step_dev.  Small batches only (B <= 128); VH_ERR_UNSUPPORTED otherwise. */
int vh_encoder_bwd_adam(const vh_encoder_desc* e, const vh_encoder_io* io, const vh_encoder_grads* g, size_t n, void* param,
                        void* grad, void* exp_avg, void* exp_avg_sq, const void* hyper, void* step, const void* guard,
                        void* stream);

/* Device conditioner (vihds/ode.py:43-58 OdeModel.device_conditioner with param = ones, :99-116 DeviceConditioner):
 * out[k][n] = (plus_one[k] ? 1 : 0) + relu(dot(w[k], dev_1hot[m % B_global] * rel[k])),  m = (b_offset + b)*IW + i,
 * n = b*IW + i.  `m % B_global` reproduces the reference's repeat([n_iwae, 1]) + reshape, which hands sample (b, i) of
 * the batch the conditioner row (b*IW + i) % B.  dev_1hot is the one-hot table of the GLOBAL batch [B_global][D]; a rank
 * holding the slab of individuals [b_offset, b_offset + B) gets exactly the rows a single process would compute
 * (one GPU: B_global = B, b_offset = 0).  rel, w: [n_cond][D];  plus_one[k] = the parameter is in data.default_devices. */
int vh_device_conditioner(int dtype, int B, int IW, int D, int n_cond, int B_global, int b_offset, const void* dev_1hot,
                          const void* rel, const void* w, const int* plus_one, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIHDS_B200_H */
