"""torch.autograd seam over the C ABI (include/vihds_b200.h): the only place where PyTorch meets the CUDA library.

``ElboProblem`` describes one (model, solver, dtype, parameter table) combination and owns the small constant device
tables (prior, clip bounds, kinds, slot map).  Three autograd Functions launch the kernels on torch's CURRENT stream
(so they can be captured into CUDA graphs and ordered with the encoder's kernels):

* ``FusedElboTerms``  -- vh_elbo_terms_fwd / vh_elbo_terms_bwd: sample+clip theta, solve, observe, log-likelihood,
  log p / log q in one launch; discrete-adjoint reverse sweep in one launch.
* ``SimulateTrace``   -- vh_simulate / vh_simulate_bwd: the narrow seam ``OdeModel.simulate`` (vihds/ode.py:66-82).
* ``IwaeCost``        -- vh_iwae_fwd / vh_iwae_bwd: the IWAE bound of ``Training.cost`` (vihds/training.py:134-148).

There is no fallback: tensors must live on a CUDA device and the library must load, otherwise these raise.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vihds_b200: the ODE+ELBO engine runs on CUDA tensors only (got a %s tensor); there is "
                               "no CPU fallback" % t.device)


def _c(t, dtype):
    """Contiguous tensor of the run dtype (no copy when it already is)."""
    if t is None:
        return None
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


class ElboProblem(object):
    """Static description of the hot path for one spec: model / solver ids, slot map, prior tables on the device.

    names      theta column names in ``u`` order (local, global-conditioned, global, constant)
    kinds      vh_kind per column;  p_mu, p_prec, clip_lo, clip_hi per column (numpy, run dtype)
    extras     names of per-trajectory inputs that are NOT sampled columns (conditioned aR/aS, ...) in ``extra`` row
               order; a name present in both wins as an extra (the conditioned value replaces the sample in the RHS,
               the sampled column keeps its log-prob terms -- what Decoder.forward + cost do in the reference)
    """

    def __init__(self, model, solver, dtype, names, kinds, p_mu, p_prec, clip_lo, clip_hi, extras=(), device="cuda",
                 n_hidden=0, n_hidden_states=0, n_latent=0, n_z=0, n_x=0, n_y=0, C_treat=2, D_dev=1,
                 init_latent_species=0.001, init_prec=0.00001, slot_alias=None):
        self.lib = L.load()
        self.model_name, self.solver_name = model, solver
        self.model, self.solver = L.model_id(model), L.solver_id(solver)
        self.dtype = dtype
        self.vh_dtype = L.VH_F64 if dtype == torch.float64 else L.VH_F32
        self.device = torch.device(device)
        self.names, self.extras = list(names), list(extras)
        self.P, self.E = len(self.names), len(self.extras)
        self.C, self.D = int(C_treat), int(D_dev)
        self.net = dict(n_hidden=n_hidden, n_hidden_states=n_hidden_states, n_latent=n_latent, n_z=n_z, n_x=n_x, n_y=n_y)
        self.init_values = (float(init_latent_species), float(init_prec))
        self.slot_alias = dict(slot_alias or {})  # library slot name -> theta name (dr_blackbox: latent0 -> z1, ...)
        if self.P > L.VH_MAX_SLOTS:
            raise ValueError("more than %d sampled parameters" % L.VH_MAX_SLOTS)
        self.slot_names = [self.slot_alias.get(nm, nm) for nm in L.slot_names(self.model)]
        self.slot_src = self._slot_map(self.names, self.extras)
        dev = self.device
        td = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(device=dev, dtype=dtype)  # noqa: E731
        self.kind = torch.as_tensor(np.asarray(kinds, np.int32)).to(dev) if self.P else None
        self.p_mu, self.p_prec = (td(p_mu), td(p_prec)) if self.P else (None, None)
        self.clip_lo, self.clip_hi = (td(clip_lo), td(clip_hi)) if self.P else (None, None)
        probe = self.problem(1, 1, 2)
        self.S = self.lib.vh_state_width(C.byref(probe))
        self.n_weights = int(self.lib.vh_num_weights(C.byref(probe)))
        self.n_species = self.S - 4 if self.n_weights > 0 else self.S
        self.dynamic_precisions = self.n_weights > 0

    def _slot_map(self, names, extras):
        src = [L.VH_SLOT_UNUSED] * L.VH_MAX_SLOTS
        for s, nm in enumerate(self.slot_names):
            if nm in extras:
                src[s] = -1 - extras.index(nm)
            elif nm in names:
                src[s] = names.index(nm)
        return src

    def problem(self, B, IW, T, P=None, E=None, slot_src=None):
        p = L.vh_problem()
        p.model, p.solver, p.dtype = self.model, self.solver, self.vh_dtype
        p.B, p.IW, p.T = int(B), int(IW), int(T)
        p.P = self.P if P is None else P
        p.E = self.E if E is None else E
        p.C, p.D = self.C, self.D
        for k, v in self.net.items():
            setattr(p, k, int(v))
        p.init_latent_species, p.init_prec = self.init_values
        src = self.slot_src if slot_src is None else slot_src
        for s in range(L.VH_MAX_SLOTS):
            p.slot_src[s] = src[s]
        return p

    def simulate_variant(self, theta_names):
        """The same model with every theta handed in as an ``extra`` row (P == 0): the OdeModel.simulate seam."""
        src = [L.VH_SLOT_UNUSED] * L.VH_MAX_SLOTS
        for s, nm in enumerate(self.slot_names):
            if nm in theta_names:
                src[s] = -1 - theta_names.index(nm)
        return src


class FusedElboTerms(torch.autograd.Function):
    """(q_mu, q_prec, u, extra, weights | batch) -> (logp_by_species [N,4], logp_theta [N], logq_theta [N],
    theta [P,N], x_states [T,S,N], x_predict [T,4,N] or None)."""

    @staticmethod
    def forward(ctx, prob, q_mu, q_prec, u, extra, weights, times, treatments, dev_1hot, observations, IW,
                want_predict):
        _need_cuda(q_mu, q_prec, u, extra, weights, times, treatments, observations)
        dt = prob.dtype
        q_mu, q_prec, u = _c(q_mu, dt), _c(q_prec, dt), _c(u, dt)
        extra, weights = _c(extra, dt), _c(weights, dt)
        times, treatments, dev_1hot, observations = _c(times, dt), _c(treatments, dt), _c(dev_1hot, dt), _c(observations, dt)
        B, P = q_mu.shape
        T = times.numel()
        N = B * IW
        assert P == prob.P and u.numel() == N * P, "u must be [B, IW, P] with P = %d" % prob.P
        dev = q_mu.device
        new = lambda *shape: torch.empty(*shape, dtype=dt, device=dev)  # noqa: E731
        theta, x_states = new(P, N), new(T, prob.S, N)
        x_predict = new(T, 4, N) if want_predict else None
        lpx, lp, lq = new(N, 4), new(N), new(N)
        p = prob.problem(B, IW, T)
        io = L.vh_fwd_io(times=_ptr(times), u=_ptr(u), q_mu=_ptr(q_mu), q_prec=_ptr(q_prec), p_mu=_ptr(prob.p_mu),
                         p_prec=_ptr(prob.p_prec), clip_lo=_ptr(prob.clip_lo), clip_hi=_ptr(prob.clip_hi),
                         kind=_ptr(prob.kind), extra=_ptr(extra), treatments=_ptr(treatments), dev_1hot=_ptr(dev_1hot),
                         observations=_ptr(observations), weights=_ptr(weights), theta=_ptr(theta),
                         x_states=_ptr(x_states), x_predict=_ptr(x_predict), logp_by_species=_ptr(lpx),
                         logp_theta=_ptr(lp), logq_theta=_ptr(lq))
        L.check(prob.lib.vh_elbo_terms_fwd(C.byref(p), C.byref(io), _stream()))
        ctx.prob, ctx.dims = prob, (B, IW, T, P, N)
        ctx.save_for_backward(q_mu, q_prec, u, extra, weights, times, treatments, dev_1hot, observations, x_states, theta)
        ctx.set_materialize_grads(False)
        if x_predict is None:
            x_predict = torch.empty(0, dtype=dt, device=dev)
            ctx.mark_non_differentiable(x_predict)
        return lpx, lp, lq, theta, x_states, x_predict

    @staticmethod
    def backward(ctx, g_lpx, g_lp, g_lq, g_theta, g_xs, g_xp):
        prob = ctx.prob
        B, IW, T, P, N = ctx.dims
        q_mu, q_prec, u, extra, weights, times, treatments, dev_1hot, observations, x_states, theta = ctx.saved_tensors
        dt, dev = prob.dtype, q_mu.device
        g_lpx, g_lp, g_lq = _c(g_lpx, dt), _c(g_lp, dt), _c(g_lq, dt)
        g_theta, g_xs = _c(g_theta, dt), _c(g_xs, dt)
        g_xp = _c(g_xp, dt) if (g_xp is not None and g_xp.numel()) else None
        d_mu = torch.empty(B, P, dtype=dt, device=dev)
        d_prec = torch.empty(B, P, dtype=dt, device=dev)
        d_extra = torch.empty_like(extra) if (extra is not None and ctx.needs_input_grad[4]) else None
        d_w = torch.empty_like(weights) if weights is not None else None
        p = prob.problem(B, IW, T)
        fio = L.vh_fwd_io(times=_ptr(times), u=_ptr(u), q_mu=_ptr(q_mu), q_prec=_ptr(q_prec), p_mu=_ptr(prob.p_mu),
                          p_prec=_ptr(prob.p_prec), clip_lo=_ptr(prob.clip_lo), clip_hi=_ptr(prob.clip_hi),
                          kind=_ptr(prob.kind), extra=_ptr(extra), treatments=_ptr(treatments), dev_1hot=_ptr(dev_1hot),
                          observations=_ptr(observations), weights=_ptr(weights), x_states=_ptr(x_states),
                          theta=_ptr(theta))
        bio = L.vh_bwd_io(fwd=fio, g_logp_by_species=_ptr(g_lpx), g_logp_theta=_ptr(g_lp), g_logq_theta=_ptr(g_lq),
                          g_theta=_ptr(g_theta), g_x_states=_ptr(g_xs), g_x_predict=_ptr(g_xp), d_q_mu=_ptr(d_mu),
                          d_q_prec=_ptr(d_prec), d_extra=_ptr(d_extra), d_weights=_ptr(d_w))
        L.check(prob.lib.vh_elbo_terms_bwd(C.byref(p), C.byref(bio), _stream()))
        return None, d_mu, d_prec, None, d_extra, d_w, None, None, None, None, None, None


class SimulateTrace(torch.autograd.Function):
    """theta planes [E,N] (already clipped / conditioned) -> x_states [T,S,N]; gradients flow back to the planes and
    to the decoder weights.  Same kernels as the fused call with P == 0."""

    @staticmethod
    def forward(ctx, prob, slot_src, planes, weights, times, treatments, dev_1hot, B, IW):
        _need_cuda(planes, weights, times, treatments)
        dt = prob.dtype
        planes, weights = _c(planes, dt), _c(weights, dt)
        times, treatments, dev_1hot = _c(times, dt), _c(treatments, dt), _c(dev_1hot, dt)
        T, N, E = times.numel(), B * IW, planes.shape[0]
        x_states = torch.empty(T, prob.S, N, dtype=dt, device=planes.device)
        p = prob.problem(B, IW, T, P=0, E=E, slot_src=slot_src)
        io = L.vh_fwd_io(times=_ptr(times), extra=_ptr(planes), treatments=_ptr(treatments), dev_1hot=_ptr(dev_1hot),
                         weights=_ptr(weights), x_states=_ptr(x_states))
        L.check(prob.lib.vh_simulate(C.byref(p), C.byref(io), _stream()))
        ctx.prob, ctx.slot_src, ctx.dims = prob, slot_src, (B, IW, T, E)
        ctx.save_for_backward(planes, weights, times, treatments, dev_1hot, x_states)
        return x_states

    @staticmethod
    def backward(ctx, g_xs):
        prob = ctx.prob
        B, IW, T, E = ctx.dims
        planes, weights, times, treatments, dev_1hot, x_states = ctx.saved_tensors
        g_xs = _c(g_xs, prob.dtype)
        d_planes = torch.zeros_like(planes)  # rows no slot reads keep a zero gradient
        d_w = torch.empty_like(weights) if weights is not None else None
        p = prob.problem(B, IW, T, P=0, E=E, slot_src=ctx.slot_src)
        fio = L.vh_fwd_io(times=_ptr(times), extra=_ptr(planes), treatments=_ptr(treatments), dev_1hot=_ptr(dev_1hot),
                          weights=_ptr(weights), x_states=_ptr(x_states))
        bio = L.vh_bwd_io(fwd=fio, g_x_states=_ptr(g_xs), d_extra=_ptr(d_planes), d_weights=_ptr(d_w))
        L.check(prob.lib.vh_simulate_bwd(C.byref(p), C.byref(bio), _stream()))
        return None, None, d_planes, d_w, None, None, None, None, None


class IwaeCost(torch.autograd.Function):
    """(logp_by_species [N,4], logp_theta [N], logq_theta [N]) -> (cost [1], log_w [N], w [N]); one block per
    individual.  ``b_total`` is the denominator of the batch mean (global batch when individuals are sharded)."""

    @staticmethod
    def forward(ctx, lpx, lp, lq, B, IW, b_total):
        _need_cuda(lpx, lp, lq)
        dt = lpx.dtype
        lpx, lp, lq = lpx.contiguous(), lp.contiguous(), lq.contiguous()
        lib = L.load()
        vdt = L.VH_F64 if dt == torch.float64 else L.VH_F32
        cost = torch.empty(1, dtype=dt, device=lpx.device)
        log_w = torch.empty(B * IW, dtype=dt, device=lpx.device)
        w = torch.empty(B * IW, dtype=dt, device=lpx.device)
        L.check(lib.vh_iwae_fwd(vdt, B, IW, b_total, _ptr(lpx), _ptr(lp), _ptr(lq), _ptr(cost), _ptr(log_w), _ptr(w), _stream()))
        ctx.dims = (vdt, B, IW, b_total)
        ctx.save_for_backward(w)
        ctx.mark_non_differentiable(log_w, w)
        return cost, log_w, w

    @staticmethod
    def backward(ctx, g_cost, _g_log_w, _g_w):
        (w,) = ctx.saved_tensors
        vdt, B, IW, b_total = ctx.dims
        N = B * IW
        g_lpx = torch.empty(N, 4, dtype=w.dtype, device=w.device)
        g_lp, g_lq = torch.empty_like(w), torch.empty_like(w)
        g = g_cost.contiguous() if g_cost is not None else None
        L.check(L.load().vh_iwae_bwd(vdt, B, IW, b_total, _ptr(w), _ptr(g), _ptr(g_lpx), _ptr(g_lp), _ptr(g_lq), _stream()))
        return g_lpx, g_lp, g_lq, None, None, None


def iw_moments(prob, w, x_states, x_predict, prec_planes, B, IW):
    """Importance-weighted trace moments of the evaluation path (vihds/utils.py:79-99) without leaving the device.
    Returns iw_predict_mu, iw_predict_std [B,4,T], iw_states [B,n_species,T], iw_variance [B,4,T]."""
    _need_cuda(w, x_states, x_predict)
    T = x_states.shape[0]
    dt, dev = prob.dtype, x_states.device
    mu, sd = torch.empty(B, 4, T, dtype=dt, device=dev), torch.empty(B, 4, T, dtype=dt, device=dev)
    st, var = torch.empty(B, prob.n_species, T, dtype=dt, device=dev), torch.empty(B, 4, T, dtype=dt, device=dev)
    p = prob.problem(B, IW, T)
    L.check(prob.lib.vh_iw_moments(C.byref(p), _ptr(w.contiguous()), _ptr(x_states), _ptr(x_predict),
                                   _ptr(prec_planes), _ptr(mu), _ptr(sd), _ptr(st), _ptr(var), _stream()))
    return mu, sd, st, var


class FlatAdam(object):
    """Adam over ONE flat fp32/fp64 vector holding every trainable parameter (views), one kernel launch per step
    (vh_adam_step_dev).  The step counter and the learning rate live on the device so that the launch can be
    captured in a CUDA graph and replayed; torch.optim.Adam defaults (vihds/training.py:82)."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8):
        params = [p for p in params if p.requires_grad]
        assert params, "no trainable parameters"
        dev, dt = params[0].device, params[0].dtype
        _need_cuda(*params)
        pad4 = lambda k: (k + 3) & ~3  # noqa: E731  every view starts 16-byte aligned (vectorised / cp.async copies)
        n = sum(pad4(p.numel()) for p in params)
        self.flat = torch.zeros(n, dtype=dt, device=dev)
        self.grad = torch.zeros(n, dtype=dt, device=dev)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)
                p.grad = self.grad[off:off + k].view_as(p)
                off += pad4(k)
        self.params = params
        self.betas, self.eps = betas, eps
        self.hyper = torch.tensor([lr, betas[0], betas[1], eps], dtype=torch.float64, device=dev)
        self.step_dev = torch.zeros(4, dtype=torch.int64, device=dev)  # [updates done, ticket scratch, skipped (NaN cost), -]
        self.vdt = L.VH_F64 if dt == torch.float64 else L.VH_F32
        self.lr = lr

    def set_lr(self, lr):
        self.lr = lr
        self.hyper[0] = lr

    def zero_grad(self):
        self.grad.zero_()

    def step_exchange(self, exchange, guard=None, lin_wgrad=None):
        """Data-parallel step: sum the ranks' gradients over NVLink peer memory and apply Adam in ONE launch
        (distributed.PeerGradientExchange); the gradient vector is cleared.  guard: this rank's cost (device tensor):
        a NaN cost on ANY rank makes every rank skip the update (vihds/training.py:331-336)."""
        if lin_wgrad is not None:  # (d_pre [B,H], pooled [B,NLIN], gradient view of the hidden-layer weight)
            d_pre, pooled, gview = lin_wgrad
            wg = L.vh_lin_wgrad(d_pre=_ptr(d_pre), pooled=_ptr(pooled), B=d_pre.shape[0], H=d_pre.shape[1], NLIN=pooled.shape[1],
                                offset=(gview.data_ptr() - self.grad.data_ptr()) // self.grad.element_size())
            L.check(L.load().vh_adam_allreduce_step_wgrad(
                self.vdt, self.flat.numel(), _ptr(self.flat), _ptr(self.grad), _ptr(self.exp_avg), _ptr(self.exp_avg_sq),
                _ptr(self.hyper), _ptr(self.step_dev), _ptr(exchange.state), exchange.rank, exchange.world, _ptr(exchange.peers),
                _ptr(guard), float(exchange.timeout_s), C.byref(wg), _stream()))
            return
        L.check(L.load().vh_adam_allreduce_step(self.vdt, self.flat.numel(), _ptr(self.flat), _ptr(self.grad),
                                                _ptr(self.exp_avg), _ptr(self.exp_avg_sq), _ptr(self.hyper),
                                                _ptr(self.step_dev), _ptr(exchange.state), exchange.rank, exchange.world,
                                                _ptr(exchange.peers), _ptr(guard), float(exchange.timeout_s), _stream()))

    def step(self, zero_grad=False, guard=None):
        """zero_grad: clear the gradient vector in the same launch (it is consumed exactly once).  guard: the step's cost
        (device tensor) -- if it is NaN the update is skipped on the device, as the reference does on the host before
        optimizer.step() (vihds/training.py:331-336), and ``skipped_steps`` counts it."""
        L.check(L.load().vh_adam_step_dev(self.vdt, self.flat.numel(), _ptr(self.flat), _ptr(self.grad), _ptr(self.exp_avg),
                                          _ptr(self.exp_avg_sq), _ptr(self.hyper), _ptr(self.step_dev), int(zero_grad),
                                          _ptr(guard), _stream()))

    def skipped_steps(self):
        """Number of updates the device-side NaN guard has refused so far (synchronises).  Non-zero means training has
        stopped: the guard is sticky, like the reference's early exit at the first NaN cost."""
        return int(self.step_dev[2].item())

    def clear_skipped(self):
        self.step_dev[2] = 0

    def state_snapshot(self):
        """Copies of (parameters, first / second moments, step counters): what a checkpoint of the optimiser holds."""
        return [t.clone() for t in (self.flat, self.exp_avg, self.exp_avg_sq, self.step_dev)]

    def restore_state(self, snap):
        """In-place restore of a ``state_snapshot`` (addresses captured in CUDA graphs stay valid); clears the gradient."""
        for t, s in zip((self.flat, self.exp_avg, self.exp_avg_sq, self.step_dev), snap):
            t.copy_(s)
        self.grad.zero_()
