"""Training: IWAE cost, the ELBO-gradient step and its CUDA-graph form.

Takes over the hot part of vihds/training.py (reference): ``Training.cost`` (:127-174), ``Training._run_batch``
(:324-340) and the Adam / MultiStepLR set-up (:82-86).  Orchestration that has no performance content (TensorBoard,
plotting, cross-validation merge) is out of scope.

Two equivalent ways to take a step:

* ``Training._run_batch(batch)``   eager, reference-shaped: ``model(batch, IW)`` -> ``cost`` -> ``backward`` -> Adam.
  Every arithmetic op of the hot path is a launch of libvihds_b200.so through engine.* autograd Functions.
* ``GraphedStep``                  the production form, and what ``Training.run`` uses on a CUDA device (one instance per
  batch shape, ``Training.graphed_step``): static device buffers and pre-built C-ABI descriptors, the
  whole step captured as two CUDA graphs -- encoder forward + device conditioner | fused forward, IWAE cost + gradient,
  fused reverse sweep, encoder backward, gradient all-reduce, Adam -- so a step costs the host two graph launches (the
  end-to-end entry slips the host-to-device copy of ``u`` under the first one).
"""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib as L
from .config import Settings
from .datasets import batch_of
from .engine import FlatAdam, IwaeCost, _ptr, _stream, iw_moments


def _nz(t):
    """Pointer of a possibly empty tensor (None when it has no elements)."""
    return None if t is None or t.numel() == 0 else _ptr(t)


class Results(object):
    """Evaluation output (vihds/utils.py:65-99): importance-weighted trace moments, reduced on the device by
    vh_iw_moments so that the [B, IW, ., T] traces never travel to the host."""

    def __init__(self):
        self.species_names = self.q_names = self.q_values = self.theta = self.elbo = None
        self.iw_predict_mu = self.iw_predict_std = self.iw_states = self.iw_variance = None

    def init(self, species_names, q, theta, elbo, normalized_iws, x_predict, x_states, precisions):
        terms = theta.terms
        prob = terms["problem"]
        B, IW = normalized_iws.shape
        prec_planes = None
        if not prob.dynamic_precisions:
            names = ["prec_x", "prec_rfp", "prec_yfp", "prec_cfp"]
            prec_planes = torch.stack([theta.samples[n].reshape(-1) for n in names]).contiguous()
        mu, sd, st, var = iw_moments(prob, normalized_iws.reshape(-1), terms["trace"], terms["predict"], prec_planes, B, IW)
        self.species_names = species_names
        self.q_names = q.get_tensor_names()
        self.q_values = np.array([x.detach().cpu().numpy() for x in q.get_tensors()], dtype=object)
        self.theta = np.array([x.detach().cpu().numpy() for x in theta.get_tensors()])
        self.elbo = elbo.detach().cpu().numpy()
        self.iw_predict_mu, self.iw_predict_std = mu.cpu().numpy(), sd.cpu().numpy()
        self.iw_states, self.iw_variance = st.cpu().numpy(), var.cpu().numpy()


def multistep_lr(base_lr, boundaries, gamma, epoch):
    """torch.optim.lr_scheduler.MultiStepLR (training.py:84-86) as a pure function of the epoch."""
    return base_lr * gamma ** sum(1 for b in boundaries if epoch >= b)


class Training(object):
    def __init__(self, args, settings, data, parameters, model):
        self.args, self.settings, self.dataset_pair, self.model = args, settings, data, model
        self.parameters = parameters
        self.optimizer = FlatAdam(model.parameters(recurse=True), lr=settings.params.learning_rate)
        self.model.n_theta = parameters.n_theta
        self.n_batch = min(settings.params.n_batch, data.n_train) if data is not None else settings.params.n_batch
        self.epoch = 0
        if data is not None:
            dev, dt = settings.device, settings.dtype
            self.train_data = batch_of(data.train.dataset, data.train.indices, dev, dt)
            self.valid_data = batch_of(data.test.dataset, data.test.indices, dev, dt)

    # -- cost (training.py:127-174) -----------------------------------------------------------------------------
    def cost(self, batch_data, batch_results, theta, q, p, full_output=False, writer=None, epoch=None, b_total=None):
        """IWAE cost from the per-sample terms the fused kernel attached to ``theta``; ``.elbo`` is the NEGATIVE
        bound, as in the reference (:174)."""
        x_states, x_predict, precisions = batch_results
        terms = getattr(theta, "terms", None)
        if terms is None:
            raise RuntimeError("cost() needs the theta returned by BaseVAE.forward / Decoder.fused (it carries the "
                               "per-sample ELBO terms computed by the CUDA kernel); there is no PyTorch fallback")
        lpx = terms["log_p_by_species"]
        B, IW = lpx.shape[0], lpx.shape[1]
        cost, log_w, w = IwaeCost.apply(lpx.reshape(B * IW, 4), p.log_prob(theta).reshape(-1), q.log_prob(theta).reshape(-1),
                                        B, IW, b_total or B)
        if not full_output:
            return Settings(elbo=cost[0])
        out = Results()
        out.init(self.model.decoder.state_names, q, theta, -cost[0], w.view(B, IW), x_predict, x_states, precisions)
        out.log_unnormalized_iws = log_w.view(B, IW)
        return out

    # -- eager step (training.py:324-340) -----------------------------------------------------------------------
    def _run_batch(self, batch, u=None):
        result, theta, q, p = self.model(batch, self.args.train_samples, u=u)
        elbo = self.cost(batch, result, theta, q, p).elbo
        if torch.isnan(elbo):
            print("\nELBO is NaN. Stopping training.")
            return False
        self.optimizer.zero_grad()
        elbo.backward()
        self.optimizer.step()
        self.last_cost = elbo.detach()
        return True

    def set_epoch(self, epoch):
        self.epoch = epoch
        p = self.settings.params
        self.optimizer.set_lr(multistep_lr(p.learning_rate, p.learning_boundaries, p.learning_gamma, epoch))

    def evaluate(self, data, samples):
        """training.py:267-300 without plotting: full-output cost on a whole data set."""
        with torch.no_grad():
            result, theta, q, p = self.model(data, samples)
            return self.cost(data, result, theta, q, p, full_output=True)

    def graphed_step(self, B, IW, T):
        """The CUDA-graph form of the step for one batch shape, built on first use and reused (the ragged last
        mini-batch of an epoch gets its own capture)."""
        cache = self.__dict__.setdefault("_graphed", {})
        key = (int(B), int(IW), int(T))
        if key not in cache:
            want = self.model.want_predict
            self.model.want_predict = False  # nothing reads the x_predict trace in a training step
            cache[key] = GraphedStep(self, *key)
            self.model.want_predict = want
        return cache[key]

    def _run_batch_graphed(self, batch, u=None):
        """Same step as ``_run_batch`` through ``GraphedStep``: two graph launches, no host synchronisation.  The NaN
        check of training.py:331-333 happens ON THE DEVICE (the Adam kernel refuses the update and every later one, see
        vh_adam_step_dev); ``run`` looks at the refusal counter when it next synchronises."""
        IW = self.args.train_samples
        gs = self.graphed_step(len(batch.inputs), IW, batch.times.numel())
        gs.load_batch(batch)
        if u is None:
            u = self.model.sample_u(len(batch.inputs), IW)  # numpy's global RNG, as vae.py:22-24
        gs.load_u(u.to(device=gs.u.device, dtype=gs.u.dtype, non_blocking=True) if not u.is_cuda else u)
        gs.draw_conditioner()  # torch CPU RNG, draw for draw what condition_theta consumes in the reference
        self.last_cost = gs.step()
        return gs

    def run(self, epochs=None, loader=None, verbose=True, graphed=None):
        """Epoch loop (training.py:342-383) over shuffled mini-batches drawn with numpy's global RNG.  On a CUDA device
        the steps go through the CUDA-graph form (``graphed=False`` forces the eager, reference-shaped ``_run_batch``);
        ``self.path_taken`` records which one ran and ``self.costs`` the cost of every step of the last epoch."""
        epochs = epochs or self.args.epochs
        ds, ids = self.dataset_pair.train.dataset, np.asarray(self.dataset_pair.train.indices)
        if graphed is None:
            graphed = torch.device(self.settings.device).type == "cuda"
        self.path_taken = "graphed" if graphed else "eager"
        for epoch in range(1, epochs + 1):
            self.set_epoch(epoch - 1)
            order = torch.randperm(len(ids)).numpy()
            costs = []
            for s in range(0, len(ids), self.n_batch):
                batch = batch_of(ds, ids[order[s:s + self.n_batch]], self.settings.device, self.settings.dtype)
                if graphed:
                    self._run_batch_graphed(batch)
                    costs.append(self.last_cost.clone())  # device tensors: no synchronisation inside the epoch
                else:
                    if not self._run_batch(batch):
                        return False
                    costs.append(self.last_cost)
            self.costs = [float(c) for c in costs]  # one synchronisation per epoch
            if graphed:
                for gs in self._graphed.values():
                    gs.check_health()
                if self.optimizer.skipped_steps() > 0:  # the first NaN cost froze the parameters, like the reference's early exit
                    print("\nELBO is NaN. Stopping training.")
                    return False
            test_epoch = getattr(self.args, "test_epoch", 0)
            if test_epoch and epoch % test_epoch == 0:
                out = self.evaluate(self.valid_data, self.args.test_samples)
                if verbose:
                    print("epoch %4d | valid iwae-elbo = %0.4f" % (epoch, float(out.elbo)))
        return True


class GraphedStep(object):
    """One ELBO-gradient step over static buffers (see module docstring).

    b_total       denominator of the batch mean (global batch when individuals are sharded over ranks)
    b_offset      index of this rank's first individual in the global batch.  The device conditioner reproduces the
                  reference's repeat / reshape quirk on GLOBAL sample indices (vihds/ode.py:46-58), so a sharded run
                  needs the global one-hot table (``load_global_devices``) to give rank-count-independent results
    process_group torch.distributed group of the data-parallel ranks (None: one GPU)
    exchange      "peer" (default; env VIHDS_ALLREDUCE): gradient exchange + Adam in one kernel over NVLink peer memory;
                  "nccl": ncclAllReduce of the flat gradient, then the Adam kernel
    """

    def __init__(self, training, B, IW, T, b_total=None, process_group=None, use_graphs=True, exchange=None, b_offset=0):
        self.tr, self.model = training, training.model
        m = self.model
        dev, dt = training.settings.device, training.settings.dtype
        self.B, self.IW, self.T, self.N = B, IW, T, B * IW
        self.b_total = b_total or B
        self.b_offset = int(b_offset)
        self.pg = process_group
        self.use_graphs = use_graphs
        enc, ode = m.encoder, m.decoder.ode_model
        P = enc.parameters.n_theta
        self.P = P
        z = lambda *s: torch.zeros(*s, dtype=dt, device=dev)  # noqa: E731
        Cn, Dn = ode.n_treatments, ode.device_depth
        # the per-individual batch tensors are views of ONE device buffer: a host batch arrives in one copy
        spans, off = {}, 0
        for k, shape in (("inputs", (B, Cn)), ("dev_1hot", (B, Dn)), ("observations", (B, 4, T))):
            spans[k] = (off, off + int(np.prod(shape)), shape)
            off += int(np.prod(shape))
        self._batch_spans = spans
        self.extras = list(ode.conditioned) if m.decoder.condition_on_device else []
        # TWO sets of input buffers (per-individual batch, u, conditioner weights): the end-to-end entry copies the next
        # step's inputs into one set while the device still computes on the other (``step_from_host``); each set has its
        # own pair of captured graphs.  ``step()`` and the ``load_*`` calls act on the selected set (set 0 unless
        # ``step_from_host`` has been used).  The time grid belongs to the data set and is shared.
        times = z(T)
        self._slots = []
        for _ in range(2):
            bd = z(off)
            self._slots.append(Settings(batch_dev=bd, u=z(self.N, P), cond_w=z(max(1, len(self.extras)), Dn),
                                        batch=Settings(times=times, **{k: bd[a:b].view(shape) for k, (a, b, shape) in spans.items()})))
        self._select(0)
        self.rel = [torch.as_tensor(np.asarray(ode.relevance[n])).to(device=dev, dtype=dt) for n in self.extras
                    if n in ode.relevance]
        if self.rel:
            self.rel_mat = torch.stack(self.rel).contiguous()
            self.plus_one = torch.tensor([int(n in ode.default_devices) for n in self.extras], dtype=torch.int32, device=dev)
            self.extra_static = z(len(self.extras), self.N)
            # sharded: one-hot table of the global batch (rows of all ranks); one GPU: the batch's own table
            self.dev_1hot_global = z(self.b_total, Dn) if self.b_total != B else None
            self._global_devices_loaded = self.dev_1hot_global is None
        # fused encoder: static activations + direct gradient pointers (no autograd on the encoder side)
        self.fused_encoder = bool(getattr(enc, "fused", False))
        if self.fused_encoder:
            self.enc_desc = enc.descriptor(B, dt)
            self.q_mu, self.q_prec = z(B, P), z(B, P)
            self.enc_pooled = z(B, enc.conditional.lin.weight.shape[1])
            self.enc_feats, self.enc_dpre = z(B, enc.conditional.n_outputs), z(B, enc.conditional.n_outputs)
            # large batches: workspace for the hidden layer's backward as GEMMs (cotangent of the pooled features)
            self.enc_dpool = z(B, enc.conditional.lin.weight.shape[1]) if B >= 256 else None
        prior = m.prior_tables(dt)
        self.prob = ode.problem(enc.names, enc.kinds, prior, self.extras, dev, dt)
        S = self.prob.S
        N = self.N
        self.buf = Settings(theta=z(P, N), x_states=z(T, S, N), lpx=z(N, 4), lp=z(N), lq=z(N), log_w=z(N), w=z(N),
                            g_lpx=z(N, 4), g_lp=z(N), g_lq=z(N), d_q_cost=z(2 * B * P + 1))
        # d_q_mu | d_q_prec | cost adjacent: the reverse launch clears all three with one memset node
        self.buf.d_q = self.buf.d_q_cost[:2 * B * P].view(2, B, P)
        self.buf.cost = self.buf.d_q_cost[2 * B * P:]
        self.buf.d_q_mu, self.buf.d_q_prec = self.buf.d_q[0], self.buf.d_q[1]
        # IWAE reduction inside the reverse launch where the latency-form kernel runs (one launch less per step)
        self.fuse_iwae = (os.environ.get("VIHDS_FUSE_IWAE", "1") != "0" and N <= 148 * 4 * 32 and ode.kernel_model != "dr_blackbox"
                          and not (self.prob.dynamic_precisions and self.prob.net.get("n_hidden", 0) > 0))
        self.one_graph = os.environ.get("VIHDS_ONE_GRAPH", "1") != "0"
        self.outputs_cleared = ode.kernel_model != "dr_blackbox"  # see _pre
        self.buf.cost_sum = z(1)  # NCCL path: sum of the ranks' costs (guard of the Adam update)
        self.d_weights = z(self.prob.n_weights) if self.prob.n_weights else None
        self.d_extra = z(len(self.extras), N) if (self.extras and hasattr(ode, "offset_layer")) else None
        self.exchange = None
        if self.pg is not None and (exchange or os.environ.get("VIHDS_ALLREDUCE", "peer")) != "nccl":
            from .distributed import ExchangeUnavailable, PeerGradientExchange
            opt = training.optimizer
            try:
                self.exchange = PeerGradientExchange(opt.flat.numel(), opt.flat.dtype, opt.flat.device, self.pg)
            except ExchangeUnavailable as e:  # same verdict on every rank: ncclAllReduce + Adam instead
                import warnings
                warnings.warn("peer gradient exchange unavailable (%s): using ncclAllReduce" % e)
        self.ready = False
        self.steps_done = 0
        self.ev_hot = None  # optional (start, end) CUDA events around the reverse-sweep kernel (forces the eager path)

    def _select(self, slot):
        """Make input set ``slot`` the one the ``load_*`` calls fill and ``step()`` computes on."""
        sl = self._slots[slot]
        self._slot = slot
        self._batch_dev, self.batch, self.u, self.cond_w = sl.batch_dev, sl.batch, sl.u, sl.cond_w
        self._apply(slot, "split")

    def _apply(self, slot, kind):
        """Point the introspection attributes (q_mu, weights, argument structs ...) at what the captured launches of
        (input set, graph kind) read and write; kind: "split" = g_pre + g_rest, "all" = the one-graph form."""
        for k, v in getattr(self, "_slot_attrs", {}).get((slot, kind), {}).items():
            setattr(self, k, v)

    _PER_SLOT = ("q_mu", "q_prec", "extra", "extra_grad", "weights", "_enc_io", "_p", "_fio", "_bio", "_iwae_args", "_p_ref",
                 "_fio_ref", "_bio_ref")

    @property
    def g_pre(self):
        return self._graphs[self._slot][0]

    @property
    def g_rest(self):
        return self._graphs[self._slot][1]

    # -- the three segments -------------------------------------------------------------------------------------
    def _pre(self):
        """encoder forward + device conditioning + weight packing.  Fused path: two launches of libvihds_b200.so on
        static buffers; otherwise stock PyTorch (captured either way)."""
        enc, ode = self.model.encoder, self.model.decoder.ode_model
        lib, s = self.prob.lib, _stream()
        # the device conditioner does not depend on the encoder: it runs on a forked branch (a parallel node of the
        # captured graph) and joins before the ODE kernel
        fork = bool(self.extras) and getattr(self, "extras_override", None) is None and not hasattr(ode, "offset_layer")
        cur = torch.cuda.current_stream()
        if not hasattr(self, "_side"):
            self._side = torch.cuda.Stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            # outputs the reverse launch accumulates into, cleared HERE (a parallel branch of the captured graph) instead of
            # by a memset node between the forward and the reverse kernel: vh_bwd_io.outputs_cleared
            if self.outputs_cleared:
                for t in (self.buf.d_q_cost, self.d_weights):
                    if t is not None:
                        L.check(lib.vh_zero_async(t.data_ptr(), t.numel() * t.element_size(), _stream()))
        if fork:
            with torch.cuda.stream(self._side):
                if not self._global_devices_loaded:
                    raise RuntimeError("sharded step with a device conditioner: call load_global_devices(dev_1hot of the "
                                       "GLOBAL batch) first -- the conditioner row of sample (b, i) is individual "
                                       "(b*IW + i) % B_global (vihds/ode.py:46-58)")
                table = self.dev_1hot_global if self.dev_1hot_global is not None else self.batch.dev_1hot
                L.check(lib.vh_device_conditioner(self.prob.vh_dtype, self.B, self.IW, self.cond_w.shape[1], len(self.extras),
                                                  table.shape[0], self.b_offset if self.dev_1hot_global is not None else 0,
                                                  _ptr(table), _ptr(self.rel_mat), _ptr(self.cond_w),
                                                  _ptr(self.plus_one), _ptr(self.extra_static), _stream()))
        if self.fused_encoder:
            pr = enc.fused_parameters()
            self._enc_io = L.vh_encoder_io(
                observations=_ptr(self.batch.observations), inputs=_ptr(self.batch.inputs), dev_1hot=_ptr(self.batch.dev_1hot),
                conv_w=_ptr(pr[0]), conv_b=_ptr(pr[1]), lin_w=_ptr(pr[2]), lin_b=_ptr(pr[3]), local_w=_nz(pr[4]),
                local_b=_nz(pr[5]), gcond_w=_nz(pr[6]), global_free=_nz(pr[7]), const_values=_nz(enc.const_values),
                q_mu=_ptr(self.q_mu), q_prec=_ptr(self.q_prec), pooled=_ptr(self.enc_pooled), enc=_ptr(self.enc_feats))
            L.check(lib.vh_encoder_fwd(C.byref(self.enc_desc), C.byref(self._enc_io), s))
        else:
            self.q_mu, self.q_prec = enc.q_table_reference(self.batch)
            self.q_mu, self.q_prec = self.q_mu.contiguous(), self.q_prec.contiguous()
        self.extra_grad = False
        if self.extras and getattr(self, "extras_override", None) is not None:
            self.extra = self.extras_override  # tests: pin the (random, per-call) conditioner to recorded values
        elif self.extras and hasattr(ode, "offset_layer"):
            self.extra = ode.conditioned_extras(self.B, self.IW, self.batch.dev_1hot)  # trainable: gradient flows back
            self.extra_grad = True
        elif self.extras:
            self.extra = self.extra_static
        else:
            self.extra = None
        cur.wait_stream(self._side)
        w = ode.flat_weights()
        self.weights = w.contiguous() if w is not None else None

    def _build_descriptors(self):
        pr, b, bt = self.prob, self.buf, self.batch
        self._p = pr.problem(self.B, self.IW, self.T)
        self._fio = L.vh_fwd_io(
            times=_ptr(bt.times), u=_ptr(self.u), q_mu=_ptr(self.q_mu), q_prec=_ptr(self.q_prec), p_mu=_ptr(pr.p_mu),
            p_prec=_ptr(pr.p_prec), clip_lo=_ptr(pr.clip_lo), clip_hi=_ptr(pr.clip_hi), kind=_ptr(pr.kind),
            extra=_ptr(self.extra), treatments=_ptr(bt.inputs), dev_1hot=_ptr(bt.dev_1hot),
            observations=_ptr(bt.observations), weights=_ptr(self.weights), theta=_ptr(b.theta),
            x_states=_ptr(b.x_states), x_predict=None, logp_by_species=_ptr(b.lpx), logp_theta=_ptr(b.lp),
            logq_theta=_ptr(b.lq))
        if self.fuse_iwae:
            self._bio = L.vh_bwd_io(fwd=self._fio, d_q_mu=_ptr(b.d_q_mu), d_q_prec=_ptr(b.d_q_prec), d_extra=_ptr(self.d_extra),
                                    d_weights=_ptr(self.d_weights), iwae_cost=_ptr(b.cost), iwae_b_total=self.b_total,
                                    outputs_cleared=int(self.outputs_cleared))
        else:
            self._bio = L.vh_bwd_io(fwd=self._fio, g_logp_by_species=_ptr(b.g_lpx), g_logp_theta=_ptr(b.g_lp),
                                    g_logq_theta=_ptr(b.g_lq), d_q_mu=_ptr(b.d_q_mu), d_q_prec=_ptr(b.d_q_prec),
                                    d_extra=_ptr(self.d_extra), d_weights=_ptr(self.d_weights),
                                    outputs_cleared=int(self.outputs_cleared))
        vdt = self.prob.vh_dtype
        self._iwae_args = (vdt, self.B, self.IW, self.b_total, _ptr(b.lpx), _ptr(b.lp), _ptr(b.lq), _ptr(b.cost),
                           _ptr(b.log_w), _ptr(b.w), _ptr(b.g_lpx), _ptr(b.g_lp), _ptr(b.g_lq))
        self._p_ref, self._fio_ref, self._bio_ref = C.byref(self._p), C.byref(self._fio), C.byref(self._bio)

    def _hot(self):
        """The hot path: three launches of libvihds_b200.so on the current stream (arguments pre-built)."""
        lib, s = self.prob.lib, _stream()
        L.check(lib.vh_elbo_terms_fwd(self._p_ref, self._fio_ref, s))
        if not self.fuse_iwae:
            L.check(lib.vh_iwae_fwd_bwd(*self._iwae_args, s))
        if self.ev_hot is not None:
            self.ev_hot[0].record()
        L.check(lib.vh_elbo_terms_bwd(self._p_ref, self._bio_ref, s))
        if self.ev_hot is not None:
            self.ev_hot[1].record()

    def _post(self):
        """encoder backward, ONE gradient all-reduce, fused Adam (captured)."""
        opt = self.tr.optimizer  # its gradient vector is clean here: prepare() cleared it once, every step clears it again
        outs, grads = [], []
        if not self.fused_encoder:
            outs, grads = [self.q_mu, self.q_prec], [self.buf.d_q_mu, self.buf.d_q_prec]
        if self.weights is not None and self.weights.requires_grad:
            outs.append(self.weights)
            grads.append(self.d_weights)
        if self.extra_grad:
            outs.append(self.extra)
            grads.append(self.d_extra)
        if outs:
            torch.autograd.backward(outs, grads)
        gr = None
        if self.fused_encoder:
            g = [p.grad for p in self.model.encoder.fused_parameters()]
            gr = L.vh_encoder_grads(d_q_mu=_ptr(self.buf.d_q_mu), d_q_prec=_ptr(self.buf.d_q_prec), g_conv_w=_ptr(g[0]),
                                    g_conv_b=_ptr(g[1]), g_lin_w=_ptr(g[2]), g_lin_b=_ptr(g[3]), g_local_w=_nz(g[4]),
                                    g_local_b=_nz(g[5]), g_gcond_w=_nz(g[6]), g_global_free=_nz(g[7]), d_pre=_ptr(self.enc_dpre),
                                    dpool=_ptr(self.enc_dpool))
        # the cost guards the update ON THE DEVICE: a NaN cost (NaN gradients) must not reach the parameters or the Adam
        # moments before the host has looked at it (vihds/training.py:331-336 checks before optimizer.step())
        if gr is not None and self.pg is None and self.B <= 128 and os.environ.get("VIHDS_FUSE_ADAM", "1") != "0":
            # one GPU, small batch: encoder backward, then the hidden-layer weight gradient + Adam in ONE launch
            L.check(self.prob.lib.vh_encoder_bwd_adam(
                C.byref(self.enc_desc), C.byref(self._enc_io), C.byref(gr), opt.flat.numel(), _ptr(opt.flat), _ptr(opt.grad),
                _ptr(opt.exp_avg), _ptr(opt.exp_avg_sq), _ptr(opt.hyper), _ptr(opt.step_dev), _ptr(self.buf.cost), _stream()))
            return
        fuse_wg = (gr is not None and self.exchange is not None and self.B <= 128 and os.environ.get("VIHDS_FUSE_ADAM", "1") != "0")
        if gr is not None:
            gr.skip_lin_wgrad = int(fuse_wg)  # formed inside the exchange launch instead
            L.check(self.prob.lib.vh_encoder_bwd(C.byref(self.enc_desc), C.byref(self._enc_io), C.byref(gr), _stream()))
        if self.exchange is not None:
            # gradient exchange over NVLink peer memory + Adam (+ the hidden-layer weight gradient): one launch
            opt.step_exchange(self.exchange, guard=self.buf.cost,
                              lin_wgrad=(self.enc_dpre, self.enc_pooled, self.model.encoder.fused_parameters()[2].grad) if fuse_wg else None)
        else:
            if self.pg is not None:
                torch.distributed.all_reduce(opt.grad, group=self.pg)
                self.buf.cost_sum.copy_(self.buf.cost)
                torch.distributed.all_reduce(self.buf.cost_sum, group=self.pg)  # a NaN on any rank reaches every rank
            opt.step(zero_grad=True, guard=self.buf.cost_sum if self.pg is not None else self.buf.cost)

    # -- capture / replay ---------------------------------------------------------------------------------------
    def prepare(self):
        """Warm up eagerly on a side stream, then capture the pre and post segments."""
        if self.ready:
            return
        self.tr.optimizer.zero_grad()
        if not self.use_graphs:
            self.ready = True
            return
        opt = self.tr.optimizer
        snap = [t.clone() for t in (opt.flat, opt.exp_avg, opt.exp_avg_sq, opt.step_dev)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._pre()
                self._build_descriptors()
                self._hot()
                self._post()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # two graphs: everything up to the ODE kernels' inputs (the end-to-end entry slips the copy of u underneath it),
        # and the rest -- ODE forward, IWAE, reverse sweep, encoder backward, all-reduce, Adam.  With the hot launches
        # issued eagerly between two graphs the host needed ~155 us per step, as long as the device.
        # One pair per input set (the kernel arguments hold the set's addresses).
        self._graphs, self._slot_attrs, pool, keep = [None] * len(self._slots), {}, None, self._slot
        for slot in sorted(range(len(self._slots)), key=lambda k: k == keep):  # the selected set last
            self._select(slot)
            if slot != keep:  # the capture passes compute on this set's buffers: give them the loaded set's contents
                for k in ("batch_dev", "u", "cond_w"):
                    self._slots[slot][k].copy_(self._slots[keep][k])
            g_pre, g_rest = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_pre, pool=pool):
                self._pre()
            pool = g_pre.pool()
            self._build_descriptors()
            g_pre.replay()
            with torch.cuda.graph(g_rest, pool=pool):
                self._hot()
                self._post()
            # what the captured launches of this set read and wrote (graph-pool tensors, argument structs)
            self._slot_attrs[(slot, "split")] = {k: getattr(self, k, None) for k in self._PER_SLOT}
            g_all = None
            if self.one_graph:
                # the whole step as ONE graph (``step()``, and ``step_from_host`` when the copy of u has already landed): no
                # graph-to-graph boundary between the encoder forward and the ODE kernels
                g_all = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_all, pool=pool):
                    self._pre()
                    self._build_descriptors()
                    self._hot()
                    self._post()
                self._slot_attrs[(slot, "all")] = {k: getattr(self, k, None) for k in self._PER_SLOT}
            self._graphs[slot] = (g_pre, g_rest, g_all)
            torch.cuda.synchronize()
        self._apply(keep, "split")
        for t, s in zip((opt.flat, opt.exp_avg, opt.exp_avg_sq, opt.step_dev), snap):
            t.copy_(s)  # warm-up and capture passes must not count as training steps
        self.ready = True
        self.check_health()  # rank skew during lazy module loading / capture is where an exchange time-out would show first

    def check_health(self):
        """Synchronises.  Raises if the gradient exchange has timed out (the replicas are then no longer in step)."""
        if self.exchange is not None:
            self.exchange.check()

    def skipped_steps(self):
        """Updates refused by the device-side NaN guard so far (synchronises)."""
        return self.tr.optimizer.skipped_steps()

    def load_global_devices(self, dev_1hot_global):
        """Sharded runs: the device one-hot rows of the GLOBAL batch [b_total, D] (same on every rank)."""
        if getattr(self, "dev_1hot_global", None) is None:
            return
        self.dev_1hot_global.copy_(dev_1hot_global.to(self.dev_1hot_global.dtype), non_blocking=True)
        self._global_devices_loaded = True

    def load_batch(self, batch, non_blocking=True):
        src = batch["times"]  # the time grid belongs to the data set: skip the copy while the same storage is handed in
        if self._times_differ(src):
            self._times_tag = self._tag_of(src)
            self.batch.times.copy_(src, non_blocking=non_blocking)
        keys = ("inputs", "dev_1hot", "observations")
        if any(batch[k].is_cuda for k in keys):
            for k in keys:
                self.batch[k].copy_(batch[k], non_blocking=non_blocking)
            return
        # host batch: packed into a pinned staging buffer (two of them, alternating: the previous copy may be in flight)
        if not hasattr(self, "_bhost"):
            self._bhost = [torch.empty(self._batch_dev.numel(), dtype=self._batch_dev.dtype).pin_memory() for _ in range(2)]
            self._bnp = [t.numpy() for t in self._bhost]
            self._bev, self._bslot = [torch.cuda.Event(), torch.cuda.Event()], 0
        sl = self._bslot
        self._bev[sl].synchronize()
        for k in keys:
            a, b, _ = self._batch_spans[k]
            self._bnp[sl][a:b] = batch[k].numpy().reshape(-1)
        self._h2d(self._batch_dev, self._bhost[sl])
        self._bev[sl].record()
        self._bslot = sl ^ 1

    @staticmethod
    def _tag_of(src):
        return (src.data_ptr(), src._version, src.numel(), tuple(src.stride()))

    def _times_differ(self, src):
        return getattr(self, "_times_tag", None) != self._tag_of(src)

    def _h2d(self, dst, src):
        """Asynchronous copy of a contiguous (pinned) host tensor into a static device buffer on the current stream: one
        C call (a torch ``copy_`` costs the host 6-8 us, three of them sit in front of every end-to-end step)."""
        L.check(self.prob.lib.vh_copy_async(dst.data_ptr(), src.data_ptr(), dst.numel() * dst.element_size(), _stream()))

    def load_u(self, u, non_blocking=True):
        # raw asynchronous copy for contiguous pinned host buffers of the right dtype / size; only the answer of
        # is_pinned() (a driver query) is cached per host allocation, everything else is re-checked on every call
        fast = False
        if not u.is_cuda and u.dtype == self.u.dtype and u.numel() == self.u.numel() and u.is_contiguous():
            pins = getattr(self, "_u_pinned", None)
            if pins is None:
                pins = self._u_pinned = {}
            key = (u.untyped_storage().data_ptr(), u.untyped_storage().nbytes())
            if key not in pins:
                if len(pins) > 64:
                    pins.clear()
                pins[key] = u.is_pinned()
            fast = pins[key]
        if fast:
            self._h2d(self.u, u)
        else:
            self.u.copy_(u.reshape(self.N, self.P), non_blocking=non_blocking)

    def _cond_fill(self, sl):
        """Per parameter two uniform fills and one normal fill of a [1, D] row: the draws
        models._draw_conditioner_weight makes (reference quirk, vihds/ode.py:48), in place on pinned staging buffer sl."""
        self._cond_ev[sl].synchronize()  # the copy that last read this buffer has finished
        for row in self._cond_rows[sl]:
            row.uniform_(-1.0, 1.0)
            row.uniform_(-1.0, 1.0)
            row.normal_(mean=2.0, std=1.5)

    def draw_conditioner(self):
        """Fresh conditioner weights per step from the torch CPU RNG stream, then one asynchronous copy.  If
        ``predraw_conditioner`` has already made this step's draw (same stream, same order) only the copy is left."""
        if not self.rel:
            return
        if not hasattr(self, "_cond_host"):
            self._cond_host = [torch.empty(len(self.extras), self.cond_w.shape[1]).pin_memory() for _ in range(2)]
            self._cond_rows = [[h[k:k + 1] for k in range(len(self.extras))] for h in self._cond_host]
            self._cond_ev, self._cond_slot, self._cond_drawn = [torch.cuda.Event(), torch.cuda.Event()], 0, False
        sl = self._cond_slot
        if not self._cond_drawn:
            self._cond_fill(sl)
        self._h2d(self.cond_w, self._cond_host[sl])
        self._cond_ev[sl].record()
        self._cond_drawn, self._cond_slot = False, sl ^ 1

    def predraw_conditioner(self):
        """Host only: make the NEXT step's draw now, while the device is busy with this one (~40 us of torch RNG calls
        that would otherwise sit in front of the next step's first launch).  The RNG stream is consumed in the same
        order; re-seeding torch between two steps takes effect one step later."""
        if self.rel and hasattr(self, "_cond_host") and not self._cond_drawn:
            self._cond_fill(self._cond_slot)
            self._cond_drawn = True

    def step_from_host(self, batch, u):
        """The public end-to-end step: ``batch`` (times, inputs, dev_1hot, observations) and ``u`` [B, IW, P] are HOST
        tensors (pinned for asynchronous copies).  Returns the cost (device tensor, no sync; valid until the next step).

        The inputs go into the input set the previous call did NOT use, on a copy stream: in a stream of calls the copies
        of step i + 1 (1 MB of u at the icml size, 18 MB at the synthetic one) run while the device computes step i.
        The small batch tensors and the conditioner weights go first and release the encoder graph; u -- which only the
        ODE kernel needs -- releases the second graph."""
        self.prepare()
        cur = torch.cuda.current_stream()
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream()
            n = len(self._slots)
            self._in_ready = [torch.cuda.Event() for _ in range(n)]
            self._u_ready = [torch.cuda.Event() for _ in range(n)]
            self._set_free = [torch.cuda.Event() for _ in range(n)]
            for e in self._set_free:
                e.record(cur)  # whatever the caller queued on the compute stream so far (load_* / step()) comes first
        slot = self._slot ^ 1
        self._select(slot)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._set_free[slot])  # the step that last computed on this set has finished
            if self._times_differ(batch["times"]):
                self._copy_stream.wait_stream(cur)  # the (shared) time grid is about to change: nothing may still read it
            self.load_batch(batch)
            self.draw_conditioner()
            self._in_ready[slot].record(self._copy_stream)
            self.load_u(u)
            self._u_ready[slot].record(self._copy_stream)
        cur.wait_event(self._in_ready[slot])
        if self.use_graphs and self._graphs[slot][2] is not None and not self._set_free[slot ^ 1].query():
            # a stream of steps (the previous one is still running): this step's copies run under it -- the one-graph form
            cur.wait_event(self._u_ready[slot])
            self._apply(slot, "all")
            self._graphs[slot][2].replay()
        elif self.use_graphs:  # the device is idle: start the encoder while u is still on its way
            self._apply(slot, "split")
            self.g_pre.replay()
            cur.wait_event(self._u_ready[slot])
            self.g_rest.replay()
        else:
            self._pre()
            self._build_descriptors()
            cur.wait_event(self._u_ready[slot])
            self._hot()
            self._post()
        self._set_free[slot].record(cur)
        self.predraw_conditioner()
        self.steps_done += 1
        return self.buf.cost

    def step(self):
        """One step on whatever the static buffers hold.  Returns the cost (device tensor, no sync)."""
        self.prepare()
        if self.use_graphs and self.ev_hot is None:
            g_all = self._graphs[self._slot][2]
            if g_all is not None:
                self._apply(self._slot, "all")
                g_all.replay()
            else:
                self.g_pre.replay()
                self.g_rest.replay()
        else:  # no graphs, or instrumented (events around the reverse-sweep launch): the same calls, issued eagerly
            self._pre()
            self._build_descriptors()
            self._hot()
            self._post()
        if hasattr(self, "_set_free"):  # step_from_host is in use as well: its copies must not overtake this step
            self._set_free[self._slot].record()
        self.steps_done += 1
        return self.buf.cost
