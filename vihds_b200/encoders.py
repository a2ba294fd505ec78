"""Amortised encoder q(theta | x, d) producing the dense q table the fused kernel consumes.

Takes over vihds/encoders.py (reference).  On the product path the whole encoder -- Conv1d -> AvgPool1d -> Linear ->
tanh feature extractor (encoders.py:16-55) and the 2 x n_param ``Linear(., 1)`` heads that the reference evaluates one
by one (encoders.py:126-253, :383-404) -- is ONE launch of libvihds_b200.so forward (vh_encoder_fwd) and two backward
(vh_encoder_bwd), bound here as ``FusedEncoder``; the heads are packed into one weight matrix per group and the result
is written straight into ``q_mu`` / ``q_prec`` of shape [B, P].  ``q_table_reference`` is the same computation in stock
PyTorch ops: the fp32 reference the kernels are tested against (and what the CPU-only tests use); it is never taken
silently -- ``q_table`` raises on CPU tensors.

Initialisation draws from the torch RNG in exactly the reference's order (conv, hidden layer, then per local
parameter the ``mu`` head and the ``log_prec`` head, then the global-conditioned heads), so that with the same seed
the packed weights equal the reference's per-parameter layers -- tests/test_host_package.py pins this against the q
parameters recorded from the reference.
"""
import ctypes as C

import torch
from torch import nn

from . import _lib as L
from .distributions import ChainedDistribution


def _ptr(t):
    return None if t is None or t.numel() == 0 else C.c_void_p(t.data_ptr())


class FusedEncoder(torch.autograd.Function):
    """vh_encoder_fwd / vh_encoder_bwd behind autograd: (batch tensors, the encoder's 8 parameter tensors) ->
    (q_mu, q_prec) [B, P].  One launch forward, two backward (see csrc/vh_encoder.cu)."""

    @staticmethod
    def forward(ctx, enc, obs, inputs, dev_1hot, conv_w, conv_b, lin_w, lin_b, local_w, local_b, gcond_w, global_free):
        lib = L.load()
        dt, dev = obs.dtype, obs.device
        B = obs.shape[0]
        desc = enc.descriptor(B, dt)
        P = desc.n_local + desc.n_gcond + desc.n_global + desc.n_const
        q_mu, q_prec = torch.empty(B, P, dtype=dt, device=dev), torch.empty(B, P, dtype=dt, device=dev)
        pooled = torch.empty(B, lin_w.shape[1], dtype=dt, device=dev)
        feats = torch.empty(B, lin_w.shape[0], dtype=dt, device=dev)
        obs, inputs, dev_1hot = obs.contiguous(), inputs.contiguous(), dev_1hot.contiguous()
        io = L.vh_encoder_io(observations=_ptr(obs), inputs=_ptr(inputs), dev_1hot=_ptr(dev_1hot), conv_w=_ptr(conv_w),
                             conv_b=_ptr(conv_b), lin_w=_ptr(lin_w), lin_b=_ptr(lin_b), local_w=_ptr(local_w),
                             local_b=_ptr(local_b), gcond_w=_ptr(gcond_w), global_free=_ptr(global_free),
                             const_values=_ptr(enc.const_values), q_mu=_ptr(q_mu), q_prec=_ptr(q_prec), pooled=_ptr(pooled),
                             enc=_ptr(feats))
        L.check(lib.vh_encoder_fwd(C.byref(desc), C.byref(io), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        ctx.enc_module, ctx.desc = enc, desc
        ctx.save_for_backward(obs, inputs, dev_1hot, conv_w, conv_b, lin_w, lin_b, local_w, local_b, gcond_w, global_free,
                              q_prec, pooled, feats)
        return q_mu, q_prec

    @staticmethod
    def backward(ctx, d_mu, d_prec):
        (obs, inputs, dev_1hot, conv_w, conv_b, lin_w, lin_b, local_w, local_b, gcond_w, global_free, q_prec, pooled,
         feats) = ctx.saved_tensors
        lib = L.load()
        z = torch.zeros_like
        g = [z(conv_w), z(conv_b), z(lin_w), z(lin_b), z(local_w), z(local_b), z(gcond_w), z(global_free)]
        d_mu = torch.zeros_like(q_prec) if d_mu is None else d_mu.contiguous()
        d_prec = torch.zeros_like(q_prec) if d_prec is None else d_prec.contiguous()
        d_pre = torch.empty_like(feats)
        dpool = torch.empty_like(pooled) if feats.shape[0] >= 256 else None  # large batches: backward through GEMMs
        io = L.vh_encoder_io(observations=_ptr(obs), inputs=_ptr(inputs), dev_1hot=_ptr(dev_1hot), conv_w=_ptr(conv_w),
                             conv_b=_ptr(conv_b), lin_w=_ptr(lin_w), lin_b=_ptr(lin_b), local_w=_ptr(local_w),
                             local_b=_ptr(local_b), gcond_w=_ptr(gcond_w), global_free=_ptr(global_free),
                             const_values=_ptr(ctx.enc_module.const_values), q_prec=_ptr(q_prec), pooled=_ptr(pooled),
                             enc=_ptr(feats))
        gr = L.vh_encoder_grads(d_q_mu=_ptr(d_mu), d_q_prec=_ptr(d_prec), g_conv_w=_ptr(g[0]), g_conv_b=_ptr(g[1]),
                                g_lin_w=_ptr(g[2]), g_lin_b=_ptr(g[3]), g_local_w=_ptr(g[4]), g_local_b=_ptr(g[5]),
                                g_gcond_w=_ptr(g[6]), g_global_free=_ptr(g[7]), d_pre=_ptr(d_pre), dpool=_ptr(dpool))
        L.check(lib.vh_encoder_bwd(C.byref(ctx.desc), C.byref(io), C.byref(gr), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return (None, None, None, None) + tuple(g)


class ConditionalEncoder(nn.Module):
    """delta-observations [B, 4, T-1] -> features [B, n_hidden] (encoders.py:16-55)."""

    def __init__(self, n_channels, n_obs, params):
        super().__init__()
        if params.transfer_func != "tanh":
            raise Exception("Unknown activation layer %s" % params.transfer_func)
        n_pool = n_obs - (params.filter_size - 1) - (params.pool_size - 1)
        self.n_outputs = params.n_hidden
        self.conv = nn.Conv1d(n_channels, params.n_filters, params.filter_size)
        nn.init.orthogonal_(self.conv.weight)
        self.pool = nn.AvgPool1d(params.pool_size, stride=1)
        self.lin = nn.Linear(n_pool * params.n_filters, self.n_outputs)
        nn.init.orthogonal_(self.lin.weight)

    def forward(self, x):
        h = self.pool(self.conv(x))
        return torch.tanh(self.lin(h.flatten(1)))


class _PackedHeads(nn.Module):
    """All (mu, log_prec) heads of one conditioned group as one affine map: rows 2k / 2k+1 = mu / log_prec of
    parameter k.  Each row is initialised by constructing the nn.Linear the reference would have constructed."""

    def __init__(self, n_params, n_inputs, bias):
        super().__init__()
        rows_w, rows_b = [], []
        for _ in range(2 * n_params):
            lin = nn.Linear(n_inputs, 1, bias)
            rows_w.append(lin.weight.detach())
            if bias:
                rows_b.append(lin.bias.detach())
        self.weight = nn.Parameter(torch.cat(rows_w, 0) if rows_w else torch.zeros(0, n_inputs))
        self.bias = nn.Parameter(torch.cat(rows_b, 0) if rows_b else torch.zeros(0)) if bias else None

    def forward(self, x):
        return torch.nn.functional.linear(x, self.weight, self.bias)


def _dims_of(dataset):
    """(n_species, n_times, n_conditions, depth) from a dataset pair (datasets.TimeSeriesDatasetPair) or a tuple."""
    if isinstance(dataset, (tuple, list)):
        return tuple(dataset)
    ds = dataset.train.dataset
    return ds.n_species, ds.n_times, dataset.n_conditions, dataset.depth


class Encoder(nn.Module):
    """``Encoder(parameters, dataset, verbose)``; ``forward(data) -> ChainedDistribution q``; ``.p`` is the prior."""

    def __init__(self, parameters, dataset, verbose=False):
        super().__init__()
        self.verbose = verbose
        self.parameters = parameters
        self.n_species, self.n_times, n_conditions, depth = _dims_of(dataset)
        params = parameters.params_dict
        self.conditional = ConditionalEncoder(self.n_species, self.n_times - 1, params)
        self.local = parameters.group("local")
        self.gcond = parameters.group("global_conditioned")
        self.glob = parameters.group("global")
        self.const = parameters.group("constant")

        def cond_of(group, default):
            """One conditioning per group: the packed heads share their inputs.  The reference builds every Q_Local /
            Q_Global_Cond from its own description.conditioning (encoders.py:125-175); a spec that mixes conditionings
            inside a group has no packed equivalent here and is refused rather than computed wrongly."""
            conds = [(bool((sp.conditioning or default).get("treatments", False)), bool((sp.conditioning or default).get("devices", False)))
                     for sp in group]
            if len(set(conds)) > 1:
                raise NotImplementedError("parameters of one group with different `conditioning` entries (%s) are not supported "
                                          "by the packed encoder heads" % ", ".join(sp.name for sp in group))
            return conds[0] if conds else (bool(default.get("treatments", False)), bool(default.get("devices", False)))

        self.local_cond = cond_of(self.local, {"treatments": False, "devices": False})
        self.gcond_cond = cond_of(self.gcond, {"treatments": False, "devices": False})
        n_in_local = self.conditional.n_outputs + n_conditions * self.local_cond[0] + depth * self.local_cond[1]
        n_in_gcond = n_conditions * self.gcond_cond[0] + depth * self.gcond_cond[1]
        self.local_heads = _PackedHeads(len(self.local), n_in_local, True)
        self.gcond_heads = _PackedHeads(len(self.gcond), n_in_gcond, False)
        # global q: free (mu, log_prec) pairs initialised from the prior (encoders.py:201-206; parameters.py:30-58)
        g0 = []
        for s in self.glob:
            g0 += [s.mu, s.init_log_prec]
        self.global_free = nn.Parameter(torch.tensor(g0, dtype=torch.get_default_dtype()))
        self.register_buffer("const_values", torch.tensor([s.value for s in self.const], dtype=torch.get_default_dtype()))
        specs = parameters.specs
        self.names = [s.name for s in specs]
        self.kinds = [s.kind for s in specs]
        self.per_individual = [s.group in ("local", "global_conditioned") for s in specs]
        self.n_free = len(self.local) + len(self.gcond) + len(self.glob)
        self.n_conditions, self.depth = n_conditions, depth
        self.fused = True
        self._p = None

    # prior ------------------------------------------------------------------------------------------------------
    @property
    def p(self):
        dev, dt = self.global_free.device, self.global_free.dtype
        if self._p is None or self._p.mu.device != dev or self._p.mu.dtype != dt:
            import numpy as np

            npdt = np.float64 if dt == torch.float64 else np.float32
            mu, prec, _, _ = self.parameters.prior_arrays(npdt)
            self._p = ChainedDistribution("p", self.names, self.kinds, torch.as_tensor(mu).to(dev), torch.as_tensor(prec).to(dev))
        return self._p

    # q ----------------------------------------------------------------------------------------------------------
    def descriptor(self, B, dtype):
        cv, pr = self.conditional, self.parameters.params_dict
        return L.vh_encoder_desc(
            dtype=L.VH_F64 if dtype == torch.float64 else L.VH_F32, B=B, T=self.n_times, n_signals=self.n_species,
            n_filters=cv.conv.out_channels, filter_size=cv.conv.kernel_size[0], pool_size=cv.pool.kernel_size[0],
            n_hidden=cv.n_outputs, C=self.n_conditions, D=self.depth, n_local=len(self.local), n_gcond=len(self.gcond),
            n_global=len(self.glob), n_const=len(self.const), local_cond_treatments=int(self.local_cond[0]),
            local_cond_devices=int(self.local_cond[1]), gcond_cond_treatments=int(self.gcond_cond[0]),
            gcond_cond_devices=int(self.gcond_cond[1]))

    def fused_parameters(self):
        """The 8 parameter tensors in the order FusedEncoder / vh_encoder_* take them."""
        cv = self.conditional
        return (cv.conv.weight, cv.conv.bias, cv.lin.weight, cv.lin.bias, self.local_heads.weight, self.local_heads.bias,
                self.gcond_heads.weight, self.global_free)

    # reference checkpoints -------------------------------------------------------------------------------------
    def reference_parameter_map(self):
        """[(reference parameter name (vihds/encoders.py module tree), tensor view of this module)]: the reference keeps
        one nn.Linear pair per local / global-conditioned parameter (``q_local_defs.<name>.layers.{mu,log_prec}``,
        encoders.py:126-175) and one free (mu, log_prec) pair per global one (``q_global_defs.<name>.free_params``,
        :201-206); here they are rows of the packed heads / entries of ``global_free``."""
        cv = self.conditional
        out = [("conditional.conv.weight", cv.conv.weight), ("conditional.conv.bias", cv.conv.bias),
               ("conditional.lin.weight", cv.lin.weight), ("conditional.lin.bias", cv.lin.bias)]
        for k, sp in enumerate(self.local):
            for j, part in enumerate(("mu", "log_prec")):
                out.append(("q_local_defs.%s.layers.%s.weight" % (sp.name, part), self.local_heads.weight[2 * k + j:2 * k + j + 1]))
                out.append(("q_local_defs.%s.layers.%s.bias" % (sp.name, part), self.local_heads.bias[2 * k + j:2 * k + j + 1]))
        for k, sp in enumerate(self.gcond):
            for j, part in enumerate(("mu", "log_prec")):
                out.append(("q_global_cond_defs.%s.layers.%s.weight" % (sp.name, part), self.gcond_heads.weight[2 * k + j:2 * k + j + 1]))
        for k, sp in enumerate(self.glob):
            for j, part in enumerate(("mu", "log_prec")):
                out.append(("q_global_defs.%s.free_params.%s" % (sp.name, part), self.global_free[2 * k + j:2 * k + j + 1]))
        return out

    def load_reference_state_dict(self, sd, prefix="encoder."):
        """Copy a state_dict of the reference's Encoder (names as in vihds/encoders.py) into the packed parameters."""
        with torch.no_grad():
            for name, view in self.reference_parameter_map():
                src = torch.as_tensor(sd[prefix + name]).to(device=view.device, dtype=view.dtype)
                view.copy_(src.reshape(view.shape))

    def reference_state_dict(self, prefix="encoder."):
        """The trainable parameters under the reference's names (a checkpoint the reference's Encoder can load)."""
        return {prefix + name: view.detach().clone() for name, view in self.reference_parameter_map()}

    def q_table(self, data):
        """(q_mu [B,P], q_prec [B,P]).  On a CUDA device: the fused encoder kernels (one launch forward, two backward);
        ``q_table_reference`` is the same computation in stock PyTorch ops (any device) -- the fp32 reference the fused
        kernels are tested against."""
        if not data.observations.is_cuda:
            raise RuntimeError("vihds_b200: Encoder.q_table runs the fused CUDA encoder kernels and needs CUDA tensors (got %s); "
                               "q_table_reference is the stock-PyTorch restatement the kernels are tested against" % data.observations.device)
        if self.fused:
            return FusedEncoder.apply(self, data.observations, data.inputs, data.dev_1hot, *self.fused_parameters())
        return self.q_table_reference(data)

    def q_table_reference(self, data):
        """One GEMM per conditioned group, no per-parameter Python loop (stock PyTorch)."""
        obs = data.observations
        B = obs.shape[0]
        delta = obs[:, :, 1:self.n_times] - obs[:, :, :self.n_times - 1]
        feats = [self.conditional(delta)]
        if self.local_cond[0]:
            feats.append(data.inputs)
        if self.local_cond[1]:
            feats.append(data.dev_1hot)
        free = [self.local_heads(torch.cat(feats, 1) if len(feats) > 1 else feats[0])] if len(self.local) else []
        if len(self.gcond):
            g = ([data.inputs] if self.gcond_cond[0] else []) + ([data.dev_1hot] if self.gcond_cond[1] else [])
            free.append(self.gcond_heads(torch.cat(g, 1) if len(g) > 1 else g[0]))
        if len(self.glob):
            free.append(self.global_free.unsqueeze(0).expand(B, -1))
        free = torch.cat(free, 1).view(B, self.n_free, 2)
        mu, prec = free[:, :, 0], free[:, :, 1].exp()
        if len(self.const):
            mu = torch.cat([mu, self.const_values.unsqueeze(0).expand(B, -1)], 1)
            prec = torch.cat([prec, torch.ones(B, len(self.const), dtype=prec.dtype, device=prec.device)], 1)
        return mu, prec

    def forward(self, data):
        mu, prec = self.q_table(data)
        return ChainedDistribution("q", self.names, self.kinds, mu, prec, self.per_individual)
