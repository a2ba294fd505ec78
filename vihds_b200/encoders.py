"""Amortised encoder q(theta | x, d) producing the dense q table the fused kernel consumes.

Takes over vihds/encoders.py (reference): the Conv1d -> AvgPool1d -> Linear -> tanh feature extractor
(encoders.py:16-55) is kept as stock PyTorch (cuDNN / cuBLAS; SURVEY.md section 2 row 9: out of scope for hand-written
kernels), but the 2 x n_param ``Linear(., 1)`` heads that the reference evaluates one by one (encoders.py:126-253,
:383-404) are packed into ONE weight matrix per group so that all local heads are a single GEMM, all
global-conditioned heads another, and the result is written straight into ``q_mu`` / ``q_prec`` of shape [B, P].

Initialisation draws from the torch RNG in exactly the reference's order (conv, hidden layer, then per local
parameter the ``mu`` head and the ``log_prec`` head, then the global-conditioned heads), so that with the same seed
the packed weights equal the reference's per-parameter layers -- tests/test_host_package.py pins this against the q
parameters recorded from the reference.
"""
import torch
from torch import nn

from . import _lib as L
from .distributions import ChainedDistribution


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
            c = group[0].conditioning if group else None
            c = c or default
            return bool(c.get("treatments", False)), bool(c.get("devices", False))

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
    def q_table(self, data):
        """(q_mu [B,P], q_prec [B,P]) -- one GEMM per conditioned group, no per-parameter Python loop."""
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
