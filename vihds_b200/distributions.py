"""q / p containers and the theta container, with the reference's access surface.

Takes over vihds/distributions.py (reference): ``ChainedDistribution`` (:58-189), ``TfNormal`` / ``TfLogNormal`` /
``TfConstant`` (:221-381) and ``DotOperatorSamples`` (:29-55).  The reference keeps one Python object per parameter
and loops over them; here a chain is ONE dense table -- ``mu`` and ``prec`` of shape [B, P] (q) or [P] (prior) plus a
``kinds`` vector -- because that is what the fused kernel consumes.  Per-parameter objects are thin views into the
table for code that does ``q.distributions["r"].mu``.

The hot path never calls ``sample`` / ``clip`` / ``log_prob`` below: ``BaseVAE.forward`` obtains theta, log q and
log p from the fused CUDA kernel (engine.FusedElboTerms) and attaches them to the theta container, and ``log_prob``
returns those.  The table-wide torch expressions here serve user-constructed theta only (same device as the table).
"""
import math
from collections import OrderedDict

import torch

from . import _lib as L

LOG2PI = math.log(2.0 * math.pi)


class DotOperatorSamples(object):
    """theta container: ``.samples`` (name -> [B, IW]), ``.keys``, ``.values`` and attribute access
    (distributions.py:29-55).  ``planes`` optionally holds the kernel's [P, N] buffer the entries are views of, and
    ``terms`` the fused per-sample terms (logp_by_species, logp_theta, logq_theta) computed for exactly these samples."""

    def __init__(self):
        self.samples = OrderedDict()
        self.keys = []
        self.values = []
        self.planes = None
        self.terms = None

    def add(self, name, sample):
        assert name not in self.samples, "DotOperatorSamples already has %s" % name
        self.samples[name] = sample
        self.keys.append(name)
        self.values.append(sample)
        setattr(self, name, sample)

    @classmethod
    def from_planes(cls, names, planes, B, IW):
        out = cls()
        out.planes = planes
        for k, nm in enumerate(names):
            out.add(nm, planes[k].view(B, IW))
        return out

    def get_n_batch(self):
        return self.values[0].shape[0]

    def get_n_samples(self):
        return self.values[0].shape[1]

    def get_tensors(self):
        return self.values

    def __str__(self):
        return "".join("%s = %s\n" % kv for kv in self.samples.items())


class _Column(object):
    """One parameter of a chain: a view into the table (``.mu``, ``.prec``, ``.sigma`` / ``.value``)."""

    def __init__(self, chain, k):
        self._chain, self._k = chain, k
        self.kind = int(chain.kinds[k])
        self.variable = chain.mu.dim() == 2 and bool(chain.per_individual[k])

    @property
    def mu(self):
        return self._chain.mu[..., self._k:self._k + 1]

    @property
    def value(self):
        return self.mu

    @property
    def prec(self):
        return self._chain.prec[..., self._k:self._k + 1]

    @property
    def sigma(self):
        return 1.0 / self.prec.sqrt()

    def get_tensors(self):
        return [self.mu] if self.kind == L.KIND_CONSTANT else [self.mu, self.prec]

    def sample(self, u, stop_grad=False):
        return self._chain.sample_columns(u.unsqueeze(-1), [self._k], stop_grad)[..., 0]

    def clip(self, x, stddevs=3):
        return self._chain.clip_columns(x.unsqueeze(-1), [self._k], stddevs)[..., 0]

    def log_prob(self, x, stop_grad=False):
        return self._chain.log_prob_columns(x.unsqueeze(-1), [self._k], stop_grad)[..., 0]


class ChainedDistribution(object):
    """Dense chain of P independent Constant / Normal / LogNormal factors.

    mu, prec : [B, P] (variational q; rows of global parameters are identical) or [P] (prior)
    kinds    : list of vh_kind;  per_individual[k] is True for local / global-conditioned columns
    """

    def __init__(self, name, names, kinds, mu, prec, per_individual=None):
        self.name = name
        self.names = list(names)
        self.kinds = list(int(k) for k in kinds)
        self.mu, self.prec = mu, prec
        self.per_individual = list(per_individual) if per_individual is not None else [False] * len(self.names)
        self.distributions = OrderedDict((nm, _Column(self, k)) for k, nm in enumerate(self.names))
        self.slot_dependencies = OrderedDict((nm, {}) for nm in self.names)

    def __getattr__(self, item):
        d = self.__dict__.get("distributions")
        if d is not None and item in d:
            return d[item]
        raise AttributeError(item)

    # -- table-wide expressions (any device) --------------------------------------------------------------------
    def _mask(self, kind, cols, like):
        return torch.tensor([self.kinds[k] == kind for k in cols], device=like.device)

    def _rows(self, t, cols, stop_grad=False):
        t = t[..., cols]
        if stop_grad:
            t = t.detach()
        return t.unsqueeze(-2) if t.dim() == 2 else t  # [B,1,k] against [B,IW,k]

    def sample_columns(self, u, cols, stop_grad=False):
        """distributions.py:327-330, :369-371, :242-243: mu + sigma*u, exp'd for LogNormal, the value for Constant."""
        mu, prec = self._rows(self.mu, cols, stop_grad), self._rows(self.prec, cols, stop_grad)
        s = mu + u / prec.sqrt()
        s = torch.where(self._mask(L.KIND_LOGNORMAL, cols, u), s.exp(), s)
        return torch.where(self._mask(L.KIND_CONSTANT, cols, u), mu + torch.zeros_like(u), s)

    def clip_columns(self, x, cols, stddevs):
        """distributions.py:332-336, :377-381: clamp to mu +- stddevs*sigma of THIS chain (detached bounds)."""
        mu, prec = self._rows(self.mu, cols, True), self._rows(self.prec, cols, True)
        sig = 1.0 / prec.sqrt()
        lo, hi = mu - stddevs * sig, mu + stddevs * sig
        ln = self._mask(L.KIND_LOGNORMAL, cols, x)
        lo, hi = torch.where(ln, lo.exp(), lo), torch.where(ln, hi.exp(), hi)
        clipped = torch.maximum(torch.minimum(x, hi), lo)
        return torch.where(self._mask(L.KIND_CONSTANT, cols, x), x, clipped)

    def log_prob_columns(self, x, cols, stop_grad=False):
        """distributions.py:338-345 (note -LOG2PI and the +1e-12 inside both logs), :373-375, :245-246."""
        mu, prec = self._rows(self.mu, cols, stop_grad), self._rows(self.prec, cols, stop_grad)
        ln = self._mask(L.KIND_LOGNORMAL, cols, x)
        lx = torch.where(ln, (x + 1e-12).log(), x)
        lp = -LOG2PI + 0.5 * (prec + 1e-12).log() - 0.5 * prec * (mu - lx).pow(2) - torch.where(ln, lx, torch.zeros_like(lx))
        return torch.where(self._mask(L.KIND_CONSTANT, cols, x), torch.zeros_like(lp), lp)

    # -- reference surface --------------------------------------------------------------------------------------
    def sample(self, list_of_u, device=None, stop_grad=False):
        assert list_of_u.shape[-1] == len(self.names), (
            "ChainedDistribution (%s #= %d):: must give a list of u's, one for each distribution." % (self.name, list_of_u.shape[-1]))
        th = self.sample_columns(list_of_u.to(self.mu.device), list(range(len(self.names))), stop_grad)
        out = DotOperatorSamples()
        for k, nm in enumerate(self.names):
            out.add(nm, th[..., k])
        return out

    def _cols_of(self, theta):
        return [(nm, self.names.index(nm)) for nm in theta.samples if nm in self.distributions]

    def clip(self, theta, stddevs=3, skip=None):
        out = DotOperatorSamples()
        for nm, value in theta.samples.items():
            if skip is not None and nm in skip:
                out.add(nm, value)
            else:
                out.add(nm, self.distributions[nm].clip(value, stddevs))
        return out

    def log_prob(self, theta, stop_grad=False):
        fused = getattr(theta, "terms", None)
        if fused is not None and self in fused and not stop_grad:
            return fused[self]
        pairs = self._cols_of(theta)
        if not pairs:
            return 0.0
        x = torch.stack([theta.samples[nm] for nm, _ in pairs], -1)
        return self.log_prob_columns(x, [k for _, k in pairs], stop_grad).sum(-1)

    def log_prob_mat(self, theta, stop_grad=False):
        pairs = self._cols_of(theta)
        x = torch.stack([theta.samples[nm] for nm, _ in pairs], -1)
        return self.log_prob_columns(x, [k for _, k in pairs], stop_grad)

    def get_tensors(self):
        return [t for d in self.distributions.values() for t in d.get_tensors()]

    def get_theta_names(self):
        return list(self.names)

    def get_tensor_names(self):
        out = []
        for nm, d in self.distributions.items():
            out += ["%s.value" % nm] if d.kind == L.KIND_CONSTANT else ["%s.mu" % nm, "%s.prec" % nm]
        return out

    def attach_summaries(self, writer, epoch, plot_histograms=False):
        for nm, d in self.distributions.items():
            if d.kind != L.KIND_CONSTANT:
                writer.add_scalar("%s/mu" % nm, d.mu.mean(), epoch)
                writer.add_scalar("%s/prec" % nm, d.prec.mean(), epoch)

    def __str__(self):
        return "".join("%s = kind %d mu %s prec %s\n" % (nm, d.kind, d.mu.flatten()[:3].tolist(), d.prec.flatten()[:3].tolist())
                       for nm, d in self.distributions.items())
