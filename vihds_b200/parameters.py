"""Parameter layout of a spec: which theta exist, how they are grouped, their priors and their initial q.

Takes over vihds/parameters.py (reference): the ``params.{shared,global,global_conditioned,local,constant}`` blocks of
a spec (parameters.py:341-453) become ONE ordered table -- local, global-conditioned, global, constant, the order in
which the reference concatenates its q and p chains (encoders.py:72-84) and therefore the column order of ``u``
(distributions.py:119-142).  The table is what the fused kernel consumes (kind / prior mu / prior precision / clip
bounds per column); the reference's per-parameter ``DistributionDescription`` objects have no counterpart.

Only Normal / LogNormal / Constant have kernels.  TruncNormal and Kumaraswamy are unusable in the reference as well
(distributions.py:442-446, :498-507 raise NotImplementedError) and are rejected here when the spec is parsed.
"""
import math
from collections import OrderedDict

import numpy as np

from . import _lib as L

GROUPS = ("local", "global_conditioned", "global", "constant")
KIND_OF = {"Constant": L.KIND_CONSTANT, "Normal": L.KIND_NORMAL, "LogNormal": L.KIND_LOGNORMAL}


class ParamSpec(object):
    """One theta: group, distribution kind, prior (mu, sigma, prec) and the initial free parameters of q."""

    __slots__ = ("name", "group", "kind", "mu", "sigma", "prec", "value", "conditioning", "given_prec")

    def __init__(self, name, group, kind, mu=0.0, sigma=None, prec=None, value=0.0, conditioning=None):
        self.name, self.group, self.kind, self.conditioning = name, group, kind, conditioning
        self.value = float(value)
        self.mu = float(mu)
        self.given_prec = None if prec is None else float(prec)
        # distributions.py:283-297: a given sigma wins, prec = 1/sigma^2; otherwise sigma = 1/sqrt(prec)
        if sigma is not None:
            self.sigma, self.prec = float(sigma), 1.0 / (float(sigma) * float(sigma))
        elif prec is not None:
            self.prec, self.sigma = float(prec), 1.0 / math.sqrt(float(prec))
        else:
            self.prec, self.sigma = 1.0, 1.0

    @property
    def init_log_prec(self):
        """Initial free parameter of q's precision.  Reference quirk (parameters.py:41-53): the description's defaults
        always carry a ``prec`` key, so the ``sigma`` branch is unreachable -- q starts at log(prec) only when the
        spec states ``prec``; with ``sigma`` alone it starts at log(1) = 0."""
        return float(np.log(self.given_prec)) if self.given_prec is not None else 0.0

    def __repr__(self):
        if self.kind == L.KIND_CONSTANT:
            return "ParamSpec(%s, %s, Constant=%g)" % (self.name, self.group, self.value)
        return "ParamSpec(%s, %s, %s, mu=%g, sigma=%g)" % (
            self.name, self.group, "LogNormal" if self.kind == L.KIND_LOGNORMAL else "Normal", self.mu, self.sigma)


def _resolve(block, shared):
    """A distribution entry may name a shared distribution (parameters.py:382-385, :415-419, :443-444)."""
    dist = block["distribution"]
    if dist in shared:
        block = shared[dist]
        dist = block["distribution"]
    if dist not in KIND_OF:
        raise NotImplementedError("distribution '%s' has no kernel (the reference raises for it too)" % dist)
    return dist, block


class Parameters(object):
    """``Parameters(settings.params)``: ordered ParamSpec table + the reference's counting interface."""

    def __init__(self, params_dict):
        self.params_dict = params_dict
        shared = params_dict.get("shared", {}) or {}
        self.specs = []
        for group in GROUPS:
            block = params_dict.get(group)
            if not block:
                continue
            cond = block.get("conditioning") if group in ("local", "global_conditioned") else None
            if group == "global_conditioned" and cond is None:
                raise Exception("global_cond MUST have conditioning")
            for name, entry in block.items():
                if name == "conditioning":
                    continue
                if group == "constant":
                    self.specs.append(ParamSpec(name, group, L.KIND_CONSTANT, value=entry))
                    continue
                dist, e = _resolve(entry, shared)
                if dist == "Constant":
                    self.specs.append(ParamSpec(name, group, L.KIND_CONSTANT, value=e.get("value", 0.0)))
                else:
                    self.specs.append(ParamSpec(name, group, KIND_OF[dist], mu=e.get("mu", 0.0), sigma=e.get("sigma"),
                                                prec=e.get("prec"), conditioning=cond))
        names = [s.name for s in self.specs]
        assert len(set(names)) == len(names), "duplicate parameter names in spec"
        self.by_name = OrderedDict((s.name, s) for s in self.specs)

    # -- reference-compatible queries (parameters.py:250-256) ---------------------------------------------------
    def group(self, group):
        return [s for s in self.specs if s.group == group]

    def get_parameter_counts(self):
        return tuple(len(self.group(g)) for g in GROUPS)

    @property
    def n_theta(self):
        return len(self.specs)

    @property
    def names(self):
        return [s.name for s in self.specs]

    def is_local(self, name):
        return name in self.by_name and self.by_name[name].group == "local"

    def is_global_cond(self, name):
        return name in self.by_name and self.by_name[name].group == "global_conditioned"

    def is_global(self, name):
        return name in self.by_name and self.by_name[name].group == "global"

    def is_constant(self, name):
        return name in self.by_name and self.by_name[name].group == "constant"

    # -- kernel-side tables -------------------------------------------------------------------------------------
    def kinds(self):
        return np.array([s.kind for s in self.specs], np.int32)

    def prior_arrays(self, np_dtype=np.float32, stddevs=4.0):
        """(p_mu, p_prec, clip_lo, clip_hi) as the reference forms them in the run dtype: torch.tensor([value]) of
        the spec number (encoders.py:283-295), prec = 1/(sigma*sigma) (distributions.py:294), bounds
        mu +- stddevs*sigma, exp'd for LogNormal (distributions.py:332-336, :377-381)."""
        t = np_dtype
        P = len(self.specs)
        mu, prec = np.zeros(P, t), np.ones(P, t)
        lo, hi = np.full(P, -np.inf, t), np.full(P, np.inf, t)
        for k, s in enumerate(self.specs):
            if s.kind == L.KIND_CONSTANT:
                mu[k] = t(s.value)
                continue
            m = t(s.mu)
            if "sigma" in self._given(s):
                sig = t(self._given(s)["sigma"])
                pr = t(1.0) / (sig * sig)
            else:
                pr = t(s.prec)
                sig = t(1.0) / np.sqrt(pr)
            mu[k], prec[k] = m, pr
            a, b = m - t(stddevs) * sig, m + t(stddevs) * sig
            if s.kind == L.KIND_LOGNORMAL:
                a, b = np.exp(a), np.exp(b)
            lo[k], hi[k] = a, b
        return mu, prec, lo, hi

    def _given(self, s):
        block = self.params_dict[s.group][s.name]
        shared = self.params_dict.get("shared", {}) or {}
        if block["distribution"] in shared:
            block = shared[block["distribution"]]
        return {k: v for k, v in block.items() if v is not None}

    def pretty_print(self):
        for g in GROUPS:
            rows = self.group(g)
            if rows:
                print("-----------------\n%s parameters\n-----------------" % g.upper())
                for s in rows:
                    print(s)
