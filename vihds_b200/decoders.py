"""Decoder: theta -> (x_states, x_predict, precisions), reference signature (vihds/decoders.py:11-45).

``Decoder.forward(theta, data, writer, epoch)`` serves callers that already hold a clipped theta: it conditions theta
on the device (host-side, models.OdeModel.condition_theta), runs the solve through the narrow kernel seam
(``OdeModel.simulate`` -> vh_simulate) and expands precisions / observations as strided views.  The training path does
not come through here: ``BaseVAE.forward`` uses ``Decoder.fused`` (sample + clip + solve + observe + log-likelihood +
log p / log q in one launch).
"""
import torch
from torch import nn

from . import models
from .distributions import DotOperatorSamples
from .engine import FusedElboTerms


class Decoder(nn.Module):
    def __init__(self, config, condition_on_device):
        super().__init__()
        if config.model not in models.LOOKUP:
            raise NotImplementedError("model '%s' has no kernel in vihds_b200 (models.LOOKUP: %s)" % (
                config.model, ", ".join(sorted(models.LOOKUP))))
        self.ode_model = models.LOOKUP[config.model](config)
        self.state_names = self.ode_model.species
        self.condition_on_device = condition_on_device
        self.config = config

    def forward(self, theta, data, writer=None, epoch=None):
        m = self.ode_model
        theta_conditioned = m.condition_theta(theta, data.dev_1hot, writer, epoch) if self.condition_on_device else theta
        solution = m.simulate(self.config, data.times, theta_conditioned, data.inputs, data.dev_1hot,
                              condition_on_device=self.condition_on_device)
        x_states, precisions = m.expand_precisions(theta_conditioned, data.times, solution)
        x_predict = m.observe(x_states, theta_conditioned)
        return (x_states, x_predict, precisions), theta_conditioned

    def fused(self, q, p, prior_tables, u, data, want_predict=True):
        """One fused launch for q.sample -> p.clip(4 sigma) -> condition -> simulate -> observe -> per-sample ELBO
        terms.  Returns ((x_states, x_predict, precisions), theta) where theta carries ``terms`` for ``cost``."""
        m = self.ode_model
        B, P = q.mu.shape
        IW = u.shape[1]
        N = B * IW
        extras = list(m.conditioned) if self.condition_on_device else []
        prob = m.problem(q.names, q.kinds, prior_tables, extras, q.mu.device, q.mu.dtype)
        extra = m.conditioned_extras(B, IW, data.dev_1hot) if extras else None
        lpx, lp, lq, planes, xs, xp = FusedElboTerms.apply(
            prob, q.mu, q.prec, u.reshape(N, P), extra, m.flat_weights(), data.times, data.inputs, data.dev_1hot,
            data.observations, IW, want_predict)
        T = data.times.numel()
        theta = DotOperatorSamples.from_planes(q.names, planes, B, IW)
        if extras:
            m.attach_conditioned(theta, extras, extra, B, IW)
        theta.terms = {q: lq.view(B, IW), p: lp.view(B, IW), "log_p_by_species": lpx.view(B, IW, 4), "problem": prob,
                       "trace": xs, "predict": xp if want_predict else None}
        sol = xs.view(T, prob.S, B, IW).permute(2, 3, 1, 0)
        x_states, precisions = m.expand_precisions(theta, data.times, sol)
        x_predict = xp.view(T, 4, B, IW).permute(2, 3, 1, 0) if want_predict else None
        return (x_states, x_predict, precisions), theta
