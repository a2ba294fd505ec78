"""VAE assembly with the reference's surface (vihds/vae.py:13-51): ``BaseVAE(encoder, decoder, device)``,
``sample_u``, ``forward(data, samples) -> (result, conditioned_theta, q, p)``, ``build_model``.

``forward`` is where the hot path starts: the fused encoder kernel (vh_encoder_fwd) produces the dense q table, and
everything from ``q.sample`` to the per-sample ELBO terms is ONE launch of the fused CUDA kernel (Decoder.fused)."""
import numpy as np
import torch
from torch import nn

from .decoders import Decoder
from .encoders import Encoder


class BaseVAE(nn.Module):
    def __init__(self, encoder, decoder, device):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.device = torch.device(device)
        self.n_theta = encoder.parameters.n_theta if hasattr(encoder, "parameters") else None
        self.want_predict = True  # the training step switches the x_predict trace off (nothing reads it there)
        self._prior_tables = {}

    def sample_u(self, n_batch, n_samples, device=None):
        """vae.py:22-24: the RNG contract of the reference -- numpy's GLOBAL generator, float32 draws."""
        return torch.tensor(np.random.randn(n_batch, n_samples, self.n_theta).astype(np.float32))

    def prior_tables(self, dtype):
        if dtype not in self._prior_tables:
            npdt = np.float64 if dtype == torch.float64 else np.float32
            self._prior_tables[dtype] = self.encoder.parameters.prior_arrays(npdt, stddevs=4.0)
        return self._prior_tables[dtype]

    def forward(self, data, samples, writer=None, epoch=None, u=None):
        """``u`` may be passed (already on the device, [B, IW, P]) by callers that pre-stage it; otherwise it is
        drawn on the host exactly like the reference and copied."""
        q = self.encoder(data)
        if u is None:
            u = self.sample_u(len(data.inputs), samples).to(device=q.mu.device, dtype=q.mu.dtype, non_blocking=True)
        p = self.encoder.p
        result, theta = self.decoder.fused(q, p, self.prior_tables(q.mu.dtype), u, data, want_predict=self.want_predict)
        return result, theta, q, p


def build_model(args, settings, dataset, parameters):
    """vae.py:39-51: decoder conditions on the device only when there is more than one device group level."""
    encoder = Encoder(parameters, dataset, getattr(args, "verbose", False))
    condition = settings.data.device_depth > 1
    decoder = Decoder(settings, condition)
    model = BaseVAE(encoder, decoder, settings.device)
    return model.to(device=settings.device, dtype=settings.dtype)
