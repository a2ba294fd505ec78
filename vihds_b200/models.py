"""ODE model plugins: ``LOOKUP[name] -> OdeModel subclass`` (reference: models/__init__.py:19-35, vihds/ode.py:28-96).

In the reference a plugin carries the right-hand side as PyTorch code (models/dr_constant.py:14-112, ...).  Here a
plugin is a DESCRIPTOR: it names the CUDA kernel family that inlines its right-hand side (``kernel_model``), lists the
species, owns the decoder's trainable weights (NeuralPrecisions / black-box MLPs, packed into the flat layout the
kernel reads) and does the host-side conditioning of theta.  ``simulate`` / ``observe`` / ``initialize_state`` keep
the reference's signatures; all arithmetic happens in libvihds_b200.so.
"""
import torch
from torch import nn

from . import _lib as L
from .engine import ElboProblem, SimulateTrace


# ---------------------------------------------------------------------------------------------------------------
# precisions (vihds/precisions.py)
# ---------------------------------------------------------------------------------------------------------------
class ConstantPrecisions(nn.Module):
    """precisions.py:18-38: the four observation precisions are sampled global parameters, constant in time."""

    dynamic = False

    def __init__(self, precision_vars):
        super().__init__()
        self.precision_vars = list(precision_vars)

    def expand(self, theta, n_times, x_states):
        planes = torch.stack([getattr(theta, v) for v in self.precision_vars], dim=2)  # [B, IW, 4]
        return x_states, planes.unsqueeze(3).expand(-1, -1, -1, n_times)

    def flat_weights(self):
        return None

    def summaries(self, _writer, _epoch):
        pass


class NeuralPrecisions(nn.Module):
    """precisions.py:41-103: four extra ODE states driven by a small net on [t, species(, constants)].  The layers
    keep the reference's names (state_dict compatible) and initialisation order; ``flat_weights`` packs them in the
    kernel's layout: [hidden W, b,] production W, b, degradation W, b."""

    dynamic = True

    def __init__(self, n_inputs, n_hidden_precisions, n_outputs=4, inverse=False):
        super().__init__()
        if inverse:
            raise NotImplementedError("inverse NeuralPrecisions has no kernel (unused by every spec of the reference)")
        self.n_inputs, self.n_outputs, self.n_hidden = n_inputs, n_outputs, int(n_hidden_precisions)
        n_in = n_inputs + 1
        if self.n_hidden < 1:
            self.prec_production = nn.Linear(n_in, n_outputs)
            nn.init.xavier_uniform_(self.prec_production.weight)
            self.prec_degradation = nn.Linear(n_in, n_outputs)
            nn.init.xavier_uniform_(self.prec_degradation.weight)
        else:
            self.prec_hidden = nn.Linear(n_in, self.n_hidden)
            nn.init.xavier_uniform_(self.prec_hidden.weight)
            self.prec_production = nn.Linear(self.n_hidden, n_outputs)
            nn.init.xavier_uniform_(self.prec_production.weight, gain=0.5)
            self.prec_degradation = nn.Linear(self.n_hidden, n_outputs)
            nn.init.xavier_uniform_(self.prec_degradation.weight, gain=1)

    def layers(self):
        names = (["prec_hidden"] if self.n_hidden >= 1 else []) + ["prec_production", "prec_degradation"]
        return [getattr(self, n) for n in names]

    def flat_weights(self):
        return torch.cat([t.reshape(-1) for lin in self.layers() for t in (lin.weight, lin.bias)])

    def expand(self, theta, _n_times, x_states):
        return x_states[:, :, :-self.n_outputs, :], x_states[:, :, -self.n_outputs:, :]

    def summaries(self, writer, epoch):
        pass


# ---------------------------------------------------------------------------------------------------------------
# plugin base
# ---------------------------------------------------------------------------------------------------------------
def _draw_conditioner_weight(n_inputs):
    """The reference builds a FRESH ``DeviceConditioner`` on every call (vihds/ode.py:48, :99-116): a bias-free
    Linear(n_inputs, 1) whose weight ends up N(2, 1.5) after the default (kaiming-uniform) and the xavier-uniform
    initialisers have each consumed n_inputs uniform draws of the torch CPU RNG stream.  Reproduced draw for draw --
    two uniform_ fills and one normal_ -- without constructing the module (tests/test_host_package.py pins the
    equality with the nn.Linear construction), so that a seeded run matches the reference."""
    w = torch.empty(1, n_inputs)
    w.uniform_(-1.0, 1.0)
    w.uniform_(-1.0, 1.0)
    return w.normal_(mean=2.0, std=1.5)


class OdeModel(nn.Module):
    """vihds/ode.py:28-96.  Subclasses set ``kernel_model`` (a models.LOOKUP key with a kernel), ``species``,
    ``n_species``, ``precisions`` and ``conditioned`` (names handed to the kernel as per-trajectory extras)."""

    kernel_model = None
    conditioned = ()

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.device_depth = config.data.device_depth
        self.n_treatments = len(config.data.conditions)
        self.use_laplace = bool(config.params.get("use_laplace", False))
        if self.use_laplace:
            raise NotImplementedError("use_laplace: log_prob_laplace raises in the reference (training.py:36-38)")
        self.relevance = config.data.relevance_vectors
        self.default_devices = config.data.default_devices
        self.precisions = None
        self.species = None
        self.n_species = None
        self._problems = {}

    # -- engine plumbing ----------------------------------------------------------------------------------------
    def net_dims(self):
        return dict(n_hidden=getattr(self.precisions, "n_hidden", 0) if self.precisions is not None else 0)

    def attach_conditioned(self, theta, extras, extra, B, IW):
        """Expose the conditioned values on theta the way condition_theta does in the reference (plain attributes)."""
        for k, name in enumerate(extras):
            setattr(theta, name, extra[k].view(B, IW))

    def problem(self, names, kinds, priors, extras, device, dtype):
        """ElboProblem for this model under the current solver / dtype (cached per signature)."""
        key = (tuple(names), tuple(extras), self.config.params.solver, str(device), dtype)
        if key not in self._problems:
            p_mu, p_prec, lo, hi = priors
            self._problems[key] = ElboProblem(
                self.kernel_model, self.config.params.solver, dtype, names, kinds, p_mu, p_prec, lo, hi, extras=extras,
                device=device, C_treat=self.n_treatments, D_dev=self.device_depth, **self.net_dims())
        return self._problems[key]

    def flat_weights(self):
        return self.precisions.flat_weights() if self.precisions is not None else None

    # -- reference surface --------------------------------------------------------------------------------------
    def device_conditioner(self, param, param_name, dev_1hot, use_bias=False, activation="relu"):
        """vihds/ode.py:43-58 including its quirks: random weights per call and the ``repeat([n_iwae, 1])`` +
        reshape that assigns row (b*IW + i) % B of the conditioner output to sample (b, i)."""
        n_batch, n_iwae = param.shape
        w = _draw_conditioner_weight(dev_1hot.shape[1]).to(dev_1hot)
        rel = torch.as_tensor(self.relevance[param_name]).to(dev_1hot)
        cond = torch.relu((dev_1hot * rel) @ w.t())
        cond = cond.repeat([n_iwae, 1])
        flat = param.reshape(n_iwae * n_batch, 1)
        out = flat * (1.0 + cond) if param_name in self.default_devices else flat * cond
        return out.reshape(n_batch, n_iwae)

    def condition_theta(self, theta, dev_1hot, writer, epoch):
        raise NotImplementedError("TODO: write your condition_theta")

    def conditioned_extras(self, B, IW, dev_1hot):
        """[E, N] planes of the conditioned values, in ``self.conditioned`` order (None when E == 0)."""
        return None

    def initialize_state(self, theta, treatments):
        raise NotImplementedError("TODO: write your initialize_state")

    def simulate(self, config, times, theta, conditions, dev_1hot, condition_on_device=True):
        """vihds/ode.py:66-82: theta (DotOperatorSamples of [B, IW] tensors, already clipped / conditioned) ->
        [B, IW, S, T], S = species + dynamic-precision states; a strided view of the kernel's [T][S][N] trace."""
        names = [k for k in theta.keys]
        for extra in self.conditioned:
            if hasattr(theta, extra) and extra not in names:
                names.append(extra)
        # conditioned values replace the samples of the same name (condition_theta overwrites the attribute)
        first = theta.values[0]
        B, IW = first.shape
        dtype = first.dtype
        prob = self.problem((), (), (None, None, None, None), (), first.device, dtype)
        src = prob.simulate_variant(names)
        planes = torch.stack([getattr(theta, n).reshape(B * IW) for n in names])
        xs = SimulateTrace.apply(prob, src, planes, self.flat_weights(), times, conditions, dev_1hot, B, IW)
        return xs.view(times.numel(), prob.S, B, IW).permute(2, 3, 1, 0)

    @classmethod
    def observe(cls, x_sample, _theta):
        """vihds/ode.py:84-93 -- provided for callers that hold a trace; the fused path takes x_predict from the
        kernel instead."""
        od = x_sample[:, :, 0, :]
        return torch.stack([od, od * x_sample[:, :, 1, :], od * (x_sample[:, :, 2, :] + x_sample[:, :, 4, :]),
                            od * (x_sample[:, :, 3, :] + x_sample[:, :, 5, :])], dim=2)

    def expand_precisions(self, theta, times, x_states):
        return self.precisions.expand(theta, len(times), x_states)

    def summaries(self, writer, epoch):
        pass


# ---------------------------------------------------------------------------------------------------------------
# double-receiver family (models/dr_constant.py:114-215)
# ---------------------------------------------------------------------------------------------------------------
class DR_Constant(OdeModel):
    kernel_model = "dr_constant"
    conditioned = ("aR", "aS")
    version = 1

    def __init__(self, config):
        super().__init__(config)
        self.precisions = ConstantPrecisions(["prec_x", "prec_rfp", "prec_yfp", "prec_cfp"])
        self.species = ["OD", "RFP", "YFP", "CFP", "F530", "F480", "LuxR", "LasR"]
        self.n_species = 8
        self.device = config.device

    def condition_theta(self, theta, dev_1hot, writer, epoch):
        """models/dr_constant.py:124-131: aR, aS = conditioner(ones) -- plain attributes, not sampled entries."""
        B, IW = theta.get_n_batch(), theta.get_n_samples()
        ones = torch.ones(B, IW, dtype=dev_1hot.dtype, device=dev_1hot.device)
        theta.aR = self.device_conditioner(ones, "aR", dev_1hot)
        theta.aS = self.device_conditioner(ones, "aS", dev_1hot)
        return theta

    def conditioned_extras(self, B, IW, dev_1hot):
        ones = torch.ones(B, IW, dtype=dev_1hot.dtype, device=dev_1hot.device)
        aR = self.device_conditioner(ones, "aR", dev_1hot)
        aS = self.device_conditioner(ones, "aS", dev_1hot)
        return torch.stack([aR.reshape(-1), aS.reshape(-1)])

    def initialize_state(self, theta, _treatments):
        zero = torch.zeros_like(theta.init_x)
        rows = [theta.init_x, theta.init_rfp, theta.init_yfp, theta.init_cfp, zero, zero, theta.init_luxR, theta.init_lasR]
        if self.precisions.dynamic:
            rows += [theta.init_prec_x, theta.init_prec_rfp, theta.init_prec_yfp, theta.init_prec_cfp]
        return torch.stack(rows, dim=2)


class DR_Constant_V2(DR_Constant):
    kernel_model = "dr_constant_v2"
    version = 2


class DR_Constant_Precisions(DR_Constant):
    kernel_model = "dr_constant_precisions"

    def __init__(self, config):
        super().__init__(config)
        self.precisions = NeuralPrecisions(self.n_species, config.params.n_hidden_decoder_precisions, 4)


class DR_Constant_Precisions_V2(DR_Constant_Precisions):
    kernel_model = "dr_constant_precisions_v2"
    version = 2


# ---------------------------------------------------------------------------------------------------------------
# relay (models/relay_constant.py:137-263; the reference classes are broken as shipped -- SURVEY.md section 8c --
# the behaviour reproduced is that of the two-line monkeypatch recorded in tests/golden/_ref_harness.py)
# ---------------------------------------------------------------------------------------------------------------
class Relay_Constant(DR_Constant):
    kernel_model = "relay_constant"

    def __init__(self, config):
        super().__init__(config)
        self.species = ["OD", "RFP", "YFP", "CFP", "F530", "F480", "LuxR", "LasR", "LuxI", "LasI", "C6", "C12"]
        self.n_species = 12

    def initialize_state(self, theta, treatments):
        zero = torch.zeros_like(theta.init_x)
        c = torch.clamp(torch.exp(treatments) - 1.0, 1e-12, 1e6)
        c6, c12 = c[:, 0:1].expand_as(zero), c[:, 1:2].expand_as(zero)
        rows = [theta.init_x, theta.init_rfp, theta.init_yfp, theta.init_cfp, zero, zero, theta.init_luxR, theta.init_lasR,
                theta.init_luxI, theta.init_lasI, c6, c12]
        if self.precisions.dynamic:
            rows += [theta.init_prec_x, theta.init_prec_rfp, theta.init_prec_yfp, theta.init_prec_cfp]
        return torch.stack(rows, dim=2)


class Relay_Constant_Precisions(Relay_Constant):
    kernel_model = "relay_constant_precisions"

    def __init__(self, config):
        super().__init__(config)
        self.precisions = NeuralPrecisions(self.n_species, config.params.n_hidden_decoder_precisions, 4)


# ---------------------------------------------------------------------------------------------------------------
# growth-only models (models/auto_constant.py:63-132, models/prpr_constant.py:61-130)
# ---------------------------------------------------------------------------------------------------------------
class Auto_Constant(OdeModel):
    kernel_model = "auto_constant"

    def __init__(self, config):
        super().__init__(config)
        self.precisions = ConstantPrecisions(["prec_x", "prec_rfp", "prec_yfp", "prec_cfp"])
        self.species = ["OD", "RFP", "F530", "F480"]
        self.n_species = 4

    def initialize_state(self, theta, _treatments):
        zero = torch.zeros_like(theta.init_x)
        rows = [theta.init_x, theta.init_rfp, zero, zero]
        if self.precisions.dynamic:
            rows += [theta.init_prec_x, theta.init_prec_rfp, theta.init_prec_yfp, theta.init_prec_cfp]
        return torch.stack(rows, dim=2)

    @classmethod
    def observe(cls, x_sample, _theta):
        od = x_sample[:, :, 0, :]
        return torch.stack([od, od * x_sample[:, :, 1, :], od * x_sample[:, :, 2, :], od * x_sample[:, :, 3, :]], dim=2)


class Auto_Constant_Precisions(Auto_Constant):
    kernel_model = "auto_constant_precisions"

    def __init__(self, config):
        super().__init__(config)
        self.precisions = NeuralPrecisions(self.n_species, config.params.n_hidden_decoder_precisions, 4)


class PRPR_Constant(OdeModel):
    kernel_model = "prpr_constant"

    def __init__(self, config):
        super().__init__(config)
        self.precisions = ConstantPrecisions(["prec_x", "prec_rfp", "prec_yfp", "prec_cfp"])
        self.species = ["OD", "RFP", "YFP", "CFP", "F530", "F480"]
        self.n_species = 6

    def initialize_state(self, theta, _treatments):
        zero = torch.zeros_like(theta.init_x)
        rows = [theta.init_x, theta.init_rfp, theta.init_yfp, theta.init_cfp, zero, zero]
        if self.precisions.dynamic:
            rows += [theta.init_prec_x, theta.init_prec_rfp, theta.init_prec_yfp, theta.init_prec_cfp]
        return torch.stack(rows, dim=2)


class PRPR_Constant_Precisions(PRPR_Constant):
    kernel_model = "prpr_constant_precisions"

    def __init__(self, config):
        super().__init__(config)
        self.precisions = NeuralPrecisions(self.n_species, config.params.n_hidden_decoder_precisions, 4)


# ---------------------------------------------------------------------------------------------------------------
# inducer (models/inducer_constant.py:82-151) and degrader (models/degrader_constant.py:146-269).  Like relay, the
# reference classes do not construct as shipped (init_with_params / OdeFunc.__init__ arity); the
# behaviour reproduced is that of the monkeypatch in oracle/ref_harness.py, pinned by golden cases minted under it.
# ---------------------------------------------------------------------------------------------------------------
class Inducer_Constant(OdeModel):
    kernel_model = "inducer_constant"

    def __init__(self, config):
        super().__init__(config)
        self.precisions = ConstantPrecisions(["prec_x", "prec_rfp", "prec_yfp", "prec_cfp"])
        self.species = ["OD", "RFP", "YFP", "F530", "F480"]
        self.n_species = 5

    def initialize_state(self, theta, _treatments):
        zero = torch.zeros_like(theta.init_x)
        rows = [theta.init_x, theta.init_rfp, theta.init_yfp, zero, zero]
        if self.precisions.dynamic:
            rows += [theta.init_prec_x, theta.init_prec_rfp, theta.init_prec_yfp, theta.init_prec_cfp]
        return torch.stack(rows, dim=2)

    @classmethod
    def observe(cls, x_sample, _theta):
        od = x_sample[:, :, 0, :]
        return torch.stack([od, od * x_sample[:, :, 1, :], od * (x_sample[:, :, 2, :] + x_sample[:, :, 3, :]),
                            od * x_sample[:, :, 4, :]], dim=2)


class Inducer_Constant_Precisions(Inducer_Constant):
    kernel_model = "inducer_constant_precisions"

    def __init__(self, config):
        super().__init__(config)
        self.precisions = NeuralPrecisions(self.n_species, config.params.n_hidden_decoder_precisions, 4)


class Degrader_Constant(OdeModel):
    """Double receiver + AiiA degrader; aR / aS are sampled parameters here (no device conditioning: the reference has
    it commented out, degrader_constant.py:56-69)."""

    kernel_model = "degrader_constant"

    def __init__(self, config):
        super().__init__(config)
        self.precisions = ConstantPrecisions(["prec_x", "prec_rfp", "prec_yfp", "prec_cfp"])
        self.species = ["OD", "RFP", "YFP", "CFP", "F530", "F480", "LuxR", "LasR", "AiiA", "C6", "C12"]
        self.n_species = 11

    def initialize_state(self, theta, treatments):
        zero = torch.zeros_like(theta.init_x)
        c = torch.clamp(torch.exp(treatments) - 1.0, 1e-12, 1e6)
        c6, c12 = c[:, 0:1].expand_as(zero), c[:, 1:2].expand_as(zero)
        rows = [theta.init_x, theta.init_rfp, theta.init_yfp, theta.init_cfp, zero, zero, theta.init_luxR, theta.init_lasR,
                theta.init_aiiA, c6, c12]
        if self.precisions.dynamic:
            rows += [theta.init_prec_x, theta.init_prec_rfp, theta.init_prec_yfp, theta.init_prec_cfp]
        return torch.stack(rows, dim=2)


class Degrader_Constant_Precisions(Degrader_Constant):
    kernel_model = "degrader_constant_precisions"

    def __init__(self, config):
        super().__init__(config)
        self.precisions = NeuralPrecisions(self.n_species, config.params.n_hidden_decoder_precisions, 4)


# ---------------------------------------------------------------------------------------------------------------
# black box (models/dr_blackbox.py:61-125, vihds/ode.py:119-138)
# ---------------------------------------------------------------------------------------------------------------
class NeuralStates(nn.Module):
    """ode.py:119-138: same layer names and initialisation order as the reference (state_dict compatible)."""

    def __init__(self, n_inputs, n_hidden, n_states, n_latents):
        super().__init__()
        self.n_latents, self.n_states = n_latents, n_states
        self.states_hidden = nn.Linear(n_inputs, n_hidden)
        nn.init.xavier_uniform_(self.states_hidden.weight)
        self.states_production = nn.Linear(n_hidden, n_states)
        nn.init.xavier_uniform_(self.states_production.weight)
        self.states_degradation = nn.Linear(n_hidden, n_states)
        nn.init.xavier_uniform_(self.states_degradation.weight)

    def layers(self):
        return [self.states_hidden, self.states_production, self.states_degradation]


class DR_Blackbox(OdeModel):
    """MLP right-hand side over [states, latent parameters, treatments, device one-hot] + a NeuralPrecisions net with
    a hidden layer; y_k are conditioned on the device through a trainable offset layer (dr_blackbox.py:86-96)."""

    kernel_model = "dr_blackbox"

    def __init__(self, config):
        super().__init__(config)
        p = config.params
        self.n_x, self.n_y, self.n_z = p.n_x, p.n_y, p.n_z
        n_latents = self.n_x + self.n_y + self.n_z
        self.n_species = 4
        self.n_latent_species = p.n_latent_species
        self.n_hidden_precisions = p.n_hidden_decoder_precisions
        self.n_states = self.n_species + self.n_latent_species
        n_inputs = self.n_states + n_latents + self.n_treatments + self.device_depth
        self.precisions = NeuralPrecisions(n_inputs, self.n_hidden_precisions, 4)
        self.species = ["OD", "RFP", "YFP", "CFP"]
        self.n_hidden = p.n_hidden_decoder
        self.init_latent_species = p.get("init_latent_species", 0.001)
        self.init_prec = p.get("init_prec", 0.00001)
        self.offset_layer = nn.Linear(self.device_depth, self.n_y)
        self.neural_states = NeuralStates(n_inputs, p.n_hidden_decoder, self.n_states, n_latents)
        self.latent_names = (["z%d" % (i + 1) for i in range(self.n_z)] + ["x%d" % (i + 1) for i in range(self.n_x)] +
                             ["y%d" % (i + 1) for i in range(self.n_y)])
        self.conditioned = tuple("offset:y%d" % (i + 1) for i in range(self.n_y))

    def net_dims(self):
        return dict(n_hidden=self.n_hidden_precisions, n_hidden_states=self.n_hidden, n_latent=self.n_latent_species,
                    n_z=self.n_z, n_x=self.n_x, n_y=self.n_y, init_latent_species=self.init_latent_species,
                    init_prec=self.init_prec, slot_alias={"latent%d" % k: nm for k, nm in enumerate(self.latent_names)})

    def flat_weights(self):
        layers = self.neural_states.layers() + self.precisions.layers()
        return torch.cat([t.reshape(-1) for lin in layers for t in (lin.weight, lin.bias)])

    def conditioned_extras(self, B, IW, dev_1hot):
        """[n_y, N] planes of offset_layer(dev_1hot)[:, k]; the kernel adds them to the sampled y_k."""
        off = self.offset_layer(dev_1hot)
        return off.t().repeat_interleave(IW, dim=1).contiguous()

    def attach_conditioned(self, theta, extras, extra, B, IW):
        for k in range(self.n_y):
            name = "y%d" % (k + 1)
            setattr(theta, name, theta.samples[name] + extra[k].view(B, IW))

    def condition_theta(self, theta, dev_1hot, writer, epoch):
        off = self.offset_layer(dev_1hot)
        for k in range(self.n_y):
            name = "y%d" % (k + 1)
            setattr(theta, name, getattr(theta, name) + off[:, k:k + 1])
        return theta

    def initialize_state(self, theta, _treatments):
        B, IW = theta.get_n_batch(), theta.get_n_samples()
        x0 = torch.stack([theta.init_x, theta.init_rfp, theta.init_yfp, theta.init_cfp], dim=2)
        h0 = torch.full([B, IW, self.n_latent_species], self.init_latent_species, dtype=x0.dtype, device=x0.device)
        prec0 = torch.full([B, IW, 4], self.init_prec, dtype=x0.dtype, device=x0.device)
        return torch.cat([x0, h0, prec0], dim=2)

    @classmethod
    def observe(cls, x_sample, _theta):
        od = x_sample[:, :, 0, :]
        return torch.stack([od, od * x_sample[:, :, 1, :], od * x_sample[:, :, 2, :], od * x_sample[:, :, 3, :]], dim=2)


LOOKUP = {
    "auto_constant": Auto_Constant,
    "auto_constant_precisions": Auto_Constant_Precisions,
    "prpr_constant": PRPR_Constant,
    "prpr_constant_precisions": PRPR_Constant_Precisions,
    "dr_blackbox": DR_Blackbox,
    "dr_constant": DR_Constant,
    "dr_constant_v2": DR_Constant_V2,
    "dr_constant_precisions": DR_Constant_Precisions,
    "dr_constant_precisions_v2": DR_Constant_Precisions_V2,
    "relay_constant": Relay_Constant,
    "relay_constant_precisions": Relay_Constant_Precisions,
    "inducer_constant": Inducer_Constant,
    "inducer_constant_precisions": Inducer_Constant_Precisions,
    "degrader_constant": Degrader_Constant,
    "degrader_constant_precisions": Degrader_Constant_Precisions,
}


def kernel_available(name):
    return name in LOOKUP and L.load().vh_model_id(LOOKUP[name].kernel_model.encode()) >= 0
