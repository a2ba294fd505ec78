"""CPU ORACLE for the vi-hds hot path (sample -> clip -> ODE solve -> observe -> log-lik -> log p / log q -> IWAE cost).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package ``vihds_b200`` imports this file; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only as
the checker / the timed CPU baseline.  The product path fails loudly if its CUDA library is missing.

It is an independent restatement (written from the maths, not copied) of the reference's algorithm in plain PyTorch
**CPU** tensors, deliberately at the reference's own granularity -- one small ATen op per arithmetic operation on
``[B, IW]`` tensors, autograd for the gradient -- so that (a) gradients come for free for parity checks and (b) timing
it reproduces the cost structure of the reference's CPU path (~83 k ATen dispatches per step).

Pinning (tests/test_oracle_golden.py): every function here is checked against golden vectors minted by running the
reference itself (tests/golden/make_golden.py).  The in-repo solvers ``modeuler``/``modeulerwhile`` are pinned to the
reference bit-for-bit-class (fp32 round-off).  ``midpoint``/``rk4``/``euler`` go through torchdiffeq==0.1 in the
reference (ode.py:80-81), a third-party dependency that is neither vendored nor installable here: their scheme is
restated from torchdiffeq 0.1's published fixed-grid algorithm and is therefore **PARITY UNPINNED** with respect to
the third-party arithmetic (the reference's own pin at that boundary is only CV < 5 %, tests/test_ode_solvers.py:89).

Reference lines followed by each function are cited in its docstring (paths relative to /root/reference).
"""
import math

import torch

LOG2PI = math.log(2.0 * math.pi)
KIND_CONSTANT, KIND_NORMAL, KIND_LOGNORMAL = 0, 1, 2


# ----------------------------------------------------------------------------------------------------------------
# distributions: sample / clip / log_prob
# ----------------------------------------------------------------------------------------------------------------
def sample_theta(u, q_mu, q_prec, kinds):
    """vihds/distributions.py:119-142 (ChainedDistribution.sample), :327-330, :369-371, :242-243.

    u [B,IW,P]; q_mu,q_prec [B,P] (global rows are constant over B); returns a list of P tensors [B,IW]."""
    out = []
    for k, kind in enumerate(kinds):
        uk = u[:, :, k]
        if kind == KIND_CONSTANT:
            out.append(torch.zeros_like(uk) + q_mu[:, k : k + 1])
            continue
        sigma = 1.0 / q_prec[:, k : k + 1].sqrt()
        s = q_mu[:, k : k + 1] + sigma * uk
        out.append(s.exp() if kind == KIND_LOGNORMAL else s)
    return out


def clip_theta(theta, p_mu, p_sigma, kinds, stddevs=4.0):
    """vihds/distributions.py:76-85, :332-336, :377-381 -- clamp to the PRIOR's mu +- stddevs*sigma (bounds detached)."""
    out = []
    for k, kind in enumerate(kinds):
        if kind == KIND_CONSTANT:
            out.append(theta[k])
            continue
        lo = (p_mu[k] - stddevs * p_sigma[k]).detach()
        hi = (p_mu[k] + stddevs * p_sigma[k]).detach()
        if kind == KIND_LOGNORMAL:
            lo, hi = lo.exp(), hi.exp()
        out.append(theta[k].clamp(float(lo), float(hi)))
    return out


def log_prob_theta(theta, mu, prec, kinds):
    """vihds/distributions.py:64-74, :338-345 (note -LOG2PI, not -LOG2PI/2, and +1e-12 inside the logs), :373-375.

    mu, prec: [B,P] (q) or [P] (prior).  Returns [B,IW]."""
    terms = []
    for k, kind in enumerate(kinds):
        x = theta[k]
        if kind == KIND_CONSTANT:
            terms.append(torch.zeros_like(x))
            continue
        m = mu[..., k : k + 1] if mu.dim() == 2 else mu[k]
        pr = prec[..., k : k + 1] if prec.dim() == 2 else prec[k]
        if kind == KIND_LOGNORMAL:
            lx = (x + 1e-12).log()
            terms.append(-LOG2PI + 0.5 * (pr + 1e-12).log() - 0.5 * pr * (m - lx).pow(2) - lx)
        else:
            terms.append(-LOG2PI + 0.5 * (pr + 1e-12).log() - 0.5 * pr * (m - x).pow(2))
    return torch.stack(terms, -1).sum(-1)


# ----------------------------------------------------------------------------------------------------------------
# neural precisions and neural states (dynamic-precision ODE states, black-box RHS)
# ----------------------------------------------------------------------------------------------------------------
class NeuralPrecisionsOracle:
    """vihds/precisions.py:44-94.  hidden == 0: sigma(W . act(x) + b); hidden > 0: sigma(W2 . act(W1 x + b1) + b2)."""

    def __init__(self, weights, act):
        self.w = weights
        self.act = act
        self.hidden = "prec_hidden.weight" in weights

    def __call__(self, t, state, constants):
        species = state[:, :, :-4]
        v = state[:, :, -4:]
        tt = t.reshape(1, 1, 1).expand(state.shape[0], state.shape[1], 1)
        parts = [tt, species] + ([constants] if constants is not None else [])
        x = torch.cat(parts, dim=2)
        w = self.w
        if self.hidden:
            h = self.act(torch.nn.functional.linear(x, w["prec_hidden.weight"], w["prec_hidden.bias"]))
        else:
            h = self.act(x)
        prod = torch.sigmoid(torch.nn.functional.linear(h, w["prec_production.weight"], w["prec_production.bias"]))
        degr = torch.sigmoid(torch.nn.functional.linear(h, w["prec_degradation.weight"], w["prec_degradation.bias"]))
        return prod - degr * v


def _treatments(inputs, n_iwae):
    """models/dr_constant.py:26-29: c = clamp(exp(x)-1, 1e-12, 1e6) tiled over the IW axis."""
    tr = torch.clamp(torch.exp(inputs) - 1.0, 1e-12, 1e6)
    c6 = tr[:, 0:1].expand(-1, n_iwae)
    c12 = tr[:, 1:2].expand(-1, n_iwae)
    return c6, c12


def _pbad(ara, th):
    """models/inducer_constant.py:48-55, models/degrader_constant.py:79-86: arabinose promoter activity."""
    nA = torch.clamp(th["nA"], 0.5, 3.0)
    return (ara.pow(nA) + th["eA"] * th["KAra"].pow(nA)) / (ara.pow(nA) + th["KAra"].pow(nA))


# ----------------------------------------------------------------------------------------------------------------
# white-box right-hand sides
# ----------------------------------------------------------------------------------------------------------------
class DrConstantRHS:
    """models/dr_constant.py:14-112 (v1 :62-68, v2 :69-73); relay extension models/relay_constant.py:13-134."""

    def __init__(self, th, inputs, version=1, precisions=None, relay=False, degrader=False):
        n_iwae = th["r"].shape[1]
        c6, c12 = _treatments(inputs, n_iwae)
        cl = torch.clamp
        self.r, self.K = cl(th["r"], 0.0, 4.0), cl(th["K"], 0.0, 4.0)
        self.tlag, self.rc, self.a530, self.a480 = th["tlag"], th["rc"], th["a530"], th["a480"]
        self.drfp, self.dyfp, self.dcfp = (cl(th[n], 1e-12, 2.0) for n in ("drfp", "dyfp", "dcfp"))
        self.dR, self.dS = cl(th["dR"], 1e-12, 5.0), cl(th["dS"], 1e-12, 5.0)
        for n in ("e76", "e81", "aCFP", "aYFP", "KGR_76", "KGS_76", "KGR_81", "KGS_81", "aR", "aS"):
            setattr(self, n, th[n])
        nR, nS = cl(th["nR"], 0.5, 3.0), cl(th["nS"], 0.5, 3.0)
        if version == 1:
            KR6, KR12, KS6, KS12 = (cl(th[n], 1e-12, 1.0) for n in ("KR6", "KR12", "KS6", "KS12"))
            self.fracLuxR = ((KR6 * c6).pow(nR) + (KR12 * c12).pow(nR)) / (1.0 + KR6 * c6 + KR12 * c12).pow(nR)
            self.fracLasR = ((KS6 * c6).pow(nS) + (KS12 * c12).pow(nS)) / (1.0 + KS6 * c6 + KS12 * c12).pow(nS)
        else:
            eS6, eR12 = cl(th["eS6"], 1e-12, 1.0), cl(th["eR12"], 1e-12, 1.0)
            self.fracLuxR = c6.pow(nR) + (eR12 * c12).pow(nR)
            self.fracLasR = (eS6 * c6).pow(nS) + c12.pow(nS)
        self.relay = relay
        if relay:
            self.dlasI, self.dluxI = cl(th["dlasI"], 1e-12, 5.0), cl(th["dluxI"], 1e-12, 5.0)
            self.KC6, self.KC12, self.Klux, self.Klas = th["KC6"], th["KC12"], th["Klux"], th["Klas"]
        self.degrader = degrader
        if degrader:  # models/degrader_constant.py:72-88
            ara = torch.clamp(torch.exp(inputs[:, 2:3]) - 1.0, 1e-12, 1e6).expand(-1, n_iwae)
            self.aI, self.daiiA = th["aI"], th["daiiA"]
            self.PBAD = _pbad(ara, th)
            self.rC6, self.rC12 = th["dA6"] * c6, th["dA12"] * c12
        self.precisions = precisions
        self.n_species = 12 if relay else (11 if degrader else 8)

    def __call__(self, t, state):
        s = state
        x, rfp, yfp, cfp, f530, f480, luxR, lasR = (s[:, :, i] for i in range(8))
        gr = self.r * torch.sigmoid(4.0 * (t - self.tlag))
        gamma = gr * (1.0 - x / self.K)
        bR = luxR * luxR * self.fracLuxR
        bS = lasR * lasR * self.fracLasR
        P76 = (self.e76 + self.KGR_76 * bR + self.KGS_76 * bS) / (1.0 + self.KGR_76 * bR + self.KGS_76 * bS)
        P81 = (self.e81 + self.KGR_81 * bR + self.KGS_81 * bS) / (1.0 + self.KGR_81 * bR + self.KGS_81 * bS)
        d = [
            gamma * x,
            self.rc - (gamma + self.drfp) * rfp,
            self.rc * self.aYFP * P81 - (gamma + self.dyfp) * yfp,
            self.rc * self.aCFP * P76 - (gamma + self.dcfp) * cfp,
            self.rc * self.a530 - gamma * f530,
            self.rc * self.a480 - gamma * f480,
            self.rc * self.aR - (gamma + self.dR) * luxR,
            self.rc * self.aS - (gamma + self.dS) * lasR,
        ]
        if self.relay:
            luxI, lasI = s[:, :, 8], s[:, :, 9]
            d += [
                self.rc * P81 - (gamma + self.dluxI) * luxI,
                self.rc * P76 - (gamma + self.dlasI) * lasI,
                (self.KC6 * self.rc * x * luxI) / (1.0 + luxI / self.Klux),
                (self.KC12 * self.rc * x * lasI) / (1.0 + lasI / self.Klas),
            ]
        if self.degrader:  # models/degrader_constant.py:128-131, verbatim (the constant loss is not multiplied by AiiA)
            aiiA = s[:, :, 8]
            d += [self.rc * self.aI * self.PBAD - (self.daiiA + (gamma * aiiA)), x * self.rC6 * aiiA, x * self.rC12 * aiiA]
        dX = torch.stack(d, dim=2)
        if self.precisions is not None:
            return torch.cat([dX, self.precisions(t, state, None)], dim=2)
        return dX


class GrowthRHS:
    """models/auto_constant.py:11-60 (4 species: OD, RFP, F530, F480) and models/prpr_constant.py:11-58 (6 species:
    + constitutively expressed YFP, CFP): growth + dilution without receivers."""

    def __init__(self, th, prpr, precisions=None):
        cl = torch.clamp
        self.r, self.K = cl(th["r"], 0.0, 4.0), cl(th["K"], 0.0, 4.0)
        self.tlag, self.rc, self.a530, self.a480 = th["tlag"], th["rc"], th["a530"], th["a480"]
        self.drfp = cl(th["drfp"], 1e-12, 2.0)
        self.prpr = prpr
        if prpr:
            self.dyfp, self.dcfp = cl(th["dyfp"], 1e-12, 2.0), cl(th["dcfp"], 1e-12, 2.0)
            self.aYFP, self.aCFP = th["aYFP_PR"], th["aCFP_PR"]
        self.precisions = precisions

    def __call__(self, t, state):
        x, rfp = state[:, :, 0], state[:, :, 1]
        gamma = self.r * torch.sigmoid(4.0 * (t - self.tlag)) * (1.0 - x / self.K)
        d = [gamma * x, self.rc - (gamma + self.drfp) * rfp]
        if self.prpr:
            yfp, cfp, f530, f480 = (state[:, :, i] for i in (2, 3, 4, 5))
            d += [self.rc * self.aYFP - (gamma + self.dyfp) * yfp, self.rc * self.aCFP - (gamma + self.dcfp) * cfp]
        else:
            f530, f480 = state[:, :, 2], state[:, :, 3]
        d += [self.rc * self.a530 - gamma * f530, self.rc * self.a480 - gamma * f480]
        dX = torch.stack(d, dim=2)
        if self.precisions is not None:
            return torch.cat([dX, self.precisions(t, state, None)], dim=2)
        return dX


class InducerRHS:
    """models/inducer_constant.py:12-84: growth + dilution, YFP expressed from the arabinose promoter PBAD
    (5 species: OD, RFP, YFP, F530, F480)."""

    def __init__(self, th, inputs, precisions=None):
        cl = torch.clamp
        n_iwae = th["r"].shape[1]
        ara = cl(torch.exp(inputs[:, 0:1]) - 1.0, 1e-12, 1e6).expand(-1, n_iwae)
        self.r, self.K = cl(th["r"], 0.0, 4.0), cl(th["K"], 0.0, 4.0)
        self.tlag, self.rc, self.a530, self.a480 = th["tlag"], th["rc"], th["a530"], th["a480"]
        self.drfp, self.dyfp = cl(th["drfp"], 1e-12, 2.0), cl(th["dyfp"], 1e-12, 2.0)
        self.aYFP = th["aYFP_Inducer"]
        self.PBAD = _pbad(ara, th)
        self.precisions = precisions

    def __call__(self, t, state):
        x, rfp, yfp, f530, f480 = (state[:, :, i] for i in range(5))
        gamma = self.r * torch.sigmoid(4.0 * (t - self.tlag)) * (1.0 - x / self.K)
        d = [gamma * x, self.rc - (gamma + self.drfp) * rfp, self.rc * self.aYFP * self.PBAD - (gamma + self.dyfp) * yfp,
             self.rc * self.a530 - gamma * f530, self.rc * self.a480 - gamma * f480]
        dX = torch.stack(d, dim=2)
        if self.precisions is not None:
            return torch.cat([dX, self.precisions(t, state, None)], dim=2)
        return dX


class BlackboxRHS:
    """models/dr_blackbox.py:15-58 + vihds/ode.py:119-138 (NeuralStates)."""

    def __init__(self, th, inputs, dev_1hot, weights, n_z, n_x, n_y):
        n_iwae = th["z1"].shape[1]
        lat = [th["z%d" % (i + 1)] for i in range(n_z)] + [th["x%d" % (i + 1)] for i in range(n_x)]
        parts = [torch.stack(lat, dim=-1)]
        if n_y > 0:
            parts.append(torch.stack([th["y%d" % (i + 1)] for i in range(n_y)], dim=-1))
        parts.append(inputs.unsqueeze(1).expand(-1, n_iwae, -1))
        parts.append(dev_1hot.unsqueeze(1).expand(-1, n_iwae, -1))
        self.constants = torch.cat(parts, dim=2)
        self.w = weights
        self.precisions = NeuralPrecisionsOracle(
            {k[len("precisions."):]: v for k, v in weights.items() if k.startswith("precisions.")}, torch.relu)

    def __call__(self, t, state):
        F = torch.nn.functional
        w = self.w
        x = state[:, :, :-4]
        aug = torch.cat([x, self.constants], dim=2)
        h = torch.relu(F.linear(aug, w["neural_states.states_hidden.weight"], w["neural_states.states_hidden.bias"]))
        prod = torch.sigmoid(F.linear(h, w["neural_states.states_production.weight"], w["neural_states.states_production.bias"]))
        degr = torch.sigmoid(F.linear(h, w["neural_states.states_degradation.weight"], w["neural_states.states_degradation.bias"]))
        dx = prod - degr * x
        return torch.cat([dx, self.precisions(t, state, self.constants)], dim=2)


# ----------------------------------------------------------------------------------------------------------------
# fixed-step solvers
# ----------------------------------------------------------------------------------------------------------------
def integrate(f, x0, times, solver):
    """vihds/solvers.py:9-41 (modeuler: CONSTANT h = t1-t0 for all steps; modeulerwhile: per-step h) and the
    fixed-grid schemes of torchdiffeq==0.1 called at vihds/ode.py:80-81 (parity unpinned, see module docstring).
    Returns [T,B,IW,S]."""
    xs = [x0]
    x = x0
    h0 = times[1] - times[0]
    for t0, t1 in zip(times[:-1], times[1:]):
        dt = t1 - t0
        if solver == "modeuler":
            f1 = f(t0, x)
            f2 = f(t1, x + h0 * f1)
            x = x + 0.5 * h0 * (f1 + f2)
        elif solver == "modeulerwhile":
            f1 = f(t0, x)
            f2 = f(t1, x + dt * f1)
            x = x + 0.5 * dt * (f1 + f2)
        elif solver == "euler":
            x = x + dt * f(t0, x)
        elif solver == "midpoint":
            xm = x + f(t0, x) * dt / 2
            x = x + dt * f(t0 + dt / 2, xm)
        elif solver == "rk4":  # torchdiffeq 0.1 "rk4" is the 3/8 rule
            k1 = f(t0, x)
            k2 = f(t0 + dt / 3, x + dt * k1 / 3)
            k3 = f(t0 + dt * 2 / 3, x + dt * (k1 / -3 + k2))
            k4 = f(t0 + dt, x + dt * (k1 - k2 + k3))
            x = x + (k1 + 3 * k2 + 3 * k3 + k4) * (dt / 8)
        else:
            raise NotImplementedError(solver)
        xs.append(x)
    return torch.stack(xs)


# ----------------------------------------------------------------------------------------------------------------
# model glue: initial state, observe, precisions, log-likelihood, IWAE cost
# ----------------------------------------------------------------------------------------------------------------
MODEL_FAMILY = {
    "dr_constant": ("dr", 1, False), "dr_constant_v2": ("dr", 2, False),
    "dr_constant_precisions": ("dr", 1, True), "dr_constant_precisions_v2": ("dr", 2, True),
    "relay_constant": ("relay", 1, False), "relay_constant_precisions": ("relay", 1, True),
    "dr_blackbox": ("blackbox", 0, True),
    "auto_constant": ("auto", 1, False), "auto_constant_precisions": ("auto", 1, True),
    "prpr_constant": ("prpr", 1, False), "prpr_constant_precisions": ("prpr", 1, True),
    "inducer_constant": ("inducer", 1, False), "inducer_constant_precisions": ("inducer", 1, True),
    "degrader_constant": ("degrader", 1, False), "degrader_constant_precisions": ("degrader", 1, True),
}


def decode(model, solver, th, times, inputs, dev_1hot, weights=None, params=None):
    """vihds/decoders.py:28-45 minus condition_theta (the caller supplies conditioned entries in ``th``):
    initialize_state (dr_constant.py:133-150, :178-199; relay_constant.py:220-250; dr_blackbox.py:98-104),
    simulate (ode.py:66-82), expand_precisions (precisions.py:31-35, :89-94), observe (ode.py:84-93; dr_blackbox.py:112-121).

    Returns x_states [B,IW,S,T], x_predict [B,IW,4,T], precisions [B,IW,4,T]."""
    family, version, dyn_prec = MODEL_FAMILY[model]
    any_th = next(iter(th.values()))
    B, IW = any_th.shape
    zero = torch.zeros(B, IW, dtype=any_th.dtype)
    weights = weights or {}
    if family in ("dr", "relay", "degrader"):
        prec = None
        if dyn_prec:
            prec = NeuralPrecisionsOracle({k[len("precisions."):]: v for k, v in weights.items()}, torch.tanh)
        x0 = [th["init_x"], th["init_rfp"], th["init_yfp"], th["init_cfp"], zero, zero, th["init_luxR"], th["init_lasR"]]
        if family == "relay":
            c6, c12 = _treatments(inputs, IW)
            x0 += [th["init_luxI"], th["init_lasI"], c6, c12]
        if family == "degrader":  # models/degrader_constant.py:168-196
            c6, c12 = _treatments(inputs, IW)
            x0 += [th["init_aiiA"], c6, c12]
        if dyn_prec:
            x0 += [th["init_prec_x"], th["init_prec_rfp"], th["init_prec_yfp"], th["init_prec_cfp"]]
        f = DrConstantRHS(th, inputs, version=version, precisions=prec, relay=(family == "relay"), degrader=(family == "degrader"))
    elif family == "inducer":  # models/inducer_constant.py:92-98, :123-142
        prec = None
        if dyn_prec:
            prec = NeuralPrecisionsOracle({k[len("precisions."):]: v for k, v in weights.items()}, torch.tanh)
        x0 = [th["init_x"], th["init_rfp"], th["init_yfp"], zero, zero]
        if dyn_prec:
            x0 += [th["init_prec_x"], th["init_prec_rfp"], th["init_prec_yfp"], th["init_prec_cfp"]]
        f = InducerRHS(th, inputs, precisions=prec)
    elif family in ("auto", "prpr"):
        prec = None
        if dyn_prec:
            prec = NeuralPrecisionsOracle({k[len("precisions."):]: v for k, v in weights.items()}, torch.tanh)
        x0 = [th["init_x"], th["init_rfp"]] + ([th["init_yfp"], th["init_cfp"]] if family == "prpr" else []) + [zero, zero]
        if dyn_prec:
            x0 += [th["init_prec_x"], th["init_prec_rfp"], th["init_prec_yfp"], th["init_prec_cfp"]]
        f = GrowthRHS(th, family == "prpr", precisions=prec)
    else:
        n_lat = params["n_latent_species"]
        x0 = [th["init_x"], th["init_rfp"], th["init_yfp"], th["init_cfp"]]
        x0 += [torch.full_like(zero, params.get("init_latent_species", 0.001))] * n_lat
        x0 += [torch.full_like(zero, params.get("init_prec", 0.00001))] * 4
        f = BlackboxRHS(th, inputs, dev_1hot, weights, params["n_z"], params["n_x"], params["n_y"])
    sol = integrate(f, torch.stack(x0, dim=2), times, solver).permute(1, 2, 3, 0)
    if dyn_prec:
        x_states, precisions = sol[:, :, :-4, :], sol[:, :, -4:, :]
    else:
        x_states = sol
        precisions = torch.stack([th[n] for n in ("prec_x", "prec_rfp", "prec_yfp", "prec_cfp")], dim=-1)
        precisions = precisions.unsqueeze(3).repeat(1, 1, 1, len(times))
    od = x_states[:, :, 0, :]
    if family in ("blackbox", "auto"):  # models/dr_blackbox.py:112-121, models/auto_constant.py:81-89
        obs = [od, od * x_states[:, :, 1, :], od * x_states[:, :, 2, :], od * x_states[:, :, 3, :]]
    elif family == "inducer":  # models/inducer_constant.py:102-110
        obs = [od, od * x_states[:, :, 1, :], od * (x_states[:, :, 2, :] + x_states[:, :, 3, :]), od * x_states[:, :, 4, :]]
    else:
        obs = [od, od * x_states[:, :, 1, :], od * (x_states[:, :, 2, :] + x_states[:, :, 4, :]),
               od * (x_states[:, :, 3, :] + x_states[:, :, 5, :])]
    x_predict = torch.stack(obs, dim=2)
    return x_states, x_predict, precisions


def log_prob_observations(x_predict, observations, precisions):
    """vihds/training.py:24-33, :41-44: Gaussian log-density summed over time -> [B,IW,4]."""
    x_obs = observations.unsqueeze(1)
    lp = -0.5 * (LOG2PI - precisions.log() + precisions * (x_predict - x_obs).pow(2))
    return lp.sum(3)


def iwae_cost(log_p_by_species, log_p_theta, log_q_theta):
    """vihds/training.py:134-148: cost = -mean_b(logsumexp_i log w - log IW); the reference returns it as ``.elbo``."""
    log_w = log_p_by_species.sum(dim=2) + log_p_theta - log_q_theta
    lse = log_w.logsumexp(dim=1, keepdim=True)
    return -(lse - math.log(log_w.shape[1])).mean(), log_w


def condition_blackbox_y(th, dev_1hot, weights, n_y):
    """models/dr_blackbox.py:86-96: y_k += offset_layer(dev_1hot)[:, k]."""
    off = torch.nn.functional.linear(dev_1hot, weights["offset_layer.weight"], weights["offset_layer.bias"])
    th = dict(th)
    for i in range(n_y):
        th["y%d" % (i + 1)] = th["y%d" % (i + 1)] + off[:, i : i + 1]
    return th


def elbo_step(case, requires_grad=True):
    """One full hot-path pass on a golden-style case dict (numpy arrays; see tests/golden/make_golden.py).

    Returns a dict of torch tensors: theta [P,B,IW], x_states, x_predict, precisions, log_p_by_species, log_p_theta,
    log_q_theta, loss and (if requires_grad) grad_q_mu/grad_q_prec [B,P] plus decoder-weight grads ``gw:<name>``."""
    dt = torch.float64 if str(case["dtype"]) == "float64" else torch.float32
    T_ = lambda a: torch.as_tensor(a, dtype=dt)  # noqa: E731
    names = [str(n) for n in case["names"]]
    kinds = [int(k) for k in case["kinds"]]
    model, solver = str(case["model"]), str(case["solver"])
    u, times, inputs, dev_1hot, obs = (T_(case[k]) for k in ("u", "times", "inputs", "dev_1hot", "observations"))
    q_mu, q_prec = T_(case["q_mu"]).requires_grad_(requires_grad), T_(case["q_prec"]).requires_grad_(requires_grad)
    p_mu, p_prec, p_sigma = T_(case["p_mu"]), T_(case["p_prec"]), T_(case["p_sigma"])
    weights = {k[len("w:ode_model."):]: T_(case[k]).requires_grad_(requires_grad) for k in case.keys() if k.startswith("w:")}
    params = case.get("params", None)

    theta = clip_theta(sample_theta(u, q_mu, q_prec, kinds), p_mu, p_sigma, kinds, 4.0)
    th = dict(zip(names, theta))
    thc = dict(th)
    for extra in ("aR", "aS"):
        if "cond_" + extra in case.keys():
            thc[extra] = T_(case["cond_" + extra])
    if model == "dr_blackbox":
        thc = condition_blackbox_y(thc, dev_1hot, weights, params["n_y"])
    x_states, x_predict, precisions = decode(model, solver, thc, times, inputs, dev_1hot, weights, params)
    lpx = log_prob_observations(x_predict, obs, precisions)
    lq = log_prob_theta(theta, q_mu, q_prec, kinds)
    lp = log_prob_theta(theta, p_mu, p_prec, kinds)
    loss, log_w = iwae_cost(lpx, lp, lq)
    out = {"theta": torch.stack(theta), "x_states": x_states, "x_predict": x_predict, "precisions": precisions,
           "log_p_by_species": lpx, "log_p_theta": lp, "log_q_theta": lq, "loss": loss, "log_w": log_w}
    if requires_grad:
        loss.backward()
        out["grad_q_mu"], out["grad_q_prec"] = q_mu.grad, q_prec.grad
        for k, w in weights.items():
            out["gw:ode_model." + k] = w.grad
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
