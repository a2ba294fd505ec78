"""Shared test helpers: golden case -> kernel-layout arrays, slot maps, IWAE upstream gradients (numpy)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np

from vihds_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MODEL_IDS = {"dr_constant": 0, "dr_constant_v2": 1, "dr_constant_precisions": 2, "dr_constant_precisions_v2": 3,
             "relay_constant": 4, "relay_constant_precisions": 5, "dr_blackbox": 6, "auto_constant": 7,
             "auto_constant_precisions": 8, "prpr_constant": 9, "prpr_constant_precisions": 10, "inducer_constant": 11,
             "inducer_constant_precisions": 12, "degrader_constant": 13, "degrader_constant_precisions": 14}
SOLVER_IDS = {"euler": 0, "midpoint": 1, "rk4": 2, "modeuler": 3, "modeulerwhile": 4}


def clip_bounds(case, stddevs=4.0):
    """Prior mu +- 4 sigma (exp'd for LogNormal), computed in the case dtype like the reference does."""
    dt = case["p_mu"].dtype
    lo = (case["p_mu"] - dt.type(stddevs) * case["p_sigma"]).astype(dt)
    hi = (case["p_mu"] + dt.type(stddevs) * case["p_sigma"]).astype(dt)
    ln = case["kinds"] == 2
    lo = np.where(ln, np.exp(lo), lo).astype(dt)
    hi = np.where(ln, np.exp(hi), hi).astype(dt)
    const = case["kinds"] == 0
    lo[const], hi[const] = -np.inf, np.inf
    return lo, hi


def bb_latent_names(params):
    return (["z%d" % (i + 1) for i in range(params["n_z"])] + ["x%d" % (i + 1) for i in range(params["n_x"])] +
            ["y%d" % (i + 1) for i in range(params["n_y"])])


def slot_map(case, slot_names):
    """slot_src for a golden case: theta column if the name is sampled, else an extra row (conditioned aR/aS).
    dr_blackbox: slots 4.. are the latent parameters z, x, y; the extra rows are the device offsets added to y."""
    names = [str(n) for n in case["names"]]
    if str(case["model"]) == "dr_blackbox":
        par = case["params"]
        src = [L.VH_SLOT_UNUSED] * L.VH_MAX_SLOTS
        for s, nm in enumerate(["init_x", "init_rfp", "init_yfp", "init_cfp"] + bb_latent_names(par)):
            src[s] = names.index(nm)
        W, b = case["w:ode_model.offset_layer.weight"], case["w:ode_model.offset_layer.bias"]
        off = case["dev_1hot"].astype(np.float64) @ W.astype(np.float64).T + b  # [B, n_y]
        IW = case["u"].shape[1]
        extra = np.stack([np.repeat(off[:, k], IW) for k in range(par["n_y"])]).astype(case["u"].dtype)
        return src, extra
    src = [L.VH_SLOT_UNUSED] * L.VH_MAX_SLOTS
    extras = []
    for s, name in enumerate(slot_names):
        if name in names:
            src[s] = names.index(name)
        elif ("cond_" + name) in case:
            src[s] = -1 - len(extras)
            extras.append(case["cond_" + name].reshape(-1))
    extra = np.stack(extras).astype(case["u"].dtype) if extras else None
    return src, extra


BB_LAYERS = ["neural_states.states_hidden", "neural_states.states_production", "neural_states.states_degradation",
             "precisions.prec_hidden", "precisions.prec_production", "precisions.prec_degradation"]


def flat_weights(case):
    """Flat decoder weight vector in the library layout (LinPrecNet: Wp, bp, Wd, bd; dr_blackbox: vh_bb.cuh)."""
    if str(case["model"]) == "dr_blackbox":
        keys = ["%s.%s" % (l, t) for l in BB_LAYERS for t in ("weight", "bias")]
        w = np.concatenate([case["w:ode_model." + k].reshape(-1) for k in keys])
        gw = np.concatenate([case["gw:ode_model." + k].reshape(-1) for k in keys])
        return w, gw
    pre = "w:ode_model.precisions."
    if (pre + "prec_production.weight") not in case:
        return None, None
    order = ["prec_production.weight", "prec_production.bias", "prec_degradation.weight", "prec_degradation.bias"]
    if (pre + "prec_hidden.weight") in case:  # NeuralPrecisions with a hidden layer (HidPrecNet layout)
        order = ["prec_hidden.weight", "prec_hidden.bias"] + order
    w = np.concatenate([case[pre + k].reshape(-1) for k in order])
    gw = np.concatenate([case["gw:ode_model.precisions." + k].reshape(-1) for k in order])
    return w, gw


def iwae_upstream(lpx, lp, lq, B, IW):
    """Gradients of the IWAE cost (vihds/training.py:134-148) w.r.t. the three per-sample term arrays."""
    lw = lpx.reshape(B, IW, 4).astype(np.float64).sum(2) + lp.reshape(B, IW) - lq.reshape(B, IW)
    m = lw.max(1, keepdims=True)
    e = np.exp(lw - m)
    w = e / e.sum(1, keepdims=True)
    lse = m[:, 0] + np.log(e.sum(1))
    cost = -(lse - math.log(IW)).mean()
    g = (-w / B).reshape(-1)
    return cost, g


def build_hostcheck():
    so = os.path.join(ROOT, "tests", "hostcheck", "_hostcheck.so")
    src = os.path.join(ROOT, "tests", "hostcheck", "vh_hostcheck.cpp")
    deps = [src] + [os.path.join(ROOT, "vihds_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "vihds_b200", "csrc"))
                    if f.endswith(".cuh")] + [os.path.join(ROOT, "include", "vihds_b200.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)
    lib.hc_fwd.argtypes = [C.POINTER(L.vh_problem), C.POINTER(L.vh_fwd_io)]
    lib.hc_bwd.argtypes = [C.POINTER(L.vh_problem), C.POINTER(L.vh_bwd_io)]
    lib.hc_slot_name.argtypes = [C.c_int, C.c_int]
    lib.hc_slot_name.restype = C.c_char_p
    return lib


def make_problem(case, src, E, n_weights_hidden=0):
    p = L.vh_problem()
    p.model = MODEL_IDS[str(case["model"])]
    p.solver = SOLVER_IDS[str(case["solver"])]
    p.dtype = L.VH_F64 if str(case["dtype"]) == "float64" else L.VH_F32
    B, IW, P = case["u"].shape
    p.B, p.IW, p.T, p.P = B, IW, len(case["times"]), P
    p.C, p.D, p.E = case["inputs"].shape[1], case["dev_1hot"].shape[1], E
    for s in range(L.VH_MAX_SLOTS):
        p.slot_src[s] = src[s]
    if "w:ode_model.precisions.prec_hidden.weight" in case and str(case["model"]) != "dr_blackbox":
        p.n_hidden = case["w:ode_model.precisions.prec_hidden.weight"].shape[0]  # --precision_hidden_layers of the run
    if str(case["model"]) == "dr_blackbox":
        par = case["params"]
        p.n_hidden, p.n_hidden_states, p.n_latent = par["n_hidden_decoder_precisions"], par["n_hidden_decoder"], par["n_latent_species"]
        p.n_z, p.n_x, p.n_y = par["n_z"], par["n_x"], par["n_y"]
        p.init_latent_species, p.init_prec = par.get("init_latent_species", 0.001), par.get("init_prec", 0.00001)
    return p


def offset_layer_grads(case, d_extra):
    """Gradient of the black-box offset layer from the kernel's d_extra [n_y][N] (chain rule done by torch in the product)."""
    B, IW, _ = case["u"].shape
    d_off = d_extra.reshape(d_extra.shape[0], B, IW).sum(2).T.astype(np.float64)  # [B, n_y]
    return d_off.T @ case["dev_1hot"].astype(np.float64), d_off.sum(0)
