"""-m gpu: the CUDA kernels, called through the C ABI (ctypes), against the golden vectors minted from the reference
and against the CPU oracle on the same inputs.  Tolerances: fp32 <= 1e-4 relative on per-timepoint states (relative
to each species' peak) and on the ELBO terms / cost (BASELINE.json north_star); gradients <= 1e-4 of the largest
component per tensor (measured: 2e-6 white-box, 3e-5 black-box); fp64 cases much tighter.  Every golden case runs through
BOTH forms of the white-box kernels: the default pick (latency forms at these sizes: team prologue, producer / consumer
reverse sweep) and the throughput forms that large batches take (VIHDS_FWD_TEAM=0, VIHDS_BWD_WS=0)."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import golden_cases, load_case
import helpers as H
from vihds_b200 import _lib as L

pytestmark = pytest.mark.gpu

DR_CASES = golden_cases()


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def _dev(a, dt=None):
    if a is None:
        return None
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dt is not None and t.is_floating_point():
        t = t.to(dt)
    return t.cuda().contiguous()


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def run_case_on_gpu(case):
    lib = L.load()
    f64 = str(case["dtype"]) == "float64"
    dt = torch.float64 if f64 else torch.float32
    model = H.MODEL_IDS[str(case["model"])]
    src, extra = H.slot_map(case, L.slot_names(model))
    w, gw_ref = H.flat_weights(case)
    p = H.make_problem(case, src, 0 if extra is None else extra.shape[0])
    B, IW, P, T = p.B, p.IW, p.P, p.T
    N = B * IW
    S = lib.vh_state_width(C.byref(p))
    lo, hi = H.clip_bounds(case)
    dev = dict(
        times=_dev(case["times"], dt), u=_dev(case["u"].reshape(N, P), dt), q_mu=_dev(case["q_mu"], dt),
        q_prec=_dev(case["q_prec"], dt), p_mu=_dev(case["p_mu"], dt), p_prec=_dev(case["p_prec"], dt), clip_lo=_dev(lo, dt),
        clip_hi=_dev(hi, dt), kind=_dev(case["kinds"].astype(np.int32)), extra=_dev(extra, dt),
        treatments=_dev(case["inputs"], dt), dev_1hot=_dev(case["dev_1hot"], dt), observations=_dev(case["observations"], dt),
        weights=_dev(w, dt),
        theta=torch.zeros(P, N, dtype=dt, device="cuda"), x_states=torch.zeros(T, S, N, dtype=dt, device="cuda"),
        x_predict=torch.zeros(T, 4, N, dtype=dt, device="cuda"), logp_by_species=torch.zeros(N, 4, dtype=dt, device="cuda"),
        logp_theta=torch.zeros(N, dtype=dt, device="cuda"), logq_theta=torch.zeros(N, dtype=dt, device="cuda"))
    io = L.vh_fwd_io(**{k: _p(v) for k, v in dev.items()})
    L.check(lib.vh_elbo_terms_fwd(C.byref(p), C.byref(io), None))
    cost = torch.zeros(1, dtype=dt, device="cuda")
    log_w = torch.zeros(N, dtype=dt, device="cuda")
    wts = torch.zeros(N, dtype=dt, device="cuda")
    L.check(lib.vh_iwae_fwd(p.dtype, B, IW, B, _p(dev["logp_by_species"]), _p(dev["logp_theta"]), _p(dev["logq_theta"]),
                            _p(cost), _p(log_w), _p(wts), None))
    g = dict(g_logp_by_species=torch.zeros(N, 4, dtype=dt, device="cuda"), g_logp_theta=torch.zeros(N, dtype=dt, device="cuda"),
             g_logq_theta=torch.zeros(N, dtype=dt, device="cuda"))
    L.check(lib.vh_iwae_bwd(p.dtype, B, IW, B, _p(wts), None, _p(g["g_logp_by_species"]), _p(g["g_logp_theta"]),
                            _p(g["g_logq_theta"]), None))
    out = dict(d_q_mu=torch.full((B, P), 7.0, dtype=dt, device="cuda"), d_q_prec=torch.full((B, P), 7.0, dtype=dt, device="cuda"),
               d_weights=None if w is None else torch.full((len(w),), 7.0, dtype=dt, device="cuda"),
               d_extra=None if extra is None else torch.zeros(extra.shape, dtype=dt, device="cuda"))
    bio = L.vh_bwd_io(fwd=io, **{k: _p(v) for k, v in {**g, **out}.items()})
    L.check(lib.vh_elbo_terms_bwd(C.byref(p), C.byref(bio), None))
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in {**dev, **out, "cost": cost, "log_w": log_w, "w": wts}.items() if v is not None}
    res["dims"] = (B, IW, P, T, S)
    res["gw_ref"] = gw_ref
    return res


@pytest.mark.parametrize("form", ["default", "throughput", "latency_ws"])
@pytest.mark.parametrize("name", DR_CASES)
def test_cuda_matches_reference_golden(name, form, monkeypatch):
    case = load_case(name)
    if form == "latency_ws":
        # dr_constant + midpoint in fp32 takes the matrix-form reverse kernel by default (vh_bwd_mx.cuh); this form keeps the
        # producer / consumer kernel it replaced under the same goldens
        if not (str(case["model"]).startswith("dr_constant") and "precisions" not in str(case["model"]) and str(case["solver"]) == "midpoint" and str(case["dtype"]) == "float32"):
            pytest.skip("same kernels as the default form")
        monkeypatch.setenv("VIHDS_BWD_MX", "0")
        monkeypatch.setenv("VIHDS_FWD_SCRIBE", "0")  # ... and the one-warp time loop of the team forward kernel
    if form == "throughput":
        if str(case["model"]) == "dr_blackbox":
            pytest.skip("the black-box kernels have one form per implementation (tests/test_gpu_bb_mma.py compares those)")
        monkeypatch.setenv("VIHDS_FWD_TEAM", "0")
        monkeypatch.setenv("VIHDS_BWD_WS", "0")
    r = run_case_on_gpu(case)
    B, IW, P, T, S = r["dims"]
    f64 = str(case["dtype"]) == "float64"
    tol = 1e-9 if f64 else 1e-4  # north-star tolerance for fp32
    assert _rel(r["theta"].reshape(P, B, IW), case["theta"]) < tol
    xs = r["x_states"].reshape(T, S, B, IW).transpose(2, 3, 1, 0)
    ns = case["x_states_last"].shape[2] if "x_states_last" in case else case["x_states"].shape[2]
    if "x_states" in case:
        for s in range(ns):
            assert _rel(xs[:, :, s], case["x_states"][:, :, s]) < tol, "species %d" % s
        xp = r["x_predict"].reshape(T, 4, B, IW).transpose(2, 3, 1, 0)
        assert _rel(xp, case["x_predict"]) < tol
        if S > ns:
            assert _rel(xs[:, :, ns:], case["precisions"]) < tol
    else:
        assert _rel(xs[:, :, :ns, -1], case["x_states_last"]) < tol
    assert _rel(r["logp_by_species"].reshape(B, IW, 4), case["log_p_by_species"]) < tol
    assert _rel(r["logp_theta"].reshape(B, IW), case["log_p_theta"]) < tol
    assert _rel(r["logq_theta"].reshape(B, IW), case["log_q_theta"]) < tol
    assert abs(float(r["cost"][0]) - float(case["loss"])) <= tol * abs(float(case["loss"]))
    gtol = 1e-6 if f64 else 1e-4
    per_ind = case["per_individual"].astype(bool)
    sel = case["kinds"] != 0
    for got, ref in ((r["d_q_mu"], case["grad_q_mu"]), (r["d_q_prec"], case["grad_q_prec"])):
        ref_tot = np.where(per_ind, ref.sum(0), ref[0])
        tot = got.sum(0)
        assert _rel(tot[sel], ref_tot[sel]) < gtol
        for k in np.nonzero(sel)[0]:
            assert abs(tot[k] - ref_tot[k]) <= 10 * gtol * abs(ref_tot[k]) + 1e-6 * np.max(np.abs(ref_tot)) * gtol, (
                "column %d (%s)" % (k, case["names"][k]))
        if per_ind.any():
            assert _rel(got[:, per_ind], ref[:, per_ind]) < gtol
    if r["gw_ref"] is not None:
        assert _rel(r["d_weights"], r["gw_ref"]) < gtol
    if str(case["model"]) == "dr_blackbox":
        dW, db = H.offset_layer_grads(case, r["d_extra"])
        assert _rel(dW, case["gw:ode_model.offset_layer.weight"]) < gtol
        assert _rel(db, case["gw:ode_model.offset_layer.bias"]) < gtol


@pytest.mark.parametrize("name", ["dr_constant_icml_midpoint_f32_iw8", "dr_constant_one_modeuler_f32_iw5"])
def test_cuda_matches_oracle(name):
    """Same inputs through the CPU oracle (not the stored outputs): catches a stale fixture as well as a kernel bug."""
    import vihds_oracle as O

    case = load_case(name)
    ref = O.elbo_step(case)
    r = run_case_on_gpu(case)
    B, IW, P, T, S = r["dims"]
    xs = r["x_states"].reshape(T, S, B, IW).transpose(2, 3, 1, 0)
    for s in range(S):
        assert _rel(xs[:, :, s], ref["x_states"][:, :, s].numpy()) < 1e-4
    assert _rel(r["log_w"].reshape(B, IW), ref["log_w"].numpy()) < 1e-4
    assert abs(float(r["cost"][0]) - float(ref["loss"])) <= 1e-4 * abs(float(ref["loss"]))
    assert _rel(r["d_q_mu"], ref["grad_q_mu"].numpy()) < 1e-4
    assert _rel(r["d_q_prec"], ref["grad_q_prec"].numpy()) < 1e-4


@pytest.mark.parametrize("form", ["default", "forced_latency"])
def test_throughput_regime_matches_oracle(form, monkeypatch):
    """N = 32,768 trajectories (4 individuals x 8,192 samples): the launcher picks the throughput-form kernels
    (block >= 64) by itself -- every output and the FULL gradient (all columns) against the CPU oracle on the same inputs;
    then the latency forms forced onto the same batch."""
    import vihds_oracle as O

    if form == "forced_latency":
        monkeypatch.setenv("VIHDS_FWD_TEAM", "1")
        monkeypatch.setenv("VIHDS_BWD_WS", "1")
    base = load_case("dr_constant_icml_midpoint_f32_iw8")
    B, IW = 4, 8192
    rng = np.random.RandomState(3)
    case = dict(base)
    for k in ("inputs", "dev_1hot", "observations", "q_mu", "q_prec"):
        case[k] = np.ascontiguousarray(base[k][:B])
    case["u"] = rng.randn(B, IW, base["u"].shape[2]).astype(np.float32)
    case["cond_aR"] = (1 + np.abs(rng.randn(B, IW))).astype(np.float32)
    case["cond_aS"] = (1 + np.abs(rng.randn(B, IW))).astype(np.float32)
    ref = O.elbo_step(case)
    r = run_case_on_gpu(case)
    _, _, P, T, S = r["dims"]
    xs = r["x_states"].reshape(T, S, B, IW).transpose(2, 3, 1, 0)
    for s in range(S):
        assert _rel(xs[:, :, s], ref["x_states"][:, :, s].numpy()) < 1e-4, "state %d" % s
    assert _rel(r["theta"].reshape(P, B, IW), ref["theta"].numpy()) < 1e-5
    assert _rel(r["log_w"].reshape(B, IW), ref["log_w"].numpy()) < 1e-4
    assert abs(float(r["cost"][0]) - float(ref["loss"])) <= 1e-4 * abs(float(ref["loss"]))
    sel = case["kinds"] != 0
    assert _rel(r["d_q_mu"][:, sel], ref["grad_q_mu"].numpy()[:, sel]) < 1e-4
    assert _rel(r["d_q_prec"][:, sel], ref["grad_q_prec"].numpy()[:, sel]) < 1e-4


def test_simulate_seam_equals_fused():
    """vh_simulate (theta given as extra rows, P == 0) must reproduce the fused call's trace bit for bit."""
    lib = L.load()
    case = load_case("dr_constant_icml_midpoint_f32_iw8")
    r = run_case_on_gpu(case)
    B, IW, P, T, S = r["dims"]
    N = B * IW
    names = [str(n) for n in case["names"]]
    slot_names = L.slot_names(0)
    rows, src = [], [L.VH_SLOT_UNUSED] * L.VH_MAX_SLOTS
    for s, nm in enumerate(slot_names):
        if nm in names:
            src[s] = -1 - len(rows)
            rows.append(r["theta"][names.index(nm)])
        elif "cond_" + nm in case:
            src[s] = -1 - len(rows)
            rows.append(case["cond_" + nm].reshape(-1))
    extra = torch.as_tensor(np.stack(rows)).cuda()
    case2 = dict(case)
    p = H.make_problem(case2, src, extra.shape[0])
    p.P = 0
    xs = torch.zeros(T, S, N, device="cuda")
    times, tr = torch.as_tensor(case["times"]).cuda(), torch.as_tensor(case["inputs"]).cuda()
    io = L.vh_fwd_io(times=_p(times), extra=_p(extra), treatments=_p(tr), x_states=_p(xs))
    L.check(lib.vh_simulate(C.byref(p), C.byref(io), None))
    torch.cuda.synchronize()
    assert np.array_equal(xs.cpu().numpy(), r["x_states"])


def test_adaptive_solver_is_refused():
    lib = L.load()
    assert lib.vh_solver_id(b"dopri5") < 0
    with pytest.raises(NotImplementedError):
        L.solver_id("dopri5")


def test_fused_adam_matches_torch():
    lib = L.load()
    torch.manual_seed(0)
    p0 = torch.randn(4097, device="cuda")
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=0.01)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for step in range(1, 6):
        g = torch.randn_like(p0)
        ref.grad = g.clone()
        opt.step()
        L.check(lib.vh_adam_step(0, p.numel(), _p(p), _p(g), _p(m), _p(v), 0.01, 0.9, 0.999, 1e-8, step, None))
    torch.cuda.synchronize()
    assert torch.allclose(p, ref.detach(), rtol=1e-5, atol=1e-6)
