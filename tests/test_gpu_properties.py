"""-m gpu: edge cases and size-independent properties of the CUDA path (through the C ABI).

* edge shapes the reference's tests never reach: T = 2, IW = 1, B = 1, partial warps, warps straddling individuals --
  each against the CPU oracle on the same inputs;
* NaN / Inf contract (vihds/training.py:331: a NaN ELBO must surface, never be clamped away);
* at the synthetic full size (B = 1024 x IW = 128 = 131,072 trajectories, T = 500: one GPU's slab of BASELINE config 4)
  properties that need no oracle: batch-composition invariance (a trajectory's result does not depend on what else is in
  the launch), bit-exact determinism of the forward launch, and gradient = sum of sub-batch gradients (linearity of the
  reverse sweep in the upstream gradient);
* the reverse sweep against central finite differences of the forward launch in fp64.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_case
import helpers as H
import vihds_oracle as O
from test_gpu_parity import _dev, _p, _rel, run_case_on_gpu
from vihds_b200 import _lib as L

pytestmark = pytest.mark.gpu


def sub_case(case, b_idx, iw=None, t=None):
    """Slice a golden case down to individuals b_idx, the first iw samples and the first t time points."""
    out = dict(case)
    IW = case["u"].shape[1] if iw is None else iw
    for k in ("u", "inputs", "dev_1hot", "observations", "q_mu", "q_prec", "cond_aR", "cond_aS"):
        if k in out:
            out[k] = np.ascontiguousarray(case[k][b_idx])
    out["u"] = np.ascontiguousarray(out["u"][:, :IW])
    for k in ("cond_aR", "cond_aS"):
        if k in out:
            out[k] = np.ascontiguousarray(out[k][:, :IW])
    if t is not None:
        out["times"] = np.ascontiguousarray(case["times"][:t])
        out["observations"] = np.ascontiguousarray(out["observations"][:, :, :t])
    return out


def check_against_oracle(case, tol=1e-4, gtol=3e-3):
    ref = O.elbo_step(case)
    r = run_case_on_gpu(case)
    B, IW, P, T, S = r["dims"]
    xs = r["x_states"].reshape(T, S, B, IW).transpose(2, 3, 1, 0)
    full = torch.cat([ref["x_states"], ref["precisions"]], 2).numpy() if S > ref["x_states"].shape[2] else ref["x_states"].numpy()
    for s in range(S):
        assert _rel(xs[:, :, s], full[:, :, s]) < tol, "state %d" % s
    assert _rel(r["log_w"].reshape(B, IW), ref["log_w"].numpy()) < tol
    assert abs(float(r["cost"][0]) - float(ref["loss"])) <= tol * abs(float(ref["loss"]))
    assert _rel(r["d_q_mu"], ref["grad_q_mu"].numpy()) < gtol
    assert _rel(r["d_q_prec"], ref["grad_q_prec"].numpy()) < gtol
    return r


@pytest.mark.parametrize("b_idx,iw,t", [
    ([0], 1, 2),            # one trajectory, one step
    ([0, 1, 2], 1, 5),      # IW = 1: logsumexp over a single sample
    ([3], 8, 86),           # B = 1
    (list(range(11)), 3, 30),  # N = 33: a partial second warp, every warp straddles several individuals
    (list(range(36)), 7, 10),  # N = 252: warps straddle individuals at odd offsets
])
def test_edge_shapes_match_oracle(b_idx, iw, t):
    case = load_case("dr_constant_icml_midpoint_f32_iw8")
    check_against_oracle(sub_case(case, b_idx, iw, t))


def test_edge_shapes_dynamic_precisions_and_relay():
    case = load_case("relay_constant_precisions_midpoint_f32_iw8")
    r = check_against_oracle(sub_case(case, [0, 1, 2, 3, 4], 7, 12))  # N = 35
    ref = O.elbo_step(sub_case(case, [0, 1, 2, 3, 4], 7, 12))
    _, gw = H.flat_weights(case)
    order = ["prec_production.weight", "prec_production.bias", "prec_degradation.weight", "prec_degradation.bias"]
    gw_ref = np.concatenate([ref["gw:ode_model.precisions." + k].numpy().reshape(-1) for k in order])
    assert _rel(r["d_weights"], gw_ref) < 3e-3


@pytest.mark.parametrize("wgrad", ["scalar", "mma"])
def test_edge_shapes_blackbox(wgrad, monkeypatch):
    """Both weight-gradient sinks of the black-box reverse kernel (scalar FMAs / mma.sync 3xTF32 on the tensor cores).
    The mode is latched on the first black-box launch of a process, hence the subprocess for the non-default one."""
    if wgrad == "mma":
        import os
        import subprocess
        import sys

        env = dict(os.environ, VIHDS_BB_WGRAD="mma")
        code = ("import sys; sys.path[:0] = [%r, %r, %r]; import test_gpu_properties as t; "
                "t._blackbox_edge_check()" % (H.ROOT, os.path.join(H.ROOT, "tests"), os.path.join(H.ROOT, "oracle")))
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        return
    _blackbox_edge_check()


def _blackbox_edge_check():
    case = load_case("dr_blackbox_icml_midpoint_f32_iw8")
    sub = sub_case(case, [0, 1, 2, 3, 4], 7, 9)  # N = 35: the warp-level weight-gradient GEMM with idle lanes
    ref = O.elbo_step(sub)
    r = run_case_on_gpu(sub)
    B, IW, P, T, S = r["dims"]
    assert abs(float(r["cost"][0]) - float(ref["loss"])) <= 1e-4 * abs(float(ref["loss"]))
    keys = ["%s.%s" % (l, t) for l in H.BB_LAYERS for t in ("weight", "bias")]
    gw_ref = np.concatenate([ref["gw:ode_model." + k].numpy().reshape(-1) for k in keys])
    assert _rel(r["d_weights"], gw_ref) < 3e-3
    dW, db = H.offset_layer_grads(sub, r["d_extra"])
    assert _rel(dW, ref["gw:ode_model.offset_layer.weight"].numpy()) < 3e-3


def test_nan_and_inf_surface_in_the_cost():
    case = sub_case(load_case("dr_constant_icml_midpoint_f32_iw8"), [0, 1, 2], 8, 20)
    bad = dict(case)
    bad["u"] = case["u"].copy()
    bad["u"][1, 3, 0] = np.nan
    r = run_case_on_gpu(bad)
    assert np.isnan(r["cost"][0]) and np.isnan(r["log_w"].reshape(3, 8)[1, 3])
    assert np.isfinite(r["log_w"].reshape(3, 8)[0]).all()  # other individuals are untouched
    bad = dict(case)
    bad["observations"] = case["observations"].copy()
    bad["observations"][2, 1, 5] = np.inf
    r = run_case_on_gpu(bad)
    assert not np.isfinite(r["cost"][0])


def _synthetic(B, IW, T, seed=0):
    """One GPU's slab of the synthetic config, built from the icml golden case (tiled individuals, fresh u)."""
    case = load_case("dr_constant_icml_midpoint_f32_iw8")
    B0 = case["u"].shape[0]
    rng = np.random.RandomState(seed)
    idx = np.arange(B) % B0
    times = np.linspace(0, 16.5, T).astype(np.float32)
    obs = np.stack([np.stack([np.interp(times, case["times"], case["observations"][b, o]) for o in range(4)]) for b in range(B0)])
    big = dict(case)
    big.update(times=times, observations=obs[idx].astype(np.float32), u=rng.randn(B, IW, case["u"].shape[2]).astype(np.float32),
               cond_aR=(1 + np.abs(rng.randn(B, IW))).astype(np.float32), cond_aS=(1 + np.abs(rng.randn(B, IW))).astype(np.float32))
    for k in ("q_mu", "q_prec", "inputs", "dev_1hot"):
        big[k] = np.ascontiguousarray(case[k][idx])
    return big


def test_full_size_batch_composition_invariance_and_determinism():
    """131,072 trajectories x 500 time points: individuals 5, 700 and 1023 give bit-identical traces and per-sample
    terms whether they run inside the full launch or as a 3-individual launch; two full launches are bit-identical."""
    big = _synthetic(1024, 128, 500)
    r1 = run_case_on_gpu(big)
    B, IW, P, T, S = r1["dims"]
    assert np.isfinite(r1["cost"][0])
    r2 = run_case_on_gpu(big)
    assert np.array_equal(r1["x_states"], r2["x_states"]) and np.array_equal(r1["logp_by_species"], r2["logp_by_species"])
    pick = [5, 700, 1023]
    small = sub_case(big, pick)
    rs = run_case_on_gpu(small)
    xs_big = r1["x_states"].reshape(T, S, B, IW)[:, :, pick]
    assert np.array_equal(xs_big, rs["x_states"].reshape(T, S, 3, IW))
    assert np.array_equal(r1["logp_by_species"].reshape(B, IW, 4)[pick], rs["logp_by_species"].reshape(3, IW, 4))
    assert np.array_equal(r1["theta"].reshape(P, B, IW)[:, pick], rs["theta"].reshape(P, 3, IW))
    # per-individual gradient rows only depend on the individual's own trajectories (and on b_total via the IWAE weights)
    per_ind = big["per_individual"].astype(bool)
    assert _rel(r1["d_q_mu"][pick][:, per_ind] * (B / 3.0), rs["d_q_mu"][:, per_ind]) < 1e-4


def test_reverse_sweep_is_linear_in_the_upstream_gradient():
    """d_q(g1 + g2) = d_q(g1) + d_q(g2): checked at T = 500 on 4,096 trajectories with random upstream gradients."""
    lib = L.load()
    big = _synthetic(32, 128, 500, seed=1)
    model = H.MODEL_IDS[str(big["model"])]
    src, extra = H.slot_map(big, L.slot_names(model))
    p = H.make_problem(big, src, extra.shape[0])
    B, IW, P, T = p.B, p.IW, p.P, p.T
    N = B * IW
    lo, hi = H.clip_bounds(big)
    dt = torch.float32
    dev = dict(times=_dev(big["times"]), u=_dev(big["u"].reshape(N, P)), q_mu=_dev(big["q_mu"]), q_prec=_dev(big["q_prec"]),
               p_mu=_dev(big["p_mu"]), p_prec=_dev(big["p_prec"]), clip_lo=_dev(lo), clip_hi=_dev(hi),
               kind=_dev(big["kinds"].astype(np.int32)), extra=_dev(extra), treatments=_dev(big["inputs"]),
               dev_1hot=_dev(big["dev_1hot"]), observations=_dev(big["observations"]),
               x_states=torch.zeros(T, 8, N, device="cuda"), logp_by_species=torch.zeros(N, 4, device="cuda"),
               logp_theta=torch.zeros(N, device="cuda"), logq_theta=torch.zeros(N, device="cuda"))
    io = L.vh_fwd_io(**{k: _p(v) for k, v in dev.items()})
    L.check(lib.vh_elbo_terms_fwd(C.byref(p), C.byref(io), None))
    g = torch.Generator(device="cuda").manual_seed(0)

    def sweep(glpx, glp, glq):
        d_mu, d_prec = torch.zeros(B, P, device="cuda"), torch.zeros(B, P, device="cuda")
        bio = L.vh_bwd_io(fwd=io, g_logp_by_species=_p(glpx), g_logp_theta=_p(glp), g_logq_theta=_p(glq), d_q_mu=_p(d_mu),
                          d_q_prec=_p(d_prec))
        L.check(lib.vh_elbo_terms_bwd(C.byref(p), C.byref(bio), None))
        torch.cuda.synchronize()
        return d_mu.double(), d_prec.double()

    a = [torch.randn(N, 4, device="cuda", generator=g) / N, torch.randn(N, device="cuda", generator=g) / N, torch.randn(N, device="cuda", generator=g) / N]
    b = [torch.randn(N, 4, device="cuda", generator=g) / N, torch.randn(N, device="cuda", generator=g) / N, torch.randn(N, device="cuda", generator=g) / N]
    da, db, dab = sweep(*a), sweep(*b), sweep(*[x + y for x, y in zip(a, b)])
    for k in range(2):
        ref = (da[k] + db[k]).cpu().numpy()
        assert _rel(dab[k].cpu().numpy(), ref) < 1e-4


@pytest.mark.parametrize("name", ["dr_constant_one_midpoint_f64_iw5", "relay_constant_precisions_midpoint_f64_iw8"])
def test_reverse_sweep_matches_finite_differences_f64(name):
    """Central differences of the fp64 forward launch (cost) against the reverse sweep, for a handful of q entries."""
    full = load_case(name)
    case = sub_case(full, [0, 1], 3, 15)
    base = run_case_on_gpu(case)
    rng = np.random.RandomState(0)
    cols = [k for k in range(case["q_mu"].shape[1]) if case["kinds"][k] != 0]
    for k in rng.choice(cols, size=6, replace=False):
        for key, grad in (("q_mu", base["d_q_mu"]), ("q_prec", base["d_q_prec"])):
            b = int(rng.randint(0, 2))
            eps = 1e-6 * max(1.0, abs(case[key][b, k]))
            costs = []
            for sgn in (+1, -1):
                c2 = dict(case)
                c2[key] = case[key].copy()
                c2[key][b, k] += sgn * eps
                costs.append(float(run_case_on_gpu(c2)["cost"][0]))
            fd = (costs[0] - costs[1]) / (2 * eps)
            assert abs(fd - grad[b, k]) <= 2e-5 * max(abs(fd), np.abs(grad).max() * 1e-3) + 1e-9, (key, b, k, fd, grad[b, k])


def test_iwae_kernel_edges():
    lib = L.load()
    for B, IW in ((1, 1), (3, 1000), (2, 33)):
        g = torch.Generator().manual_seed(B * 1000 + IW)
        lpx = torch.randn(B * IW, 4, generator=g).cuda() * 50
        lp, lq = torch.randn(B * IW, generator=g).cuda() * 10, torch.randn(B * IW, generator=g).cuda() * 10
        cost, log_w, w = torch.zeros(1, device="cuda"), torch.zeros(B * IW, device="cuda"), torch.zeros(B * IW, device="cuda")
        L.check(lib.vh_iwae_fwd(0, B, IW, B, _p(lpx), _p(lp), _p(lq), _p(cost), _p(log_w), _p(w), None))
        lw = (lpx.double().sum(1) + lp.double() - lq.double()).view(B, IW)
        ref = -(torch.logsumexp(lw, 1) - np.log(IW)).mean()
        assert abs(float(cost) - float(ref)) <= 1e-5 * abs(float(ref)) + 1e-5
        assert torch.allclose(w.view(B, IW).sum(1).cpu(), torch.ones(B), atol=1e-5)
    # a row whose every sample has log w = -inf: logsumexp = -inf, cost = +inf (torch.logsumexp semantics), no NaN
    lpx = torch.full((4, 4), -float("inf"), device="cuda")
    z = torch.zeros(4, device="cuda")
    cost = torch.zeros(1, device="cuda")
    L.check(lib.vh_iwae_fwd(0, 1, 4, 1, _p(lpx), _p(z), _p(z), _p(cost), None, None, None))
    assert float(cost) == float("inf")


def test_iwae_fused_matches_the_two_launches():
    """vh_iwae_fwd_bwd (one launch, the training step's entry) against vh_iwae_fwd + vh_iwae_bwd with g = 1: fp32 and
    fp64, IW below / above the block size, outputs overwritten (not accumulated), NaN contract."""
    lib = L.load()
    for vdt, dt in ((0, torch.float32), (1, torch.float64)):
        for B, IW in ((1, 1), (3, 1000), (36, 200), (70, 33)):
            g = torch.Generator().manual_seed(B * 1000 + IW)
            lpx = (torch.randn(B * IW, 4, generator=g, dtype=dt) * 50).cuda()
            lp, lq = (torch.randn(B * IW, generator=g, dtype=dt) * 10).cuda(), (torch.randn(B * IW, generator=g, dtype=dt) * 10).cuda()
            z = lambda *s: torch.full(s, 7.0, device="cuda", dtype=dt)  # noqa: E731  (outputs must be overwritten)
            c0, lw0, w0, g0 = z(1), z(B * IW), z(B * IW), (z(B * IW, 4), z(B * IW), z(B * IW))
            c1, lw1, w1, g1 = z(1), z(B * IW), z(B * IW), (z(B * IW, 4), z(B * IW), z(B * IW))
            L.check(lib.vh_iwae_fwd(vdt, B, IW, B + 1, _p(lpx), _p(lp), _p(lq), _p(c0), _p(lw0), _p(w0), None))
            L.check(lib.vh_iwae_bwd(vdt, B, IW, B + 1, _p(w0), None, _p(g0[0]), _p(g0[1]), _p(g0[2]), None))
            L.check(lib.vh_iwae_fwd_bwd(vdt, B, IW, B + 1, _p(lpx), _p(lp), _p(lq), _p(c1), _p(lw1), _p(w1), _p(g1[0]),
                                        _p(g1[1]), _p(g1[2]), None))
            tol = 1e-5 if dt == torch.float32 else 1e-12
            assert abs(float(c0) - float(c1)) <= tol * abs(float(c0)), (B, IW, float(c0), float(c1))
            assert torch.equal(lw0, lw1)
            for a_, b_ in ((w0, w1), (g0[0], g1[0]), (g0[1], g1[1]), (g0[2], g1[2])):
                assert torch.allclose(a_, b_, rtol=tol * 10, atol=1e-30), (B, IW)
    lpx = torch.zeros(8, 4, device="cuda")
    lpx[5, 2] = float("nan")
    zz, cost = torch.zeros(8, device="cuda"), torch.zeros(1, device="cuda")
    L.check(lib.vh_iwae_fwd_bwd(0, 2, 4, 2, _p(lpx), _p(zz), _p(zz), _p(cost), None, None, None, None, None, None))
    assert np.isnan(float(cost))


def test_adam_dev_counts_steps_and_clears_the_gradient():
    """vh_adam_step_dev: same update as vh_adam_step, the device-side step counter advances by one per call (bumped by
    the last thread block, ticket scratch back to zero) and zero_grad clears the consumed gradient."""
    lib = L.load()
    for n in (1, 255, 256, 100003):
        g = torch.Generator().manual_seed(n)
        p0 = torch.randn(n, generator=g).cuda()
        pa, pb = p0.clone(), p0.clone()
        ma, va, mb, vb = (torch.zeros(n, device="cuda") for _ in range(4))
        hyper = torch.tensor([0.01, 0.9, 0.999, 1e-8], dtype=torch.float64, device="cuda")
        step = torch.zeros(4, dtype=torch.int64, device="cuda")
        good = torch.tensor([1.5], device="cuda")
        for it in range(1, 4):
            gr = torch.randn(n, generator=g).cuda()
            ga = gr.clone()
            L.check(lib.vh_adam_step_dev(0, n, _p(pa), _p(ga), _p(ma), _p(va), _p(hyper), _p(step), it % 2,
                                         _p(good) if it == 2 else None, None))
            L.check(lib.vh_adam_step(0, n, _p(pb), _p(gr), _p(mb), _p(vb), 0.01, 0.9, 0.999, 1e-8, it, None))
            assert step.tolist() == [it, 0, 0, 0]
            assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-7)
            assert torch.equal(ga, torch.zeros_like(ga) if it % 2 else gr)


def test_adam_dev_nan_cost_leaves_parameters_and_moments_untouched():
    """The device-side guard of the graphed step (vihds/training.py:331-336 checks isnan BEFORE optimizer.step()): with a
    NaN cost -- and therefore NaN gradients -- the call must not touch parameters or moments, must not count as a step,
    must still clear the gradient and must count the refusal; the next good step continues with the old step count."""
    lib = L.load()
    n = 70001
    g = torch.Generator().manual_seed(5)
    pa = torch.randn(n, generator=g).cuda()
    ma, va = torch.rand(n, generator=g).cuda(), torch.rand(n, generator=g).cuda()
    p0, m0, v0 = pa.clone(), ma.clone(), va.clone()
    hyper = torch.tensor([0.01, 0.9, 0.999, 1e-8], dtype=torch.float64, device="cuda")
    step = torch.tensor([7, 0, 0, 0], dtype=torch.int64, device="cuda")
    bad = torch.tensor([float("nan")], device="cuda")
    ga = torch.full((n,), float("nan"), device="cuda")
    L.check(lib.vh_adam_step_dev(0, n, _p(pa), _p(ga), _p(ma), _p(va), _p(hyper), _p(step), 1, _p(bad), None))
    torch.cuda.synchronize()
    assert torch.equal(pa, p0) and torch.equal(ma, m0) and torch.equal(va, v0)
    assert step.tolist() == [7, 0, 1, 0] and not ga.any()
    # one-rank exchange kernel: same contract
    from vihds_b200.distributed import PeerGradientExchange
    ex = PeerGradientExchange(n, torch.float32, torch.device("cuda", 0))
    ga = torch.full((n,), float("nan"), device="cuda")
    L.check(lib.vh_adam_allreduce_step(0, n, _p(pa), _p(ga), _p(ma), _p(va), _p(hyper), _p(step), _p(ex.state), 0, 1,
                                       _p(ex.peers), _p(bad), 0.0, None))
    torch.cuda.synchronize()
    assert torch.equal(pa, p0) and torch.equal(ma, m0) and torch.equal(va, v0)
    assert step.tolist() == [7, 0, 2, 0] and ex.state.tolist() == [1, 0, 0, 1] and not ga.any()
    # the refusal is sticky (the reference stops training at the first NaN): a good cost does not re-open the update ...
    gr = torch.randn(n, generator=g).cuda()
    ga, pb, mb, vb = gr.clone(), p0.clone(), m0.clone(), v0.clone()
    good = torch.tensor([3.0], device="cuda")
    L.check(lib.vh_adam_allreduce_step(0, n, _p(pa), _p(ga), _p(ma), _p(va), _p(hyper), _p(step), _p(ex.state), 0, 1,
                                       _p(ex.peers), _p(good), 0.0, None))
    torch.cuda.synchronize()
    assert torch.equal(pa, p0) and step.tolist() == [7, 0, 3, 0] and ex.state.tolist() == [2, 0, 0, 2]
    # ... until the host clears the counters
    step[2] = 0
    ex.state[3] = 0
    ga = gr.clone()
    L.check(lib.vh_adam_allreduce_step(0, n, _p(pa), _p(ga), _p(ma), _p(va), _p(hyper), _p(step), _p(ex.state), 0, 1,
                                       _p(ex.peers), _p(good), 0.0, None))
    L.check(lib.vh_adam_step(0, n, _p(pb), _p(gr), _p(mb), _p(vb), 0.01, 0.9, 0.999, 1e-8, 8, None))
    torch.cuda.synchronize()
    assert step.tolist() == [8, 0, 0, 0] and torch.allclose(pa, pb, rtol=1e-6, atol=1e-7)
    ex.close()


def test_fused_allreduce_adam_single_rank_equals_adam():
    """vh_adam_allreduce_step with a one-rank exchange (the kernel pushes into its own inbox, publishes and waits on its
    own flag): same update as vh_adam_step, gradient cleared, epoch / step counters advance, inbox parity alternates."""
    from vihds_b200.distributed import PeerGradientExchange
    lib = L.load()
    for n in (1, 1000, 100003):
        g = torch.Generator().manual_seed(n)
        p0 = torch.randn(n, generator=g).cuda()
        pa, pb = p0.clone(), p0.clone()
        ma, va, mb, vb = (torch.zeros(n, device="cuda") for _ in range(4))
        hyper = torch.tensor([0.01, 0.9, 0.999, 1e-8], dtype=torch.float64, device="cuda")
        step = torch.zeros(4, dtype=torch.int64, device="cuda")
        ex = PeerGradientExchange(n, torch.float32, torch.device("cuda", 0))  # closed at the end of the loop body
        for it in range(1, 5):
            gr = torch.randn(n, generator=g).cuda()
            ga = gr.clone()
            L.check(lib.vh_adam_allreduce_step(0, n, _p(pa), _p(ga), _p(ma), _p(va), _p(hyper), _p(step), _p(ex.state),
                                               0, 1, _p(ex.peers), None, 0.0, None))
            L.check(lib.vh_adam_step(0, n, _p(pb), _p(gr), _p(mb), _p(vb), 0.01, 0.9, 0.999, 1e-8, it, None))
            assert step.tolist()[0] == it and ex.state.tolist() == [it, 0, 0, 0]
            assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-7)
            assert not ga.any()
        torch.cuda.synchronize()
        ex.close()
