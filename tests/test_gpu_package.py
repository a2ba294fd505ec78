"""-m gpu: the host package (BaseVAE.forward -> Training.cost -> backward; OdeModel.simulate; GraphedStep) driving
the CUDA library, against the golden vectors minted from the reference.  Tolerances as in test_gpu_parity.py."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_case
from vihds_b200.config import Config, Settings
from vihds_b200.datasets import TimeSeriesDataset, build_datasets
from vihds_b200.parameters import Parameters
from vihds_b200.training import GraphedStep, Training
from vihds_b200.vae import build_model

pytestmark = pytest.mark.gpu
# the stock-PyTorch encoder is the fp32 REFERENCE here: cuDNN would otherwise run its conv weight-gradient in TF32
# (2.7e-4 relative error on conv.weight.grad, measured), which is the reference's imprecision, not the fused kernel's
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


class Args:
    seed, gpu, precision_hidden_layers, yaml, verbose = 0, None, None, None, False
    folds, split, heldout, train_samples, test_samples = 4, 1, None, 8, 8


FIXTURE = {"dr_constant_icml": "dataset_dr_icml", "relay_constant_precisions": "dataset_relay",
           "dr_blackbox_icml": "dataset_dr_icml"}


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def build(spec, dims=None, iw=8):
    with open(os.path.join(GOLDEN, "specs", spec + ".json")) as f:
        settings = Config(Args(), spec=json.load(f), device="cuda")
    par = Parameters(settings.params)
    pair = None
    if spec in FIXTURE:
        ds = TimeSeriesDataset.from_npz(os.path.join(GOLDEN, FIXTURE[spec] + ".npz"), settings.data)
        pair = build_datasets(Args(), settings, dataset=ds)
    torch.manual_seed(0)
    model = build_model(Args(), settings, pair if pair is not None else dims, par)
    a = Args()
    a.train_samples = iw
    return settings, par, model, Training(a, settings, pair, par, model)


def batch_from_case(case, dtype=torch.float32):
    d = lambda k: torch.as_tensor(case[k]).to("cuda", dtype)  # noqa: E731
    return Settings(times=d("times"), inputs=d("inputs"), dev_1hot=d("dev_1hot"), observations=d("observations"))


def pin_conditioner(model, case):
    """The reference draws fresh random conditioner weights per call; the golden case recorded the result."""
    ode = model.decoder.ode_model
    if "cond_aR" in case:
        planes = torch.stack([torch.as_tensor(case["cond_" + n]).reshape(-1) for n in ode.conditioned]).cuda()
        ode.conditioned_extras = lambda B, IW, dev_1hot: planes


@pytest.mark.parametrize("case_name,spec,dims", [
    ("dr_constant_icml_midpoint_f32_iw8", "dr_constant_icml", None),
    ("dr_constant_one_midpoint_f32_iw5", "dr_constant_one", (4, 100, 2, 1)),
    ("relay_constant_precisions_midpoint_f32_iw8", "relay_constant_precisions", None),
    ("dr_blackbox_icml_midpoint_f32_iw8", "dr_blackbox_icml", None),
    ("auto_constant_precisions_midpoint_f32_iw8", "auto_constant_precisions", (4, 100, 1, 1)),
    ("prpr_constant_midpoint_f32_iw8", "prpr_constant", (4, 200, 1, 1)),
    ("inducer_constant_precisions_midpoint_f32_iw8", "inducer_constant_precisions", (4, 100, 1, 1)),
    ("degrader_constant_precisions_midpoint_f32_iw8", "degrader_constant_precisions", (4, 135, 3, 1)),
])
def test_model_forward_cost_backward_match_reference(case_name, spec, dims):
    case = load_case(case_name)
    settings, par, model, training = build(spec, dims)
    if any(k.startswith("w:") for k in case):  # decoder weights: recorded from the reference run
        sd = {k[2:]: torch.as_tensor(case[k]).cuda() for k in case if k.startswith("w:")}
        model.decoder.load_state_dict(sd)
    pin_conditioner(model, case)
    batch = batch_from_case(case)
    u = torch.as_tensor(case["u"]).cuda()
    result, theta, q, p = model(batch, u.shape[1], u=u)
    x_states, x_predict, precisions = result
    q.mu.retain_grad()
    q.prec.retain_grad()
    cost = training.cost(batch, result, theta, q, p).elbo
    cost.backward()
    tol = 1e-4
    assert abs(float(cost) - float(case["loss"])) <= tol * abs(float(case["loss"]))
    for s in range(case["x_states"].shape[2]):
        assert _rel(x_states[:, :, s].detach().cpu().numpy(), case["x_states"][:, :, s]) < tol
    assert _rel(x_predict.detach().cpu().numpy(), case["x_predict"]) < tol
    assert _rel(precisions.detach().cpu().numpy(), case["precisions"]) < tol
    assert x_states.shape == case["x_states"].shape and precisions.shape == case["precisions"].shape
    assert _rel(theta.terms["log_p_by_species"].detach().cpu().numpy(), case["log_p_by_species"]) < tol
    assert _rel(q.log_prob(theta).detach().cpu().numpy(), case["log_q_theta"]) < tol
    assert _rel(p.log_prob(theta).detach().cpu().numpy(), case["log_p_theta"]) < tol
    per_ind = case["per_individual"].astype(bool)
    sel = case["kinds"] != 0
    for got, ref in ((q.mu.grad.cpu().numpy(), case["grad_q_mu"]), (q.prec.grad.cpu().numpy(), case["grad_q_prec"])):
        ref_tot = np.where(per_ind, ref.sum(0), ref[0])
        assert _rel(got.sum(0)[sel], ref_tot[sel]) < 3e-3
    for k in case:
        if k.startswith("gw:"):
            g = dict(model.decoder.named_parameters())[k[3:]].grad
            assert _rel(g.cpu().numpy(), case[k]) < 3e-3, k
    # no NaN gradients on any trainable tensor (tests/test_grad_dr.py of the reference)
    assert all(torch.isfinite(p_.grad).all() for p_ in model.parameters() if p_.grad is not None)


def test_simulate_seam_matches_reference_and_differentiates():
    """OdeModel.simulate (tests/test_ode_solvers.py of the reference calls it directly)."""
    case = load_case("dr_constant_one_midpoint_f32_iw5")
    settings, par, model, training = build("dr_constant_one", (4, 100, 2, 1), iw=5)
    from vihds_b200.distributions import DotOperatorSamples

    theta = DotOperatorSamples()
    for k, nm in enumerate(case["names"]):
        theta.add(str(nm), torch.as_tensor(case["theta"][k]).cuda().requires_grad_(True))
    batch = batch_from_case(case)
    ode = model.decoder.ode_model
    sol = ode.simulate(settings, batch.times, theta, batch.inputs, batch.dev_1hot, condition_on_device=False)
    assert sol.shape == case["x_states"].shape
    for s in range(8):
        assert _rel(sol[:, :, s].detach().cpu().numpy(), case["x_states"][:, :, s]) < 1e-4
    sol[:, :, 1, -1].sum().backward()
    assert torch.isfinite(theta.r.grad).all() and float(theta.r.grad.abs().sum()) > 0
    (x_states, x_predict, precisions), _ = model.decoder(theta, batch)
    assert _rel(x_predict.detach().cpu().numpy(), case["x_predict"]) < 1e-4
    assert _rel(precisions.detach().cpu().numpy(), case["precisions"]) < 1e-4


@pytest.mark.parametrize("solver", ["modeuler", "rk4", "euler", "modeulerwhile"])
def test_solver_switch_via_config(solver):
    case = load_case("dr_constant_one_%s_f32_iw5" % solver)
    settings, par, model, training = build("dr_constant_one", (4, 100, 2, 1), iw=5)
    settings.params.solver = solver
    batch = batch_from_case(case)
    u = torch.as_tensor(case["u"]).cuda()
    result, theta, q, p = model(batch, 5, u=u)
    cost = training.cost(batch, result, theta, q, p).elbo
    assert abs(float(cost) - float(case["loss"])) <= 1e-4 * abs(float(case["loss"]))


def test_adaptive_solver_raises():
    settings, par, model, training = build("dr_constant_one", (4, 100, 2, 1), iw=5)
    settings.params.solver = "dopri5"
    case = load_case("dr_constant_one_midpoint_f32_iw5")
    with pytest.raises(NotImplementedError):
        model(batch_from_case(case), 5, u=torch.as_tensor(case["u"]).cuda())


@pytest.mark.parametrize("use_graphs", [False, True])
def test_graphed_step_equals_eager_steps(use_graphs):
    """Three Adam steps through GraphedStep (static buffers, captured encoder fwd/bwd + fused Adam) land on the same
    parameters as three reference-shaped eager steps (model -> cost -> backward -> Adam)."""
    case = load_case("dr_constant_icml_midpoint_f32_iw8")
    batch = batch_from_case(case)
    B, IW, P = case["u"].shape
    us = [torch.randn(B, IW, P, generator=torch.Generator().manual_seed(i)).cuda() for i in range(3)]
    planes = torch.stack([torch.as_tensor(case["cond_" + n]).reshape(-1) for n in ("aR", "aS")]).cuda()

    _, _, model_a, tr_a = build("dr_constant_icml")
    model_a.decoder.ode_model.conditioned_extras = lambda B_, IW_, d: planes
    model_a.want_predict = False
    costs_a = []
    for u in us:
        assert tr_a._run_batch(batch, u=u)
        costs_a.append(float(tr_a.last_cost))

    _, _, model_b, tr_b = build("dr_constant_icml")
    gs = GraphedStep(tr_b, B, IW, batch.times.numel(), use_graphs=use_graphs)
    gs.extras_override = planes
    gs.load_batch(batch)
    costs_b = []
    for u in us:
        gs.load_u(u)
        costs_b.append(float(gs.step().item()))
    assert np.allclose(costs_a, costs_b, rtol=1e-5), (costs_a, costs_b)
    assert _rel(tr_b.optimizer.flat.cpu().numpy(), tr_a.optimizer.flat.cpu().numpy()) < 1e-4
    assert tr_b.optimizer.step_dev.tolist() == [3, 0, 0, 0]


def test_evaluation_moments_match_numpy_reduction():
    """Results.init (vihds/utils.py:79-99): device-side IW moments equal the reference's numpy formulas."""
    case = load_case("dr_constant_icml_midpoint_f32_iw8")
    settings, par, model, training = build("dr_constant_icml")
    pin_conditioner(model, case)
    batch = batch_from_case(case)
    u = torch.as_tensor(case["u"]).cuda()
    with torch.no_grad():
        result, theta, q, p = model(batch, u.shape[1], u=u)
        out = training.cost(batch, result, theta, q, p, full_output=True)
    x_states, x_predict, precisions = (t.cpu().numpy().astype(np.float64) for t in result)
    lw = out.log_unnormalized_iws.cpu().numpy().astype(np.float64)
    w = np.exp(lw - lw.max(1, keepdims=True))
    w = (w / w.sum(1, keepdims=True))[:, :, None, None]
    mu = (w * x_predict).sum(1)
    sd = np.sqrt((w * (x_predict ** 2 + 1.0 / precisions)).sum(1) - mu ** 2)
    assert _rel(out.iw_predict_mu, mu) < 1e-4 and _rel(out.iw_predict_std, sd) < 1e-3
    assert _rel(out.iw_states, (w * x_states).sum(1)) < 1e-4 and _rel(out.iw_variance, (w / precisions).sum(1)) < 1e-4
    assert abs(float(out.elbo) + float(case["loss"])) <= 1e-4 * abs(float(case["loss"]))


@pytest.mark.parametrize("case_name,spec,dims", [
    ("dr_constant_icml_midpoint_f32_iw8", "dr_constant_icml", None),           # local heads conditioned on devices
    ("dr_constant_one_midpoint_f32_iw5", "dr_constant_one", (4, 100, 2, 1)),     # + global-conditioned heads
    ("dr_blackbox_icml_midpoint_f32_iw8", "dr_blackbox_icml", None),
])
def test_fused_encoder_matches_pytorch_reference(case_name, spec, dims):
    """csrc/vh_encoder.cu (one launch forward, two backward) against the same encoder in stock PyTorch ops: q tables
    and the gradient of a random linear functional of them w.r.t. all eight parameter tensors.  fp32, rtol 2e-5."""
    case = load_case(case_name)
    settings, par, model, training = build(spec, dims)
    enc = model.encoder
    batch = batch_from_case(case)
    mu1, pr1 = enc.q_table(batch)
    mu0, pr0 = enc.q_table_reference(batch)
    assert torch.allclose(mu1, mu0, rtol=2e-5, atol=1e-6) and torch.allclose(pr1, pr0, rtol=2e-5, atol=1e-6)
    g = torch.Generator(device="cuda").manual_seed(0)
    g_mu, g_pr = torch.randn(mu0.shape, device="cuda", generator=g), torch.randn(mu0.shape, device="cuda", generator=g)
    params = [p for p in enc.fused_parameters() if p.numel()]
    got = torch.autograd.grad([mu1, pr1], params, [g_mu, g_pr], allow_unused=True)
    ref = torch.autograd.grad([mu0, pr0], params, [g_mu, g_pr], allow_unused=True)
    for a, b, p in zip(got, ref, params):
        assert _rel(a.cpu().numpy(), b.cpu().numpy()) < 2e-5, tuple(p.shape)


def test_graphed_step_fused_encoder_equals_pytorch_encoder():
    """Same three steps with the fused encoder kernels and with the stock-PyTorch encoder inside GraphedStep."""
    case = load_case("dr_constant_icml_midpoint_f32_iw8")
    batch = batch_from_case(case)
    B, IW, P = case["u"].shape
    us = [torch.randn(B, IW, P, generator=torch.Generator().manual_seed(i)).cuda() for i in range(3)]
    flats, costs = [], []
    for fused in (True, False):
        _, _, model, tr = build("dr_constant_icml")
        model.encoder.fused = fused
        gs = GraphedStep(tr, B, IW, batch.times.numel())
        assert gs.fused_encoder == fused
        gs.load_batch(batch)
        gs.cond_w.copy_(torch.tensor([[2.0, 1.0, 3.0, 0.5, 2.5, 1.5, 0.2], [1.0, 2.0, 0.3, 3.0, 0.7, 2.2, 1.1]]).cuda())
        c = []
        for u in us:
            gs.load_u(u)
            c.append(float(gs.step().item()))
        costs.append(c)
        flats.append(tr.optimizer.flat.cpu().numpy())
    assert np.allclose(costs[0], costs[1], rtol=1e-5)
    assert _rel(flats[0], flats[1]) < 1e-4


def test_device_conditioner_kernel_matches_reference_quirk():
    """vh_device_conditioner == OdeModel.device_conditioner(ones) including the repeat/reshape row scramble."""
    import ctypes as C

    from vihds_b200 import _lib as L

    settings, par, model, training = build("dr_constant_icml")
    ode = model.decoder.ode_model
    case = load_case("dr_constant_icml_midpoint_f32_iw8")
    dev_1hot = torch.as_tensor(case["dev_1hot"]).cuda()
    B, IW, D = dev_1hot.shape[0], 5, dev_1hot.shape[1]
    torch.manual_seed(3)
    ones = torch.ones(B, IW, device="cuda")
    ref = torch.stack([ode.device_conditioner(ones, n, dev_1hot).reshape(-1) for n in ("aR", "aS")])
    torch.manual_seed(3)
    from vihds_b200.models import _draw_conditioner_weight

    w = torch.cat([_draw_conditioner_weight(D) for _ in range(2)], 0).cuda()
    rel = torch.stack([torch.as_tensor(ode.relevance[n]) for n in ("aR", "aS")]).cuda().contiguous()
    plus = torch.tensor([1, 1], dtype=torch.int32, device="cuda")
    out = torch.zeros(2, B * IW, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    L.check(L.load().vh_device_conditioner(0, B, IW, D, 2, B, 0, p(dev_1hot), p(rel), p(w), p(plus), p(out), None))
    torch.cuda.synchronize()
    assert torch.allclose(out, ref, rtol=1e-6, atol=1e-6)
    # sharded over individuals: every slab, computed from the GLOBAL one-hot table and its offset, must reproduce the
    # rows of the single-process result (sample (b, i) of the global batch gets individual (b*IW + i) % B_global)
    from vihds_b200.distributed import shard_bounds
    for world in (2, 5):
        for rank in range(world):
            lo, hi = shard_bounds(B, world, rank)
            part = torch.zeros(2, (hi - lo) * IW, device="cuda")
            L.check(L.load().vh_device_conditioner(0, hi - lo, IW, D, 2, B, lo, p(dev_1hot), p(rel), p(w), p(plus), p(part), None))
            torch.cuda.synchronize()
            assert torch.equal(part, out[:, lo * IW:hi * IW])


@pytest.mark.parametrize("spec", ["dr_constant_icml", "dr_blackbox_icml", "relay_constant_precisions"])
def test_training_run_prints_finite_elbo(spec, capsys):
    """tests/test_run_xval.py of the reference: a short run (2 epochs, IW = 10, evaluation every epoch) completes, prints
    one 'iwae-elbo' line per evaluation, and the ELBO is finite."""
    settings, par, model, training = build(spec, iw=10)
    training.args.epochs, training.args.test_epoch, training.args.test_samples = 2, 1, 10
    np.random.seed(0)
    torch.manual_seed(0)
    assert training.run(verbose=True) is True
    lines = [l for l in capsys.readouterr().out.splitlines() if "iwae-elbo" in l]
    assert len(lines) == 2
    assert all(np.isfinite(float(l.split("=")[-1])) for l in lines)


@pytest.mark.parametrize("B", [300, 601, 130])
def test_fused_encoder_large_batch_paths(B):
    """Against the stock-PyTorch encoder: B = 300 and 601 take the GEMM path of the hidden layer (conv + pool | GEMM | tanh +
    heads forward; heads | two GEMMs | pool + conv backward), with one individual per CTA resp. four (>= 4 x 148, ragged last
    group); B = 130 the monolithic kernels with the B-split weight-gradient kernel."""
    case = load_case("dr_constant_icml_midpoint_f32_iw8")
    settings, par, model, training = build("dr_constant_icml")
    enc = model.encoder
    small = batch_from_case(case)
    idx = torch.arange(B, device="cuda") % small.inputs.shape[0]
    g = torch.Generator(device="cuda").manual_seed(1)
    batch = Settings(times=small.times, inputs=small.inputs[idx].contiguous(), dev_1hot=small.dev_1hot[idx].contiguous(),
                     observations=(small.observations[idx] + 0.01 * torch.randn(B, 4, small.times.numel(), device="cuda", generator=g)).contiguous())
    mu1, pr1 = enc.q_table(batch)
    mu0, pr0 = enc.q_table_reference(batch)
    assert torch.allclose(mu1, mu0, rtol=2e-5, atol=1e-6) and torch.allclose(pr1, pr0, rtol=2e-5, atol=1e-6)
    g_mu, g_pr = torch.randn(mu0.shape, device="cuda", generator=g), torch.randn(mu0.shape, device="cuda", generator=g)
    params = [p for p in enc.fused_parameters() if p.numel()]
    got = torch.autograd.grad([mu1, pr1], params, [g_mu, g_pr], allow_unused=True)
    ref = torch.autograd.grad([mu0, pr0], params, [g_mu, g_pr], allow_unused=True)
    for a, b, p in zip(got, ref, params):
        assert _rel(a.cpu().numpy(), b.cpu().numpy()) < 5e-5, tuple(p.shape)


@pytest.mark.parametrize("one_graph", ["1", "0"])
def test_step_from_host_equals_step(one_graph, monkeypatch):
    """The end-to-end entry (host buffers; two input sets filled on a copy stream, alternating) takes the same steps as
    load_* + step().  Every step has its own batch contents and u, the calls are issued back to back without a
    synchronisation, and a ``step()`` in the middle of the stream (mode "mixed") computes on the set in use."""
    monkeypatch.setenv("VIHDS_ONE_GRAPH", one_graph)  # "0": the two-graph form of the step everywhere
    case = load_case("dr_constant_icml_midpoint_f32_iw8")
    batch = batch_from_case(case)
    B, IW, P = case["u"].shape
    n = 7
    g = torch.Generator().manual_seed(3)
    hosts, us = [], []
    for i in range(n):
        h = Settings(**{k: v.cpu().clone() for k, v in batch.items()})
        h.observations = h.observations + 0.02 * torch.randn(h.observations.shape, generator=g)
        hosts.append(Settings(**{k: v.pin_memory() for k, v in h.items()}))
        hosts[-1].times = hosts[0].times  # the data set's time grid: one storage
        us.append(torch.randn(B, IW, P, generator=torch.Generator().manual_seed(i)).pin_memory())
    flats, costs = [], []
    for mode in ("host", "mixed", "device"):
        _, _, model, tr = build("dr_constant_icml")
        gs = GraphedStep(tr, B, IW, batch.times.numel())
        torch.manual_seed(5)  # conditioner weights are drawn from the torch CPU stream inside every path
        cs = []
        for i, (h, u) in enumerate(zip(hosts, us)):
            if mode == "host" or (mode == "mixed" and i not in (2, 3)):
                c = gs.step_from_host(h, u)
            else:
                gs.load_batch(Settings(**{k: v.cuda() for k, v in h.items()}))
                gs.draw_conditioner()
                gs.load_u(u.cuda())
                c = gs.step()
            cs.append(c.clone())
        torch.cuda.synchronize()
        assert all(torch.isfinite(c).all() for c in cs)
        costs.append(torch.cat(cs).cpu().numpy())
        flats.append(tr.optimizer.flat.cpu().numpy())
    # not bit-equal: the gradient sums use atomics (summation order varies run to run, ~1e-7 relative per step); a step
    # that computed on the wrong input set would be off by O(1) (every step has its own u)
    for k in (1, 2):
        assert np.allclose(costs[0], costs[k], rtol=1e-5, atol=0), (costs[0], costs[k])
        assert _rel(flats[0], flats[k]) < 2e-5
