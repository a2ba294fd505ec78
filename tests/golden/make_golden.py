"""Mint the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF (read-only, /root/reference).

Run in the build container only:  ``python tests/golden/make_golden.py``  (about a minute on 8 cores).

Every ``<case>.npz`` holds the frozen inputs of one hot-path invocation (batch tensors, ``u``, the encoder's q
parameters, prior parameters, conditioned aR/aS or MLP weights) and what the reference computed from them
(x_states, x_predict, precisions, log_p_by_species, log_p_theta, log_q_theta, the IWAE cost and its gradient w.r.t.
the q parameters / decoder weights).  The parsed YAML of each spec is stored as JSON (``specs/<name>.json``) so the
GPU box, which has no /root/reference, can rebuild the same parameter set.  See _ref_harness.py for the shims
(notably: torchdiffeq is restated, so midpoint/rk4/euler cases are "parity unpinned"; modeuler* are pure reference).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_harness as H  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def _slice_batch(batch, n):
    from munch import munchify

    return munchify({
        "devices": batch.devices[:n],
        "dev_1hot": batch.dev_1hot[:n],
        "inputs": batch.inputs[:n],
        "observations": batch.observations[:n],
        "times": batch.times,
    })


def run_case(spec, solver, dtype, iw, n_batch=None, traces=True, tag=None, extra_args=()):
    args, settings, data, parameters, model, training = H.build_reference(spec, samples=iw, dtype=dtype, solver=solver,
                                                                          extra_args=extra_args)
    from vihds.distributions import TfConstant, TfLogNormal, TfNormal

    batch = next(iter(training.train_loader))
    if n_batch is not None:
        batch = _slice_batch(batch, n_batch)
    B = len(batch.inputs)

    captured = {}
    orig_sample_u = model.sample_u

    def sample_u(n_batch_, n_samples, device=None):
        u = orig_sample_u(n_batch_, n_samples)
        if dtype == "float64":
            u = u.double()
        captured["u"] = u
        return u

    model.sample_u = sample_u
    model.train()
    result, theta, q, p = model(batch, iw)
    x_states, x_predict, precisions = result
    names = list(q.distributions.keys())
    P = len(names)
    for d in q.distributions.values():
        for t in d.get_tensors():
            if torch.is_tensor(t) and t.requires_grad and not t.is_leaf:
                t.retain_grad()
    dec_params = [(n, w) for n, w in model.decoder.named_parameters()]
    # per-sample terms exactly as Training.cost forms them (training.py:130-136)
    from vihds.training import log_prob_observations

    log_p_by_species = log_prob_observations(model, x_predict, batch.observations, precisions, False)
    log_q_theta = q.log_prob(theta)
    log_p_theta = p.log_prob(theta)
    loss = training.cost(batch, result, theta, q, p).elbo
    loss.backward()

    kinds = np.zeros(P, np.int32)
    q_mu = np.zeros((B, P), np.float64)
    q_prec = np.ones((B, P), np.float64)
    p_mu = np.zeros(P, np.float64)
    p_prec = np.ones(P, np.float64)
    p_sigma = np.ones(P, np.float64)
    g_mu = np.zeros((B, P), np.float64)
    g_prec = np.zeros((B, P), np.float64)
    per_individual = np.zeros(P, np.int32)
    for k, name in enumerate(names):
        d = q.distributions[name]
        pd_ = p.distributions[name]
        if isinstance(d, TfConstant):
            kinds[k] = 0
            q_mu[:, k] = float(d.value)
            p_mu[k] = float(pd_.value)
            continue
        kinds[k] = 2 if isinstance(d, TfLogNormal) else 1
        assert isinstance(d, TfNormal)
        mu, prec = _np(d.mu).reshape(-1), _np(d.prec).reshape(-1)
        per_individual[k] = int(mu.size == B and B > 1)
        q_mu[:, k] = mu
        q_prec[:, k] = prec
        p_mu[k] = float(pd_.mu)
        p_prec[k] = float(pd_.prec)
        p_sigma[k] = float(pd_.sigma)
        gm = d.mu.grad
        gp = d.prec.grad
        if gm is not None:
            gm = _np(gm).reshape(-1)
            g_mu[: gm.size, k] = gm  # global parameters: total gradient lands in row 0
        if gp is not None:
            gp = _np(gp).reshape(-1)
            g_prec[: gp.size, k] = gp
    fdt = np.float64 if dtype == "float64" else np.float32
    out = {
        "spec": spec, "solver": settings.params.solver, "dtype": dtype, "model": settings.model,
        "names": np.array(names), "kinds": kinds, "per_individual": per_individual,
        "times": _np(batch.times).astype(fdt), "inputs": _np(batch.inputs).astype(fdt),
        "dev_1hot": _np(batch.dev_1hot).astype(fdt), "observations": _np(batch.observations).astype(fdt),
        "devices": np.asarray(batch.devices).astype(np.int32),
        "u": _np(captured["u"]).astype(fdt),
        "q_mu": q_mu.astype(fdt), "q_prec": q_prec.astype(fdt), "p_mu": p_mu.astype(fdt), "p_prec": p_prec.astype(fdt),
        "p_sigma": p_sigma.astype(fdt),
        "theta": np.stack([_np(theta.samples[n]) for n in names]).astype(fdt),  # clipped, unconditioned [P,B,IW]
        "log_p_by_species": _np(log_p_by_species).astype(fdt),
        "log_p_theta": _np(log_p_theta).astype(fdt), "log_q_theta": _np(log_q_theta).astype(fdt),
        "loss": np.array(float(loss), np.float64),
        "grad_q_mu": g_mu.astype(fdt), "grad_q_prec": g_prec.astype(fdt),
    }
    # conditioned extras (quirk a4: fresh random conditioner per call => recorded, not recomputed)
    for extra in ("aR", "aS"):
        if extra not in names and hasattr(theta, extra):
            out["cond_" + extra] = _np(getattr(theta, extra)).astype(fdt)
    for i in (1, 2):
        n = "y%d" % i
        if n in names and settings.model == "dr_blackbox":
            out["cond_" + n] = _np(getattr(theta, n)).astype(fdt)
    for n, w in dec_params:
        out["w:" + n] = _np(w).astype(fdt)
        if w.grad is not None:
            out["gw:" + n] = _np(w.grad).astype(fdt)
    if traces:
        out["x_states"] = _np(x_states).astype(fdt)
        out["x_predict"] = _np(x_predict).astype(fdt)
        out["precisions"] = _np(precisions).astype(fdt)
    else:
        out["x_states_last"] = _np(x_states[:, :, :, -1]).astype(fdt)
    name = tag or "%s_%s_%s_iw%d" % (spec, settings.params.solver, "f64" if dtype == "float64" else "f32", iw)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("==> %s: loss=%.6f  B=%d IW=%d P=%d" % (name, float(loss), B, iw, P))
    torch.set_default_dtype(torch.float32)
    return settings, data


def dump_spec(spec):
    import yaml

    os.makedirs(os.path.join(HERE, "specs"), exist_ok=True)
    with open(os.path.join(H.REFERENCE_ROOT, "specs", spec + ".yaml")) as f:
        d = yaml.safe_load(f)
    with open(os.path.join(HERE, "specs", spec + ".json"), "w") as f:
        json.dump(d, f, indent=1, sort_keys=False)


def dump_dataset(spec, tag):
    """Whole pre-processed dataset of a spec (what TimeSeriesDataset holds after scale_data), for bench/tests."""
    args, settings, data, parameters, model, training = H.build_reference(spec, samples=2)
    ds = data.train.dataset
    np.savez_compressed(
        os.path.join(HERE, tag + ".npz"),
        times=_np(ds.times).astype(np.float32), inputs=_np(ds.inputs).astype(np.float32),
        dev_1hot=_np(ds.dev_1hot).astype(np.float32), observations=_np(ds.observations).astype(np.float32),
        devices=np.asarray(ds.devices).astype(np.int32), train_ids=np.asarray(data.train.indices),
        test_ids=np.asarray(data.test.indices),
    )
    print("==> %s: %d individuals, T=%d" % (tag, len(ds), ds.n_times))


def run_training_steps(spec="dr_constant_icml", iw=20, k=5, tag=None, epochs=None):
    """k consecutive training steps of the reference's own ``Training._run_batch`` (vihds/training.py:324-340) on ONE
    mini-batch: records what is random per step (u from numpy's RNG, the device-conditioner output) and what the
    reference made of it (the cost of every step, every trainable parameter before the first and after the last step).
    ``epochs``: instead of k steps on one batch, run that many passes over the reference's own shuffling train_loader
    (vihds/training.py:360-366) and record the mini-batch of every step."""
    import time

    args, settings, data, parameters, model, training = H.build_reference(spec, samples=iw)
    batch = next(iter(training.train_loader))
    us, conds, losses = [], [], []
    orig_sample_u = model.sample_u

    def sample_u(n_batch_, n_samples, device=None):
        u = orig_sample_u(n_batch_, n_samples)
        us.append(_np(u).copy())
        return u

    model.sample_u = sample_u
    ode = model.decoder.ode_model
    orig_dc = ode.device_conditioner

    def device_conditioner(param, name, dev_1hot, *a, **kw):
        out = orig_dc(param, name, dev_1hot, *a, **kw)
        conds.append((name, _np(out).copy()))
        return out

    ode.device_conditioner = device_conditioner
    orig_cost = training.cost

    def cost(*a, **kw):
        r = orig_cost(*a, **kw)
        losses.append(float(r.elbo))
        return r

    training.cost = cost

    class Log(object):
        batch_feed_time = batch_train_time = 0.0

    model.train()
    init = {n: _np(w).copy() for n, w in model.named_parameters()}
    batches = []
    if epochs:
        for _ in range(epochs):
            for b in training.train_loader:
                batches.append(b)
                assert training._run_batch(time.time(), b, Log())
        k = len(batches)
        assert len({len(b.inputs) for b in batches}) == 1, "ragged batches are not recorded by this fixture"
    else:
        for _ in range(k):
            assert training._run_batch(time.time(), batch, Log())
    final = {n: _np(w).copy() for n, w in model.named_parameters()}
    names = sorted({n for n, _ in conds}, key=[n for n, _ in conds].index)
    out = {
        "spec": spec, "solver": settings.params.solver, "dtype": "float32", "model": settings.model, "steps": np.array(k),
        "learning_rate": np.array(float(settings.params.learning_rate)),
        "times": _np(batch.times).astype(np.float32),
        # one batch for all steps, or [K, ...] stacks when the loader was iterated
        "inputs": (np.stack([_np(b.inputs) for b in batches]) if batches else _np(batch.inputs)).astype(np.float32),
        "dev_1hot": (np.stack([_np(b.dev_1hot) for b in batches]) if batches else _np(batch.dev_1hot)).astype(np.float32),
        "observations": (np.stack([_np(b.observations) for b in batches]) if batches else _np(batch.observations)).astype(np.float32),
        "per_step_batches": np.array(bool(batches)),
        "u": np.stack(us).astype(np.float32), "cond_names": np.array(names),
        "cond": np.stack([np.stack([c for n, c in conds[i * len(names):(i + 1) * len(names)]]) for i in range(k)]).astype(np.float32)
        if names else np.zeros((k, 0), np.float32),
        "losses": np.array(losses[:k], np.float64),
    }
    for n, w in init.items():
        out["init:" + n] = w
    for n, w in final.items():
        out["final:" + n] = w
    name = tag or "%s_train%d_iw%d" % (spec, k, iw)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("==> %s: losses %s" % (name, " ".join("%.4f" % v for v in losses[:k])))


SPECS = ("dr_constant_one", "dr_constant_icml", "dr_blackbox_icml", "relay_constant_precisions", "dr_constant_v2",
         "dr_constant_precisions", "dr_constant_precisions_v2", "auto_constant", "auto_constant_precisions",
         "prpr_constant", "prpr_constant_precisions", "inducer_constant_precisions", "degrader_constant_precisions")


def small_models():
    """auto_* / prpr_*: the remaining runnable specs of the reference (10 of 16 run)."""
    for spec in ("auto_constant", "auto_constant_precisions", "prpr_constant", "prpr_constant_precisions"):
        run_case(spec, "midpoint", "float32", 8, n_batch=12)
    run_case("auto_constant_precisions", "midpoint", "float64", 8, n_batch=6)
    run_case("prpr_constant", "modeuler", "float32", 8, n_batch=12)


def extra_models():
    """inducer / degrader: broken as shipped like relay, runnable under the same monkeypatch."""
    run_case("inducer_constant_precisions", "midpoint", "float32", 8, n_batch=12)
    run_case("inducer_constant_precisions", "midpoint", "float64", 8, n_batch=6)
    run_case("degrader_constant_precisions", "midpoint", "float32", 8, n_batch=12)
    run_case("degrader_constant_precisions", "midpoint", "float64", 8, n_batch=6)
    run_case("degrader_constant_precisions", "modeuler", "float32", 8, n_batch=12)
    # NeuralPrecisions WITH a hidden layer on white-box models (--precision_hidden_layers, run_xval.py:38)
    hid = ("--precision_hidden_layers=5",)
    run_case("dr_constant_precisions", "midpoint", "float32", 8, n_batch=12, extra_args=hid, tag="dr_constant_precisions_hidden5_midpoint_f32_iw8")
    run_case("dr_constant_precisions", "midpoint", "float64", 8, n_batch=6, extra_args=hid, tag="dr_constant_precisions_hidden5_midpoint_f64_iw8")
    run_case("relay_constant_precisions", "modeuler", "float32", 8, n_batch=12, extra_args=hid, tag="relay_constant_precisions_hidden5_modeuler_f32_iw8")


def main(group="all"):
    for spec in SPECS:
        dump_spec(spec)
    if group == "small":
        return small_models()
    if group == "extra":
        return extra_models()
    if group == "full":  # BASELINE sizes (B = 36 x IW = 200) of configs 3 and 5, last-timepoint states only
        run_case("dr_blackbox_icml", "midpoint", "float32", 200, traces=False)
        run_case("relay_constant_precisions", "midpoint", "float32", 200, traces=False)
        return
    if group == "train":
        run_training_steps("dr_constant_icml", 20, 5)
        # BASELINE config 1: dr_constant_one, IW = 5, whole epochs through the reference's shuffling loader
        return run_training_steps("dr_constant_one", 5, None, tag="dr_constant_one_epochs4_iw5_train", epochs=4)
    # config 1: every fixed-step solver, fp32; fp64 for the default and the in-repo solver
    for solver in ("midpoint", "rk4", "euler", "modeuler", "modeulerwhile"):
        run_case("dr_constant_one", solver, "float32", 5, n_batch=8)
    for solver in ("midpoint", "modeuler"):
        run_case("dr_constant_one", solver, "float64", 5, n_batch=8)
    # config 2
    run_case("dr_constant_icml", "midpoint", "float32", 8)
    run_case("dr_constant_icml", "modeuler", "float32", 8, n_batch=12)
    run_case("dr_constant_icml", "midpoint", "float32", 200, traces=False)
    run_case("dr_constant_icml", "midpoint", "float64", 8, n_batch=12)
    # config 3 / 5 and the remaining dr_constant variants
    run_case("dr_blackbox_icml", "midpoint", "float32", 8, n_batch=12)
    run_case("dr_blackbox_icml", "midpoint", "float64", 8, n_batch=12)
    run_case("relay_constant_precisions", "midpoint", "float32", 8, n_batch=12)
    run_case("relay_constant_precisions", "midpoint", "float64", 8, n_batch=12)
    run_case("dr_constant_v2", "midpoint", "float32", 8, n_batch=12)
    run_case("dr_constant_precisions", "midpoint", "float32", 8, n_batch=12)
    run_case("dr_constant_precisions_v2", "midpoint", "float32", 8, n_batch=12)
    small_models()
    extra_models()
    run_case("dr_blackbox_icml", "midpoint", "float32", 200, traces=False)
    run_case("relay_constant_precisions", "midpoint", "float32", 200, traces=False)
    run_training_steps("dr_constant_icml", 20, 5)
    run_training_steps("dr_constant_one", 5, None, tag="dr_constant_one_epochs4_iw5_train", epochs=4)
    dump_dataset("dr_constant_icml", "dataset_dr_icml")
    dump_dataset("relay_constant_precisions", "dataset_relay")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "all")
