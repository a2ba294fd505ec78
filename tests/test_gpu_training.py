"""-m gpu: sequences of training steps.

* the 5-step golden minted from the reference's own ``Training._run_batch`` (tests/golden/make_golden.py
  run_training_steps; vihds/training.py:324-340): same initial parameters, same u, same recorded device-conditioner
  output per step -> the CUDA-graph step must reproduce the reference's cost of every step and every trained parameter;
* ``Training.run`` takes the CUDA-graph path on a GPU and gives what the eager, reference-shaped path gives;
* a NaN cost freezes the parameters on the device (graphed path), as the reference's early exit does.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from test_gpu_package import build
from vihds_b200.datasets import batch_of
from vihds_b200.training import GraphedStep

pytestmark = pytest.mark.gpu


def _load_reference_parameters(model, z, prefix):
    sd = {k[len(prefix):]: torch.as_tensor(z[k]) for k in z.files if k.startswith(prefix)}
    model.encoder.load_reference_state_dict(sd)
    return sd


FIXTURES = {
    # 5 steps on one icml mini-batch (device conditioner recorded per step)
    "dr_constant_icml_train5_iw20": ("dr_constant_icml", None),
    # BASELINE config 1: dr_constant_one, IW = 5, four epochs through the reference's shuffling loader (one 36-individual
    # mini-batch per epoch; no device conditioner: aR / aS are sampled global-conditioned parameters)
    "dr_constant_one_epochs4_iw5_train": ("dr_constant_one", (4, 100, 2, 1)),
}


@pytest.mark.parametrize("use_graphs", [True, False])
@pytest.mark.parametrize("fixture", sorted(FIXTURES))
def test_training_steps_match_the_reference(fixture, use_graphs):
    z = np.load(os.path.join(GOLDEN, fixture + ".npz"))
    spec, dims = FIXTURES[fixture]
    settings, par, model, training = build(spec, dims)
    _load_reference_parameters(model, z, "init:")
    assert abs(float(z["learning_rate"]) - training.optimizer.lr) < 1e-12
    K, B, IW, P = z["u"].shape
    T = len(z["times"])
    per_step = bool(z["per_step_batches"]) if "per_step_batches" in z.files else False
    model.want_predict = False
    gs = GraphedStep(training, B, IW, T, use_graphs=use_graphs)
    assert list(z["cond_names"]) == list(gs.extras)
    dev = lambda a: torch.as_tensor(a).cuda()  # noqa: E731
    planes = torch.zeros(max(1, len(gs.extras)), B * IW, device="cuda")
    if gs.extras:
        gs.extras_override = planes  # the conditioner output the reference drew at each step (fresh random weights per call)
    costs = []
    for i in range(K):
        pick = (lambda a: a[i]) if per_step else (lambda a: a)
        gs.load_batch({"times": dev(z["times"]), "inputs": dev(pick(z["inputs"])), "dev_1hot": dev(pick(z["dev_1hot"])),
                       "observations": dev(pick(z["observations"]))})
        if gs.extras:
            planes.copy_(torch.as_tensor(z["cond"][i]).reshape(len(gs.extras), B * IW))
        gs.load_u(torch.as_tensor(z["u"][i]).cuda())
        costs.append(float(gs.step().item()))
    ref = z["losses"]
    assert np.allclose(costs, ref, rtol=1e-4), (costs, ref.tolist())
    # every trained parameter, under the reference's names.  Adam moves each parameter by ~lr = 0.01 per step whatever
    # the size of its gradient, so a relative error e in a (tiny, cancellation-dominated) gradient entry shows up as
    # ~e * lr per step: measured on B200 after 5 steps: max |difference| 6.6e-5 (0.13 % of the 0.05 a parameter travels),
    # 98.8 % of the 36,974 entries within 1e-5.  Bars: 2e-4 on every entry, 1e-5 on at least 97 % of them.
    final = {k[len("final:"):]: z[k] for k in z.files if k.startswith("final:")}
    got = model.encoder.reference_state_dict()
    assert set(got) == set(final)
    worst, n_bad, n_all = 0.0, 0, 0
    for k, v in got.items():
        d = np.abs(v.cpu().numpy().reshape(-1) - final[k].reshape(-1))
        worst = max(worst, float(d.max()))
        n_bad += int((d > 1e-5).sum())
        n_all += d.size
    assert worst < 2e-4 and n_bad <= 0.03 * n_all, (worst, n_bad, n_all)
    assert training.optimizer.step_dev.tolist()[0] == K


def test_training_run_takes_the_graphed_path_and_matches_eager():
    costs, flats = {}, {}
    for graphed in (False, True):
        torch.manual_seed(0)
        np.random.seed(0)
        settings, par, model, training = build("dr_constant_icml")
        training.args.train_samples, training.args.epochs, training.args.test_epoch = 16, 1, 0
        torch.manual_seed(5)
        np.random.seed(5)
        assert training.run(epochs=1, verbose=False, graphed=None if graphed else False)
        assert training.path_taken == ("graphed" if graphed else "eager")
        costs[graphed], flats[graphed] = training.costs, training.optimizer.flat.clone()
        n_steps = -(-training.dataset_pair.n_train // training.n_batch)
        assert len(training.costs) == n_steps and training.optimizer.step_dev.tolist()[0] == n_steps
        if graphed:  # the ragged last mini-batch has its own captured step
            assert len(training._graphed) == (2 if training.dataset_pair.n_train % training.n_batch else 1)
    assert np.allclose(costs[True], costs[False], rtol=2e-4), (costs[True], costs[False])
    assert float((flats[True] - flats[False]).abs().max()) < 5e-3


def test_nan_cost_freezes_the_graphed_step():
    settings, par, model, training = build("dr_constant_icml")
    ds = training.dataset_pair.train.dataset
    batch = batch_of(ds, np.asarray(training.dataset_pair.train.indices)[:6], settings.device, settings.dtype)
    IW = 8
    model.want_predict = False
    gs = GraphedStep(training, 6, IW, batch.times.numel())
    gs.load_batch(batch)
    u = torch.randn(6, IW, par.n_theta, device="cuda")
    gs.load_u(u)
    gs.draw_conditioner()
    assert np.isfinite(float(gs.step().item()))
    before = [t.clone() for t in (training.optimizer.flat, training.optimizer.exp_avg, training.optimizer.exp_avg_sq)]
    bad = u.clone()
    bad[2, 3, 0] = float("nan")
    gs.load_u(bad)
    assert np.isnan(float(gs.step().item()))
    gs.load_u(u)
    gs.step()  # a later good step is refused as well: training has stopped
    torch.cuda.synchronize()
    after = (training.optimizer.flat, training.optimizer.exp_avg, training.optimizer.exp_avg_sq)
    assert all(torch.equal(a, b) for a, b in zip(after, before))
    assert gs.skipped_steps() == 2 and training.optimizer.step_dev.tolist()[0] == 1
    assert not training.optimizer.grad.any()
