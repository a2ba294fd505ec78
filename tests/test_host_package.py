"""CPU: host-side mirror of the reference interface (config, parameter table, encoder initialisation, data loading,
dense distribution tables) against values recorded from the reference (tests/golden) and against the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_case
import vihds_oracle as O
from vihds_b200 import datasets as D
from vihds_b200.config import Config, Settings
from vihds_b200.distributions import ChainedDistribution
from vihds_b200.encoders import Encoder
from vihds_b200.parameters import Parameters
from vihds_b200.training import multistep_lr


class Args:
    seed, gpu, precision_hidden_layers, yaml = 0, None, None, None
    folds, split, heldout = 4, 1, None


def _config(spec):
    with open(os.path.join(GOLDEN, "specs", spec + ".json")) as f:
        return Config(Args(), spec=json.load(f), device="cpu")


CASES = [("dr_constant_icml_midpoint_f32_iw8", "dr_constant_icml", (4, 86, 2, 7)),
         ("dr_constant_one_midpoint_f32_iw5", "dr_constant_one", (4, 100, 2, 1)),
         ("relay_constant_precisions_midpoint_f32_iw8", "relay_constant_precisions", (4, 99, 2, 1)),
         ("dr_blackbox_icml_midpoint_f32_iw8", "dr_blackbox_icml", (4, 86, 2, 7)),
         ("inducer_constant_precisions_midpoint_f32_iw8", "inducer_constant_precisions", (4, 100, 1, 1)),
         ("degrader_constant_precisions_midpoint_f32_iw8", "degrader_constant_precisions", (4, 135, 3, 1))]


def test_device_bookkeeping_icml():
    cfg = _config("dr_constant_icml")
    assert cfg.data.device_depth == 7 and cfg.params.solver == "midpoint" and cfg.params.n_batch == 36
    assert cfg.data.relevance_vectors["aR"].tolist() == [0, 1, 1, 0, 0, 0, 0]
    assert cfg.data.relevance_vectors["aS"].tolist() == [0, 0, 0, 0, 1, 1, 1]
    assert _config("dr_constant_one").data.device_depth == 1


@pytest.mark.parametrize("case_name,spec,dims", CASES)
def test_parameter_table_and_encoder_match_reference(case_name, spec, dims):
    """Column order, kinds, prior tables and -- with the same seed -- the encoder's q(theta | x) equal what the
    reference produced (q_mu, q_prec recorded by tests/golden/make_golden.py)."""
    case = load_case(case_name)
    cfg = _config(spec)
    par = Parameters(cfg.params)
    assert par.names == [str(n) for n in case["names"]]
    assert (par.kinds() == case["kinds"]).all()
    mu, prec, lo, hi = par.prior_arrays(np.float32)
    sel = case["kinds"] != 0
    assert np.allclose(mu[sel], case["p_mu"][sel], rtol=0, atol=0)
    assert np.allclose(prec[sel], case["p_prec"][sel], rtol=1e-7)
    torch.manual_seed(0)
    enc = Encoder(par, dims)
    data = Settings(observations=torch.as_tensor(case["observations"]), inputs=torch.as_tensor(case["inputs"]),
                    dev_1hot=torch.as_tensor(case["dev_1hot"]))
    q_mu, q_prec = enc.q_table_reference(data)  # the stock-PyTorch restatement; the product path needs CUDA tensors
    assert np.abs(q_mu.detach().numpy() - case["q_mu"]).max() < 1e-6
    assert (np.abs(q_prec.detach().numpy() - case["q_prec"]) / case["q_prec"]).max() < 1e-6
    assert enc.p.mu.shape == (len(par.names),)
    with pytest.raises(RuntimeError):  # no silent CPU path on the product surface
        enc(data)


def test_encoder_loads_and_exports_reference_state_dict():
    """The packed heads <-> the reference's per-parameter layers (names of vihds/encoders.py): a reference checkpoint
    loads into the packed encoder and comes back out unchanged, and the trained reference parameters of the 5-step
    training golden give the q table the reference itself would compute from them."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "dr_constant_icml_train5_iw20.npz"))
    cfg = _config("dr_constant_icml")
    par = Parameters(cfg.params)
    enc = Encoder(par, (4, 86, 2, 7))
    sd = {k[len("final:"):]: torch.as_tensor(z[k]) for k in z.files if k.startswith("final:")}
    enc.load_reference_state_dict(sd)
    out = enc.reference_state_dict()
    assert set(out) == {k for k in sd if k.startswith("encoder.")}
    for k, v in out.items():
        assert torch.equal(v.reshape(-1), sd[k].reshape(-1).to(v.dtype)), k
    # a row of the packed local heads is the reference's nn.Linear(., 1) of that parameter
    k = [s.name for s in enc.local].index("tlag")
    assert torch.equal(enc.local_heads.weight[2 * k + 1], sd["encoder.q_local_defs.tlag.layers.log_prec.weight"][0])


def test_dense_distribution_table_matches_oracle():
    case = load_case("dr_constant_one_midpoint_f32_iw5")
    kinds = [int(k) for k in case["kinds"]]
    names = [str(n) for n in case["names"]]
    T_ = torch.as_tensor
    q = ChainedDistribution("q", names, kinds, T_(case["q_mu"]), T_(case["q_prec"]))
    p = ChainedDistribution("p", names, kinds, T_(case["p_mu"]), T_(case["p_prec"]))
    u = T_(case["u"])
    theta = p.clip(q.sample(u), stddevs=4)
    ref = O.clip_theta(O.sample_theta(u, T_(case["q_mu"]), T_(case["q_prec"]), kinds), T_(case["p_mu"]), T_(case["p_sigma"]), kinds, 4.0)
    for k, nm in enumerate(names):
        assert torch.allclose(theta.samples[nm], ref[k], rtol=2e-6, atol=1e-30), nm
        assert np.allclose(theta.samples[nm].numpy(), case["theta"][k], rtol=1e-5), nm
    assert torch.allclose(q.log_prob(theta), O.log_prob_theta(ref, T_(case["q_mu"]), T_(case["q_prec"]), kinds), rtol=1e-5, atol=1e-4)
    assert np.allclose(p.log_prob(theta).numpy(), case["log_p_theta"], rtol=1e-5, atol=1e-3)
    assert q.distributions["r"].mu.shape == (case["q_mu"].shape[0], 1)


@pytest.mark.skipif(not os.path.isdir("/root/reference/data"), reason="plate-reader CSVs only exist in the build container")
@pytest.mark.parametrize("spec,fixture", [("dr_constant_icml", "dataset_dr_icml"), ("relay_constant_precisions", "dataset_relay")])
def test_csv_loader_matches_reference_preprocessing(spec, fixture, monkeypatch):
    monkeypatch.setenv("INFERENCE_DATA_DIR", "/root/reference/data")
    cfg = _config(spec)
    pair = D.build_datasets(Args(), cfg)
    ds = pair.train.dataset
    z = np.load(os.path.join(GOLDEN, fixture + ".npz"))
    for k in ("times", "inputs", "dev_1hot", "observations", "devices"):
        assert np.array_equal(np.asarray(getattr(ds, k)), z[k]), k
    assert np.array_equal(np.asarray(pair.train.indices), z["train_ids"])
    assert np.array_equal(np.asarray(pair.test.indices), z["test_ids"])


def test_npz_dataset_and_shapes():
    """tests/test_shapes.py of the reference: 312 individuals -> 234 / 78 at 4 folds; batch shapes."""
    cfg = _config("dr_constant_icml")
    ds = D.TimeSeriesDataset.from_npz(os.path.join(GOLDEN, "dataset_dr_icml.npz"), cfg.data)
    pair = D.build_datasets(Args(), cfg, dataset=ds)
    assert (pair.n_train, pair.n_test, pair.depth, pair.n_conditions) == (234, 78, 7, 2)
    b = D.batch_of(ds, np.asarray(pair.train.indices)[:36], "cpu")
    assert b.dev_1hot.shape == (36, 7) and b.inputs.shape == (36, 2) and b.observations.shape == (36, 4, 86)


def test_multistep_lr():
    assert multistep_lr(0.01, [250, 1000], 0.2, 0) == 0.01
    assert abs(multistep_lr(0.01, [250, 1000], 0.2, 250) - 0.002) < 1e-12
    assert abs(multistep_lr(0.01, [250, 1000], 0.2, 1000) - 0.0004) < 1e-12


def test_engine_refuses_cpu_tensors():
    from vihds_b200 import models
    from vihds_b200.distributions import DotOperatorSamples

    cfg = _config("dr_constant_one")
    m = models.LOOKUP["dr_constant"](cfg)
    th = DotOperatorSamples()
    th.add("r", torch.ones(2, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.simulate(cfg, torch.linspace(0, 1, 5), th, torch.zeros(2, 2), torch.zeros(2, 1))


def test_conditioner_weight_draw_matches_reference_construction():
    """models._draw_conditioner_weight consumes the torch CPU RNG exactly like building the reference's
    DeviceConditioner (vihds/ode.py:99-116): nn.Linear default init, xavier_uniform_, normal_(2, 1.5)."""
    from torch import nn

    from vihds_b200.models import _draw_conditioner_weight

    for d in (7, 1, 12):
        torch.manual_seed(11)
        got = [_draw_conditioner_weight(d) for _ in range(3)]
        after = torch.rand(1)
        torch.manual_seed(11)
        ref = []
        for _ in range(3):
            lin = nn.Linear(d, 1, False)
            nn.init.xavier_uniform_(lin.weight)
            nn.init.normal_(lin.weight, mean=2.0, std=1.5)
            ref.append(lin.weight.detach().clone())
        assert all(torch.equal(a, b) for a, b in zip(got, ref)) and torch.equal(after, torch.rand(1))


def test_in_place_conditioner_draw_consumes_the_same_stream():
    """GraphedStep.draw_conditioner fills rows of a staging buffer in place; same RNG consumption as the module build."""
    from vihds_b200.models import _draw_conditioner_weight

    torch.manual_seed(4)
    ref = torch.cat([_draw_conditioner_weight(7) for _ in range(2)], 0)
    torch.manual_seed(4)
    host = torch.empty(2, 7)
    for row in (host[0:1], host[1:2]):
        row.uniform_(-1.0, 1.0)
        row.uniform_(-1.0, 1.0)
        row.normal_(mean=2.0, std=1.5)
    assert torch.equal(host, ref)
