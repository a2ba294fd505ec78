"""CPU: inference-graph staging and prior propagation (vihds/inference_graph.py, vihds/run_inference_graph.py:28-67).
Where the reference tree is available (build container: /root/reference, or the vendored oracle/_ref copy) the
reference's own ``propagate_params`` / ``create_inference_graph`` run on the same files and must agree exactly; the
numbers are also pinned here so the test means something on a box without it."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from vihds_b200 import inference_graph as G
from vihds_b200.config import Config

GRAPH = {
    "nodes": {
        "auto": {"spec": "specs/auto_constant_precisions.yaml", "experiment": "auto_prec", "seed": 0},
        "prpr": {"spec": "specs/prpr_constant_precisions.yaml", "experiment": "prpr_prec", "seed": 0},
        "inducer": {"spec": "specs/inducer_constant_precisions.yaml", "experiment": "inducer_prec", "seed": 0},
        "degrader": {"spec": "specs/degrader_constant_precisions.yaml", "experiment": "degrader_prec", "seed": 0},
    },
    "edges": [
        {"from": {"node": "auto", "parameter": "a480"}, "to": {"node": "prpr", "parameter": "a480"}},
        {"from": {"node": "auto", "parameter": "drfp"}, "to": {"node": "prpr", "parameter": "drfp"}},
        {"from": {"node": "prpr", "parameter": "dyfp"}, "to": {"node": "inducer", "parameter": "dyfp"}},
        {"from": {"node": "inducer", "parameter": "nA"}, "to": {"node": "degrader", "parameter": "nA"}},
        {"from": {"node": "inducer", "parameter": "KAra"}, "to": {"node": "degrader", "parameter": "KAra"}},
    ],
}


class Args:
    seed, gpu, precision_hidden_layers, yaml = 0, None, None, None
    folds, split, heldout = 4, 1, None


def _write_results(folder):
    names = ["nA.mu", "nA.prec", "KAra.mu", "KAra.prec", "r.mu", "r.prec"]
    rng = np.random.RandomState(0)
    vals = [rng.rand(4) + 0.5, rng.rand(4) * 3 + 1, rng.rand(4) + 2, rng.rand(4) * 5 + 0.1,
            rng.rand(4, 7), rng.rand(4, 7)]  # local parameters are ragged rows in the reference's object array
    G.save_q_results(folder, names, vals)
    return names, vals


def test_staging_follows_the_edges():
    nodes = G.create_inference_graph(GRAPH, "g")
    assert {k: n.stage for k, n in nodes.items()} == {"auto": 0, "prpr": 1, "inducer": 2, "degrader": 3}
    stages = G.arrange_by_stage(nodes.values())
    assert [n.name for n in stages[3]] == ["degrader"] and nodes["prpr"].args.experiment == "g/prpr_prec"
    assert [e.sourceParam for e in nodes["degrader"].incoming] == ["nA", "KAra"]
    with pytest.raises(ValueError):
        G.create_inference_graph({"nodes": {"x": {"spec": "a.yaml"}}, "edges": []})


def test_pooled_precision_and_propagation(tmp_path):
    assert abs(G.pooled_prec([1.0, 2.0, 4.0]) - 3 / (1 + 0.5 + 0.25)) < 1e-15
    nodes = G.create_inference_graph(GRAPH, "g")
    folder = str(tmp_path / "inducer_run")
    names, vals = _write_results(folder)
    with open(os.path.join(GOLDEN, "specs", "degrader_constant_precisions.json")) as f:
        settings = Config(Args(), spec=json.load(f), device="cpu")
    before = dict(settings.params["global"]["eA"])
    G.propagate_params(nodes["degrader"], settings, {"inducer": folder})
    nA = settings.params["global"]["nA"]
    assert nA["distribution"] == "LogNormal" and abs(nA["mu"] - np.mean(vals[0])) < 1e-15
    assert abs(nA["sigma"] - 4 / np.sum(1 / vals[1])) < 1e-12  # the pooled precision, stored under `sigma` as in the reference
    assert abs(settings.params["global"]["KAra"]["mu"] - np.mean(vals[2])) < 1e-15
    assert dict(settings.params["global"]["eA"]) == before  # parameters without an incoming edge are untouched
    # pinned numbers (RandomState(0) above)
    assert abs(nA["mu"] - 1.1029123573) < 1e-9 and abs(nA["sigma"] - 2.6932133145) < 1e-9


def test_propagation_equals_the_reference(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_harness as H

    if not os.path.isdir(os.path.join(H.REFERENCE_ROOT, "vihds")):
        pytest.skip("reference tree not present")
    H.import_reference()
    import munch
    import vihds.run_inference_graph as R

    nodes = G.create_inference_graph(GRAPH, "g")
    folder = str(tmp_path / "inducer_run")
    _write_results(folder)
    with open(os.path.join(GOLDEN, "specs", "degrader_constant_precisions.json")) as f:
        spec = json.load(f)
    mine = Config(Args(), spec=json.loads(json.dumps(spec)), device="cpu")
    G.propagate_params(nodes["degrader"], mine, {"inducer": folder})
    theirs = munch.munchify({"params": json.loads(json.dumps(spec))["params"]})
    R.propagate_params(nodes["degrader"], theirs, {"inducer": folder})
    for key in ("nA", "KAra"):
        a, b = mine.params["global"][key], theirs.params["global"][key]
        assert a["distribution"] == b["distribution"] and a["mu"] == b["mu"] and a["sigma"] == b["sigma"]
