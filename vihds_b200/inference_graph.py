"""Inference graphs: the posterior of one node's parameter becomes the prior of another node's.

Host-side mirror of vihds/inference_graph.py (Edge / Node / staging, :10-126) and of the prior propagation of
vihds/run_inference_graph.py:28-67 (``pooled_prec``, ``propagate_params``).  No kernel content: this is the data format
either side of the hot path -- a node's run leaves ``xval_q_values.npy`` / ``xval_q_names.txt``
(vihds/xval.py:107-111), the next node's spec gets its priors rewritten from them before its encoder / prior tables are
built.  The reference's behaviour is kept as it is, including that the propagated dictionary stores the pooled
PRECISION under the key ``sigma`` and always names the distribution ``LogNormal`` (run_inference_graph.py:48-66).
"""
import os

import numpy as np
import yaml

from .config import Settings


class Edge(object):
    def __init__(self, source, sourceParam, target, targetParam):
        self.source, self.sourceParam, self.target, self.targetParam = source, sourceParam, target, targetParam


class Node(object):
    """``args``: the node's run arguments (spec, experiment, seed, ... as given in the graph yaml, inference_graph.py:17-60)."""

    def __init__(self, name, yamlargs, graph_name):
        for need in ("spec", "experiment"):
            if need not in yamlargs:
                raise ValueError("Node " + name + " missing " + need + " property")
        self.name, self.stage, self.incoming, self.outgoing = name, None, [], []
        args = dict(yamlargs)
        args["experiment"] = graph_name + "/" + str(yamlargs["experiment"])
        args["yaml"] = args.pop("spec")
        self.args = Settings(**args)

    def addIncomingEdge(self, edge):
        self.incoming.append(edge)

    def addOutgoingEdge(self, edge):
        self.outgoing.append(edge)


def set_stage(node):
    """A node without incoming edges runs at stage 0, any other one stage after its latest source (inference_graph.py:81-94)."""
    if node.stage is None:
        for e in node.incoming:
            set_stage(e.source)
        node.stage = 1 + max(e.source.stage for e in node.incoming) if node.incoming else 0
    return node.stage


def create_inference_graph(graph, graph_name="unnamed"):
    """``graph``: path of a graph yaml (inferencegraphs/*.yaml) or the parsed dict.  Returns {name: Node}."""
    if isinstance(graph, str):
        with open(graph) as f:
            graph = yaml.safe_load(f)
    nodemap = {k: Node(k, v, graph_name) for k, v in graph["nodes"].items()}
    for edge in graph.get("edges") or []:
        source, target = nodemap[edge["from"]["node"]], nodemap[edge["to"]["node"]]
        e = Edge(source, edge["from"]["parameter"], target, edge["to"]["parameter"])
        source.addOutgoingEdge(e)
        target.addIncomingEdge(e)
    for node in nodemap.values():
        set_stage(node)
    return nodemap


def arrange_by_stage(nodes):
    """{stage: [nodes that can run in parallel at that stage]} (inference_graph.py:115-126)."""
    stagemap = {}
    for node in nodes:
        stagemap.setdefault(node.stage, []).append(node)
    return stagemap


def pooled_prec(xarr):
    """Harmonic mean of the folds' precisions (run_inference_graph.py:28-33)."""
    den = 0
    for x in xarr:
        den = den + (1 / x)
    return len(xarr) / den


def save_q_results(folder, q_names, q_values):
    """What a node's run leaves for its successors (vihds/xval.py:107-111): one row of q_values per name, one column
    per fold."""
    os.makedirs(folder, exist_ok=True)
    with open(os.path.join(folder, "xval_q_names.txt"), "w") as f:
        for n in q_names:
            f.write("%s\n" % n)
    arr = np.empty(len(q_values), dtype=object)  # ragged rows (local parameters have one value per individual)
    for i, v in enumerate(q_values):
        arr[i] = np.asarray(v)
    np.save(os.path.join(folder, "xval_q_values.npy"), arr, allow_pickle=True)


def propagate_params(node, settings, resultmap, verbose=False):
    """Rewrite the priors of ``node``'s spec from its source nodes' posteriors (run_inference_graph.py:36-67): mean of the
    folds' ``<param>.mu``, pooled ``<param>.prec``; the target entry in params.global / local / shared is replaced."""
    for incoming in node.incoming:
        folder = resultmap[incoming.source.name]
        xval = np.load(os.path.join(folder, "xval_q_values.npy"), allow_pickle=True)
        with open(os.path.join(folder, "xval_q_names.txt")) as f:
            labels = [line.rstrip() for line in f]
        avgmu = np.mean(xval[labels.index(incoming.sourceParam + ".mu")])
        prec = pooled_prec(xval[labels.index(incoming.sourceParam + ".prec")])
        for key in ("global", "local", "shared"):
            group = settings.params.get(key) if hasattr(settings.params, "get") else getattr(settings.params, key, None)
            if group is not None and incoming.targetParam in group:
                if verbose:
                    print("%s: %s <- %s.%s (in params.%s)" % (node.name, incoming.targetParam, incoming.source.name,
                                                             incoming.sourceParam, key))
                group[incoming.targetParam] = Settings(distribution="LogNormal", mu=avgmu, sigma=prec)
    return settings
