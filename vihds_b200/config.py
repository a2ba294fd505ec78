"""Spec loading: the YAML files of the reference (`specs/*.yaml`) are accepted unchanged.

Mirrors the *meaning* of vihds/config.py (reference, read-only): parameter defaults (config.py:56-88), data defaults
(config.py:124-140), the device-group bookkeeping that yields ``device_depth`` and the relevance vectors
(config.py:95-121), seeding (config.py:30-35) and the dtype switch ``data.dtype`` (config.py:164-178).  The GPU box
has no copy of the reference tree, so specs can also be given as the JSON dumps under tests/golden/specs/.
"""
import json
import os
from collections import OrderedDict

import numpy as np
import torch


class Settings(dict):
    """Attribute-access dictionary (the reference uses ``munch``; only attribute and item access are relied upon)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        del self[key]


def to_settings(x):
    if isinstance(x, dict):
        return Settings((k, to_settings(v)) for k, v in x.items())
    if isinstance(x, (list, tuple)):
        return type(x)(to_settings(v) for v in x)
    return x


munchify = to_settings  # drop-in name used by reference-style drivers

PARAM_DEFAULTS = OrderedDict([
    ("solver", "midpoint"), ("adjoint_solver", False), ("use_laplace", False), ("n_filters", 10), ("filter_size", 10),
    ("pool_size", 5), ("lambda_l2", 0.001), ("lambda_l2_hidden", 0.001), ("n_hidden", 50), ("n_hidden_decoder", 50),
    ("n_batch", 36), ("data_format", "channels_last"), ("precision_type", "constant"), ("precision_alpha", 1000.0),
    ("precision_beta", 1.0), ("init_prec", 0.00001), ("init_latent_species", 0.001), ("transfer_func", "tanh"),
    ("n_hidden_decoder_precisions", 20), ("n_growth_layers", 4), ("tb_gradients", False), ("plot_histograms", False),
    ("learning_boundaries", [250, 500]), ("learning_rate", 0.01), ("learning_gamma", 0.2),
])


def _n_levels(values):
    return len({v for v in values if v is not None})


def with_param_defaults(params):
    out = to_settings(dict(PARAM_DEFAULTS))
    for k, v in params.items():
        out[k] = v
    return out


def with_data_defaults(data):
    """Data block + derived device bookkeeping.  One one-hot block per entry of ``groups`` (in order); the relevance
    vector of a group masks the other groups' blocks and the group's default level (config.py:102-115)."""
    out = to_settings({"groups": {"default": [0] * len(data["devices"])}, "default_devices": {}, "normalize": None,
                       "merge": True, "subtract_background": True, "separate_conditions": False, "dtype": "float32"})
    for k, v in data.items():
        out[k] = v
    out.data_dir = os.getenv("INFERENCE_DATA_DIR") or "data"
    out.component_maps = OrderedDict((g, OrderedDict(zip(out.devices, levels))) for g, levels in out.groups.items())
    widths = [_n_levels(cm.values()) for cm in out.component_maps.values()]
    out.device_depth = int(sum(widths))
    out.relevance_vectors = OrderedDict()
    start = 0
    for (g, _), w in zip(out.groups.items(), widths):
        rv = np.zeros(out.device_depth, np.float32)
        rv[start:start + w] = 1.0
        if g in out.default_devices:
            rv[start + out.default_devices[g]] = 0.0
        out.relevance_vectors[g] = rv
        start += w
    out.device_map = {name: float(i) for i, name in enumerate(out.devices)}
    out.device_idx_to_device_name = dict(enumerate(out.devices))
    out.device_lookup = {v: k for k, v in out.device_map.items()}
    return out


def read_spec(path):
    """Parse a spec file: YAML (reference format) or the JSON dump of one."""
    with open(path, "r") as f:
        if path.endswith(".json"):
            return json.load(f)
        import yaml

        return yaml.safe_load(f)


def seed_everything(seed):
    """config.py:30-35: numpy global RNG (sample_u, data split) and the torch RNG (weight init, device conditioner)."""
    if seed is not None:
        np.random.seed(seed)
        torch.manual_seed(seed)


def torch_dtype(name):
    if name == "float32":
        return torch.float32
    if name == "float64":
        return torch.float64
    raise Exception("Unknown dtype %s" % name)


class Config(object):
    """``Config(args)`` with the reference's attribute surface: ``.data .params .model .seed .device .trainer``.

    ``args`` needs ``yaml`` and may carry ``seed``, ``gpu``, ``precision_hidden_layers`` (run_xval.py:17-57).  The
    engine is CUDA-only: the device is ``cuda:<gpu or current>``; a missing GPU is an error at first kernel launch,
    never a CPU fallback."""

    def __init__(self, args=None, spec=None, device=None):
        seed = getattr(args, "seed", None) if args is not None else None
        seed_everything(seed)
        if spec is None:
            spec = read_spec(args.yaml)
        spec = to_settings(spec)
        self.data = with_data_defaults(spec.data)
        self.params = with_param_defaults(spec.params)
        hidden = getattr(args, "precision_hidden_layers", None) if args is not None else None
        if hidden is not None:
            self.params.n_hidden_decoder_precisions = hidden
        self.model = spec.model
        self.seed = seed
        if device is None:
            gpu = getattr(args, "gpu", None) if args is not None else None
            device = torch.device("cuda", gpu if gpu is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0))
        self.device = torch.device(device)
        self.dtype = torch_dtype(self.data.dtype)
        self.trainer = None
