"""Time-series datasets: CSV plate-reader files -> device one-hots, log1p treatments, scaled observations [L, 4, T].

Takes over vihds/datasets.py and data/procdata.py (reference).  Host-side I/O only; kept
so that a spec runs end to end on current numpy/pandas -- the reference's ``merge_observations`` builds a ragged
``np.asarray`` (datasets.py:137-138) that numpy >= 1.24 rejects for every multi-file spec.  Semantics preserved:
rows are kept when their device is listed in the spec (procdata.py:152); a row is dropped when it has a non-zero
treatment outside ``data.conditions`` (procdata.py:58-66); files are merged onto the SHORTEST time grid by
nearest-time selection (datasets.py:136-145); each signal is divided by its global maximum and each series shifted
to a zero minimum (datasets.py:48-61); treatments become log(1 + c) (datasets.py:87); folds are contiguous chunks of
one seeded permutation (datasets.py:205-216).

The GPU box has no data directory: ``TimeSeriesDataset.from_npz`` loads the pre-processed fixture
tests/golden/dataset_*.npz that tests/golden/make_golden.py dumped from the reference.
"""
import os
import re
from collections import OrderedDict

import numpy as np
import torch
from torch.utils.data import Dataset, Subset

from .config import Settings


def _np_dtype(settings):
    return np.float64 if settings.dtype == "float64" else np.float32


def parse_treatment(cell):
    """'C6=25000;EtOH=1' -> OrderedDict (procdata.py:15-29)."""
    out = OrderedDict()
    if "=" in cell:
        for item in cell.split(";"):
            k, v = item.split("=")
            out[k] = float(v)
    return out


def signal_of(header):
    """'Raw Data (OD) 3 - 0 h 23 min' -> 'OD' (procdata.py:69-80; pandas de-duplication suffixes removed first)."""
    header = header.split(".")[0]
    m = re.search(r"\(([^)]*)\)", header)
    return m.group(1) if m else header


def load_csv(csv_file, settings):
    """One plate file -> (devices [L] int, treatments [L, C], times [T], observations [L, n_signals, T])."""
    import pandas as pd

    table = pd.read_csv(os.path.join(settings.data_dir, csv_file), sep=",", na_filter=False, header=None, dtype=str)
    headers = [signal_of(h) for h in table.iloc[0, 5:]]
    all_times = table.iloc[1, 5:].to_numpy()
    body = table.iloc[2:, :]
    body = body[np.isin(body.iloc[:, 0].to_numpy(), settings.devices)]
    if len(body) == 0:
        return None
    parsed = [parse_treatment(c) for c in body.iloc[:, 4]]
    keep = [i for i, tr in enumerate(parsed) if all(v == 0.0 for k, v in tr.items() if k not in settings.conditions)]
    t = _np_dtype(settings)
    treatments = np.array([[parsed[i].get(c, 0.0) for c in settings.conditions] for i in keep], dtype=t)
    devices = np.array([settings.device_map[d] for d in body.iloc[keep, 0]], dtype=int)
    headers = np.array(headers)
    values = body.iloc[keep, 5:].to_numpy()
    obs = np.stack([values[:, headers == s].astype(np.float64) for s in settings.signals], axis=1).astype(t)
    times = all_times[headers == "OD"].astype(np.float64).astype(t)
    return devices, treatments, times, obs


def merge_on_shortest_grid(times_list, observations_list):
    """datasets.py:136-145 (ragged-safe): choose the file with the fewest time points, pick from every other file
    the sample nearest to each chosen time."""
    shortest = int(np.argmin([len(t) for t in times_list]))
    grid = times_list[shortest]
    merged = []
    for t, obs in zip(times_list, observations_list):
        nearest = [int(np.abs(np.asarray(t) - ti).argmin()) for ti in grid]
        merged.append(obs[:, :, nearest])
    return grid, np.concatenate(merged)


def device_one_hots(devices, settings):
    """datasets.py:25-46: concatenated one-hot blocks, one block per device group."""
    rows = []
    for d in devices:
        name = settings.device_idx_to_device_name[int(d)]
        blocks = []
        for cm in settings.component_maps.values():
            width = len({v for v in cm.values() if v is not None})
            block = np.zeros(width)
            if cm[name] is not None:
                block[cm[name]] = 1
            blocks.append(block)
        rows.append(np.hstack(blocks))
    return np.array(rows).astype(_np_dtype(settings))


def scale_observations(X, settings):
    """datasets.py:48-61 (in place): per-signal global max scaling, then per-series minimum subtraction."""
    n_out = X.shape[1]
    scales = [np.max(X[:, i, :]).astype(np.float32) for i in range(n_out)] if settings.normalize is None else settings.normalize
    for i, scale in enumerate(scales):
        X[:, i, :] /= scale
        if settings.subtract_background:
            X[:, i, :] -= np.min(X[:, i, :], axis=1)[:, np.newaxis]
    return X, scales


class TimeSeriesDataset(Dataset):
    """L individuals sharing one time grid.  Fields as in the reference (datasets.py:64-119)."""

    def __init__(self, data_settings=None, params=None):
        self.data_settings, self.params = data_settings, params
        self.n_times = self.n_species = None

    def _set(self, devices, dev_1hot, inputs, times, observations, scales=None):
        self.devices = np.asarray(devices)
        self.dev_1hot = torch.as_tensor(dev_1hot)
        self.inputs = torch.as_tensor(inputs)
        self.times = torch.as_tensor(times)
        self.observations = torch.as_tensor(observations)
        self.scales = scales
        self.n_times = len(times)
        self.n_species = int(self.observations.shape[1])
        return self

    def _preprocess(self, devices, treatments, times, observations):
        obs, scales = scale_observations(observations, self.data_settings)
        return self._set(devices, device_one_hots(devices, self.data_settings), np.log(1.0 + treatments), times, obs, scales)

    def init_single(self, f):
        return self._preprocess(*load_csv(f, self.data_settings))

    def init_multiple_merge(self):
        loaded = [r for r in (load_csv(f, self.data_settings) for f in self.data_settings.files) if r is not None]
        devices, treatments, times_list, obs_list = zip(*loaded)
        times, obs = merge_on_shortest_grid(times_list, obs_list)
        return self._preprocess(np.concatenate(devices), np.concatenate(treatments), times, obs)

    @classmethod
    def from_arrays(cls, devices, dev_1hot, inputs, times, observations, data_settings=None):
        return cls(data_settings)._set(devices, dev_1hot, inputs, times, observations)

    @classmethod
    def from_npz(cls, path, data_settings=None):
        z = np.load(path)
        return cls.from_arrays(z["devices"], z["dev_1hot"], z["inputs"], z["times"], z["observations"], data_settings)

    def __len__(self):
        return len(self.devices)

    def __getitem__(self, idx):
        if torch.is_tensor(idx):
            idx = idx.tolist()
        return {"devices": self.devices[idx], "dev_1hot": self.dev_1hot[idx], "inputs": self.inputs[idx],
                "observations": self.observations[idx]}


class TimeSeriesDatasetPair(object):
    """datasets.py:148-172."""

    def __init__(self, train_dataset, test_dataset, data_settings):
        self.train, self.test = train_dataset, test_dataset
        self.n_train, self.n_test = len(train_dataset), len(test_dataset)
        self.depth = data_settings.device_depth
        self.n_conditions = len(data_settings.conditions)


def split_folds(n, folds, split, seed):
    """datasets.py:205-216: one seeded permutation cut into ``folds`` chunks; chunk ``split`` (1-based) validates."""
    np.random.seed(seed)
    chunks = np.array_split(np.random.permutation(n), folds)
    val_ids = np.sort(chunks[split - 1])
    return np.setdiff1d(np.arange(n, dtype=int), val_ids), val_ids


def build_datasets(args, config, dataset=None):
    """``build_datasets(args, config)`` of the reference; ``dataset`` lets a pre-processed TimeSeriesDataset (e.g. the
    npz fixture) stand in for the CSV directory."""
    if getattr(args, "heldout", None):
        raise NotImplementedError("TODO: implement heldout device")
    if dataset is None:
        if not config.data.merge:
            raise NotImplementedError("TODO: Enable non-merged time-series data")
        dataset = TimeSeriesDataset(config.data, config.params).init_multiple_merge()
    train_ids, val_ids = split_folds(len(dataset), args.folds, args.split, args.seed)
    return TimeSeriesDatasetPair(Subset(dataset, train_ids), Subset(dataset, val_ids), config.data)


def batch_of(dataset, ids, device, dtype=None):
    """Batch container (training.py:47-68) for individuals ``ids`` of a TimeSeriesDataset, on ``device``."""
    mv = lambda t: t.to(device=device, dtype=dtype) if dtype is not None else t.to(device)  # noqa: E731
    item = dataset[ids]
    return Settings(devices=torch.as_tensor(np.asarray(item["devices"])), dev_1hot=mv(item["dev_1hot"]),
                    inputs=mv(item["inputs"]), observations=mv(item["observations"]), times=mv(dataset.times))
