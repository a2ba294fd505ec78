"""Time the reference's OWN training step -- ``Training._run_batch`` (vihds/training.py:324-340: model forward incl. the
encoder, cost, NaN check, backward, Adam step, zero_grad) -- on the host cores, through oracle/ref_harness.py.
Baseline infrastructure for ``bench.py --impl reference`` / ``cpu_baseline.kind = "reference"``; never on the product path.
"""
import os
import time

import numpy as np


class _Log(object):
    batch_feed_time = 0.0
    batch_train_time = 0.0


def available():
    import ref_harness as H

    return os.path.isdir(os.path.join(H.REFERENCE_ROOT, "vihds"))


def time_reference(spec, iw, steps, warmup, threads=None, solver=None):
    """Returns dict(sec_per_step, cores, losses, B, T, root).  One mini-batch of the spec's training loader (seed 0,
    n_batch individuals), re-used for every step -- the workload of the reference's first training iterations."""
    import torch

    import ref_harness as H

    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    args, settings, data, parameters, model, training = H.build_reference(spec, samples=iw, solver=solver)
    batch = next(iter(training.train_loader))
    model.train()
    ts, losses = [], []
    for i in range(warmup + steps):
        log = _Log()
        t0 = time.perf_counter()
        ok = training._run_batch(t0, batch, log)
        t1 = time.perf_counter()
        if not ok:
            raise RuntimeError("reference: ELBO is NaN")
        if i >= warmup:
            ts.append(t1 - t0)
    with torch.no_grad():
        res, theta, q, p = model(batch, iw)
        losses.append(float(training.cost(batch, res, theta, q, p).elbo))
    return {"sec_per_step": float(np.mean(ts)), "cores": cores, "loss": losses[-1], "B": int(len(batch.inputs)),
            "T": int(len(batch.times)), "solver": settings.params.solver, "root": H.REFERENCE_ROOT}
