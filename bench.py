#!/usr/bin/env python
"""ELBO-gradient step benchmark (BASELINE.json metric: trajectories/s and steps/s of the ELBO-grad step).

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's own Training._run_batch on the host
                                                                    # cores (oracle/_ref; the oracle port if absent)

One *step* = the body of the reference's ``Training._run_batch`` (vihds/training.py:329-337): encoder forward, theta
sampling/clipping, fixed-step ODE solve, observation log-likelihood, log p / log q, IWAE cost, backward to every
trainable parameter, gradient all-reduce (N > 1), Adam.  Workload at N = 1 (default): ``specs/dr_constant_icml.yaml``,
batch = 36 individuals x IW = 200 samples = 7,200 trajectories, T = 86, midpoint solver, fp32 (BASELINE.json
configs[1]); real pre-processed plate data (tests/golden/dataset_dr_icml.npz), random-init weights at seed 0.  For
N > 1 every rank takes its own 36 individuals (weak scaling; the only exchange is the gradient all-reduce).

Prints ONE JSON line (rank 0).  Keys beyond the base contract: ``roofline`` (reverse-sweep kernel, the dominant
launch), ``cpu_baseline`` (the reference / the oracle port on the host cores), ``e2e`` (same step fed from pinned HOST
buffers through the public API, H2D + D2H inside the timed region), ``kernels`` (per-launch device times of the hot
launches) and ``workloads``: the other BASELINE.json configs (synthetic slab B=1024 x IW=128 x T=500 per GPU,
dr_blackbox_icml, relay_constant_precisions), ~30 steps each, with ms_per_step, traj/s and the forward / reverse
launches' achieved fraction of the measured HBM peak.  ``--workload X`` makes X the top-level workload;
``--no-extra-workloads`` skips the others.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

WORKLOADS = {
    # name: (spec, dataset fixture, B per GPU, IW, T or None (= dataset grid))
    "dr_constant_icml": ("dr_constant_icml", "dataset_dr_icml", 36, 200, None),
    "relay_constant_precisions": ("relay_constant_precisions", "dataset_relay", 36, 200, None),
    "dr_blackbox_icml": ("dr_blackbox_icml", "dataset_dr_icml", 36, 200, None),
    "synthetic_dr_constant": ("dr_constant_icml", "dataset_dr_icml", 1024, 128, 500),
}


class Args(object):
    """Stand-in for the reference's argparse namespace (run_xval.py:17-57)."""

    def __init__(self, **kw):
        self.seed, self.gpu, self.precision_hidden_layers, self.yaml = 0, None, None, None
        self.folds, self.split, self.heldout, self.verbose = 4, 1, None, False
        self.train_samples, self.test_samples, self.epochs, self.test_epoch = 200, 1000, 1, 0
        self.__dict__.update(kw)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(B, IW, T, S, P, E, elem=4):
    """SURVEY.md section 8d, for the launches as the training step issues them (no x_predict trace):
    forward  reads u [N,P], obs [B,4,T], extras [E,N], q tables; writes theta [P,N], x_states [T,S,N], 6 terms per n;
    reverse  reads x_states, u, obs, extras, 6 upstream grads per n; writes d_q [B,P] x 2."""
    N = B * IW
    fwd = elem * (N * P + B * 4 * T + E * N + 2 * B * P + N * P + N * T * S + 6 * N)
    bwd = elem * (N * T * S + N * P + B * 4 * T + E * N + 2 * B * P + 6 * N + 2 * B * P)
    return fwd, bwd


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
            except ValueError:
                continue
            for nm, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# workload construction
# ---------------------------------------------------------------------------------------------------------------
def load_spec(name):
    with open(os.path.join(GOLDEN, "specs", name + ".json")) as f:
        return json.load(f)


def synthetic_individuals(ds, B, T, rng):
    """B synthetic individuals on a uniform T-point grid over [0, 16.5] h.  Treatments:
    exactly one non-zero inducer per individual, concentration in {25000/3^k} or 0, log1p-transformed; devices
    uniform over the icml devices; observations: real curves of the same device resampled onto the grid (values in
    [0, 1] after the reference's scaling) plus N(0, 0.01^2) noise."""
    import torch

    times = np.linspace(0.0, 16.5, T).astype(np.float32)
    idx = rng.randint(0, len(ds), size=B)
    conc = np.concatenate([[0.0], 25000.0 / 3.0 ** np.arange(11)])
    inputs = np.zeros((B, 2), np.float32)
    inputs[np.arange(B), rng.randint(0, 2, size=B)] = np.log1p(conc[rng.randint(0, len(conc), size=B)]).astype(np.float32)
    t0 = ds.times.numpy()
    obs0 = ds.observations.numpy()[idx]
    obs = np.empty((B, 4, T), np.float32)
    for o in range(4):
        for b in range(B):
            obs[b, o] = np.interp(times, t0, obs0[b, o])
    obs = np.clip(obs + rng.randn(B, 4, T).astype(np.float32) * 0.01, 0.0, None)
    return {"times": torch.as_tensor(times), "inputs": torch.as_tensor(inputs), "dev_1hot": ds.dev_1hot[idx].clone(),
            "observations": torch.as_tensor(obs), "devices": torch.as_tensor(ds.devices[idx])}


def build_workload(name, rank, world, device, B_override=None, IW_override=None, with_training=True):
    import torch

    from vihds_b200.config import Config
    from vihds_b200.datasets import TimeSeriesDataset, build_datasets
    from vihds_b200.parameters import Parameters
    from vihds_b200.training import Training
    from vihds_b200.vae import build_model

    spec_name, fixture, B, IW, T = WORKLOADS[name]
    B, IW = B_override or B, IW_override or IW
    args = Args(train_samples=IW)
    settings = Config(args, spec=load_spec(spec_name), device=device)
    ds = TimeSeriesDataset.from_npz(os.path.join(GOLDEN, fixture + ".npz"), settings.data)
    pair = build_datasets(args, settings, dataset=ds)
    parameters = Parameters(settings.params)
    torch.manual_seed(0)
    model = build_model(args, settings, pair, parameters)
    training = Training(args, settings, pair, parameters, model) if with_training else None
    rng = np.random.RandomState(1234 + rank)
    if T is None:
        ids = np.asarray(pair.train.indices)
        order = np.random.RandomState(0).permutation(len(ids))
        take = ids[np.take(order, np.arange(rank * B, (rank + 1) * B), mode="wrap")]
        item = ds[take]
        host = {"times": ds.times, "inputs": item["inputs"], "dev_1hot": item["dev_1hot"], "observations": item["observations"]}
    else:
        host = synthetic_individuals(ds, B, T, rng)
        if ds.n_times != T:  # the encoder's hidden layer is sized by T: rebuild it for the synthetic grid
            from vihds_b200.encoders import Encoder

            torch.manual_seed(0)
            model.encoder = Encoder(parameters, (4, T, pair.n_conditions, pair.depth)).to(device=device, dtype=settings.dtype)
            training = Training(args, settings, pair, parameters, model) if with_training else None
    host = {k: v.to(settings.dtype).contiguous() for k, v in host.items() if k != "devices"}
    return settings, parameters, model, training, host, B, IW, int(host["times"].numel()), rng


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (plain PyTorch CPU restatement of the reference algorithm at its own granularity)
# ---------------------------------------------------------------------------------------------------------------
def cpu_case(parameters, model, host, B, IW, rng, solver):
    """Golden-style case dict for oracle/vihds_oracle.py from the bench workload (q from the freshly initialised
    encoder, evaluated on the CPU copy of the batch)."""
    import torch

    from vihds_b200.config import Settings

    enc = model.encoder
    dev = enc.global_free.device  # (Encoder.parameters is the spec table, as in the reference, not nn.Module.parameters)
    with torch.no_grad():
        q_table = enc.q_table if dev.type == "cuda" else enc.q_table_reference  # CPU arm: the stock-PyTorch restatement
        mu, prec = q_table(Settings(observations=host["observations"].to(dev), inputs=host["inputs"].to(dev),
                                    dev_1hot=host["dev_1hot"].to(dev)))
    p_mu, p_prec, _, _ = parameters.prior_arrays(np.float32)
    ode = model.decoder.ode_model
    case = {
        "dtype": "float32", "model": ode.kernel_model, "solver": solver, "names": np.array(parameters.names),
        "kinds": parameters.kinds(), "u": rng.randn(B, IW, parameters.n_theta).astype(np.float32),
        "times": host["times"].numpy(), "inputs": host["inputs"].numpy(), "dev_1hot": host["dev_1hot"].numpy(),
        "observations": host["observations"].numpy(), "q_mu": mu.float().cpu().numpy(), "q_prec": prec.float().cpu().numpy(),
        "p_mu": p_mu, "p_prec": p_prec, "p_sigma": (1.0 / np.sqrt(p_prec)).astype(np.float32),
    }
    if model.decoder.condition_on_device and ode.kernel_model != "dr_blackbox":
        for nm in ode.conditioned:
            case["cond_" + nm] = (1.0 + np.abs(rng.randn(B, IW))).astype(np.float32)
    for name, w in model.decoder.named_parameters():
        case["w:" + name] = w.detach().float().cpu().numpy()
    case["params"] = dict(model.decoder.config.params)
    return case


def workload_config(name, B, IW, T, S, P, solver, world):
    """The ``config`` object of the JSON line: the workload only, identical for the CUDA arm and the reference arm."""
    return {"workload": name, "spec": WORKLOADS[name][0], "batch_per_gpu": B, "global_batch": B * world, "iw": IW,
            "trajectories_per_step": B * IW * world, "T": T, "state_width": S, "n_theta": P, "solver": solver,
            "parallelism": "dp%d (individuals sharded over ranks, one gradient sum per step)" % world,
            "l2": "CUDA arm: L2 flushed between timed steps (256 MiB memset)"}


def state_width(model):
    ode = model.decoder.ode_model
    return int(ode.n_species + getattr(ode, "n_latent_species", 0) + (4 if ode.precisions.dynamic else 0))


def time_reference_itself(spec, IW, steps, warmup):
    """The reference's own Training._run_batch from oracle/_ref (None if that copy is absent)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_timing
    except ImportError:
        return None
    if not ref_timing.available():
        return None
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):  # the reference prints its argument parsing
        return ref_timing.time_reference(spec, IW, steps, warmup)


def time_oracle(case, steps, warmup):
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vihds_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = O.elbo_step(case, requires_grad=True)
        t1 = time.perf_counter()
        if i >= warmup:
            ts.append(t1 - t0)
    return float(np.mean(ts)), cores, float(out["loss"])


# ---------------------------------------------------------------------------------------------------------------
# the CUDA arm, one workload
# ---------------------------------------------------------------------------------------------------------------
def global_dev_1hot(name, ds, pair, B, world):
    """Device one-hot rows of the GLOBAL batch (rank r holds rows [r*B, (r+1)*B)), the same on every rank: the device
    conditioner indexes the global batch (vihds/ode.py:46-58 quirk), see GraphedStep.load_global_devices."""
    if WORKLOADS[name][4] is None:
        ids = np.asarray(pair.train.indices)
        order = np.random.RandomState(0).permutation(len(ids))
        take = ids[np.take(order, np.arange(0, world * B), mode="wrap")]
        return ds[take]["dev_1hot"]
    import torch

    return torch.cat([ds.dev_1hot[np.random.RandomState(1234 + r).randint(0, len(ds), size=B)] for r in range(world)])


def run_workload(name, a, rank, local_rank, world, device, pg, primary):
    import ctypes as C

    import torch

    from vihds_b200 import _lib as L
    from vihds_b200.engine import _ptr, _stream
    from vihds_b200.training import GraphedStep

    steps, warmup = (a.steps, a.warmup) if primary else (min(a.steps, 30), min(max(a.warmup, 3), 5))
    settings, parameters, model, training, host, B, IW, T, rng = build_workload(
        name, rank, world, device, a.batch if primary else None, a.iw if primary else None)
    P, N = parameters.n_theta, B * IW
    model.want_predict = False
    gs = GraphedStep(training, B, IW, T, b_total=B * world, process_group=pg, use_graphs=not a.no_graphs, b_offset=rank * B)
    if world > 1 and gs.rel:
        gs.load_global_devices(global_dev_1hot(name, training.dataset_pair.train.dataset, training.dataset_pair, B, world).to(device))
    pinned = {k: v.pin_memory() for k, v in host.items()}
    gs.load_batch(pinned)
    n_pool = 4 if N * P * 4 < (64 << 20) else 2
    u_host = [torch.from_numpy(rng.randn(B, IW, P).astype(np.float32)).to(settings.dtype).pin_memory() for _ in range(n_pool)]
    u_dev = [u.to(device) for u in u_host]
    gs.load_u(u_dev[0])
    gs.draw_conditioner()
    gs.prepare()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2
    cost_host = torch.zeros(1, dtype=settings.dtype).pin_memory()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def run_untimed(k):
        for j in range(k):
            gs.load_u(u_dev[j % n_pool])
            gs.step()
        torch.cuda.synchronize()

    # clock ramp: an idle B200 sits at ~120 MHz SM clock and needs a few hundred ms of load to reach its boost
    # clocks; spin untimed steps for ~a.spin seconds first (on top of the W warm-up steps)
    # The clock-ramp steps train on ONE mini-batch for up to ~0.3 s; the reference's own loop on a single relay mini-batch
    # diverges geometrically from about its 30th step (ELBO doubles per step; measured with oracle/ref_timing.py), so the
    # measured steps start again from the initial parameters and optimiser state: what is timed is the first W + K
    # training steps of the run, as in the reference arm.
    snap0 = training.optimizer.state_snapshot()
    run_untimed(3)
    t_spin = time.perf_counter()
    run_untimed(5)
    per_step = (time.perf_counter() - t_spin) / 5
    n_spin = int(min(5000, max(0, (a.spin if primary else 0.3) / max(per_step, 1e-6))))
    if world > 1:  # every rank must issue the SAME number of steps (each one contains the gradient exchange)
        t = torch.tensor([n_spin], dtype=torch.int64, device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        n_spin = int(t.item())
    run_untimed(n_spin)
    training.optimizer.restore_state(snap0)
    # device-resident pass: inputs already in HBM, CUDA-event timing, L2 flushed between steps
    for i in range(warmup):
        gs.load_u(u_dev[i % n_pool])
        gs.draw_conditioner()
        gs.step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    wall0 = time.perf_counter()
    for i in range(steps):
        flush.zero_()
        ev[i][0].record()
        gs.load_u(u_dev[i % n_pool])
        gs.draw_conditioner()
        gs.step()
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    step_ms = np.array([s.elapsed_time(e) for s, e in ev])
    # the same steps once more with events around the reverse-sweep launch (the roofline's kernel, timed inside the step:
    # same stream, same cache state; the step is issued eagerly for this -- an event cannot sit inside the replayed graph)
    final_cost = float(gs.buf.cost.item())
    skipped = gs.skipped_steps()
    training.optimizer.restore_state(snap0)
    n_k = min(steps, 50)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_k)]
    for i in range(n_k):
        flush.zero_()
        gs.load_u(u_dev[i % n_pool])
        gs.draw_conditioner()
        gs.ev_hot = kev[i]
        gs.step()
    gs.ev_hot = None
    barrier()
    bwd_ms = np.array([s.elapsed_time(e) for s, e in kev])
    total_ms = float(step_ms.sum())
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / steps
    value = N * world / (ms_per_step * 1e-3)
    timed_out = bool(gs.exchange is not None and gs.exchange.timed_out())
    identical = None
    if world > 1:  # replicas must hold bit-identical parameters after the timed steps
        flat = training.optimizer.flat
        h = flat.view(torch.int32).to(torch.int64).sum().reshape(1)
        hs = [torch.empty_like(h) for _ in range(world)]
        torch.distributed.all_gather(hs, h)
        identical = all(int(x.item()) == int(hs[0].item()) for x in hs)
    if timed_out:
        raise RuntimeError("gradient exchange: a peer's flag did not arrive (vh_adam_allreduce_step timed out)")

    res = {"ms_per_step": ms_per_step, "value": value, "steps": steps, "warmup": warmup, "wall_s_timed_region": round(wall, 6),
           "cost_after_last_step": final_cost, "skipped_steps_nan_guard": skipped}
    if world > 1:
        res["params_identical_across_ranks"] = identical
        res["exchange_timed_out"] = timed_out
        res["exchange"] = "fused with Adam over NVLink peer memory" if gs.exchange is not None else "ncclAllReduce"

    # end-to-end pass (primary workload): the public step fed from pinned HOST buffers -- H2D of the batch + u, D2H of the
    # cost and the host's isnan(elbo) check for every step (training.py:331).
    #   e2e.value            a STREAM of K steps, the way an epoch runs: the inputs of step i + 1 are copied while the device
    #                        computes step i (two input sets, GraphedStep.step_from_host), the cost of step i is looked at
    #                        while step i + 1 runs; wall clock around the K steps, max over ranks
    #   e2e.latency_ms       one step alone: call -> cost on the host, device idle before and L2 flushed (nothing overlaps)
    if primary:
        h2d = sum(v.numel() * v.element_size() for v in pinned.values()) + u_host[0].numel() * u_host[0].element_size()
        h2d += gs.cond_w.numel() * gs.cond_w.element_size() if gs.extras else 0
        lat = []
        for i in range(warmup + min(steps, 30)):
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            cost = gs.step_from_host(pinned, u_host[i % n_pool])
            cost_host.copy_(cost, non_blocking=True)
            torch.cuda.synchronize()
            if torch.isnan(cost_host).any():
                raise RuntimeError("ELBO is NaN")
            t1 = time.perf_counter()
            if i >= warmup:
                lat.append(t1 - t0)
        ring = [torch.zeros(1, dtype=settings.dtype).pin_memory() for _ in range(2)]
        rev = [torch.cuda.Event(), torch.cuda.Event()]

        def look(k):  # the cost of step k is on the host once its event has passed
            rev[k & 1].synchronize()
            if torch.isnan(ring[k & 1]).any():
                raise RuntimeError("ELBO is NaN")

        t0 = 0.0
        for i in range(warmup + steps):
            if i == warmup:
                barrier()
                t0 = time.perf_counter()
            cost = gs.step_from_host(pinned, u_host[i % n_pool])
            ring[i & 1].copy_(cost, non_blocking=True)
            rev[i & 1].record()
            if i > 0 and i != warmup:
                look(i - 1)
        look(warmup + steps - 1)
        torch.cuda.synchronize()
        e2e_step = (time.perf_counter() - t0) / steps
        lat_step = float(np.mean(lat))
        if world > 1:
            t = torch.tensor([e2e_step, lat_step], dtype=torch.float64, device=device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            e2e_step, lat_step = float(t[0].item()), float(t[1].item())
        res["e2e"] = {"value": N * world / e2e_step, "unit": "traj/s", "ms_per_step": e2e_step * 1e3,
                      "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(cost_host.numel() * cost_host.element_size()),
                      "mode": "stream of %d public steps from pinned host buffers; step i+1's copies overlap step i, every cost "
                              "is read back and checked one step late; no L2 flush inside the stream (inputs arrive from host "
                              "memory every step)" % steps,
                      "latency_ms": lat_step * 1e3,
                      "latency_mode": "one public step alone, L2 flushed and device idle before it: call -> cost checked on the host"}

    # per-launch device times of the two hot launches (eager, L2 flushed) and their rooflines
    lib = gs.prob.lib

    def t_launch(fn, reps=10):
        ts = []
        for _ in range(reps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        return float(np.median(ts)) * 1e3

    fwd_us = t_launch(lambda: L.check(lib.vh_elbo_terms_fwd(C.byref(gs._p), C.byref(gs._fio), _stream())))
    bwd_us = t_launch(lambda: L.check(lib.vh_elbo_terms_bwd(C.byref(gs._p), C.byref(gs._bio), _stream())))
    bwd_in_step_us = float(bwd_ms.mean()) * 1e3
    S, E = gs.prob.S, len(gs.extras)
    bytes_fwd, bytes_bwd = algorithmic_bytes(B, IW, T, S, P, E, 8 if settings.dtype == torch.float64 else 4)
    peak, peak_src = measured_peak_gbs()
    bb = name == "dr_blackbox_icml"
    traffic, traffic_src = None, None
    prof = os.path.join(ROOT, "profiles", "traffic_%s.json" % name)
    if os.path.exists(prof):
        with open(prof) as f:
            traffic = json.load(f).get("elbo_bwd_dram_bytes")
        traffic_src = "profiles/traffic_%s.json: dram__bytes_read+write of this kernel from a committed `ncu --set full` capture, not measured in this run" % name
    note = ("latency-bound at this size: %d trajectories = %d warps on 148 SMs x 4 schedulers" % (N, (N + 31) // 32)
            if N < 148 * 4 * 32 * 4 else "FP32-issue-bound: ~440 instructions per 32 B of trace") + " (DESIGN.md section 5)"
    res["kernels"] = {"elbo_fwd_us": fwd_us, "elbo_bwd_us": bwd_us, "elbo_bwd_in_step_us": bwd_in_step_us}
    # which form of the white-box kernels this batch takes (vh_launch.cuh: pick_block / use_ws / MxOk)
    latency = N <= 148 * 4 * 32
    hidden = bool(getattr(gs.prob, "dynamic_precisions", False) and gs.prob.net.get("n_hidden", 0) > 0)
    mx = (latency and name.startswith("dr_constant") and "precisions" not in name and settings.params.solver == "midpoint"
          and settings.dtype == torch.float32 and os.environ.get("VIHDS_BWD_MX", "1") != "0")
    bwd_name = "bbm_bwd_kernel" if bb else ("elbo_bwd_mx_kernel" if mx else ("elbo_bwd_ws_kernel" if latency and not hidden else "elbo_bwd_kernel"))
    fwd_name = "bbm_fwd_kernel" if bb else ("elbo_fwd_team_kernel" if latency else "elbo_fwd_kernel")
    res["roofline"] = {
        "kernel": "%s (discrete-adjoint reverse sweep, the dominant launch)" % bwd_name,
        "bound": "hbm", "achieved": bytes_bwd / (bwd_in_step_us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": bytes_bwd / (bwd_in_step_us * 1e-6) / 1e9 / peak, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": peak_src, "algorithmic_bytes": bytes_bwd, "launch_us": bwd_in_step_us, "note": note}
    res["roofline_fwd"] = {
        "kernel": fwd_name, "bound": "hbm", "achieved": bytes_fwd / (fwd_us * 1e-6) / 1e9,
        "peak": peak, "unit": "GB/s", "frac": bytes_fwd / (fwd_us * 1e-6) / 1e9 / peak, "algorithmic_bytes": bytes_fwd,
        "launch_us": fwd_us}
    res["config"] = workload_config(name, B, IW, T, S, P, settings.params.solver, world)
    res["dtype"] = "f32" if settings.dtype == torch.float32 else "f64"
    # launches of libvihds_b200.so per step: enc_fwd, [conditioner,] elbo_fwd, [iwae_fwd_bwd unless fused into] elbo_bwd,
    # enc_bwd, [enc_lin_wgrad unless fused into] adam / exchange+adam (also clears the gradient vector, bumps the step counter)
    fused_adam = (gs.fused_encoder and (world == 1 or gs.exchange is not None) and B <= 128 and
                  os.environ.get("VIHDS_FUSE_ADAM", "1") != "0")
    res["launches_per_step"] = 7 + (1 if gs.rel else 0) - (1 if gs.fuse_iwae else 0) - (1 if fused_adam else 0)
    res["fusions"] = {"iwae_in_reverse_launch": bool(gs.fuse_iwae), "lin_wgrad_in_adam_launch": bool(fused_adam)}
    res["_objects"] = (settings, parameters, model, host, B, IW, T)
    if gs.exchange is not None:
        barrier()
        gs.exchange.close()
    return res


# ---------------------------------------------------------------------------------------------------------------
def cpu_arm(name, a, steps, warmup, device="cpu"):
    """The reference on the host cores for workload `name`: the reference's own Training._run_batch from oracle/_ref
    (kind "reference") when that copy is present and the workload is a shipped spec on its own data grid; otherwise the
    oracle port (kind "port": sample/clip/solve/log-lik/log p,q/IWAE + backward to q, no encoder / Adam)."""
    spec_name, _, B0, IW0, T0 = WORKLOADS[name]
    B, IW = a.batch or B0, a.iw or IW0
    out = {}
    if T0 is None and B == B0:
        # bounded sample: fewer importance samples if (K + W) full steps would not fit in ~4 minutes (0.6 s per full step)
        frac = min(1.0, 240.0 / max(1e-9, (steps + warmup) * 0.6 * (IW / 200.0)))
        iw_c = IW if frac >= 1.0 else max(10, int(IW * frac))
        r = time_reference_itself(spec_name, iw_c, steps, warmup)
        if r is not None:
            out.update(kind="reference", sec=r["sec_per_step"], cores=r["cores"], B=r["B"], IW=iw_c, T=r["T"], loss=r["loss"],
                       sample="%d steps of the reference's Training._run_batch (vihds/training.py:324-340: encoder, sample, clip, "
                              "solve, cost, backward, Adam) on one B=%d x IW=%d mini-batch, T=%d, torch CPU, %d threads (oracle/_ref)" % (
                                  steps, r["B"], iw_c, r["T"], r["cores"]))
    if not out:
        settings, parameters, model, training, host, B, IW, T, rng = build_workload(name, 0, 1, device, a.batch, a.iw, with_training=False)
        Bc = min(B, 64)  # bounded sample: the autograd graph of the full synthetic slab does not fit in host memory
        host_c = {k: (v[:Bc] if k != "times" else v) for k, v in host.items()}
        case = cpu_case(parameters, model, host_c, Bc, IW, rng, settings.params.solver)
        sec, cores, loss = time_oracle(case, steps, warmup)
        out.update(kind="port", sec=sec, cores=cores, B=Bc, IW=IW, T=T, loss=loss,
                   sample="%d steps of B=%d x IW=%d, T=%d: sample/clip/solve/log-lik/log p,q/IWAE + backward to q "
                          "(oracle/vihds_oracle.py, torch CPU, %d threads; no encoder, no Adam)" % (steps, Bc, IW, T, cores))
    out["value"] = out["B"] * out["IW"] / out["sec"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dr_constant_icml", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="individuals per GPU (default: the workload's)")
    ap.add_argument("--iw", type=int, default=None)
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-workloads", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--spin", type=float, default=1.5, help="seconds of untimed steps before the warm-up (clock ramp)")
    a = ap.parse_args()

    import torch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    spec_name, _, B0, IW0, T0 = WORKLOADS[a.workload]

    # ---------------- reference arm: the reference's CPU implementation on the host cores (rank 0 only) ------------
    if a.impl == "reference":
        if rank != 0:
            return
        settings, parameters, model, training, host, B, IW, T, rng = build_workload(a.workload, 0, 1, "cpu", a.batch, a.iw, with_training=False)
        c = cpu_arm(a.workload, a, a.steps, a.warmup)
        tps = c["value"]
        line = {
            "impl": "reference", "metric": "elbo_grad_trajectories_per_sec", "value": tps, "unit": "traj/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": c["sec"] * 1e3, "steps_per_sec": 1.0 / c["sec"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "real plate data, random-init weights (seed 0)",
            "config": workload_config(a.workload, B, IW, T, state_width(model), parameters.n_theta, settings.params.solver, a.gpus),
            "cpu_baseline": {"value": tps, "unit": "traj/s", "cores": c["cores"], "kind": c["kind"], "sample": c["sample"]},
            "e2e": {"value": tps, "unit": "traj/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "loss": c["loss"],
        }
        print(json.dumps(line))
        return

    # ---------------- this repo's arm ---------------------------------------------------------------------------
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the engine has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    from vihds_b200.distributed import init_from_env

    _, _, pg = init_from_env("nccl", device)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    main_res = run_workload(a.workload, a, rank, local_rank, world, device, pg, True)
    clocks = sampler.stop() if sampler is not None else None
    extra = {}
    if not a.no_extra_workloads:
        for nm in ("synthetic_dr_constant", "dr_blackbox_icml", "relay_constant_precisions", "dr_constant_icml"):
            if nm == a.workload:
                continue
            r = run_workload(nm, a, rank, local_rank, world, device, pg, False)
            r.pop("_objects")
            r["unit"] = "traj/s"
            extra[nm] = r
    if rank != 0:
        finish(world)
        return
    settings, parameters, model, host, B, IW, T = main_res.pop("_objects")
    steps = main_res.pop("steps")
    line = {
        "metric": "elbo_grad_trajectories_per_sec", "value": main_res.pop("value"), "unit": "traj/s", "n_gpus": world, "steps": steps,
        "warmup": main_res.pop("warmup"), "ms_per_step": main_res["ms_per_step"], "steps_per_sec": 1e3 / main_res.pop("ms_per_step"),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": main_res.pop("dtype"),
        "data": "real plate data (pre-processed fixture tests/golden/%s.npz)%s, random-init weights (seed 0), u ~ N(0,1)" % (
            WORKLOADS[a.workload][1], "" if T0 is None else " resampled to a synthetic T=%d grid" % T),
        "config": main_res.pop("config"), "cuda_graphs": not a.no_graphs,
        "roofline": main_res.pop("roofline"), "roofline_fwd": main_res.pop("roofline_fwd"), "kernels": main_res.pop("kernels"),
        "e2e": main_res.pop("e2e"), "gpu_launches": int(main_res.pop("launches_per_step") * steps), "clocks": clocks,
    }
    line.update(main_res)
    line["timed_region_ms"] = round(line["ms_per_step"] * steps, 3)
    if extra:
        line["workloads"] = extra
    if not a.no_cpu_baseline and world == 1:
        c = cpu_arm(a.workload, a, a.cpu_steps, 2)
        line["cpu_baseline"] = {"value": c["value"], "unit": "traj/s", "cores": c["cores"], "kind": c["kind"],
                                "ms_per_step": c["sec"] * 1e3, "sample": c["sample"]}
        if c["kind"] == "reference":  # the port as a second figure (no encoder / Adam: an upper bound of the reference's speed)
            Bc = min(B, 64)
            host_c = {k: (v[:Bc] if k != "times" else v) for k, v in host.items()}
            case = cpu_case(parameters, model, host_c, Bc, IW, np.random.RandomState(7), settings.params.solver)
            sec, cores, _ = time_oracle(case, min(a.cpu_steps, 4), 1)
            line["cpu_baseline_port"] = {"value": Bc * IW / sec, "unit": "traj/s", "cores": cores, "kind": "port", "ms_per_step": sec * 1e3}
    print(json.dumps(line), flush=True)
    finish(world)


def finish(world):
    """Leave without tearing the NCCL communicator down: destroy_process_group() blocks while captured CUDA graphs that
    contain NCCL kernels are still alive (observed on torch 2.11 / NCCL 2.28), and there is nothing left to clean up."""
    if world > 1:
        import torch

        torch.distributed.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
