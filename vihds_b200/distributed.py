"""Data parallelism over the individual (B) axis: one process per GPU, ONE gradient all-reduce per step.

The reference is single-device (no NCCL / DDP anywhere).  The trajectory axis is embarrassingly
parallel, but the IWAE bound reduces over the IW samples of ONE individual (vihds/training.py:144), so individuals
-- never samples -- are sharded: every rank owns all IW samples of its individuals, the per-individual logsumexp stays
local, ``cost = -mean_b(...)`` becomes a sum over ranks of ``-sum_{b local}(...) / B_global`` (``b_total`` of
vh_iwae_fwd/bwd), and the only exchange is a SUM all-reduce of the flat gradient (~44 k floats for
dr_constant_icml) followed by an identical Adam step on every rank.  ``u`` is drawn from numpy's global RNG for
the GLOBAL batch and sliced (vihds/vae.py:22-24 is the RNG contract), so results do not depend on the rank count.

Backend: "nccl" over NVLink on the GPUs (captured inside the post-step CUDA graph, training.GraphedStep._post);
"gloo" in the CPU tests of this module's logic (tests/test_distributed_cpu.py).
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend="nccl", device=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).  Returns (rank, world, group)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world == 1:
        return 0, 1, None
    if not dist.is_initialized():
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, dist.group.WORLD


def shard_bounds(n, world, rank):
    """Contiguous, as-even-as-possible slab of ``n`` individuals for ``rank`` (B = 36 over 8 ranks: 5,5,5,5,4,4,4,4)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch, world, rank):
    """Slice every per-individual tensor of a batch container; ``times`` is shared."""
    n = batch["inputs"].shape[0]
    lo, hi = shard_bounds(n, world, rank)
    out = type(batch)()
    for k, v in batch.items():
        out[k] = v if k == "times" else v[lo:hi]
    return out, (lo, hi)


def sample_u_global(n_batch_global, n_samples, n_theta, world, rank, dtype=np.float32):
    """Draw u for the GLOBAL batch from numpy's global RNG on every rank (same seed => same stream), keep the slab."""
    u = np.random.randn(n_batch_global, n_samples, n_theta).astype(dtype)
    lo, hi = shard_bounds(n_batch_global, world, rank)
    return torch.from_numpy(u[lo:hi])


def allreduce_gradient_(flat_grad, group):
    """SUM over ranks, in place.  The local gradients already carry the 1 / B_global factor."""
    if group is not None:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def allreduce_max(value, group, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if group is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
