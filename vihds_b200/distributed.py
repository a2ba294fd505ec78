"""Data parallelism over the individual (B) axis: one process per GPU, ONE gradient all-reduce per step.

The reference is single-device (no NCCL / DDP anywhere).  The trajectory axis is embarrassingly
parallel, but the IWAE bound reduces over the IW samples of ONE individual (vihds/training.py:144), so individuals
-- never samples -- are sharded: every rank owns all IW samples of its individuals, the per-individual logsumexp stays
local, ``cost = -mean_b(...)`` becomes a sum over ranks of ``-sum_{b local}(...) / B_global`` (``b_total`` of
vh_iwae_fwd/bwd), and the only exchange is a SUM all-reduce of the flat gradient (~44 k floats for
dr_constant_icml) followed by an identical Adam step on every rank.  ``u`` is drawn from numpy's global RNG for
the GLOBAL batch and sliced (vihds/vae.py:22-24 is the RNG contract), so results do not depend on the rank count.

Backend: "nccl" for the rendezvous and the control plane.  The gradient exchange itself is ``PeerGradientExchange``:
one kernel of libvihds_b200.so (vh_adam_allreduce_step) that pushes the gradient into every rank's inbox over NVLink
peer memory, sums in rank order and applies Adam -- captured inside the step's CUDA graph (training.GraphedStep._post).
``VIHDS_ALLREDUCE=nccl`` selects ncclAllReduce + the plain Adam kernel instead.  "gloo" in the CPU tests of this module's
host logic (tests/test_distributed_cpu.py).
"""
import ctypes as C
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


class ExchangeUnavailable(RuntimeError):
    """Raised on EVERY rank when any rank could not set the peer exchange up (the caller then takes the NCCL path)."""


class PeerGradientExchange(object):
    """Exchange buffers of the fused all-reduce + Adam kernel (include/vihds_b200.h: vh_peer_*, vh_adam_allreduce_step).

    Every rank allocates one buffer (flags + a double-buffered inbox with one slot per rank), the CUDA-IPC handles go
    round through ``all_gather_object``, and each rank maps its peers' buffers.  ``group=None`` builds a one-rank
    exchange (the kernel then only talks to itself: used by the single-GPU test)."""

    def __init__(self, n, dtype, device, group=None, timeout_s=None):
        from . import _lib as L
        # bound of the in-kernel wait for the peers' flags (seconds of %globaltimer); env VIHDS_EXCHANGE_TIMEOUT_S
        self.timeout_s = float(timeout_s if timeout_s is not None else os.environ.get("VIHDS_EXCHANGE_TIMEOUT_S", "10"))
        self.lib = lib = L.load()
        self.rank = dist.get_rank(group) if group is not None else 0
        self.world = dist.get_world_size(group) if group is not None else 1
        vdt = L.VH_F64 if dtype == torch.float64 else L.VH_F32
        nbytes = lib.vh_peer_buffer_bytes(vdt, n, self.world)
        handle = (C.c_ubyte * 64)()
        own = C.c_void_p()
        self._own, self._opened, err = None, [], None
        # Every rank runs every collective below whatever happens locally, and all ranks reach the same verdict:
        # a rank that cannot create or map a buffer (no peer access, IPC unavailable) must not leave the others waiting.
        with torch.cuda.device(device):
            try:
                L.check(lib.vh_peer_buffer_create(nbytes, C.byref(own), handle))
                self._own = own.value
            except RuntimeError as e:
                err = str(e)
            mine = bytes(handle) if err is None else None
            handles = [mine]
            if group is not None:
                handles = [None] * self.world
                dist.all_gather_object(handles, mine, group=group)
            ptrs = []
            if err is None and all(h is not None for h in handles):
                try:
                    for r in range(self.world):
                        if r == self.rank:
                            ptrs.append(own.value)
                            continue
                        q = C.c_void_p()
                        buf = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                        L.check(lib.vh_peer_buffer_open(buf, C.byref(q)))
                        ptrs.append(q.value)
                        self._opened.append(q.value)
                except RuntimeError as e:
                    err = str(e)
            elif err is None:
                err = "a peer could not create its exchange buffer"
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=device)
            if group is not None:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # also: nobody pushes before everybody has mapped
            if int(ok.item()) == 0:
                self.close()
                raise ExchangeUnavailable(err or "a peer could not map the exchange buffers")
        self.peers = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.state = torch.zeros(4, dtype=torch.int64, device=device)  # epoch, ticket, timed_out (sticky), skipped

    def timed_out(self):
        return bool(self.state[2].item())

    def check(self):
        """Raise if the exchange has ever timed out (synchronises).  A timed-out call applies no complete update and
        every later call is a no-op on this rank, so the replicas must not be used any further."""
        if self.timed_out():
            raise RuntimeError("gradient exchange: a peer's flag did not arrive within %.1f s (vh_adam_allreduce_step "
                               "timed out); the parameter replicas are no longer in step" % self.timeout_s)

    def close(self):
        """Unmap the peers' buffers and free this rank's (call on every rank, after a barrier: peers may still push)."""
        for q in self._opened:
            self.lib.vh_peer_buffer_close(C.c_void_p(q))
        self._opened = []
        if self._own is not None:
            self.lib.vh_peer_buffer_destroy(C.c_void_p(self._own))
            self._own = None
