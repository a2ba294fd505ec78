"""CPU, gloo, world_size 2: the host-side logic of the multi-GPU path (vihds_b200/distributed.py) -- sharding of
individuals, rank-count-invariant u, and that local gradients of the sharded IWAE cost (with the global-batch
denominator) all-reduce to the single-process gradient.  The per-sample terms come from the CPU oracle here (test
infrastructure); on the GPUs they come from the CUDA kernels with exactly the same b_total convention."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_case
import vihds_oracle as O
from vihds_b200 import distributed as D
from vihds_b200.config import Settings


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_cover_and_balance():
    for n, w in ((36, 8), (36, 2), (8192, 8), (5, 8)):
        parts = [D.shard_bounds(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    assert [b - a for a, b in (D.shard_bounds(36, 8, r) for r in range(8))] == [5, 5, 5, 5, 4, 4, 4, 4]


def _local_cost(case, lo, hi, b_total):
    """IWAE cost contribution of individuals [lo, hi) with the GLOBAL denominator, and its gradient w.r.t. q."""
    sub = dict(case)
    for k in ("u", "inputs", "dev_1hot", "observations", "q_mu", "q_prec", "cond_aR", "cond_aS"):
        if k in sub:
            sub[k] = case[k][lo:hi]
    dt = torch.float64
    T_ = lambda a: torch.as_tensor(a, dtype=dt)  # noqa: E731
    kinds = [int(k) for k in case["kinds"]]
    names = [str(n) for n in case["names"]]
    q_mu, q_prec = T_(sub["q_mu"]).requires_grad_(True), T_(sub["q_prec"]).requires_grad_(True)
    theta = O.clip_theta(O.sample_theta(T_(sub["u"]), q_mu, q_prec, kinds), T_(case["p_mu"]), T_(case["p_sigma"]), kinds, 4.0)
    th = dict(zip(names, theta))
    for e in ("aR", "aS"):
        if "cond_" + e in sub:
            th[e] = T_(sub["cond_" + e])
    _, xp, prec = O.decode(str(case["model"]), str(case["solver"]), th, T_(case["times"]), T_(sub["inputs"]), T_(sub["dev_1hot"]))
    lpx = O.log_prob_observations(xp, T_(sub["observations"]), prec)
    log_w = lpx.sum(2) + O.log_prob_theta(theta, T_(case["p_mu"]), T_(case["p_prec"]), kinds) - O.log_prob_theta(theta, q_mu, q_prec, kinds)
    cost = -(log_w.logsumexp(1) - np.log(log_w.shape[1])).sum() / b_total
    cost.backward()
    return cost.detach(), q_mu.grad, q_prec.grad


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    r, w, group = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    case = load_case("dr_constant_one_modeuler_f64_iw5")
    B = case["u"].shape[0]
    # rank-count-invariant u: every rank draws the global tensor from the same seeded numpy stream and keeps its slab
    np.random.seed(3)
    u_loc = D.sample_u_global(B, 5, case["u"].shape[2], world, rank)
    batch = Settings(inputs=torch.as_tensor(case["inputs"]), dev_1hot=torch.as_tensor(case["dev_1hot"]),
                     observations=torch.as_tensor(case["observations"]), times=torch.as_tensor(case["times"]))
    shard, (lo, hi) = D.shard_batch(batch, world, rank)
    assert shard.inputs.shape[0] == hi - lo and shard.times.shape == batch.times.shape
    cost, g_mu, g_prec = _local_cost(case, lo, hi, B)
    # global parameters: their gradient is the sum over individuals -> flat vector [cost, sum_b g_mu, sum_b g_prec]
    flat = torch.cat([cost.reshape(1), g_mu.sum(0), g_prec.sum(0)])
    D.allreduce_gradient_(flat, group)
    tmax = D.allreduce_max(float(rank), group, "cpu")
    if rank == 0:
        torch.save({"flat": flat, "u_lo_hi": (lo, hi), "u": u_loc, "tmax": tmax}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_allreduce_matches_single_process(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    case = load_case("dr_constant_one_modeuler_f64_iw5")
    B = case["u"].shape[0]
    cost, g_mu, g_prec = _local_cost(case, 0, B, B)
    ref = torch.cat([cost.reshape(1), g_mu.sum(0), g_prec.sum(0)])
    assert torch.allclose(got["flat"], ref, rtol=1e-9, atol=1e-12)
    assert abs(float(ref[0]) - float(case["loss"])) < 1e-8 * abs(float(case["loss"]))
    np.random.seed(3)
    u = np.random.randn(B, 5, case["u"].shape[2]).astype(np.float32)
    lo, hi = got["u_lo_hi"]
    assert np.array_equal(got["u"].numpy(), u[lo:hi]) and got["tmax"] == 1.0
