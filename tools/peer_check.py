"""N-rank check of the fused gradient exchange + Adam kernel against ncclAllReduce + vh_adam_step (torchrun).
Every rank feeds rank-dependent random gradients for a few steps; parameters must agree with the NCCL path and be
bit-identical across ranks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from vihds_b200 import _lib as L
from vihds_b200.distributed import PeerGradientExchange, init_from_env

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
rank, world, pg = init_from_env("nccl", dev)
lib = L.load()
_p = lambda t: t.data_ptr()
ok = True
for n in (1000, 44411, 1 << 20):
    g0 = torch.Generator().manual_seed(1234)
    p0 = torch.randn(n, generator=g0).to(dev)
    pa, pb = p0.clone(), p0.clone()
    ma, va, mb, vb = (torch.zeros(n, device=dev) for _ in range(4))
    hyper = torch.tensor([0.01, 0.9, 0.999, 1e-8], dtype=torch.float64, device=dev)
    step = torch.zeros(4, dtype=torch.int64, device=dev)
    ex = PeerGradientExchange(n, torch.float32, dev, pg)
    gr_gen = torch.Generator().manual_seed(100 + rank)
    for it in range(1, 8):
        gr = torch.randn(n, generator=gr_gen).to(dev)
        ga = gr.clone()
        L.check(lib.vh_adam_allreduce_step(0, n, _p(pa), _p(ga), _p(ma), _p(va), _p(hyper), _p(step), _p(ex.state),
                                           rank, world, _p(ex.peers), None, 0.0, None))
        dist.all_reduce(gr, group=pg)
        L.check(lib.vh_adam_step(0, n, _p(pb), _p(gr), _p(mb), _p(vb), 0.01, 0.9, 0.999, 1e-8, it, None))
    torch.cuda.synchronize()
    err = float((pa - pb).abs().max())
    gathered = [torch.empty_like(pa) for _ in range(world)]
    dist.all_gather(gathered, pa, group=pg)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    to = ex.timed_out()
    print("rank %d n=%d: max |peer - nccl| = %.3e, bit-identical across ranks: %s, timed_out: %s, epoch %d" % (
        rank, n, err, same, to, int(ex.state[0])), flush=True)
    ok = ok and err < 1e-5 and same and not to
# the whole data-parallel training step, three Adam steps: peer exchange vs NCCL, same seeds, same u
import bench
from vihds_b200.training import GraphedStep
flats = {}
for mode in ("peer", "nccl"):
    torch.manual_seed(0)
    settings, parameters, model, training, host, B, IW, T, rng = bench.build_workload("dr_constant_icml", rank, world, dev)
    model.want_predict = False
    gs = GraphedStep(training, B, IW, T, b_total=B * world, process_group=pg, exchange=mode)
    gs.load_batch({k: v.to(dev) for k, v in host.items()})
    gs.extras_override = torch.full((len(gs.extras), B * IW), 1.0, device=dev)  # pin the (random) conditioner
    ug = torch.Generator().manual_seed(7 + rank)
    costs = []
    for i in range(3):
        gs.load_u(torch.randn(B, IW, parameters.n_theta, generator=ug).to(dev))
        costs.append(float(gs.step().item()))
    torch.cuda.synchronize()
    flats[mode] = training.optimizer.flat.clone()
    print("rank %d %s costs %s steps %s" % (rank, mode, costs, training.optimizer.step_dev.tolist()), flush=True)
d = float((flats["peer"] - flats["nccl"]).abs().max())
print("rank %d: max |flat(peer) - flat(nccl)| after 3 steps = %.3e" % (rank, d), flush=True)
ok = ok and d < 2e-3  # Adam moves every parameter by ~lr = 0.01 per step; fp32 atomics order differs between runs
# collective NaN verdict: one rank's cost is NaN -> EVERY rank skips the update (and keeps skipping: the guard is sticky)
n = 5000
pa = torch.randn(n, generator=torch.Generator().manual_seed(1)).to(dev)
p0 = pa.clone()
ma, va = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
hyper = torch.tensor([0.01, 0.9, 0.999, 1e-8], dtype=torch.float64, device=dev)
step = torch.zeros(4, dtype=torch.int64, device=dev)
ex = PeerGradientExchange(n, torch.float32, dev, pg)
for it, bad_rank in enumerate((None, 1 % world, None)):
    ga = torch.randn(n, generator=torch.Generator().manual_seed(50 + it)).to(dev)
    cost = torch.tensor([float("nan") if rank == bad_rank else 1.0], device=dev)
    L.check(lib.vh_adam_allreduce_step(0, n, _p(pa), _p(ga), _p(ma), _p(va), _p(hyper), _p(step), _p(ex.state),
                                       rank, world, _p(ex.peers), _p(cost), 0.0, None))
    torch.cuda.synchronize()
    if it == 0:
        p1 = pa.clone()
        ok = ok and not torch.equal(p1, p0) and step.tolist() == [1, 0, 0, 0]
    else:
        ok = ok and torch.equal(pa, p1) and step.tolist() == [1, 0, it, 0] and ex.state.tolist()[2:] == [0, it]
print("rank %d: NaN on one rank -> all ranks skipped: steps %s state %s" % (rank, step.tolist(), ex.state.tolist()), flush=True)
dist.barrier()
ex.close()

# rank-count invariance WITHOUT pinning the conditioner: the individuals of a global batch sharded over the ranks must
# give what one process computes for the whole batch (device conditioner on global sample indices, u sliced, cost / B_global)
from vihds_b200.distributed import shard_bounds
Bg, IWs = 6 * world, 16
res = {}
for mode in ("global", "sharded"):
    torch.manual_seed(0)
    settings, parameters, model, training, host, B, IW, T, rng = bench.build_workload("dr_constant_icml", 0, 1, dev, Bg, IWs)
    model.want_predict = False
    lo, hi = shard_bounds(Bg, world, rank) if mode == "sharded" else (0, Bg)
    gs = GraphedStep(training, hi - lo, IWs, T, b_total=Bg, process_group=pg if mode == "sharded" else None, b_offset=lo)
    hd = {k: (v if k == "times" else v[lo:hi]).to(dev) for k, v in host.items()}
    gs.load_batch(hd)
    if mode == "sharded":
        gs.load_global_devices(host["dev_1hot"].to(dev))
    ug = torch.Generator().manual_seed(99)
    torch.manual_seed(123)  # the conditioner weights come from the torch CPU RNG: same stream on every rank and in both modes
    costs = []
    for i in range(3):
        u = torch.randn(Bg, IWs, parameters.n_theta, generator=ug)
        gs.load_u(u[lo:hi].contiguous().to(dev))
        gs.draw_conditioner()
        c = gs.step().clone()
        if mode == "sharded":
            dist.all_reduce(c, group=pg)
        costs.append(float(c.item()))
    torch.cuda.synchronize()
    res[mode] = (costs, training.optimizer.flat.clone())
    if gs.exchange is not None:
        dist.barrier()
        gs.exchange.close()
dc = max(abs(a - b) / abs(b) for a, b in zip(res["sharded"][0], res["global"][0]))
dp = float((res["sharded"][1] - res["global"][1]).abs().max())
print("rank %d: sharded vs one-process global batch: costs rel %.2e, params max abs %.2e (%s | %s)" % (
    rank, dc, dp, res["sharded"][0], res["global"][0]), flush=True)
ok = ok and dc < 1e-5 and dp < 2e-3
dist.barrier()
print("rank %d %s" % (rank, "PEER_CHECK_OK" if ok else "PEER_CHECK_FAILED"), flush=True)
os._exit(0 if ok else 1)
