"""Stage-by-stage check of the multi-GPU step (NCCL all-reduce inside the captured post-step graph)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench

def log(*a):
    print("[rank %s %.1fs]" % (os.environ.get("RANK"), time.time() - T0), *a, file=sys.stderr, flush=True)

T0 = time.time()
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
from vihds_b200.distributed import init_from_env
from vihds_b200.training import GraphedStep
_, _, pg = init_from_env("nccl", device)
log("pg up")
t = torch.ones(4, device=device)
dist.all_reduce(t)
torch.cuda.synchronize()
log("eager allreduce ok", t[0].item())
settings, parameters, model, training, host, B, IW, T, rng = bench.build_workload("dr_constant_icml", rank, world, device)
model.want_predict = False
use_graphs = os.environ.get("NOGRAPH") is None
gs = GraphedStep(training, B, IW, T, b_total=B * world, process_group=pg, use_graphs=use_graphs)
gs.load_batch({k: v.to(device) for k, v in host.items()})
gs.load_u(torch.randn(B, IW, parameters.n_theta, device=device))
gs.draw_conditioner()
log("built; preparing (graphs=%s)" % use_graphs)
gs.prepare()
torch.cuda.synchronize()
log("prepared")
for i in range(5):
    c = gs.step()
torch.cuda.synchronize()
log("5 steps ok cost", float(c.item()))
dist.barrier()
log("barrier ok")
dist.destroy_process_group()
log("done")
