"""Host-side cost of each piece of the end-to-end step (perf_counter around the calls, GPU idle at the start)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
torch.cuda.set_device(0)
settings, parameters, model, training, host, B, IW, T, rng = bench.build_workload("dr_constant_icml", 0, 1, torch.device("cuda", 0))
from vihds_b200.training import GraphedStep
model.want_predict = False
gs = GraphedStep(training, B, IW, T)
pinned = {k: v.pin_memory() for k, v in host.items()}
u = torch.randn(B, IW, parameters.n_theta).pin_memory()
gs.load_batch(pinned); gs.load_u(u); gs.draw_conditioner(); gs.prepare()
cost_host = torch.zeros(1).pin_memory()
for _ in range(20):
    gs.step_from_host(pinned, u)
torch.cuda.synchronize()
acc = {}
def tick(name, t0):
    t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0); return t1
n = 200
cur = torch.cuda.current_stream()
for _ in range(n):
    torch.cuda.synchronize()
    t = time.perf_counter(); t_start = t
    gs.load_batch(pinned); t = tick("load_batch", t)
    gs.draw_conditioner(); t = tick("draw_conditioner", t)
    gs.load_u(u); t = tick("load_u", t)
    gs.g_pre.replay(); t = tick("g_pre.replay", t)
    gs.g_rest.replay(); t = tick("g_rest.replay", t)
    cost_host.copy_(gs.buf.cost, non_blocking=True); t = tick("cost d2h enqueue", t)
    torch.cuda.synchronize(); t = tick("final sync (GPU tail)", t)
    acc["total"] = acc.get("total", 0.0) + (t - t_start)
for k, v in acc.items():
    print("%-28s %7.1f us" % (k, v / n * 1e6))
