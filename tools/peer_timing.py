"""Phase times of the fused gradient-exchange kernel inside real training steps (torchrun, N ranks): runs the bench workload's
GraphedStep for a number of steps and prints, per rank, the median time block 0 spent in each phase (vh_peer_debug_times)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from vihds_b200 import _lib as L
from vihds_b200.distributed import init_from_env
from vihds_b200.training import GraphedStep

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
rank, world, pg = init_from_env("nccl", dev)
settings, parameters, model, training, host, B, IW, T, rng = bench.build_workload("dr_constant_icml", rank, world, dev, None, None)
model.want_predict = False
gs = GraphedStep(training, B, IW, T, b_total=B * world, process_group=pg, b_offset=rank * B)
if world > 1 and gs.rel:
    gs.load_global_devices(bench.global_dev_1hot("dr_constant_icml", training.dataset_pair.train.dataset, training.dataset_pair, B, world).to(dev))
gs.load_batch({k: v.pin_memory() for k, v in host.items()})
gs.load_u(torch.randn(B, IW, parameters.n_theta, device=dev))
gs.draw_conditioner()
gs.prepare()
lib = L.load()
lib.vh_peer_debug_times.argtypes = [C.POINTER(C.c_ulonglong)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []
for i in range(60):
    flush.zero_()
    gs.step()
    torch.cuda.synchronize()
    out = (C.c_ulonglong * 8)()
    lib.vh_peer_debug_times(out)
    if i >= 10:
        rows.append([int(x) for x in out])
a = np.array(rows, dtype=np.float64)
d = np.diff(a[:, :7], axis=1) / 1e3
names = ["pdl_wait", "push (+weight gradient)", "fence + flags", "wait for peers", "vote", "sum + Adam"]
torch.distributed.barrier()
for r in range(world):
    if r == rank:
        print("rank %d of %d: median us per phase: %s | total %.1f" % (rank, world, ", ".join("%s %.1f" % (n, v) for n, v in zip(names, np.median(d, 0))),
                                                                        np.median(a[:, 6] - a[:, 0]) / 1e3), flush=True)
    torch.distributed.barrier()
os._exit(0)
