"""Where one ELBO-gradient step spends its time: CUDA events between the segments of GraphedStep.step (device time,
L2 flushed before each step as in bench.py) and the host's enqueue time for the same calls (is the loop GPU- or
CPU-bound?).  usage: python tools/step_timeline.py [workload]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from vihds_b200 import _lib as L
from vihds_b200.training import GraphedStep

wl = sys.argv[1] if len(sys.argv) > 1 else "dr_constant_icml"
torch.cuda.set_device(0)
settings, parameters, model, training, host, B, IW, T, rng = bench.build_workload(wl, 0, 1, torch.device("cuda", 0))
model.want_predict = False
gs = GraphedStep(training, B, IW, T)
pinned = {k: v.pin_memory() for k, v in host.items()}
u = torch.randn(B, IW, parameters.n_theta).cuda()
gs.load_batch(pinned); gs.load_u(u); gs.draw_conditioner(); gs.prepare()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(50):
    gs.step()
torch.cuda.synchronize()
lib = gs.prob.lib
s = torch.cuda.current_stream().cuda_stream
names = ["load_u+cond", "g_pre", "elbo_fwd", "iwae", "elbo_bwd", "post (eager)"]
n = 200
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)] for _ in range(n)]
host_t = np.zeros(len(names))
t_loop0 = time.perf_counter()
for i in range(n):
    flush.zero_()
    e = evs[i]
    e[0].record(); t = time.perf_counter()
    gs.load_u(u); gs.draw_conditioner()
    e[1].record(); t2 = time.perf_counter(); host_t[0] += t2 - t; t = t2
    gs.g_pre.replay()
    e[2].record(); t2 = time.perf_counter(); host_t[1] += t2 - t; t = t2
    L.check(lib.vh_elbo_terms_fwd(gs._p_ref, gs._fio_ref, s))
    e[3].record(); t2 = time.perf_counter(); host_t[2] += t2 - t; t = t2
    L.check(lib.vh_iwae_fwd_bwd(*gs._iwae_args, s))
    e[4].record(); t2 = time.perf_counter(); host_t[3] += t2 - t; t = t2
    L.check(lib.vh_elbo_terms_bwd(gs._p_ref, gs._bio_ref, s))
    e[5].record(); t2 = time.perf_counter(); host_t[4] += t2 - t; t = t2
    gs._post()
    e[6].record(); t2 = time.perf_counter(); host_t[5] += t2 - t; t = t2
t_enq = time.perf_counter() - t_loop0
torch.cuda.synchronize()
t_all = time.perf_counter() - t_loop0
dev = np.array([[e[k].elapsed_time(e[k + 1]) for k in range(len(names))] for e in evs]) * 1e3
print("%-14s %10s %10s" % ("segment", "device us", "host us"))
for k, nm in enumerate(names):
    print("%-14s %10.1f %10.1f" % (nm, np.median(dev[:, k]), host_t[k] / n * 1e6))
print("%-14s %10.1f %10.1f" % ("step", np.median(dev.sum(1)), host_t.sum() / n * 1e6))
print("host enqueue per iteration %.1f us, wall per iteration %.1f us (flush memset ~70 us of it)" % (t_enq / n * 1e6, t_all / n * 1e6))

# the production form: two graph replays per step
torch.cuda.synchronize()
ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
t0 = time.perf_counter()
for i in range(n):
    flush.zero_()
    ev2[i][0].record()
    gs.load_u(u); gs.draw_conditioner()
    gs.step()
    ev2[i][1].record()
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("graphed step: device %.1f us, host enqueue %.1f us / iteration, wall %.1f us / iteration" % (
    np.median([a_.elapsed_time(b_) for a_, b_ in ev2]) * 1e3, t_enq / n * 1e6, t_all / n * 1e6))
