"""Kernel-only timing of the dr_blackbox forward / reverse launches through the C ABI (golden case tiled to B x IW)."""
import argparse, ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from conftest import load_case  # noqa: E402
import helpers as H  # noqa: E402
from vihds_b200 import _lib as L  # noqa: E402
from test_gpu_parity import _dev, _p  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=36); ap.add_argument("--IW", type=int, default=200); ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
lib = L.load()
case = load_case("dr_blackbox_icml_midpoint_f32_iw8")
B0, _, P = case["u"].shape
B, IW = a.B, a.IW
N = B * IW
idx = np.arange(B) % B0
big = dict(case)
big["u"] = np.random.RandomState(0).randn(B, IW, P).astype(np.float32)
for k in ("q_mu", "q_prec", "inputs", "dev_1hot", "observations"):
    big[k] = np.ascontiguousarray(case[k][idx])
src, extra = H.slot_map(big, L.slot_names(6))
w, _ = H.flat_weights(big)
p = H.make_problem(big, src, extra.shape[0])
T, S = p.T, 10
lo, hi = H.clip_bounds(big)
dev = dict(times=_dev(big["times"]), u=_dev(big["u"].reshape(N, P)), q_mu=_dev(big["q_mu"]), q_prec=_dev(big["q_prec"]),
           p_mu=_dev(big["p_mu"]), p_prec=_dev(big["p_prec"]), clip_lo=_dev(lo), clip_hi=_dev(hi), kind=_dev(big["kinds"].astype(np.int32)),
           extra=_dev(extra), treatments=_dev(big["inputs"]), dev_1hot=_dev(big["dev_1hot"]), observations=_dev(big["observations"]),
           weights=_dev(w), theta=torch.empty(P, N, device="cuda"), x_states=torch.empty(T, S, N, device="cuda"),
           logp_by_species=torch.empty(N, 4, device="cuda"), logp_theta=torch.empty(N, device="cuda"), logq_theta=torch.empty(N, device="cuda"))
io = L.vh_fwd_io(**{k: _p(v) for k, v in dev.items()})
g = dict(g_logp_by_species=torch.full((N, 4), -1.0 / N, device="cuda"), g_logp_theta=torch.full((N,), -1.0 / N, device="cuda"),
         g_logq_theta=torch.full((N,), 1.0 / N, device="cuda"))
out = dict(d_q_mu=torch.empty(B, P, device="cuda"), d_q_prec=torch.empty(B, P, device="cuda"), d_weights=torch.empty(len(w), device="cuda"),
           d_extra=torch.empty(extra.shape, device="cuda"))
bio = L.vh_bwd_io(fwd=io, **{k: _p(v) for k, v in {**g, **out}.items()})
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn):
    ts = []
    for i in range(a.iters + 3):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
f = timeit(lambda: L.check(lib.vh_elbo_terms_fwd(C.byref(p), C.byref(io), None)))
b = timeit(lambda: L.check(lib.vh_elbo_terms_bwd(C.byref(p), C.byref(bio), None)))
print("blackbox B=%d IW=%d N=%d T=%d: fwd %.3f ms (%.2f Mtraj/s)  bwd %.3f ms (%.2f Mtraj/s)" % (B, IW, N, T, f, N / f / 1e3, b, N / b / 1e3))
