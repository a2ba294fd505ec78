"""Kernel-only timing of the fused forward / reverse kernels through the C ABI on synthetic dr_constant batches."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_case  # noqa: E402
import helpers as H  # noqa: E402
from vihds_b200 import _lib as L  # noqa: E402


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=36)
    ap.add_argument("--IW", type=int, default=200)
    ap.add_argument("--T", type=int, default=86)
    ap.add_argument("--solver", default="midpoint")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--case", default="dr_constant_icml_midpoint_f32_iw8")
    a = ap.parse_args()
    lib = L.load()
    case = load_case(a.case)
    case["solver"] = a.solver
    model = H.MODEL_IDS[str(case["model"])]
    B0, _, P = case["u"].shape
    B, IW, T = a.B, a.IW, a.T
    N = B * IW
    rng = np.random.RandomState(0)
    idx = np.arange(B) % B0
    times = np.linspace(0, 16.5, T).astype(np.float32)
    obs0 = case["observations"][idx]
    obs = np.stack([np.stack([np.interp(times, case["times"], obs0[b, o]) for o in range(4)]) for b in range(B)]).astype(np.float32)
    big = dict(case)
    big["u"] = rng.randn(B, IW, P).astype(np.float32)
    big["times"], big["observations"] = times, obs
    for k in ("q_mu", "q_prec", "inputs", "dev_1hot"):
        big[k] = case[k][idx]
    src, _ = H.slot_map(case, L.slot_names(model))
    extra = None
    if "cond_aR" in case:
        extra = (1.0 + np.abs(rng.randn(2, N))).astype(np.float32)
    p = H.make_problem(big, src, 0 if extra is None else 2)
    S = lib.vh_state_width(C.byref(p))
    lo, hi = H.clip_bounds(case)
    cu = lambda x: None if x is None else torch.as_tensor(np.ascontiguousarray(x)).cuda()  # noqa: E731
    dev = dict(times=cu(times), u=cu(big["u"].reshape(N, P)), q_mu=cu(big["q_mu"]), q_prec=cu(big["q_prec"]),
               p_mu=cu(case["p_mu"]), p_prec=cu(case["p_prec"]), clip_lo=cu(lo), clip_hi=cu(hi),
               kind=cu(case["kinds"].astype(np.int32)), extra=cu(extra), treatments=cu(big["inputs"]),
               dev_1hot=cu(big["dev_1hot"]), observations=cu(obs),
               theta=torch.empty(P, N, device="cuda"), x_states=torch.empty(T, S, N, device="cuda"),
               x_predict=torch.empty(T, 4, N, device="cuda"), logp_by_species=torch.empty(N, 4, device="cuda"),
               logp_theta=torch.empty(N, device="cuda"), logq_theta=torch.empty(N, device="cuda"))
    io = L.vh_fwd_io(**{k: _p(v) for k, v in dev.items()})
    g = dict(g_logp_by_species=torch.full((N, 4), -1.0 / N, device="cuda"), g_logp_theta=torch.full((N,), -1.0 / N, device="cuda"),
             g_logq_theta=torch.full((N,), 1.0 / N, device="cuda"))
    out = dict(d_q_mu=torch.empty(B, P, device="cuda"), d_q_prec=torch.empty(B, P, device="cuda"))
    bio = L.vh_bwd_io(fwd=io, **{k: _p(v) for k, v in {**g, **out}.items()})
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timeit(fn):
        ts = []
        for i in range(a.iters + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)), float(np.min(ts))

    f_med, f_min = timeit(lambda: L.check(lib.vh_elbo_terms_fwd(C.byref(p), C.byref(io), None)))
    b_med, b_min = timeit(lambda: L.check(lib.vh_elbo_terms_bwd(C.byref(p), C.byref(bio), None)))
    bytes_fwd = 4 * (N * P + B * 4 * T + N * T * (S + 4) + N * 6 + N * P)
    bytes_bwd = 4 * (N * T * S + B * 4 * T + N * P + N * 6)
    print("B=%d IW=%d N=%d T=%d solver=%s" % (B, IW, N, T, a.solver))
    print("fwd: median %.3f ms (min %.3f)  %.1f GB/s algorithmic  %.2f Mtraj/s" % (f_med, f_min, bytes_fwd / f_med / 1e6, N / f_med / 1e3))
    print("bwd: median %.3f ms (min %.3f)  %.1f GB/s algorithmic  %.2f Mtraj/s" % (b_med, b_min, bytes_bwd / b_med / 1e6, N / b_med / 1e3))
    print("finite:", bool(torch.isfinite(dev["logp_by_species"]).all()), bool(torch.isfinite(out["d_q_mu"]).all()))


if __name__ == "__main__":
    main()
