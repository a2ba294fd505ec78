"""GPU: dr_blackbox tensor-core kernels (vh_bb_mma.cuh) against the scalar kernels on the same inputs, output by output,
over shapes of growing complexity (T = 2 euler isolates one RHS evaluation / one VJP).  Prints relative errors; exit
code 1 if any exceeds the bar.  `python tools/bb_mma_check.py [--time]`."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

from conftest import load_case  # noqa: E402
from test_gpu_parity import run_case_on_gpu  # noqa: E402
from test_gpu_properties import sub_case  # noqa: E402

KEYS = ["theta", "x_states", "x_predict", "logp_by_species", "logp_theta", "logq_theta", "cost", "d_q_mu", "d_q_prec",
        "d_weights", "d_extra"]


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def compare(case, label, bar=2e-4):
    os.environ["VIHDS_BB_IMPL"] = "scalar"
    ref = run_case_on_gpu(case)
    os.environ["VIHDS_BB_IMPL"] = "mma"
    got = run_case_on_gpu(case)
    errs = {k: rel(got[k], ref[k]) for k in KEYS if k in ref}
    B, IW, P, T, S = ref["dims"]
    xs_g, xs_r = got["x_states"].reshape(T, S, B * IW), ref["x_states"].reshape(T, S, B * IW)
    per_state = [rel(xs_g[:, s], xs_r[:, s]) for s in range(S)]
    worst = max(errs.values())
    print("%-34s N=%-5d T=%-3d worst %.2e  %s" % (label, B * IW, T, worst, " ".join("%s=%.1e" % (k, v) for k, v in errs.items())))
    if worst > bar or not np.isfinite(worst):
        print("    per-state x_states errors:", " ".join("%.1e" % e for e in per_state))
        if "d_weights" in ref and errs["d_weights"] > bar:
            d = np.abs(got["d_weights"].astype(np.float64) - ref["d_weights"]) / (np.max(np.abs(ref["d_weights"])) + 1e-300)
            bad = np.nonzero(d > bar)[0]
            print("    d_weights: %d bad entries, first %s" % (len(bad), bad[:24]))
        for k in ("d_q_mu", "d_q_prec"):
            if errs[k] > bar:
                d = np.abs(got[k].astype(np.float64) - ref[k]) / (np.max(np.abs(ref[k])) + 1e-300)
                print("    %s bad columns:" % k, sorted(set(np.nonzero(d > bar)[1].tolist())))
    return worst <= bar


def main():
    case = load_case("dr_blackbox_icml_midpoint_f32_iw8")
    ok = True
    for solver, t, label in (("euler", 2, "euler T=2 (one RHS / one VJP)"), ("euler", 5, "euler T=5"), ("midpoint", 3, "midpoint T=3"),
                             ("midpoint", None, "midpoint full T"), ("rk4", 9, "rk4 T=9"), ("modeuler", 9, "modeuler T=9"),
                             ("modeulerwhile", 9, "modeulerwhile T=9")):
        c = sub_case(case, list(range(case["u"].shape[0])), None, t)
        c["solver"] = np.array(solver)
        ok &= compare(c, label)
    c = sub_case(case, [0, 1, 2, 3, 4], 7, 9)
    ok &= compare(c, "midpoint N=35 (ragged)")
    c = sub_case(case, [3], 1, 4)
    ok &= compare(c, "midpoint N=1")
    # bigger: tile to B=36 x IW=200 with fresh u
    rng = np.random.RandomState(0)
    B0 = case["u"].shape[0]
    idx = np.arange(36) % B0
    big = dict(case)
    for k in ("inputs", "dev_1hot", "observations", "q_mu", "q_prec"):
        big[k] = np.ascontiguousarray(case[k][idx])
    big["u"] = rng.randn(36, 200, case["u"].shape[2]).astype(np.float32)
    ok &= compare(big, "midpoint B=36 IW=200")
    if "--time" in sys.argv:
        import ctypes as C

        import torch

        from vihds_b200 import _lib as L  # noqa: F401

        for impl in ("scalar", "mma"):
            os.environ["VIHDS_BB_IMPL"] = impl
            run_case_on_gpu(big)
            torch.cuda.synchronize()
            import time

            t0 = time.perf_counter()
            for _ in range(5):
                run_case_on_gpu(big)
            torch.cuda.synchronize()
            print("impl %-6s: %.2f ms per fwd+iwae+bwd call incl. host set-up" % (impl, (time.perf_counter() - t0) / 5 * 1e3))
    print("ALL OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
