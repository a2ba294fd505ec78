"""GPU: the lane-split forward kernel (vh_lane.cuh, VIHDS_FWD_LANE=1) against the default forward kernels on the golden
cases of the dr_constant family, every forward output; then kernel-only timings of both at the icml size."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
from conftest import load_case  # noqa: E402
from test_gpu_parity import _rel, run_case_on_gpu  # noqa: E402

ok = True
for name in ("dr_constant_icml_midpoint_f32_iw8", "dr_constant_one_rk4_f32_iw5", "dr_constant_one_modeuler_f32_iw5",
             "dr_constant_v2_midpoint_f32_iw8", "dr_constant_icml_midpoint_f32_iw200"):
    case = load_case(name)
    os.environ["VIHDS_FWD_LANE"] = "0"
    ref = run_case_on_gpu(case)
    os.environ["VIHDS_FWD_LANE"] = "1"
    got = run_case_on_gpu(case)
    errs = {k: _rel(got[k], ref[k]) for k in ("theta", "x_states", "x_predict", "logp_by_species", "logp_theta", "logq_theta", "cost", "d_q_mu")}
    print(name, " ".join("%s=%.1e" % kv for kv in errs.items()))
    ok &= max(errs.values()) < 2e-5
print("LANE_OK" if ok else "LANE_MISMATCH")
